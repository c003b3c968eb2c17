"""Host-side mirror of the alignment half of the reference's ``codes/models/archs/TDAN_arch.py``:
``Align`` (TDAN_arch.py:17-72), the second in-repo consumer of
``ModulatedDeformConvPack(extra_offset_mask=True)``.

Same constructor, parameter names / state_dict keys and forward semantics.  The four deformable
packs run through this package's DCN operator (``rvsr_mdcn_pack_fwd`` at inference -- offset conv,
sigmoid, gather and contraction in one C-ABI call; ``rvsr_mdcn_fwd`` / ``rvsr_mdcn_bwd`` under autograd).
The per-frame Python loop of the reference (:55-70) is folded into the batch dimension: every layer runs
once on B*N images, the reference frame is broadcast, and the per-frame images are concatenated in the
reference's channel order.  ``Trunk`` / ``TDAN`` (the reconstruction half, plain convs + pixel shuffle)
stay with the reference."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import arch_util
from .dcn.deform_conv import ModulatedDeformConvPack as DCN


class Align(nn.Module):

    def __init__(self, channel=1, nf=64, nb=5, groups=8):
        super(Align, self).__init__()
        self.initial_conv = nn.Conv2d(channel, nf, 3, padding=1, bias=True)
        self.residual_layers = arch_util.make_layer(arch_util.ResidualBlock_noBN, nb)
        self.bottle_neck = nn.Conv2d(nf * 2, nf, 3, padding=1, bias=True)

        def dcn():
            return DCN(nf, nf, 3, stride=1, padding=1, dilation=1, deformable_groups=groups, extra_offset_mask=True)

        self.offset_conv_1 = nn.Conv2d(nf, nf, 3, padding=1, bias=True)
        self.deform_conv_1 = dcn()
        self.offset_conv_2 = nn.Conv2d(nf, nf, 3, padding=1, bias=True)
        self.deform_conv_2 = dcn()
        self.offset_conv_3 = nn.Conv2d(nf, nf, 3, padding=1, bias=True)
        self.deform_conv_3 = dcn()
        self.offset_conv = nn.Conv2d(nf, nf, 3, padding=1, bias=True)
        self.deform_conv = dcn()
        self.reconstruction = nn.Conv2d(nf, channel, 3, padding=1, bias=True)

    def forward(self, x):
        B, N, C, H, W = x.size()
        out = F.relu(self.initial_conv(x.reshape(-1, C, H, W)))
        out = self.residual_layers(out)                                   # [B*N, nf, H, W]
        nf = out.shape[1]
        nei = out
        ref = out.view(B, N, nf, H, W)[:, N // 2:N // 2 + 1].expand(B, N, nf, H, W).reshape(B * N, nf, H, W)
        fea = self.bottle_neck(torch.cat([ref, nei], dim=1))              # TDAN_arch.py:57-58
        fea = self.deform_conv_1([fea, self.offset_conv_1(fea)])          # :60-61
        fea = self.deform_conv_2([fea, self.offset_conv_2(fea)])          # :62-63
        fea = self.deform_conv_3([nei.contiguous(), self.offset_conv_3(fea)])  # :64-65 -- samples the neighbour features
        aligned = self.deform_conv([fea, self.offset_conv(fea)])          # :66-67
        im = self.reconstruction(aligned)                                 # [B*N, C, H, W]
        return im.view(B, N * C, H, W)                                    # cat over frames along channels (:70)

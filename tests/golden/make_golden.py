"""Generate the golden fixtures in tests/golden/*.npz by running the REFERENCE.

Run in the build container only (it needs /root/reference, which does not exist
on the GPU box):

    python tests/golden/make_golden.py

It imports the reference's own ``codes/models/archs/EDVR_arch.py`` and
``arch_util.py`` unmodified.  Two things the reference needs are absent here and
are stubbed, exactly as SURVEY.md Appendix B describes:
  * ``kornia`` -- imported at EDVR_arch.py:4 but never used;
  * ``models.archs.dcn.deform_conv`` -- the reference DCN is CUDA-only
    (deform_conv.py:109-110 raises on CPU tensors) and there is no GPU in this
    container, so the pack's final call is routed to
    ``torchvision.ops.deform_conv2d`` (same operator; pinned against
    oracle/dcn_oracle.c by tests/test_oracle.py).  The pack's own logic
    (conv_offset_mask, chunk, cat, sigmoid) is the reference's, restated in the shim.
The committed .npz files hold only seeds, configs and OUTPUT tensors; weights and
inputs are regenerated from the seeds by tests/golden/synth.py.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from synth import synth_input, synth_normal, synth_state_dict  # noqa: E402

REF = os.environ.get("RVSR_REFERENCE", "/root/reference")


def import_reference():
    import torchvision.ops as tvo
    from torch.nn.modules.utils import _pair

    sys.path.insert(0, os.path.join(REF, "codes"))
    sys.modules.setdefault("kornia", types.ModuleType("kornia"))

    class ModulatedDeformConvPack(nn.Module):
        # ctor / parameters as reference dcn/deform_conv.py:220-272
        def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                     groups=1, deformable_groups=1, bias=True, extra_offset_mask=False):
            super().__init__()
            self.stride, self.padding, self.dilation = stride, padding, dilation
            self.groups, self.deformable_groups = groups, deformable_groups
            self.extra_offset_mask = extra_offset_mask
            ks = _pair(kernel_size)
            self.weight = nn.Parameter(torch.zeros(out_channels, in_channels // groups, *ks))
            self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None
            self.conv_offset_mask = nn.Conv2d(in_channels, deformable_groups * 3 * ks[0] * ks[1],
                                              kernel_size=ks, stride=_pair(stride),
                                              padding=_pair(padding), bias=True)

        def forward(self, x):  # reference dcn/deform_conv.py:274-292
            if self.extra_offset_mask:
                out = self.conv_offset_mask(x[1])
                x = x[0]
            else:
                out = self.conv_offset_mask(x)
            o1, o2, mask = torch.chunk(out, 3, dim=1)
            offset = torch.cat((o1, o2), dim=1)
            mask = torch.sigmoid(mask)
            return tvo.deform_conv2d(x, offset, self.weight, self.bias, stride=self.stride,
                                     padding=self.padding, dilation=self.dilation, mask=mask)

    pkg = types.ModuleType("models.archs.dcn")
    pkg.__path__ = []
    mod = types.ModuleType("models.archs.dcn.deform_conv")
    mod.ModulatedDeformConvPack = ModulatedDeformConvPack
    sys.modules["models.archs.dcn"] = pkg
    sys.modules["models.archs.dcn.deform_conv"] = mod
    import models.archs.EDVR_arch as E  # the reference file itself
    return E, tvo


CASES = {
    # name: (class, ctor kwargs, input shape [B,N,C,H,W], weight seed, input seed)
    "edvr_tiny": ("EDVR", dict(nf=8, nc=3, nframes=5, groups=8, front_RBs=1, back_RBs=1,
                               w_TSA=True), (1, 5, 3, 32, 32), 11, 12),
    "edvr_tiny_b2_g2": ("EDVR", dict(nf=8, nc=3, nframes=3, groups=2, front_RBs=1, back_RBs=2,
                                     w_TSA=True), (2, 3, 3, 16, 24), 21, 22),
    # EDVR_NoUp hard-codes HRconv to 64 input channels (EDVR_arch.py:348), so nf must be 64
    "edvr_noup_3f": ("EDVR_NoUp", dict(nf=64, nc=3, nframes=3, groups=4, front_RBs=2, back_RBs=2,
                                       w_TSA=True), (1, 3, 3, 24, 32), 31, 32),
    "edvr_nf64_crop": ("EDVR", dict(nf=64, nc=3, nframes=5, groups=8, front_RBs=5, back_RBs=10,
                                    w_TSA=True), (1, 5, 3, 16, 24), 41, 42),
    "edvr_noup_nf64_ship": ("EDVR_NoUp", dict(nf=64, nc=3, nframes=3, groups=8, front_RBs=5,
                                              back_RBs=10, w_TSA=False), (1, 3, 3, 20, 16), 51, 52),
    "edvr_predeblur": ("EDVR", dict(nf=16, nc=3, nframes=3, groups=4, front_RBs=1, back_RBs=1,
                                    predeblur=True, w_TSA=True), (1, 3, 3, 16, 16), 61, 62),
    # HR_in=True (EDVR_arch.py:228-231, :267-274, :315-316): full-resolution input, two stride-2 convs in the stem,
    # output at the INPUT resolution, base = the centre frame itself; alone and together with predeblur (:15-59 HR_in stem)
    "edvr_hr_in": ("EDVR", dict(nf=16, nc=3, nframes=3, groups=4, front_RBs=1, back_RBs=1,
                                HR_in=True, w_TSA=True), (1, 3, 3, 32, 48), 81, 82),
    "edvr_predeblur_hr_in": ("EDVR", dict(nf=16, nc=3, nframes=3, groups=4, front_RBs=1, back_RBs=1,
                                          predeblur=True, HR_in=True, w_TSA=False), (1, 3, 3, 32, 32), 83, 84),
    # BASELINE cfg4's architecture (7 frames, 128 channels, 16 channels per deformable group) on a small crop
    "edvr_nf128_7f": ("EDVR", dict(nf=128, nc=3, nframes=7, groups=8, front_RBs=5, back_RBs=10,
                                   w_TSA=True), (1, 7, 3, 16, 24), 71, 72),
}


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    E, tvo = import_reference()
    only = sys.argv[1:]  # optional: regenerate just these cases
    for name, (cls, kw, shape, wseed, xseed) in CASES.items():
        if only and name not in only:
            continue
        net = getattr(E, cls)(**kw).eval()
        sd = synth_state_dict({k: v.shape for k, v in net.state_dict().items()}, wseed)
        net.load_state_dict(sd, strict=True)
        x = synth_input(shape, xseed)
        taps = {}
        def grab(mod, inp, outp, taps=taps):  # must return None: a value would replace the output
            taps.setdefault("aligned0", outp.detach().clone())

        hooks = [net.pcd_align.register_forward_hook(grab)]
        with torch.no_grad():
            y = net(x)
        for h in hooks:
            h.remove()
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"), cls=cls, kwargs=repr(kw), shape=np.array(shape),
            wseed=wseed, xseed=xseed, out=y.numpy(), aligned0=taps["aligned0"].numpy(),
            keys=np.array(list(sd.keys())), torch_version=torch.__version__)
        print(name, tuple(y.shape), float(y.abs().mean()), "aligned0", float(taps["aligned0"].abs().mean()))

    if not only or "edvr_tiny_grads" in only:
        # whole-network backward golden (BASELINE cfg5's step on a tiny architecture): the reference network in float64,
        # L1 loss against a seeded target, gradient of EVERY parameter (torchvision's deform_conv2d has autograd)
        kw = dict(nf=8, nc=3, nframes=5, groups=8, front_RBs=1, back_RBs=1, w_TSA=True)
        net = E.EDVR(**kw).double().train()
        sd = synth_state_dict({k: v.shape for k, v in net.state_dict().items()}, 91)
        net.load_state_dict(sd, strict=True)
        x = synth_input((2, 5, 3, 24, 24), 92).double()
        gt = synth_normal((2, 3, 96, 96), 93, std=0.3).double() + 0.5
        loss = torch.nn.functional.l1_loss(net(x), gt)
        loss.backward()
        grads = {"g:" + k: p.grad.float().numpy() for k, p in net.named_parameters()}
        np.savez_compressed(os.path.join(HERE, "edvr_tiny_grads.npz"), kwargs=repr(kw), wseed=91, xseed=92, gtseed=93,
                            shape=np.array([2, 5, 3, 24, 24]), loss=float(loss), **grads)
        print("edvr_tiny_grads loss", float(loss), len(grads), "gradients")

    if only and "dcn_unit" not in only:
        return
    # unit-level DCN fixture: large offsets (many taps leave the image), fp64 + grads
    B, C, H, W, Cout, dg = 2, 16, 11, 13, 12, 4
    x = synth_normal((B, C, H, W), 71).double().requires_grad_()
    off = synth_normal((B, dg * 18, H, W), 72, std=4.0).double().requires_grad_()
    msk = torch.sigmoid(synth_normal((B, dg * 9, H, W), 73).double()).requires_grad_()
    w = synth_normal((Cout, C, 3, 3), 74, std=0.1).double().requires_grad_()
    b = synth_normal((Cout,), 75).double().requires_grad_()
    go = synth_normal((B, Cout, H, W), 76).double()
    y = tvo.deform_conv2d(x, off, w, b, stride=1, padding=1, dilation=1, mask=msk)
    y.backward(go)
    np.savez_compressed(os.path.join(HERE, "dcn_unit.npz"), dims=np.array([B, C, H, W, Cout, dg]),
                        out=y.detach().numpy(), gx=x.grad.numpy(), goff=off.grad.numpy(),
                        gmask=msk.grad.numpy(), gw=w.grad.numpy(), gb=b.grad.numpy())
    print("dcn_unit", tuple(y.shape))


if __name__ == "__main__":
    main()

"""Host-side mirror of the reference's ``codes/models/archs/EDVR_arch.py``.

Same classes, constructor arguments, sub-module / parameter names (the state_dict contract:
reference test scripts load weights with strict=True, test_RealVSR_wi_GT.py:77) and forward
signatures:

    EDVR(nf, nc, nframes, groups, front_RBs, back_RBs, center, predeblur, HR_in, w_TSA)   ref :211-320
    EDVR_NoUp(...)                                                                          ref :323-404
    PCD_Align(nf, groups)  ref :62-132      TSA_Fusion(nf, nframes, center)  ref :135-208
    Predeblur_ResNet_Pyramid(nf, HR_in)  ref :15-59

Two execution paths, both on this package's own CUDA kernels for the deformable conv:

* engine path  -- inference (no autograd) on CUDA tensors: the whole forward is ONE call into
  the C++ engine (rvsr_engine_forward); weights are re-handed to the engine whenever a
  parameter changes.  This is the B200 hot path.
* module path  -- training / autograd: the same
  graph expressed with nn.Conv2d modules plus ``ModulatedDeformConvPack`` (our DCN operator
  with its own backward), structured like the reference so autograd works unchanged.

There is no CPU path (the reference's DCN is CUDA-only as well, deform_conv.py:109-110).
"""
import functools
import os
import threading

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import arch_util
from .dcn.deform_conv import ModulatedDeformConvPack as DCN
from .. import engine as _engine

_ENGINE_LOCK = threading.Lock()  # DataParallel runs the replicas' forwards in threads


def _conv3(cin, cout, stride=1):
    return nn.Conv2d(cin, cout, 3, stride, 1, bias=True)


def _up2(t):
    """F.interpolate(t, scale_factor=2, mode='bilinear', align_corners=False) (EDVR_arch.py:53-57, :109-121, :195-202)"""
    from .. import ops
    return ops.upsample2x(t)


class Predeblur_ResNet_Pyramid(nn.Module):
    def __init__(self, nf=128, HR_in=False):
        super(Predeblur_ResNet_Pyramid, self).__init__()
        self.HR_in = bool(HR_in)
        if self.HR_in:
            self.conv_first_1 = _conv3(3, nf)
            self.conv_first_2 = _conv3(nf, nf, 2)
            self.conv_first_3 = _conv3(nf, nf, 2)
        else:
            self.conv_first = _conv3(3, nf)
        for name in ('RB_L1_1', 'RB_L1_2', 'RB_L1_3', 'RB_L1_4', 'RB_L1_5', 'RB_L2_1', 'RB_L2_2', 'RB_L3_1'):
            setattr(self, name, arch_util.ResidualBlock_noBN(nf=nf))
        self.deblur_L2_conv = _conv3(nf, nf, 2)
        self.deblur_L3_conv = _conv3(nf, nf, 2)
        self.lrelu = nn.LeakyReLU(negative_slope=0.1, inplace=True)

    def forward(self, x):
        act = self.lrelu
        if self.HR_in:
            l1 = act(self.conv_first_3(act(self.conv_first_2(act(self.conv_first_1(x))))))
        else:
            l1 = act(self.conv_first(x))
        l2 = act(self.deblur_L2_conv(l1))
        l3 = act(self.deblur_L3_conv(l2))
        l2 = self.RB_L2_1(l2) + _up2(self.RB_L3_1(l3))
        l1 = self.RB_L1_2(self.RB_L1_1(l1)) + _up2(self.RB_L2_2(l2))
        return self.RB_L1_5(self.RB_L1_4(self.RB_L1_3(l1)))


class PCD_Align(nn.Module):
    """Pyramid (3 levels), Cascading, Deformable alignment of one neighbour frame to the reference frame."""

    def __init__(self, nf=64, groups=8):
        super(PCD_Align, self).__init__()
        dcn = functools.partial(DCN, nf, nf, 3, stride=1, padding=1, dilation=1, deformable_groups=groups,
                                extra_offset_mask=True)
        # registration order = reference order (state_dict key order)
        self.L3_offset_conv1 = _conv3(nf * 2, nf)
        self.L3_offset_conv2 = _conv3(nf, nf)
        self.L3_dcnpack = dcn()
        self.L2_offset_conv1 = _conv3(nf * 2, nf)
        self.L2_offset_conv2 = _conv3(nf * 2, nf)
        self.L2_offset_conv3 = _conv3(nf, nf)
        self.L2_dcnpack = dcn()
        self.L2_fea_conv = _conv3(nf * 2, nf)
        self.L1_offset_conv1 = _conv3(nf * 2, nf)
        self.L1_offset_conv2 = _conv3(nf * 2, nf)
        self.L1_offset_conv3 = _conv3(nf, nf)
        self.L1_dcnpack = dcn()
        self.L1_fea_conv = _conv3(nf * 2, nf)
        self.cas_offset_conv1 = _conv3(nf * 2, nf)
        self.cas_offset_conv2 = _conv3(nf, nf)
        self.cas_dcnpack = dcn()
        self.lrelu = nn.LeakyReLU(negative_slope=0.1, inplace=True)

    def forward(self, nbr_fea_l, ref_fea_l):
        """nbr_fea_l, ref_fea_l: [L1, L2, L3] features, each [B, C, H, W] -> aligned L1 feature."""
        act, cat = self.lrelu, torch.cat
        off3 = act(self.L3_offset_conv2(act(self.L3_offset_conv1(cat([nbr_fea_l[2], ref_fea_l[2]], 1)))))
        fea3 = act(self.L3_dcnpack([nbr_fea_l[2], off3]))
        off2 = act(self.L2_offset_conv1(cat([nbr_fea_l[1], ref_fea_l[1]], 1)))
        off2 = act(self.L2_offset_conv2(cat([off2, _up2(off3) * 2], 1)))  # offsets double with resolution
        off2 = act(self.L2_offset_conv3(off2))
        fea2 = self.L2_dcnpack([nbr_fea_l[1], off2])
        fea2 = act(self.L2_fea_conv(cat([fea2, _up2(fea3)], 1)))
        off1 = act(self.L1_offset_conv1(cat([nbr_fea_l[0], ref_fea_l[0]], 1)))
        off1 = act(self.L1_offset_conv2(cat([off1, _up2(off2) * 2], 1)))
        off1 = act(self.L1_offset_conv3(off1))
        fea1 = self.L1_dcnpack([nbr_fea_l[0], off1])
        fea1 = self.L1_fea_conv(cat([fea1, _up2(fea2)], 1))  # no activation here (reference :125)
        offc = act(self.cas_offset_conv2(act(self.cas_offset_conv1(cat([fea1, ref_fea_l[0]], 1)))))
        return act(self.cas_dcnpack([fea1, offc]))


class TSA_Fusion(nn.Module):
    """Temporal (correlation with the centre frame, sigmoid) and spatial (3-level pyramid) attention fusion."""

    def __init__(self, nf=64, nframes=5, center=2):
        super(TSA_Fusion, self).__init__()
        self.center = center
        c1 = lambda cin: nn.Conv2d(cin, nf, 1, 1, bias=True)  # noqa: E731
        self.tAtt_1 = _conv3(nf, nf)
        self.tAtt_2 = _conv3(nf, nf)
        self.fea_fusion = c1(nframes * nf)
        self.sAtt_1 = c1(nframes * nf)
        self.maxpool = nn.MaxPool2d(3, stride=2, padding=1)
        self.avgpool = nn.AvgPool2d(3, stride=2, padding=1)
        self.sAtt_2 = c1(nf * 2)
        self.sAtt_3 = _conv3(nf, nf)
        self.sAtt_4 = c1(nf)
        self.sAtt_5 = _conv3(nf, nf)
        self.sAtt_L1 = c1(nf)
        self.sAtt_L2 = _conv3(nf * 2, nf)
        self.sAtt_L3 = _conv3(nf, nf)
        self.sAtt_add_1 = c1(nf)
        self.sAtt_add_2 = c1(nf)
        self.lrelu = nn.LeakyReLU(negative_slope=0.1, inplace=True)

    def forward(self, aligned_fea):
        B, N, C, H, W = aligned_fea.size()
        act = self.lrelu
        emb_ref = self.tAtt_2(aligned_fea[:, self.center].clone())
        emb = self.tAtt_1(aligned_fea.reshape(-1, C, H, W)).view(B, N, -1, H, W)
        prob = torch.sigmoid((emb * emb_ref.unsqueeze(1)).sum(2, keepdim=True))  # [B, N, 1, H, W]
        weighted = (aligned_fea * prob).reshape(B, N * C, H, W)
        fea = act(self.fea_fusion(weighted))
        att = act(self.sAtt_1(weighted))
        att = act(self.sAtt_2(torch.cat([self.maxpool(att), self.avgpool(att)], 1)))
        att_l = act(self.sAtt_L1(att))
        att_l = act(self.sAtt_L2(torch.cat([self.maxpool(att_l), self.avgpool(att_l)], 1)))
        att_l = _up2(act(self.sAtt_L3(att_l)))
        att = act(self.sAtt_3(att)) + att_l
        att = self.sAtt_5(_up2(act(self.sAtt_4(att))))
        att_add = self.sAtt_add_2(act(self.sAtt_add_1(att)))
        return fea * torch.sigmoid(att) * 2 + att_add


class _EDVRBase(nn.Module):
    """Shared body of EDVR (x4 pixel-shuffle tail) and EDVR_NoUp (same-resolution tail)."""
    _upsample = True

    def _build(self, nf, nc, nframes, groups, front_RBs, back_RBs, center, predeblur, HR_in, w_TSA):
        self.nf, self.nc = nf, nc
        self.center = nframes // 2 if center is None else center
        self.is_predeblur = bool(predeblur)
        self.HR_in = bool(HR_in)
        self.w_TSA = w_TSA
        self._cfg = dict(nf=nf, nc=nc, nframes=nframes, groups=groups, front_RBs=front_RBs, back_RBs=back_RBs,
                         center=self.center, predeblur=self.is_predeblur, HR_in=self.HR_in, w_TSA=bool(w_TSA),
                         upsample=self._upsample)
        rb = functools.partial(arch_util.ResidualBlock_noBN, nf=nf)
        if self._upsample and self.is_predeblur:
            self.pre_deblur = Predeblur_ResNet_Pyramid(nf=nf, HR_in=self.HR_in)
            self.conv_1x1 = nn.Conv2d(nf, nf, 1, 1, bias=True)
        elif self._upsample and self.HR_in:
            self.conv_first_1 = _conv3(nc, nf)
            self.conv_first_2 = _conv3(nf, nf, 2)
            self.conv_first_3 = _conv3(nf, nf, 2)
        else:
            self.conv_first = _conv3(nc, nf)
        self.feature_extraction = arch_util.make_layer(rb, front_RBs)
        self.fea_L2_conv1 = _conv3(nf, nf, 2)
        self.fea_L2_conv2 = _conv3(nf, nf)
        self.fea_L3_conv1 = _conv3(nf, nf, 2)
        self.fea_L3_conv2 = _conv3(nf, nf)
        self.pcd_align = PCD_Align(nf=nf, groups=groups)
        if self.w_TSA:
            self.tsa_fusion = TSA_Fusion(nf=nf, nframes=nframes, center=self.center)
        else:
            self.tsa_fusion = nn.Conv2d(nframes * nf, nf, 1, 1, bias=True)
        self.recon_trunk = arch_util.make_layer(rb, back_RBs)
        if self._upsample:
            self.upconv1 = _conv3(nf, nf * 4)
            self.upconv2 = _conv3(nf, 64 * 4)  # the reference hard-codes 64 from here on (:250-253)
            self.pixel_shuffle = nn.PixelShuffle(2)
        self.HRconv = _conv3(64, 64)
        self.conv_last = _conv3(64, nc)
        self.lrelu = nn.LeakyReLU(negative_slope=0.1, inplace=True)
        # 'auto': engine for inference on CUDA, modules when autograd is needed.
        # 'module' forces the nn.Module graph; 'engine' raises instead of falling back.
        self.exec_path = os.environ.get("RVSR_EXEC_PATH", "auto")
        # engine arithmetic: None = follow the input dtype (fp32 in -> fp32 SIMT kernels,
        # fp16 in -> tensor-core kernels); 'fp16' runs fp32 inputs through the fp16 engine.
        self.engine_precision = os.environ.get("RVSR_ENGINE_PRECISION") or None
        # how a weight change is detected before an engine forward: 'stamp' (data_ptr + version of every weight),
        # 'checksum' (stamp + a device-side sum, catches p.data.xxx_() edits) or 'always' (see invalidate_engine)
        self.engine_weight_check = os.environ.get("RVSR_WEIGHT_CHECK", "stamp")
        # (device index, precision) -> [EDVREngine, (weight stamp, checksum)].  Shared BY REFERENCE with nn.DataParallel
        # replicas (replicate() shallow-copies __dict__): replica d finds the engine of device d built by an earlier
        # iteration instead of creating one per forward.
        self.__dict__['_engines'] = {}

    # ------------------------------------------------------------------ engine path
    def _named_weights(self):
        """(name, tensor) of every weight, in state_dict order.  Works on nn.DataParallel replicas too: torch's
        replicate() leaves a replica's ``_parameters`` empty and sets the broadcast copies as plain attributes
        (kept in ``_former_parameters``), so ``parameters()`` / ``state_dict()`` are empty there."""
        for prefix, m in self.named_modules():
            former = m.__dict__.get('_former_parameters') or {}
            for k, v in m._parameters.items():
                t = v if v is not None else former.get(k)
                if t is not None:
                    yield (prefix + '.' + k if prefix else k), t
            for k, v in former.items():
                if k not in m._parameters and v is not None:
                    yield (prefix + '.' + k if prefix else k), v

    def _on_replica(self):
        return bool(self.__dict__.get('_is_replica', False)) or '_former_parameters' in self.__dict__

    def _engine_ok(self, x):
        if self.exec_path == "module":
            return False
        # the grad check looks at the actual weight tensors (a DataParallel replica has no parameters())
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or any(t.requires_grad for _, t in self._named_weights()))
        ok = (x.is_cuda and self.nf % 8 == 0 and x.dtype in (torch.float32, torch.float16) and not needs_grad)
        if not ok and self.exec_path == "engine":
            raise RuntimeError("realvsr_b200: exec_path='engine' but this call needs the module path "
                               "(autograd enabled, non-CUDA input, or nf not a multiple of 8)")
        return ok

    def _weight_stamp(self):
        return tuple((t.data_ptr(), t._version) for _, t in self._named_weights())

    def invalidate_engine(self):
        """Force the engine to re-read the weights at the next forward.  The automatic check compares
        (data_ptr, version) of every weight, which sees ``load_state_dict``, ``.to()``, optimizer steps and any
        in-place op on the parameter -- but NOT in-place writes through ``p.data`` (``p.data.mul_(s)``,
        ``p.data.copy_(ema)``): torch does not bump ``p._version`` for those.  Call this after such an edit, or set
        ``engine_weight_check = 'always'`` (re-hand the weights on every call, a few ms) /
        ``'checksum'`` (a device-side sum of all weights is compared on every call: one small D2H sync)."""
        for slot in self._engines.values():
            slot[1] = None

    def _weights_changed(self, slot, stamp):
        mode = self.engine_weight_check
        if mode == "always" or self._on_replica():
            return True   # replicas: broadcast copies are fresh tensors every iteration; a recycled address proves nothing
        if slot[1] is None or slot[1][0] != stamp:
            return True
        if mode == "checksum":
            return slot[1][1] != self._weight_checksum()
        return False

    def _weight_checksum(self):
        ws = [t.detach() for _, t in self._named_weights()]
        return float(torch.stack(torch._foreach_norm([w.float() for w in ws])).double().sum())

    def _get_engine(self, x):
        prec = self.engine_precision or ("fp16" if x.dtype == torch.float16 else "fp32")
        key = (x.device.index, prec)
        with _ENGINE_LOCK:
            slot = self._engines.get(key)
            if slot is None:
                slot = [_engine.EDVREngine(precision=prec, device=x.device, **self._cfg), None]
                self._engines[key] = slot
        stamp = self._weight_stamp()
        if self._weights_changed(slot, stamp):
            slot[0].load_state_dict({k: t for k, t in self._named_weights()}, strict=True)
            slot[1] = (stamp, self._weight_checksum() if self.engine_weight_check == "checksum" else None)
        return slot[0]

    def load_state_dict(self, *args, **kwargs):
        r = super(_EDVRBase, self).load_state_dict(*args, **kwargs)
        self.invalidate_engine()
        return r

    def _apply(self, fn, *args, **kwargs):
        r = super(_EDVRBase, self)._apply(fn, *args, **kwargs)
        self.invalidate_engine()
        return r

    def __getstate__(self):
        # the engine cache holds ctypes handles (not picklable / deep-copyable): a copy starts with an empty cache
        d = self.__dict__.copy()
        d['_engines'] = {}
        return d

    def forward(self, x):
        if self._engine_ok(x):
            return self._get_engine(x)(x)
        if self._train_c8_ok(x):
            return self._forward_c8(x)
        return self._forward_modules(x)

    # ------------------------------------------------------------------ bf16 training path (channel-blocked tensors)
    def _train_c8_ok(self, x):
        """bf16 autograd on this library's convolution kernels (realvsr_b200/train_c8.py) instead of cuDNN: taken under
        torch.autocast(bfloat16) -- BASELINE cfg5's configuration -- or with exec_path = 'train_c8'.  Gradients are bf16
        quality (like autocast's); the fp32 module path stays the default for fp32 training."""
        want = self.exec_path == "train_c8" or (self.exec_path == "auto" and torch.is_autocast_enabled() and
                                                torch.get_autocast_dtype("cuda") == torch.bfloat16 and
                                                os.environ.get("RVSR_TRAIN_C8", "1") != "0")
        # the frames themselves get no gradient on this path (train_c8.conv_first): an input that requires one takes the module path
        ok = (x.is_cuda and self.nf == 64 and not x.requires_grad and not (self._upsample and (self.is_predeblur or self.HR_in)))
        if want and not ok and self.exec_path == "train_c8":
            raise RuntimeError("realvsr_b200: exec_path='train_c8' needs CUDA input without requires_grad, nf == 64 and the standard stem")
        return want and ok

    def _forward_c8(self, x):
        """Same graph as _forward_modules.  Every 64-channel convolution (3x3 and 1x1; stride 2 as stride 1 + subsampling), the
        residual adds, torch.cat, the x2 upsamples and PixelShuffle + lrelu run as train_c8 Functions on [N, C/8, H, W, 8] bf16
        tensors, and so does the DCN operator (train_c8.dcn_pack); TSA's pools / sigmoids / products are torch ops on the same
        tensors.  conv_first packs the NCHW frames itself (train_c8.conv_first); the x4 bilinear base is torch's."""
        from .. import train_c8 as T
        B, N, C, H, W = x.size()
        bf = torch.bfloat16
        conv = lambda m, t, act=None, residual=None, shuffle=False: T.conv(t, m.weight, m.bias, act=act, residual=residual,  # noqa: E731
                                                                           shuffle=shuffle)
        pair = lambda m1, m2, t: T.conv_pair(t, m1.weight, m1.bias, "lrelu", m2.weight, m2.bias, "lrelu")  # noqa: E731

        def trunk(blocks, t):
            for blk in blocks:  # ResidualBlock_noBN (arch_util.py:135-139): x + conv2(relu(conv1(x)))
                t = T.conv_pair(t, blk.conv1.weight, blk.conv1.bias, "relu", blk.conv2.weight, blk.conv2.bias, None, skip=True)
            return t

        def dcn(pack, t, feat, act=None):
            # ModulatedDeformConvPack.forward (deform_conv.py:274-292): the 64 -> 216 offset / mask convolution with its weights
            # zero-padded to 256 outputs (whole 64-wide tiles for the data gradient), then the operator on the same C8 tensors
            # (rvsr_c8_mdcn_fwd / _bwd: dcn_tc_kernel, dcn_bwd_tc_kernel)
            k = pack.conv_offset_mask.weight.shape[0]
            w_om = F.pad(pack.conv_offset_mask.weight, (0, 0, 0, 0, 0, 0, 0, 256 - k))
            b_om = F.pad(pack.conv_offset_mask.bias, (0, 256 - k))
            return T.dcn_pack(t, T.conv(feat, w_om, b_om), pack.weight, pack.bias, "lrelu" if act else None)

        def down2(m, t):
            # stride-2 3x3 convolution = the stride-1 convolution at the even pixels (pad 1 both ways): 4x the MMAs of two small
            # layers instead of two layout round trips through cuDNN
            return conv(m, t, "lrelu")[:, :, ::2, ::2].contiguous()

        def tsa(m, aligned):
            # TSA_Fusion.forward (EDVR_arch.py:168-208) on C8 tensors: convolutions as above, temporal attention / pools / final
            # modulation as train_c8 Functions, the one remaining add is torch's
            a6 = aligned.view(B, N, *aligned.shape[1:])                                   # [B, N, 8, H, W, 8]
            emb_ref = conv(m.tAtt_2, a6[:, self.center].contiguous())                     # [B, 8, H, W, 8]
            srcs = T.tsa_temporal(aligned, conv(m.tAtt_1, aligned), emb_ref, N)           # the N x 64 channels of the 1x1 fusions
            fea = conv(m.fea_fusion, srcs, "lrelu")
            att = conv(m.sAtt_1, srcs, "lrelu")
            att = conv(m.sAtt_2, T.pool_maxavg(att), "lrelu")
            att_l = conv(m.sAtt_L1, att, "lrelu")
            att_l = conv(m.sAtt_L2, T.pool_maxavg(att_l), "lrelu")
            att_l = T.upsample2x(conv(m.sAtt_L3, att_l, "lrelu"))
            att = conv(m.sAtt_3, att, "lrelu") + att_l
            att = conv(m.sAtt_5, T.upsample2x(conv(m.sAtt_4, att, "lrelu")))
            att_add = conv(m.sAtt_add_2, conv(m.sAtt_add_1, att, "lrelu"))
            return T.tsa_final(fea, att, att_add)

        with torch.autocast("cuda", dtype=bf):
            x_center = x[:, self.center].contiguous()
            frames = x.reshape(-1, C, H, W)
            l1 = trunk(self.feature_extraction, T.conv_first(frames, self.conv_first.weight, self.conv_first.bias, "lrelu"))
            l2 = conv(self.fea_L2_conv2, down2(self.fea_L2_conv1, l1), "lrelu")
            l3 = conv(self.fea_L3_conv2, down2(self.fea_L3_conv1, l2), "lrelu")
            pyr = [l1, l2, l3]
            ref = [lv.view(B, N, *lv.shape[1:])[:, self.center:self.center + 1].expand(B, N, *lv.shape[1:]).reshape(lv.shape)
                   for lv in pyr]
            p = self.pcd_align  # PCD_Align.forward (EDVR_arch.py:98-132), all N frames as one batch
            off3 = pair(p.L3_offset_conv1, p.L3_offset_conv2, [pyr[2], ref[2]])
            fea3 = dcn(p.L3_dcnpack, pyr[2], off3, act=True)
            off2 = conv(p.L2_offset_conv1, [pyr[1], ref[1]], "lrelu")
            off2 = pair(p.L2_offset_conv2, p.L2_offset_conv3, [off2, T.upsample2x(off3, 2.0)])
            fea2 = conv(p.L2_fea_conv, [dcn(p.L2_dcnpack, pyr[1], off2), T.upsample2x(fea3)], "lrelu")
            off1 = conv(p.L1_offset_conv1, [pyr[0], ref[0]], "lrelu")
            off1 = pair(p.L1_offset_conv2, p.L1_offset_conv3, [off1, T.upsample2x(off2, 2.0)])
            fea1 = conv(p.L1_fea_conv, [dcn(p.L1_dcnpack, pyr[0], off1), T.upsample2x(fea2)])
            offc = pair(p.cas_offset_conv1, p.cas_offset_conv2, [fea1, ref[0]])
            aligned = dcn(p.cas_dcnpack, fea1, offc, act=True)
            if self.w_TSA:
                fea = tsa(self.tsa_fusion, aligned)
            else:
                a6 = aligned.view(B, N, *aligned.shape[1:])
                fea = conv(self.tsa_fusion, [a6[:, i].contiguous() for i in range(N)])
            out = trunk(self.recon_trunk, fea)
            if self._upsample:
                out = conv(self.upconv1, out, "lrelu", shuffle=True)   # lrelu(PixelShuffle(conv)) == PixelShuffle(lrelu(conv))
                out = conv(self.upconv2, out, "lrelu", shuffle=True)
            # conv_last (64 -> nc <= 8): the output padded to 16 channels (the narrowest tile of the tcgen05 kernels)
            w_last = F.pad(self.conv_last.weight, (0, 0, 0, 0, 0, 0, 0, 16 - self.nc))
            b_last = F.pad(self.conv_last.bias, (0, 16 - self.nc))
            out = T.from_c8(T.conv(conv(self.HRconv, out, "lrelu"), w_last, b_last), self.nc, torch.float32)
        if self._upsample:
            base = F.interpolate(x_center, scale_factor=4, mode='bilinear', align_corners=False)
        else:
            base = x_center
        return out + base

    # ------------------------------------------------------------------ module path (autograd-capable)
    def _forward_modules(self, x):
        B, N, C, H, W = x.size()
        act = self.lrelu
        x_center = x[:, self.center].contiguous()
        frames = x.reshape(-1, C, H, W)
        if self._upsample and self.is_predeblur:
            l1 = self.conv_1x1(self.pre_deblur(frames))
            if self.HR_in:
                H, W = H // 4, W // 4
        elif self._upsample and self.HR_in:
            l1 = act(self.conv_first_3(act(self.conv_first_2(act(self.conv_first_1(frames))))))
            H, W = H // 4, W // 4
        else:
            l1 = act(self.conv_first(frames))
        l1 = self.feature_extraction(l1)
        l2 = act(self.fea_L2_conv2(act(self.fea_L2_conv1(l1))))
        l3 = act(self.fea_L3_conv2(act(self.fea_L3_conv1(l2))))
        # The reference aligns the N frames one by one in a Python loop (EDVR_arch.py:297-303); PCD_Align shares its weights
        # over the frames, so here the loop is folded into the batch: [B*N] neighbours against the centre frame's features
        # repeated N times -- one set of (larger) kernel launches, identical arithmetic per frame, same autograd graph shape.
        pyr = [l1, l2, l3]
        ref = [lv.view(B, N, *lv.shape[1:])[:, self.center:self.center + 1].expand(B, N, *lv.shape[1:]).reshape(lv.shape)
               for lv in pyr]
        aligned = self.pcd_align(pyr, ref).view(B, N, -1, H, W)
        fea = self.tsa_fusion(aligned if self.w_TSA else aligned.view(B, -1, H, W))
        out = self.recon_trunk(fea)
        if self._upsample:
            out = act(self.pixel_shuffle(self.upconv1(out)))
            out = act(self.pixel_shuffle(self.upconv2(out)))
        out = self.conv_last(act(self.HRconv(out)))
        if self._upsample and not self.HR_in:
            base = F.interpolate(x_center, scale_factor=4, mode='bilinear', align_corners=False)
        else:
            base = x_center
        return out + base


class EDVR(_EDVRBase):
    _upsample = True

    def __init__(self, nf=64, nc=3, nframes=5, groups=8, front_RBs=5, back_RBs=10, center=None, predeblur=False,
                 HR_in=False, w_TSA=True):
        super(EDVR, self).__init__()
        self._build(nf, nc, nframes, groups, front_RBs, back_RBs, center, predeblur, HR_in, w_TSA)


class EDVR_NoUp(_EDVRBase):
    """The variant RealVSR ships (scale 1): no pixel-shuffle tail, output = conv_last + centre frame.
    As in the reference, predeblur / HR_in are accepted and ignored by this class (:335-339, :358-404)."""
    _upsample = False

    def __init__(self, nf=64, nc=3, nframes=5, groups=8, front_RBs=5, back_RBs=10, center=None, predeblur=False,
                 HR_in=False, w_TSA=True):
        super(EDVR_NoUp, self).__init__()
        self._build(nf, nc, nframes, groups, front_RBs, back_RBs, center, False, False, w_TSA)
        self.is_predeblur, self.HR_in = bool(predeblur), bool(HR_in)  # stored like the reference, unused
        self._cfg.update(predeblur=False, HR_in=False)

    def _engine_ok(self, x):
        keep = self.is_predeblur, self.HR_in
        self.is_predeblur = self.HR_in = False
        try:
            return super(EDVR_NoUp, self)._engine_ok(x)
        finally:
            self.is_predeblur, self.HR_in = keep

    def _forward_modules(self, x):
        keep = self.is_predeblur, self.HR_in
        self.is_predeblur = self.HR_in = False
        try:
            return super(EDVR_NoUp, self)._forward_modules(x)
        finally:
            self.is_predeblur, self.HR_in = keep

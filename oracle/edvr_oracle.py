"""oracle/edvr_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference's EDVR hot path (feature pyramid -> PCD
alignment -> TSA fusion -> reconstruction), written functionally over a
``state_dict`` so it shares no module code with the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it.

What it follows in the reference (/root/reference/codes/models/archs):
  * ``EDVR.forward``            EDVR_arch.py:258-320
  * ``EDVR_NoUp.forward``       EDVR_arch.py:358-404
  * ``PCD_Align.forward``       EDVR_arch.py:98-132
  * ``TSA_Fusion.forward``      EDVR_arch.py:168-208
  * ``Predeblur_ResNet_Pyramid``EDVR_arch.py:43-59
  * ``ResidualBlock_noBN``      arch_util.py:135-139
  * ``ModulatedDeformConvPack`` dcn/deform_conv.py:274-292  (chunk / cat / sigmoid)
  * the DCNv2 op itself         dcn/src/deform_conv_cuda_kernel.cu:467-633 and
                                dcn/src/deform_conv_cuda.cpp:539-568, restated in
                                plain C in ``oracle/dcn_oracle.c`` (called here via ctypes).

Third-party arithmetic: every plain conv / pool / interpolate / pixel_shuffle in
the reference is a PyTorch ATen call (requirements.txt leaves torch unpinned);
the oracle uses the same ATen CPU ops from the torch in this image
(2.11.0) -- that *is* the reference's arithmetic for those layers.

Pinning: tests/test_oracle.py compares this file against golden tensors
produced by importing the reference's own EDVR_arch.py unmodified
(tests/golden/make_golden.py; the reference DCN is CUDA-only, so that script
routes its DCN call to torchvision.ops.deform_conv2d, which test_oracle.py in
turn pins against dcn_oracle.c).  The reference ships no tests or golden
vectors of its own.
"""
import ctypes
import os
import subprocess

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libdcn_oracle.so")
_lib = None


def build(force=False):
    """Compile oracle/dcn_oracle.c with gcc into oracle/libdcn_oracle.so."""
    src = os.path.join(_HERE, "dcn_oracle.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)):
        return _LIB_PATH
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=c99", "-fopenmp", "-o", _LIB_PATH, src, "-lm"]
    subprocess.run(cmd, check=True)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def dcn_forward(x, offset, mask, weight, bias=None, stride=1, padding=1, dilation=1, groups=1,
                deformable_groups=1):
    """DCNv2 forward on CPU through dcn_oracle.c.  fp32 or fp64 NCHW tensors."""
    lib = _load()
    assert x.dtype in (torch.float32, torch.float64)
    dt = x.dtype
    x, offset, mask, weight = [t.detach().to(dt).contiguous() for t in (x, offset, mask, weight)]
    bias = bias.detach().to(dt).contiguous() if bias is not None else None
    B, C, H, W = x.shape
    Cout, _, kh, kw = weight.shape
    Ho = (H + 2 * padding - (dilation * (kh - 1) + 1)) // stride + 1
    Wo = (W + 2 * padding - (dilation * (kw - 1) + 1)) // stride + 1
    out = torch.empty(B, Cout, Ho, Wo, dtype=dt)
    fn = lib.dcn_oracle_fwd_f64 if dt == torch.float64 else lib.dcn_oracle_fwd_f32
    rc = fn(_ptr(x), _ptr(offset), _ptr(mask), _ptr(weight), _ptr(bias), _ptr(out), B, C, H, W,
            Cout, kh, kw, stride, padding, dilation, groups, deformable_groups)
    if rc != 0:
        raise RuntimeError("dcn_oracle_fwd failed rc=%d" % rc)
    return out


def dcn_backward(x, offset, mask, weight, grad_out, stride=1, padding=1, dilation=1, groups=1,
                 deformable_groups=1, with_bias=True):
    """DCNv2 gradients on CPU. Returns (gx, goffset, gmask, gweight, gbias)."""
    lib = _load()
    dt = x.dtype
    x, offset, mask, weight, grad_out = [t.detach().to(dt).contiguous()
                                         for t in (x, offset, mask, weight, grad_out)]
    B, C, H, W = x.shape
    Cout, _, kh, kw = weight.shape
    gx = torch.zeros_like(x)
    go = torch.zeros_like(offset)
    gm = torch.zeros_like(mask)
    gw = torch.zeros_like(weight)
    gb = torch.zeros(Cout, dtype=dt) if with_bias else None
    fn = lib.dcn_oracle_bwd_f64 if dt == torch.float64 else lib.dcn_oracle_bwd_f32
    rc = fn(_ptr(x), _ptr(offset), _ptr(mask), _ptr(weight), _ptr(grad_out), _ptr(gx), _ptr(go),
            _ptr(gm), _ptr(gw), _ptr(gb), B, C, H, W, Cout, kh, kw, stride, padding, dilation,
            groups, deformable_groups)
    if rc != 0:
        raise RuntimeError("dcn_oracle_bwd failed rc=%d" % rc)
    return gx, go, gm, gw, gb


# --------------------------------------------------------------------------- network


def _conv(sd, name, x, stride=1, padding=None):
    w = sd[name + ".weight"]
    if padding is None:
        padding = w.shape[-1] // 2
    return F.conv2d(x, w, sd[name + ".bias"], stride=stride, padding=padding)


def _lrelu(x):
    return F.leaky_relu(x, 0.1)


def _up2(x):
    return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)


def resblock(sd, prefix, x):
    """arch_util.py:135-139: x + conv2(relu(conv1(x)))."""
    return x + _conv(sd, prefix + ".conv2", F.relu(_conv(sd, prefix + ".conv1", x)))


def dcn_pack(sd, prefix, x, feat, groups, dcn=dcn_forward):
    """deform_conv.py:274-292 with extra_offset_mask=True: offsets and mask are
    predicted from ``feat``; offset = first 2/3 of the channels, mask = sigmoid(last 1/3)."""
    om = _conv(sd, prefix + ".conv_offset_mask", feat)
    n = om.shape[1] // 3
    offset, mask = om[:, :2 * n], torch.sigmoid(om[:, 2 * n:])
    return dcn(x, offset.contiguous(), mask.contiguous(), sd[prefix + ".weight"],
               sd[prefix + ".bias"], 1, 1, 1, 1, groups)


def pcd_align(sd, nbr, ref, groups, p="pcd_align", dcn=dcn_forward, taps=None):
    """EDVR_arch.py:98-132.  nbr/ref = [L1, L2, L3] feature lists."""
    cat = torch.cat
    L3_off = _lrelu(_conv(sd, p + ".L3_offset_conv1", cat([nbr[2], ref[2]], 1)))
    L3_off = _lrelu(_conv(sd, p + ".L3_offset_conv2", L3_off))
    L3_fea = _lrelu(dcn_pack(sd, p + ".L3_dcnpack", nbr[2], L3_off, groups, dcn))
    L2_off = _lrelu(_conv(sd, p + ".L2_offset_conv1", cat([nbr[1], ref[1]], 1)))
    L2_off = _lrelu(_conv(sd, p + ".L2_offset_conv2", cat([L2_off, _up2(L3_off) * 2], 1)))
    L2_off = _lrelu(_conv(sd, p + ".L2_offset_conv3", L2_off))
    L2_fea = dcn_pack(sd, p + ".L2_dcnpack", nbr[1], L2_off, groups, dcn)
    L2_fea = _lrelu(_conv(sd, p + ".L2_fea_conv", cat([L2_fea, _up2(L3_fea)], 1)))
    L1_off = _lrelu(_conv(sd, p + ".L1_offset_conv1", cat([nbr[0], ref[0]], 1)))
    L1_off = _lrelu(_conv(sd, p + ".L1_offset_conv2", cat([L1_off, _up2(L2_off) * 2], 1)))
    L1_off = _lrelu(_conv(sd, p + ".L1_offset_conv3", L1_off))
    L1_fea = dcn_pack(sd, p + ".L1_dcnpack", nbr[0], L1_off, groups, dcn)
    L1_fea = _conv(sd, p + ".L1_fea_conv", cat([L1_fea, _up2(L2_fea)], 1))  # no lrelu here
    off = _lrelu(_conv(sd, p + ".cas_offset_conv1", cat([L1_fea, ref[0]], 1)))
    off = _lrelu(_conv(sd, p + ".cas_offset_conv2", off))
    out = _lrelu(dcn_pack(sd, p + ".cas_dcnpack", L1_fea, off, groups, dcn))
    if taps is not None:
        taps.update(L3_offset=L3_off, L3_fea=L3_fea, L2_offset=L2_off, L2_fea=L2_fea,
                    L1_offset=L1_off, L1_fea=L1_fea, cas_offset=off)
    return out


def tsa_fusion(sd, aligned, center, p="tsa_fusion"):
    """EDVR_arch.py:168-208.  aligned: [B, N, C, H, W]."""
    B, N, C, H, W = aligned.shape
    emb_ref = _conv(sd, p + ".tAtt_2", aligned[:, center])
    emb = _conv(sd, p + ".tAtt_1", aligned.reshape(-1, C, H, W)).view(B, N, -1, H, W)
    cor = torch.stack([(emb[:, i] * emb_ref).sum(1) for i in range(N)], 1)  # B, N, H, W
    prob = torch.sigmoid(cor).unsqueeze(2)  # broadcast over C (sigmoid, not softmax)
    ali = (aligned * prob).reshape(B, N * C, H, W)
    fea = _lrelu(_conv(sd, p + ".fea_fusion", ali))
    att = _lrelu(_conv(sd, p + ".sAtt_1", ali))
    att = _lrelu(_conv(sd, p + ".sAtt_2",
                       torch.cat([F.max_pool2d(att, 3, 2, 1), F.avg_pool2d(att, 3, 2, 1)], 1)))
    att_L = _lrelu(_conv(sd, p + ".sAtt_L1", att))
    att_L = _lrelu(_conv(sd, p + ".sAtt_L2",
                         torch.cat([F.max_pool2d(att_L, 3, 2, 1), F.avg_pool2d(att_L, 3, 2, 1)], 1)))
    att_L = _up2(_lrelu(_conv(sd, p + ".sAtt_L3", att_L)))
    att = _lrelu(_conv(sd, p + ".sAtt_3", att)) + att_L
    att = _up2(_lrelu(_conv(sd, p + ".sAtt_4", att)))
    att = _conv(sd, p + ".sAtt_5", att)
    att_add = _conv(sd, p + ".sAtt_add_2", _lrelu(_conv(sd, p + ".sAtt_add_1", att)))
    return fea * torch.sigmoid(att) * 2 + att_add


def predeblur(sd, x, HR_in, p="pre_deblur"):
    """EDVR_arch.py:43-59."""
    if HR_in:
        L1 = _lrelu(_conv(sd, p + ".conv_first_1", x))
        L1 = _lrelu(_conv(sd, p + ".conv_first_2", L1, stride=2))
        L1 = _lrelu(_conv(sd, p + ".conv_first_3", L1, stride=2))
    else:
        L1 = _lrelu(_conv(sd, p + ".conv_first", x))
    L2 = _lrelu(_conv(sd, p + ".deblur_L2_conv", L1, stride=2))
    L3 = _lrelu(_conv(sd, p + ".deblur_L3_conv", L2, stride=2))
    L3 = _up2(resblock(sd, p + ".RB_L3_1", L3))
    L2 = resblock(sd, p + ".RB_L2_1", L2) + L3
    L2 = _up2(resblock(sd, p + ".RB_L2_2", L2))
    L1 = resblock(sd, p + ".RB_L1_2", resblock(sd, p + ".RB_L1_1", L1)) + L2
    for k in (3, 4, 5):
        L1 = resblock(sd, p + ".RB_L1_%d" % k, L1)
    return L1


def _count(sd, prefix):
    n = 0
    while "%s.%d.conv1.weight" % (prefix, n) in sd:
        n += 1
    return n


def edvr_forward(sd, x, groups=8, center=None, w_TSA=True, upsample=True, is_predeblur=False,
                 HR_in=False, dcn=dcn_forward, taps=None, detach=True):
    """``EDVR.forward`` (upsample=True, EDVR_arch.py:258-320) or ``EDVR_NoUp.forward``
    (upsample=False, :358-404) from a reference-keyed ``state_dict``.  detach=False keeps the weights in the autograd graph
    (oracle/ref_gpu.py times the reference's training step that way, with a differentiable ``dcn``)."""
    sd = {k: (v.detach() if detach else v).to(x.dtype) for k, v in sd.items()}
    B, N, C, H, W = x.shape
    center = N // 2 if center is None else center
    x_center = x[:, center].contiguous()
    xf = x.reshape(-1, C, H, W)
    if upsample and is_predeblur:
        L1 = _conv(sd, "conv_1x1", predeblur(sd, xf, HR_in))
        if HR_in:
            H, W = H // 4, W // 4
    elif upsample and HR_in:
        L1 = _lrelu(_conv(sd, "conv_first_1", xf))
        L1 = _lrelu(_conv(sd, "conv_first_2", L1, stride=2))
        L1 = _lrelu(_conv(sd, "conv_first_3", L1, stride=2))
        H, W = H // 4, W // 4
    else:
        L1 = _lrelu(_conv(sd, "conv_first", xf))
    for i in range(_count(sd, "feature_extraction")):
        L1 = resblock(sd, "feature_extraction.%d" % i, L1)
    L2 = _lrelu(_conv(sd, "fea_L2_conv1", L1, stride=2))
    L2 = _lrelu(_conv(sd, "fea_L2_conv2", L2))
    L3 = _lrelu(_conv(sd, "fea_L3_conv1", L2, stride=2))
    L3 = _lrelu(_conv(sd, "fea_L3_conv2", L3))
    L1 = L1.view(B, N, -1, H, W)
    L2 = L2.view(B, N, -1, H // 2, W // 2)
    L3 = L3.view(B, N, -1, H // 4, W // 4)
    ref = [L1[:, center], L2[:, center], L3[:, center]]
    aligned = torch.stack(
        [pcd_align(sd, [L1[:, i], L2[:, i], L3[:, i]], ref, groups, dcn=dcn,
                   taps=taps if (taps is not None and i == 0) else None) for i in range(N)], 1)
    if w_TSA:
        fea = tsa_fusion(sd, aligned, center)
    else:
        fea = _conv(sd, "tsa_fusion", aligned.reshape(B, -1, H, W))
    if taps is not None:
        taps.update(L1=L1, L2=L2, L3=L3, aligned=aligned, fused=fea)
    out = fea
    for i in range(_count(sd, "recon_trunk")):
        out = resblock(sd, "recon_trunk.%d" % i, out)
    if upsample:
        out = _lrelu(F.pixel_shuffle(_conv(sd, "upconv1", out), 2))
        out = _lrelu(F.pixel_shuffle(_conv(sd, "upconv2", out), 2))
    out = _conv(sd, "conv_last", _lrelu(_conv(sd, "HRconv", out)))
    if upsample and not HR_in:
        base = F.interpolate(x_center, scale_factor=4, mode="bilinear", align_corners=False)
    else:
        base = x_center
    return out + base


def pixel_shuffle2_index(c, h, w, i, j):
    """Integer index map of nn.PixelShuffle(2): out[b,c,2h+i,2w+j] = in[b,4c+2i+j,h,w]."""
    return 4 * c + 2 * i + j, h, w


def index_generation_note():
    return "data/util.py:169-214 is a 'next' row (SURVEY.md 8f), not restated yet"

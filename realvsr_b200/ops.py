"""Functional wrappers over the operator-level C ABI (include/rvsr_b200.h).

``conv2d_fused`` is one nn.Conv2d site of the reference's EDVR_arch.py together with what
surrounds it there: the torch.cat in front, bias, activation, residual add and pixel shuffle.
NCHW CUDA tensors in and out; all compute is the library's own kernels.
"""
import ctypes

import torch

from . import _lib

ACT = {None: _lib.ACT_NONE, "none": _lib.ACT_NONE, "lrelu": _lib.ACT_LRELU, "relu": _lib.ACT_RELU}


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def conv2d_fused(x, weight, bias=None, x2=None, residual=None, stride=1, act=None, shuffle=False, use_tc=None):
    """y = [pixel_shuffle2](act(conv(cat(x, x2), weight, pad=k//2, stride) + bias) + residual)."""
    if not x.is_cuda:
        raise NotImplementedError("realvsr_b200.ops.conv2d_fused: CUDA tensors only (no CPU fallback)")
    if x.dtype not in (torch.float32, torch.float16):
        raise RuntimeError("conv2d_fused: float32/float16 only")
    dt = _lib.F16 if x.dtype == torch.float16 else _lib.F32
    if use_tc is None:
        use_tc = dt == _lib.F16
    x = x.contiguous()
    B, C1, H, W = x.shape
    C2 = 0 if x2 is None else x2.shape[1]
    Cout, Cin, ks, _ = weight.shape
    if Cin != C1 + C2:
        raise RuntimeError("conv2d_fused: weight expects %d input channels, got %d" % (Cin, C1 + C2))
    Ho, Wo = (H, W) if stride == 1 else ((H - 1) // 2 + 1, (W - 1) // 2 + 1)
    y = x.new_empty((B, Cout // 4, 2 * Ho, 2 * Wo) if shuffle else (B, Cout, Ho, Wo))
    L = _lib.lib()
    cast = lambda t: None if t is None else t.to(x.dtype).contiguous()  # noqa: E731
    w, b, x2, residual = cast(weight), cast(bias), cast(x2), cast(residual)
    with torch.cuda.device(x.device):
        ws = torch.empty(L.rvsr_conv2d_fwd_workspace_bytes(B, C1, C2, H, W, Cout, ks, dt), dtype=torch.uint8,
                         device=x.device)
        _lib.check(L.rvsr_conv2d_fwd(_p(x), _p(x2), _p(w), _p(b), _p(residual), _p(y), B, C1, C2, H, W, Cout, ks,
                                     stride, ACT[act], int(bool(shuffle)), dt, int(bool(use_tc)), _p(ws), ws.numel(),
                                     ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)),
                   "conv2d_fused")
    return y


def mdcn_pack(x, feat, w_offset_mask, b_offset_mask, weight, bias, deformable_groups, act=None):
    """ModulatedDeformConvPack.forward with extra_offset_mask=True (reference dcn/deform_conv.py:274-292)
    as ONE operator: conv_offset_mask(feat) -> split / sigmoid -> modulated deformable 3x3 conv of x
    (+ optional activation).  fp16 inputs with 64 channels run the tcgen05 kernel pair (offsets and
    mask never leave the fused format); everything else the CUDA-core kernels."""
    if not x.is_cuda:
        raise NotImplementedError("realvsr_b200.ops.mdcn_pack: CUDA tensors only (no CPU fallback)")
    if x.dtype not in (torch.float32, torch.float16):
        raise RuntimeError("mdcn_pack: float32/float16 only")
    dt = _lib.F16 if x.dtype == torch.float16 else _lib.F32
    x, feat = x.contiguous(), feat.to(x.dtype).contiguous()
    B, C, H, W = x.shape
    Cout = weight.shape[0]
    if tuple(w_offset_mask.shape) != (27 * deformable_groups, C, 3, 3) or tuple(weight.shape[1:]) != (C, 3, 3):
        raise RuntimeError("mdcn_pack: expects 3x3 kernels, conv_offset_mask with 27*deformable_groups outputs")
    y = x.new_empty(B, Cout, H, W)
    L = _lib.lib()
    cast = lambda t: None if t is None else t.to(x.dtype).contiguous()  # noqa: E731
    wom, bom, w, b = cast(w_offset_mask), cast(b_offset_mask), cast(weight), cast(bias)
    with torch.cuda.device(x.device):
        ws = torch.empty(max(1, L.rvsr_mdcn_pack_fwd_workspace_bytes(B, C, H, W, Cout, deformable_groups, dt)),
                         dtype=torch.uint8, device=x.device)
        _lib.check(L.rvsr_mdcn_pack_fwd(_p(x), _p(feat), _p(wom), _p(bom), _p(w), _p(b), _p(y), B, C, H, W, Cout,
                                        deformable_groups, ACT[act], dt, _p(ws), ws.numel(),
                                        ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)),
                   "mdcn_pack")
    return y


class _Upsample2xNCHW(torch.autograd.Function):
    """F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False) on a CUDA NCHW tensor (fp32 / fp16 / bf16), with its
    exact adjoint as backward.  torch's NCHW kernel for this op gives one thread an output pixel and loops over all images x
    channels inside it (a [80, 64, 16, 16] tensor runs on 1024 threads): 9.7 ms of the 60 ms fp32 training step."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        N, C, H, W = x.shape
        y = x.new_empty((N, C, 2 * H, 2 * W))
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().rvsr_upsample2x_nchw(_p(x), _p(y), N * C, H, W, 1.0, 0, _DT[x.dtype],
                                                       ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)), "upsample2x_nchw")
        return y

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        N, C, H2, W2 = g.shape
        gx = g.new_empty((N, C, H2 // 2, W2 // 2))
        with torch.cuda.device(g.device):
            _lib.check(_lib.lib().rvsr_upsample2x_nchw(_p(g), _p(gx), N * C, H2 // 2, W2 // 2, 1.0, 1, _DT[g.dtype],
                                                       ctypes.c_void_p(torch.cuda.current_stream(g.device).cuda_stream)), "upsample2x_nchw (adjoint)")
        return gx


_DT = {torch.float32: _lib.F32, torch.float16: _lib.F16, torch.bfloat16: _lib.BF16}


def upsample2x(x):
    """x2 bilinear upsample (align_corners=False) of an NCHW tensor: this library's kernel on CUDA tensors of a supported dtype,
    torch's op otherwise (CPU tensors: the reference's own arithmetic)."""
    if x.is_cuda and x.dim() == 4 and x.dtype in _DT:
        return _Upsample2xNCHW.apply(x)
    return torch.nn.functional.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)

"""Host-side mirror of the reference's ``codes/models/archs/dcn/deform_conv.py``.

Same public names, constructor arguments, parameter names and error behaviour; the
native call goes to the C ABI in include/rvsr_b200.h instead of the reference's pybind
module ``deform_conv_cuda``:

  ModulatedDeformConvFunction.forward   (reference deform_conv.py:99-119)  -> rvsr_mdcn_fwd
  ModulatedDeformConvFunction.backward  (reference deform_conv.py:121-141) -> rvsr_mdcn_bwd
  ModulatedDeformConv / ...Pack         (reference deform_conv.py:220-292)

DCN v1 (``DeformConv``, ``DeformConvPack``, ``deform_conv``; reference :15-94, :160-217) is
exported by the reference but never instantiated anywhere in it; the names are kept
importable and raise NotImplementedError when used.
"""
import ctypes
import math

import torch
import torch.nn as nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.modules.utils import _pair

from ... import _lib


def _dtype_code(t):
    if t.dtype == torch.float32:
        return _lib.F32
    if t.dtype == torch.float16:
        return _lib.F16
    if t.dtype == torch.bfloat16:
        return _lib.BF16   # bf16 tensors, fp32 arithmetic (torch.autocast(bfloat16) training); the reference has no bf16 dispatch
    # the reference dispatches double/float/half only (AT_DISPATCH_FLOATING_TYPES_AND_HALF)
    raise RuntimeError("modulated_deform_conv: unsupported dtype %s (float32/float16/bfloat16 only)" % t.dtype)


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


class ModulatedDeformConvFunction(Function):

    @staticmethod
    def forward(ctx, input, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1, groups=1,
                deformable_groups=1):
        ctx.conv = (int(stride), int(padding), int(dilation), int(groups), int(deformable_groups))
        ctx.with_bias = bias is not None
        if not input.is_cuda:
            raise NotImplementedError  # like the reference (deform_conv.py:109-110): CUDA only
        if not (input.is_contiguous() and weight.is_contiguous()):
            # reference: TORCH_CHECK(input.is_contiguous()) deform_conv_cuda.cpp:497-498
            raise RuntimeError("input tensor has to be contiguous")
        offset, mask = offset.contiguous(), mask.contiguous()
        if any(t.requires_grad for t in (weight, mask, offset, input)):
            ctx.save_for_backward(input, offset, mask, weight, bias if bias is not None else input.new_empty(1))
        B, C, H, W = input.shape
        Cout, _, kh, kw = weight.shape
        s, p, d, g, dg = ctx.conv
        output = input.new_empty(ModulatedDeformConvFunction._infer_shape(ctx, input, weight))
        L, dt = _lib.lib(), _dtype_code(input)
        dims = (B, C, H, W, Cout, kh, kw, s, p, d, g, dg)
        with torch.cuda.device(input.device):
            ws = input.new_empty(max(1, L.rvsr_mdcn_fwd_workspace_bytes(*dims, dt)), dtype=torch.uint8)
            b = bias.to(input.dtype).contiguous() if bias is not None else None
            _lib.check(L.rvsr_mdcn_fwd(_p(input), _p(offset.to(input.dtype)), _p(mask.to(input.dtype)),
                                       _p(weight.to(input.dtype)), _p(b), _p(output), *dims, dt, _p(ws),
                                       ws.numel(), _stream(input)), "modulated_deform_conv forward")
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        if not grad_output.is_cuda:
            raise NotImplementedError
        input, offset, mask, weight, bias = ctx.saved_tensors
        s, p, d, g, dg = ctx.conv
        # gradients are computed in fp32 (the reference's half path accumulates atomics in half).  bf16 tensors go through
        # the ABI as they are (RVSR_BF16: widened and rounded inside the library); fp16 / mixed dtypes are widened here.
        native_bf16 = all(t.dtype == torch.bfloat16 for t in (input, offset, mask, grad_output))  # weight: fp32 master under autocast
        wd = torch.bfloat16 if native_bf16 else torch.float32
        code = _lib.BF16 if native_bf16 else _lib.F32
        xi, of, mk, wt, go = [t.detach().to(wd).contiguous() for t in (input, offset, mask, weight, grad_output)]
        grad_input, grad_offset, grad_mask = torch.empty_like(xi), torch.empty_like(of), torch.empty_like(mk)
        grad_weight = torch.zeros_like(wt)
        grad_bias = torch.zeros(weight.shape[0], dtype=wd, device=xi.device) if ctx.with_bias else None
        B, C, H, W = xi.shape
        Cout, _, kh, kw = wt.shape
        dims = (B, C, H, W, Cout, kh, kw, s, p, d, g, dg)
        L = _lib.lib()
        with torch.cuda.device(xi.device):
            ws = xi.new_empty(max(1, L.rvsr_mdcn_bwd_workspace_bytes(*dims, code)), dtype=torch.uint8)
            _lib.check(L.rvsr_mdcn_bwd(_p(xi), _p(of), _p(mk), _p(wt), _p(go), _p(grad_input), _p(grad_offset),
                                       _p(grad_mask), _p(grad_weight), _p(grad_bias), *dims, code, _p(ws),
                                       ws.numel(), _stream(xi)), "modulated_deform_conv backward")
        cast = lambda t, like: None if t is None else t.to(like.dtype)  # noqa: E731
        return (cast(grad_input, input), cast(grad_offset, offset), cast(grad_mask, mask),
                cast(grad_weight, weight), cast(grad_bias, bias) if ctx.with_bias else None,
                None, None, None, None, None)

    @staticmethod
    def _infer_shape(ctx, input, weight):
        s, p, d = ctx.conv[:3]
        n, (h, w), (kh, kw) = input.size(0), input.shape[2:4], weight.shape[2:4]
        return (n, weight.size(0), (h + 2 * p - (d * (kh - 1) + 1)) // s + 1,
                (w + 2 * p - (d * (kw - 1) + 1)) // s + 1)


modulated_deform_conv = ModulatedDeformConvFunction.apply


def deform_conv(*args, **kwargs):
    raise NotImplementedError("DCN v1 (deform_conv) is exported but unused by the reference; not built here")


class ModulatedDeformConv(nn.Module):

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 deformable_groups=1, bias=True):
        super(ModulatedDeformConv, self).__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = _pair(kernel_size)
        self.stride, self.padding, self.dilation = stride, padding, dilation
        self.groups, self.deformable_groups = groups, deformable_groups
        self.with_bias = bias
        self.weight = nn.Parameter(torch.Tensor(out_channels, in_channels // groups, *self.kernel_size))
        if bias:
            self.bias = nn.Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self):
        # U(-1/sqrt(fan_in), 1/sqrt(fan_in)), zero bias -- reference deform_conv.py:243-250
        fan_in = self.in_channels * self.kernel_size[0] * self.kernel_size[1]
        bound = 1. / math.sqrt(fan_in)
        with torch.no_grad():
            self.weight.uniform_(-bound, bound)
            if self.bias is not None:
                self.bias.zero_()

    def forward(self, x, offset, mask):
        return modulated_deform_conv(x, offset, mask, self.weight, self.bias, self.stride, self.padding,
                                     self.dilation, self.groups, self.deformable_groups)


class ModulatedDeformConvPack(ModulatedDeformConv):
    """DCNv2 that predicts its own offsets and mask (from ``x`` or, with
    ``extra_offset_mask=True``, from a second feature map: ``forward([x, feat])``)."""

    def __init__(self, *args, extra_offset_mask=False, **kwargs):
        super(ModulatedDeformConvPack, self).__init__(*args, **kwargs)
        self.extra_offset_mask = extra_offset_mask
        k = self.kernel_size[0] * self.kernel_size[1]
        self.conv_offset_mask = nn.Conv2d(self.in_channels, self.deformable_groups * 3 * k,
                                          kernel_size=self.kernel_size, stride=_pair(self.stride),
                                          padding=_pair(self.padding), bias=True)
        self.init_offset()

    def init_offset(self):
        # zero init => offsets 0, mask 0.5 at start (reference deform_conv.py:270-272)
        nn.init.zeros_(self.conv_offset_mask.weight)
        nn.init.zeros_(self.conv_offset_mask.bias)

    def _fusable(self, x, feat):
        return (x.is_cuda and not torch.is_grad_enabled() and self.kernel_size == (3, 3) and self.stride == 1
                and self.padding == 1 and self.dilation == 1 and self.groups == 1 and self.bias is not None
                and x.dtype in (torch.float32, torch.float16) and feat.shape == x.shape)

    def forward(self, x):
        feat = x
        if self.extra_offset_mask:
            x, feat = x[0], x[1]
        if self._fusable(x, feat):
            # inference: the whole pack is one C-ABI call (rvsr_mdcn_pack_fwd)
            from ... import ops
            return ops.mdcn_pack(x, feat, self.conv_offset_mask.weight, self.conv_offset_mask.bias, self.weight,
                                 self.bias, self.deformable_groups)
        x = x.contiguous()   # channels_last activations (a user-side cuDNN setting) reach the NCHW operator as a copy
        om = self.conv_offset_mask(feat)
        third = om.shape[1] // 3
        # chunk(3) then cat(o1, o2) == the first two thirds; the reference also computes a
        # mean(|offset|) here and throws it away (deform_conv.py:285) -- not reproduced.
        offset, mask = om[:, :2 * third], torch.sigmoid(om[:, 2 * third:])
        return modulated_deform_conv(x, offset, mask, self.weight.contiguous(), self.bias, self.stride, self.padding,
                                     self.dilation, self.groups, self.deformable_groups)


class DeformConv(nn.Module):
    def __init__(self, *args, **kwargs):
        super(DeformConv, self).__init__()
        raise NotImplementedError("DCN v1 (DeformConv) is exported but unused by the reference; not built here")


class DeformConvPack(DeformConv):
    pass

"""oracle/ref_gpu.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The reference's own GPU path, as far as it can travel to the GPU box: the op sequence of
``EDVR.forward`` (EDVR_arch.py:258-320: one PCD pass per frame in a Python loop, torch.cat, separate
activations -- restated in oracle/edvr_oracle.py) executed with torch CUDA ops (cuDNN convolutions, ATen
pool / interpolate / pixel_shuffle: what the reference's nn.Module calls dispatch to) and, for the DCN,
the reference's OWN CUDA extension compiled unmodified into oracle/_ref/ (oracle/build_ref.py), called
exactly as ``ModulatedDeformConvFunction.forward`` calls it (deform_conv.py:113-118: output = new_empty,
two empty dummy buffers, kernel / stride / pad / dilation / groups / deformable groups, with_bias).

Used by
  * tests/test_gpu_fullsize.py -- "is our fp16 engine as close to the fp32 oracle as the reference's own
    fp16 GPU path is?" (the honest reading of north_star's 1e-3 for an fp16 configuration);
  * bench.py's ``gpu_reference`` key -- SURVEY.md 8(d) "the reference's own compiled CUDA extension (fp32
    and fp16) on 1xB200 as the existing GPU kernel", timed AFTER the product's timed regions.
Nothing under realvsr_b200/ imports this file.
"""
import torch

from . import edvr_oracle as O
from .build_ref import load_ref

_ext = None


def available():
    global _ext
    if _ext is None:
        _ext = load_ref() or False
    return bool(_ext) and torch.cuda.is_available()


def ref_dcn(x, offset, mask, weight, bias, stride=1, padding=1, dilation=1, groups=1, deformable_groups=1):
    """modulated_deform_conv_cuda_forward of the reference extension (deform_conv_cuda.cpp:490-569)."""
    assert available(), "oracle/_ref/deform_conv_cuda.so not built (python oracle/build_ref.py)"
    x, weight = x.contiguous(), weight.contiguous()
    B, C, H, W = x.shape
    Cout, _, kh, kw = weight.shape
    Ho = (H + 2 * padding - (dilation * (kh - 1) + 1)) // stride + 1
    Wo = (W + 2 * padding - (dilation * (kw - 1) + 1)) // stride + 1
    out = x.new_empty(B, Cout, Ho, Wo)
    _ext.modulated_deform_conv_cuda_forward(x, weight, bias if bias is not None else x.new_empty(1), x.new_empty(0),
                                            offset.contiguous(), mask.contiguous(), out, x.new_empty(0), kh, kw,
                                            stride, stride, padding, padding, dilation, dilation, groups,
                                            deformable_groups, bias is not None)
    return out


def edvr_forward(sd, x, **kw):
    """The reference network on the GPU in x's dtype (fp32 or fp16): sd / x are moved to x's device and dtype."""
    sd = {k: v.to(device=x.device, dtype=x.dtype) for k, v in sd.items()}
    with torch.no_grad():
        return O.edvr_forward(sd, x, dcn=ref_dcn, **kw)


def time_forward(sd, x, steps=5, warmup=2, **kw):
    """CUDA-event time (ms, median) of one reference forward on x."""
    sd = {k: v.to(device=x.device, dtype=x.dtype) for k, v in sd.items()}
    ts = []
    with torch.no_grad():
        for i in range(warmup + steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            O.edvr_forward(sd, x, dcn=ref_dcn, **kw)
            e1.record()
            torch.cuda.synchronize()
            if i >= warmup:
                ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


class _RefDcnTrain(torch.autograd.Function):
    """The reference extension's forward AND backward wired into autograd the way the reference's own
    ``ModulatedDeformConvFunction`` does it (deform_conv.py:97-141): zero-initialised gradient tensors, two empty dummy
    buffers, one ``modulated_deform_conv_cuda_backward`` call (deform_conv_cuda.cpp:571-685)."""

    @staticmethod
    def forward(ctx, x, offset, mask, weight, bias, stride, padding, dilation, groups, deformable_groups):
        ctx.cfg = (stride, padding, dilation, groups, deformable_groups)
        x, offset, mask, weight = x.contiguous(), offset.contiguous(), mask.contiguous(), weight.contiguous()
        ctx.save_for_backward(x, offset, mask, weight, bias)
        return ref_dcn(x, offset, mask, weight, bias, stride, padding, dilation, groups, deformable_groups)

    @staticmethod
    def backward(ctx, grad_output):
        x, offset, mask, weight, bias = ctx.saved_tensors
        stride, padding, dilation, groups, dg = ctx.cfg
        gx, go, gm, gw, gb = (torch.zeros_like(t) for t in (x, offset, mask, weight, bias))
        _ext.modulated_deform_conv_cuda_backward(x, weight, bias, x.new_empty(0), offset, mask, x.new_empty(0), gx, gw, gb, go, gm,
                                                 grad_output.contiguous(), weight.shape[2], weight.shape[3], stride, stride, padding,
                                                 padding, dilation, dilation, groups, dg, True)
        return gx, go, gm, gw, gb, None, None, None, None, None


def ref_dcn_train(x, offset, mask, weight, bias, stride=1, padding=1, dilation=1, groups=1, deformable_groups=1):
    assert available() and bias is not None
    return _RefDcnTrain.apply(x, offset, mask, weight, bias, stride, padding, dilation, groups, deformable_groups)


def time_train_step(sd, x, gt, steps=3, warmup=2, **kw):
    """CUDA-event time (ms, median) of one fp32 training step of the reference network (forward, L1 loss, backward) on torch
    CUDA ops + the reference extension -- the reference trains in fp32 and its extension has no bfloat16 dispatch."""
    params = {k: v.to(device=x.device, dtype=x.dtype).requires_grad_() for k, v in sd.items()}
    ts = []
    for i in range(warmup + steps):
        for p_ in params.values():
            p_.grad = None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loss = torch.nn.functional.l1_loss(O.edvr_forward(params, x, dcn=ref_dcn_train, detach=False, **kw), gt)
        loss.backward()
        e1.record()
        torch.cuda.synchronize()
        if i >= warmup:
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], float(loss.detach())


def train_grads(sd, x, gt, **kw):
    """(loss, {name: gradient}) of one fp32 training step of the reference network through the reference extension's backward."""
    params = {k: v.to(device=x.device, dtype=x.dtype).requires_grad_() for k, v in sd.items()}
    loss = torch.nn.functional.l1_loss(O.edvr_forward(params, x, dcn=ref_dcn_train, detach=False, **kw), gt)
    loss.backward()
    return float(loss.detach()), {k: v.grad.detach().clone() for k, v in params.items()}

"""bf16 training path on channel-blocked ("C8") tensors: torch.autograd.Functions over the rvsr_c8_* C ABI.

BASELINE cfg5 is a bf16 training step.  Under torch.autocast every nn.Conv2d of EDVR_arch.py runs on cuDNN, which
spends more time converting NCHW <-> NHWC, reducing bias gradients and launching elementwise kernels than convolving
(profiles/r02_cfg5_profile.txt, lower half).  Here activations stay in the inference engine's layout -- [N, C/8, H, W, 8]
bfloat16, an ordinary 5-D torch tensor -- between layers, and each layer is one Function whose forward AND backward are
this library's kernels:

    conv / conv_pair / conv_first   forward: rvsr_c8_conv_fwd (bias, activation, residual add, torch.cat of the inputs,
                                    PixelShuffle(2) fused); data gradient: the same kernel with the transposed, flipped weights
                                    (one launch per concatenated input); weight + bias gradient: rvsr_c8_conv_wgrad (pixel-K
                                    GEMM on tcgen05, fixed summation order); activation gradient from the saved OUTPUT
    dcn_pack                        ModulatedDeformConvPack on dcn_tc_kernel / dcn_bwd_tc_kernel
    upsample2x, tsa_temporal, pool_maxavg, tsa_final, to_c8 / from_c8     one CUDA-core kernel per direction
    GraphedStep                     forward + loss + backward as ONE CUDA graph

torch.autograd still owns the graph: it sums the gradients of a tensor with several consumers and calls these
backwards in order.  Parameters stay fp32 OIHW nn.Parameters (the reference's state_dict contract); they are re-packed
into the kernels' operand layout on every call because an optimizer step changes them.  No CPU or cuDNN fallback inside
a Function: an unsupported shape raises NotImplementedError (the caller, EDVR._forward_c8, only routes supported layers
here).
"""
import contextlib
import ctypes
import os

import torch

from . import _lib

ACT = {None: _lib.ACT_NONE, "none": _lib.ACT_NONE, "lrelu": _lib.ACT_LRELU, "relu": _lib.ACT_RELU}


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream(dev):
    # the raw handle of torch's current stream on `dev` (torch.cuda.current_stream(dev).cuda_stream costs ~8 us of Python per call,
    # and a training step makes ~1000 calls)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(idx))


class _OnDevice(object):
    """`with _OnDevice(dev)` only when dev is not already the current device (the context manager costs ~10 us of Python)."""
    __slots__ = ("ctx",)

    def __init__(self, dev):
        idx = dev.index
        self.ctx = None if (idx is None or idx == torch.cuda.current_device()) else torch.cuda.device(dev)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            return self.ctx.__exit__(*exc)
        return False


def channels(t):
    return t.shape[1] * 8


def _check_c8(t, what):
    if t.dim() != 5 or t.shape[-1] != 8 or t.dtype != torch.bfloat16 or not t.is_cuda:
        raise RuntimeError("%s: expected a CUDA bfloat16 [N, C/8, H, W, 8] tensor, got %s %s" % (what, tuple(t.shape), t.dtype))
    return t.contiguous()


# ---------------------------------------------------------------- layout conversion
def _to_c8_raw(x, planes=None):
    if not x.is_cuda:
        raise NotImplementedError("realvsr_b200.train_c8: CUDA tensors only (no CPU fallback)")
    if x.dtype not in (torch.bfloat16, torch.float32):
        x = x.float()
    x = x.contiguous()
    N, C, H, W = x.shape
    planes = (C + 7) // 8 if planes is None else planes
    y = torch.empty((N, planes, H, W, 8), dtype=torch.bfloat16, device=x.device)
    if planes > (C + 7) // 8:
        y[:, (C + 7) // 8:].zero_()
    with _OnDevice(x.device):
        _lib.check(_lib.lib().rvsr_c8_from_nchw(_p(x), _lib.BF16 if x.dtype == torch.bfloat16 else _lib.F32, _p(y), N, C, H, W,
                                                planes, _stream(x.device)), "c8_from_nchw")
    return y


def _from_c8_raw(x, C, dtype):
    N, planes, H, W, _ = x.shape
    y = torch.empty((N, C, H, W), dtype=dtype, device=x.device)
    with _OnDevice(x.device):
        _lib.check(_lib.lib().rvsr_c8_to_nchw(_p(x), _p(y), _lib.BF16 if dtype == torch.bfloat16 else _lib.F32, N, C, H, W,
                                              planes, _stream(x.device)), "c8_to_nchw")
    return y


class _ToC8(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.meta = (x.shape[1], x.dtype)
        return _to_c8_raw(x)

    @staticmethod
    def backward(ctx, g):
        C, dtype = ctx.meta
        return _from_c8_raw(_check_c8(g, "to_c8 backward"), C, dtype)


class _FromC8(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, C, dtype):
        x = _check_c8(x, "from_c8")
        ctx.planes = x.shape[1]
        return _from_c8_raw(x, C, dtype)

    @staticmethod
    def backward(ctx, g):
        return _to_c8_raw(g, ctx.planes), None, None  # channels beyond C (padding, or planes the caller dropped) get zero gradient


def to_c8(x):
    """NCHW (bf16 / fp32) -> [N, ceil(C/8), H, W, 8] bf16."""
    return _ToC8.apply(x)


def from_c8(x, C=None, dtype=torch.bfloat16):
    """[N, C/8, H, W, 8] bf16 -> NCHW `dtype` (the first C channels)."""
    return _FromC8.apply(x, channels(x) if C is None else C, dtype)


# ---------------------------------------------------------------- side streams of a captured step
class _Defer:
    """Side streams of GraphedStep: parallel branches of the captured graph.
    * Weight gradients.  Nothing downstream of a layer's backward reads its weight gradient, so inside a captured step the
      weight-gradient launches go to a SIDE stream: their under-filled grids and tails overlap the data-gradient chain on the
      main stream.  autograd cannot express that (AccumulateGrad runs on the main stream as soon as backward returns), so in
      this mode backward returns None for the parameters and `finish` stores the gradients in parameter.grad after the join.
      Inputs of a side-stream launch are kept alive until the main stream has waited for it (LAG launches later), so the
      allocator cannot hand their memory to a main-stream kernel in between.
    * Operand packing.  The packed operands depend on the parameters only: their launches go to a second stream that runs
      ahead of the layers (each convolution waits for its own operand's event)."""
    active = False     # weight gradients on `side` (during backward)
    packing = False    # operand packing on `pack_stream` (whole step)
    side = pack_stream = None
    wgrad_wanted = False
    pending = []       # [event after the launch on the side stream, input references or None, [(parameter, gradient)]]
    keep = []          # packed operands of this step
    LAG = 3

    @classmethod
    def begin(cls, dev, wgrad=True, packing=True):
        if cls.side is None or cls.side.device != dev:
            cls.side, cls.pack_stream = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        cls.pending, cls.keep = [], []
        cls.wgrad_wanted, cls.packing = wgrad, packing
        if packing:
            cls.pack_stream.wait_stream(torch.cuda.current_stream(dev))

    @classmethod
    def begin_backward(cls, dev):
        cls.active = cls.wgrad_wanted
        if cls.active:
            cls.side.wait_stream(torch.cuda.current_stream(dev))

    @classmethod
    def finish(cls, dev):
        """Join the side streams and deliver the gradients: parameter.grad = gradient (summed when a parameter has several)."""
        main = torch.cuda.current_stream(dev)
        if cls.active:
            main.wait_stream(cls.side)
        if cls.packing:
            main.wait_stream(cls.pack_stream)
        cls.active = cls.packing = False
        pending, cls.pending, cls.keep = cls.pending, [], []
        for _, _, outs in pending:
            for prm, grad in outs:
                prm.grad = grad if prm.grad is None else prm.grad + grad


# ---------------------------------------------------------------- convolution
_PACK_INFO = {}


def _pack_weights(weight, views):
    """fp32 OIHW parameter -> the kernels' operand layout, several views in ONE launch.  views: (Cout, Cin, ks, shuffle, mode,
    cin_total, c0, geom) with mode 0 = forward operand over input channels [c0, c0 + Cin), 1 = operand of the data gradient of
    input channels [c0, c0 + Cout); geom = (nsrc, C, N, H, W) of the launch that will read it -- the library decides which of
    its two layouts that launch reads (rvsr_c8_conv_layouts) and only that one is packed."""
    L = _lib.lib()
    n = len(views)
    spec = (ctypes.c_int * (8 * n))()
    sizes = []
    for k, (Cout, Cin, ks, shuffle, mode, cin_total, c0, geom) in enumerate(views):
        key = (Cout, Cin, ks, bool(shuffle), geom)
        hit = _PACK_INFO.get(key)
        if hit is None:  # pure functions of the shape: asked once per shape, not once per step
            nbytes = L.rvsr_c8_conv_weight_bytes(Cout, Cin, ks, int(shuffle))
            layouts = L.rvsr_c8_conv_layouts(geom[0], geom[1], geom[2], geom[3], geom[4], Cout, ks, int(shuffle)) if nbytes else 0
            hit = _PACK_INFO[key] = (nbytes, layouts)
        nbytes, layouts = hit
        if nbytes == 0 or layouts == 0:
            raise NotImplementedError("train_c8: convolution %d <- %d (k=%d) is not covered by the tcgen05 kernels" % (Cout, Cin, ks))
        spec[8 * k:8 * k + 8] = [Cout, Cin, ks, int(shuffle), mode, cin_total, c0, layouts]
        sizes.append(nbytes)
    dev = weight.device
    ahead = _Defer.packing and isinstance(weight, torch.nn.Parameter)   # a derived weight is produced on the main stream
    with torch.cuda.stream(_Defer.pack_stream) if ahead else contextlib.nullcontext():
        dsts = [torch.empty(b, dtype=torch.uint8, device=dev) for b in sizes]
        ptrs = (ctypes.c_void_p * n)(*[d.data_ptr() for d in dsts])
        _lib.check(L.rvsr_c8_conv_pack_weights(_p(weight), n, spec, ptrs, _stream(dev)), "c8_conv_pack_weights")
        if ahead:
            ev = torch.cuda.Event()
            ev.record(_Defer.pack_stream)
    if ahead:
        torch.cuda.current_stream(dev).wait_event(ev)
        _Defer.keep.append(dsts)   # the pack stream's allocator must not reuse them while main-stream kernels read them
    return dsts


def _pack_weight(weight, Cout, Cin, ks, shuffle, mode, cin_total, c0, geom):
    return _pack_weights(weight, [(Cout, Cin, ks, shuffle, mode, cin_total, c0, geom)])[0]


def _dgrad_views(nsrc, Cout, ks, N, H, W, needs):
    """views of the data-gradient operands of the sources that need a gradient (see _dgrad)"""
    ng = Cout // 64 if Cout >= 64 else 1
    return [(64, Cout, ks, False, 1, nsrc * 64, i * 64, (ng, min(Cout, 64), N, H, W)) for i in range(nsrc) if needs[i]]


def _conv_launch(xs, w_packed, bias, residual, N, H, W, C, Cout, ks, act, shuffle, res_mode=0, res_slope=0.0):
    dev = xs[0].device
    n = len(xs)
    ptrs = (ctypes.c_void_p * n)(*[x.data_ptr() for x in xs])
    strides = (ctypes.c_longlong * n)(*[x.stride(0) for x in xs])
    if shuffle:
        y = torch.empty((N, Cout // 32, 2 * H, 2 * W, 8), dtype=torch.bfloat16, device=dev)
    else:
        y = torch.empty((N, (Cout + 7) // 8, H, W, 8), dtype=torch.bfloat16, device=dev)
    _lib.check(_lib.lib().rvsr_c8_conv_fwd(ptrs, strides, n, C, _p(w_packed), _p(bias), _p(residual), _p(y), N, H, W, Cout, ks, 1,
                                           act, int(shuffle), res_mode, res_slope, _stream(dev)), "c8_conv_fwd")
    return y


def _src_ok(x):
    """A source may be a channel slice [:, a:b] of a contiguous C8 tensor (image stride > its own size)."""
    s = x.stride()
    H, W = x.shape[2], x.shape[3]
    return s[4] == 1 and s[3] == 8 and s[2] == 8 * W and s[1] == 8 * W * H


_SLOPE = {_lib.ACT_LRELU: 0.1, _lib.ACT_RELU: 0.0}


_WGRAD_WS = {}


def _wgrad_jobs(jobs, N, H, W, Cout, ks):
    """Up to 8 weight gradients of one geometry in ONE launch.  jobs: (x, g, gw, cin_total, c0, gb or None): gw[:, c0:c0 + 64] and
    gb are written from source x and output gradient g."""
    L, n = _lib.lib(), len(jobs)
    dev = jobs[0][1].device
    arr = lambda ctype, vals: (ctype * n)(*vals)  # noqa: E731
    key = (n, N, H, W, Cout)
    nbytes = _WGRAD_WS.get(key)
    if nbytes is None:
        nbytes = _WGRAD_WS[key] = L.rvsr_c8_conv_wgrad_workspace_bytes(n, N, H, W, Cout)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    _lib.check(L.rvsr_c8_conv_wgrad(n, arr(ctypes.c_void_p, [j[0].data_ptr() for j in jobs]), arr(ctypes.c_longlong, [j[0].stride(0) for j in jobs]),
                                    arr(ctypes.c_void_p, [j[1].data_ptr() for j in jobs]), arr(ctypes.c_void_p, [j[2].data_ptr() for j in jobs]),
                                    arr(ctypes.c_void_p, [0 if j[5] is None else j[5].data_ptr() for j in jobs]),
                                    arr(ctypes.c_int, [j[3] for j in jobs]), arr(ctypes.c_int, [j[4] for j in jobs]),
                                    N, H, W, 64, Cout, ks, _p(ws), ws.numel(), _stream(dev)), "c8_conv_wgrad")


def _wgrad_alloc(nsrc, Cout, ks, need_bias, dev):
    gw = torch.empty((Cout, nsrc * 64, ks, ks), dtype=torch.float32, device=dev)
    return gw, (torch.empty(Cout, dtype=torch.float32, device=dev) if need_bias else None)


def _wgrad_launch(jobs, N, H, W, Cout, ks, params, outs):
    """Launch the weight-gradient jobs (8 per launch).  params: the tensors that receive a gradient; outs(): [(parameter,
    gradient tensor)] the jobs produce, called on the launch stream after the launches.  Returns True when the gradients travel
    through _Defer (the caller then returns None to autograd); a derived (non-leaf) weight, e.g. the permuted offset / mask
    convolution, always takes autograd's route."""
    if not _Defer.active or any(q is not None and not q.is_leaf for q in params):
        for k in range(0, len(jobs), 8):
            _wgrad_jobs(jobs[k:k + 8], N, H, W, Cout, ks)
        return False
    dev = jobs[0][1].device
    main, side = torch.cuda.current_stream(dev), _Defer.side
    side.wait_stream(main)
    with torch.cuda.stream(side):
        for k in range(0, len(jobs), 8):
            _wgrad_jobs(jobs[k:k + 8], N, H, W, Cout, ks)
        produced = outs()
        ev = torch.cuda.Event()
        ev.record(side)
    _Defer.pending.append([ev, jobs, produced])
    live = [q for q in _Defer.pending if q[1] is not None]
    if len(live) > _Defer.LAG:
        main.wait_event(live[0][0])
        live[0][1] = None
    return True


def _wgrad(xs, gp, N, H, W, Cout, ks, params):
    """Weight (+ bias) gradient of conv(cat(xs)) given the gradient gp of its (pre-activation) output: one job per source.
    params = (weight or None, bias or None): the parameters that need a gradient.  Returns (gw, gb), None where not needed or
    where _Defer delivers it."""
    gw, gb = _wgrad_alloc(len(xs), Cout, ks, params[1] is not None, gp.device)
    jobs = [(x, gp, gw, len(xs) * 64, i * 64, gb if i == 0 else None) for i, x in enumerate(xs)]
    outs = [(p, t) for p, t in zip(params, (gw, gb)) if p is not None]
    if _wgrad_launch(jobs, N, H, W, Cout, ks, params, lambda: outs):
        return None, None
    return (gw if params[0] is not None else None), gb


def _dgrad(weight, gp, i, nsrc, Cout, ks, N, H, W, residual=None, mask=None, slope=0.0, wp=None):
    """Gradient of source i (64 channels) of conv(cat(xs), weight): conv(gp, W[:, slice_i]^T flipped) [+ residual] [* act'(mask)].
    wp: the operand packed in the forward pass (one launch with the forward operand), else it is packed here."""
    gsrc = [gp[:, 8 * k:8 * k + 8] for k in range(Cout // 64)] if Cout >= 64 else [gp]
    if wp is None:
        wp = _pack_weight(weight, 64, Cout, ks, False, 1, nsrc * 64, i * 64, (len(gsrc), min(Cout, 64), N, H, W))
    if mask is not None:
        return _conv_launch(gsrc, wp, None, mask, N, H, W, min(Cout, 64), 64, ks, _lib.ACT_NONE, False, 2, slope)
    return _conv_launch(gsrc, wp, None, residual, N, H, W, min(Cout, 64), 64, ks, _lib.ACT_NONE, False)


class _ConvC8(torch.autograd.Function):
    """y = [PixelShuffle2](act(conv(cat(xs), weight) + bias)) [+ residual]; stride 1, pad ks // 2."""

    @staticmethod
    def forward(ctx, weight, bias, residual, act, shuffle, *xs):
        xs = [x if _src_ok(x) else x.contiguous() for x in xs]
        for x in xs:
            if x.dim() != 5 or x.dtype != torch.bfloat16 or not x.is_cuda:
                raise RuntimeError("conv_c8: sources must be CUDA bfloat16 [N, C/8, H, W, 8] tensors")
        N, C8, H, W, _ = xs[0].shape
        C = C8 * 8
        Cout, Cin, ks, _ = weight.shape
        if Cin != C * len(xs) or any(tuple(x.shape) != tuple(xs[0].shape) for x in xs):
            raise RuntimeError("conv_c8: weight expects %d input channels, sources give %d x %d" % (Cin, len(xs), C))
        if weight.dtype != torch.float32 or (bias is not None and bias.dtype != torch.float32):
            raise RuntimeError("conv_c8: parameters must be fp32 (the reference's state_dict dtype)")
        ctx.params = (weight, bias)
        weight = weight.contiguous()
        if residual is not None:
            residual = _check_c8(residual, "conv_c8 residual")
        needs = ctx.needs_input_grad
        dg_ok = C == 64 and (Cout % 64 == 0 or (Cout < 64 and Cout % 16 == 0))
        with _OnDevice(weight.device):
            views = [(Cout, Cin, ks, shuffle, 0, Cin, 0, (len(xs), C, N, H, W))]
            if dg_ok:
                views += _dgrad_views(len(xs), Cout, ks, N, H, W, needs[5:5 + len(xs)])
            packs = _pack_weights(weight, views)   # the forward operand and those of the data gradients: one launch
            y = _conv_launch(xs, packs[0], bias, residual, N, H, W, C, Cout, ks, act, shuffle)
        ctx.dgrad_packs = packs[1:] if dg_ok else None
        ctx.meta = (act, shuffle, N, H, W, C, Cout, ks, len(xs), residual is not None)
        ctx.saved_x = bool(needs[0] or needs[1])
        ctx.save_for_backward(weight, *(xs if ctx.saved_x else []), *([y] if act != _lib.ACT_NONE else []))
        return y

    @staticmethod
    def backward(ctx, g):
        act, shuffle, N, H, W, C, Cout, ks, nsrc, has_res = ctx.meta
        saved = ctx.saved_tensors
        weight = saved[0]
        xs = saved[1:1 + nsrc] if ctx.saved_x else None
        y = saved[-1] if act != _lib.ACT_NONE else None
        L = _lib.lib()
        dev = weight.device
        g = _check_c8(g, "conv_c8 backward")
        needs = ctx.needs_input_grad
        with _OnDevice(dev):
            s = _stream(dev)
            g_res = g if has_res and needs[2] else None  # the residual joins after the activation: its gradient is g itself
            if has_res and act != _lib.ACT_NONE:
                raise RuntimeError("conv_c8: residual with an activation is not a layer of this network")
            if shuffle:  # g, y are in the shuffled geometry [N, Cout/4 ch, 2H, 2W] -> gradient of the conv output
                gp = torch.empty((N, Cout // 8, H, W, 8), dtype=torch.bfloat16, device=dev)
                _lib.check(L.rvsr_c8_unshuffle2_act_bwd(_p(g), _p(y), _p(gp), N, Cout, H, W, act, s), "c8_unshuffle2_act_bwd")
            elif act != _lib.ACT_NONE:
                gp = torch.empty_like(g)
                _lib.check(L.rvsr_c8_act_bwd(_p(g), _p(y), _p(gp), g.numel(), act, s), "c8_act_bwd")
            else:
                gp = g
            gw = gb = None
            if needs[0] or needs[1]:
                if C != 64:
                    raise NotImplementedError("conv_c8: the weight gradient is built for 64-channel sources")
                gw, gb = _wgrad(xs, gp, N, H, W, Cout, ks, (ctx.params[0] if needs[0] else None, ctx.params[1] if needs[1] else None))
            gxs = [None] * nsrc
            if any(needs[5:5 + nsrc]):
                # dX_i = conv(dY, W[:, slice_i]^T flipped): Cout gradient channels enter as 64-channel sources
                if C != 64 or (Cout % 64 != 0 and not (Cout < 64 and Cout % 16 == 0)):
                    raise NotImplementedError("conv_c8: the data gradient needs 64-channel sources and Cout %% 64 == 0 or Cout in "
                                              "{16, 32, 48} (got %d <- %d x %d)" % (Cout, nsrc, C))
                packs = iter(ctx.dgrad_packs)
                gxs = [_dgrad(weight, gp, i, nsrc, Cout, ks, N, H, W, wp=next(packs)) if needs[5 + i] else None for i in range(nsrc)]
        return (gw, gb, g_res, None, None, *gxs)


class _ConvFirstC8(torch.autograd.Function):
    """conv_first (EDVR_arch.py:233, :276): 3x3 convolution of an NCHW image with <= 16 channels + LeakyReLU, C8 output.
    The image is packed to 16 channels (the tensor-core K granularity), the weight's input channels zero-padded to match.
    Weight gradient on conv_wgrad_tc_kernel, which reads 64-channel sources: the packed image is allocated with three images
    of zero slack behind it and handed over as "8 channel blocks per image, image stride 2 blocks" -- blocks 2..7 of an
    image alias the following images (finite values), their rows of the gradient are discarded.  The input gets no gradient
    (frames are data)."""

    @staticmethod
    def forward(ctx, x, weight, bias, act):
        if not x.is_cuda:
            raise NotImplementedError("realvsr_b200.train_c8: CUDA tensors only (no CPU fallback)")
        N, C, H, W = x.shape
        Cout, Cin, ks, _ = weight.shape
        if Cin != C or C > 16 or ks != 3 or Cout % 64 != 0 or weight.dtype != torch.float32:
            raise NotImplementedError("conv_first_c8: 3x3, <= 16 input channels, Cout %% 64 == 0, fp32 parameters")
        xp = torch.zeros((N + 3, 2, H, W, 8), dtype=torch.bfloat16, device=x.device)
        xin = x if x.dtype in (torch.bfloat16, torch.float32) else x.float()
        xin = xin.contiguous()
        w16 = torch.nn.functional.pad(weight.detach(), (0, 0, 0, 0, 0, 16 - C)).contiguous()
        with _OnDevice(x.device):
            _lib.check(_lib.lib().rvsr_c8_from_nchw(_p(xin), _lib.BF16 if xin.dtype == torch.bfloat16 else _lib.F32, _p(xp), N, C, H, W, 2,
                                                    _stream(x.device)), "c8_from_nchw")
            wp = _pack_weight(w16, Cout, 16, 3, False, 0, 16, 0, (1, 16, N, H, W))
            y = _conv_launch([xp[:N]], wp, bias, None, N, H, W, 16, Cout, 3, act, False)
        ctx.meta = (act, N, C, H, W, Cout)
        ctx.params = (weight, bias)
        ctx.save_for_backward(xp, *([y] if act != _lib.ACT_NONE else []))
        return y

    @staticmethod
    def backward(ctx, g):
        act, N, C, H, W, Cout = ctx.meta
        xp = ctx.saved_tensors[0]
        g = _check_c8(g, "conv_first_c8 backward")
        L, dev = _lib.lib(), g.device
        with _OnDevice(dev):
            s = _stream(dev)
            if act != _lib.ACT_NONE:
                gp = torch.empty_like(g)
                _lib.check(L.rvsr_c8_act_bwd(_p(g), _p(ctx.saved_tensors[1]), _p(gp), g.numel(), act, s), "c8_act_bwd")
            else:
                gp = g
            needs = ctx.needs_input_grad
            gw, gb = _wgrad_alloc(1, Cout, 3, needs[2], dev)
            res = []

            def outs():
                res.append(gw[:, :C].contiguous())
                return ([(ctx.params[0], res[0])] if needs[1] else []) + ([(ctx.params[1], gb)] if needs[2] else [])
            if _wgrad_launch([(xp, gp, gw, 64, 0, gb)], N, H, W, Cout, 3, ctx.params, outs):
                return None, None, None, None
            outs()
        return None, (res[0] if needs[1] else None), gb, None


def conv_first(x, weight, bias=None, act=None):
    """First convolution of the network on an NCHW image (<= 16 channels): C8 bf16 output, weight / bias gradients."""
    return _ConvFirstC8.apply(x, weight, bias, ACT[act])


class _ConvPairC8(torch.autograd.Function):
    """y = act2(conv2(act1(conv1(cat(xs))))) [+ xs[0]]: two chained 64-channel 3x3 convolutions whose intermediate has no other
    consumer -- ResidualBlock_noBN (arch_util.py:135-139, skip = True) and the conv -> conv pairs of PCD_Align
    (EDVR_arch.py:104-105, :112-113, :121-122, :128-129).  Knowing that, the backward fuses what autograd would run as separate
    kernels: the gradient through act1 rides in the epilogue of conv2's data gradient (mask mode), and the skip connection's
    gradient in the epilogue of conv1's data gradient (residual add)."""

    @staticmethod
    def forward(ctx, w1, b1, w2, b2, act1, act2, skip, *xs):
        xs = [x if _src_ok(x) else x.contiguous() for x in xs]
        N, C8, H, W, _ = xs[0].shape
        if C8 != 8 or tuple(w1.shape) != (64, 64 * len(xs), 3, 3) or tuple(w2.shape) != (64, 64, 3, 3) or act1 == _lib.ACT_NONE:
            raise NotImplementedError("conv_pair_c8: 64-channel 3x3 convolutions with an activation in between")
        if skip and (act2 != _lib.ACT_NONE or len(xs) != 1):
            raise NotImplementedError("conv_pair_c8: the skip connection is ResidualBlock_noBN's (one input, no final activation)")
        ctx.params = (w1, b1, w2, b2)
        w1, w2 = w1.contiguous(), w2.contiguous()
        needs = ctx.needs_input_grad
        with _OnDevice(w1.device):
            p1 = _pack_weights(w1, [(64, 64 * len(xs), 3, False, 0, 64 * len(xs), 0, (len(xs), 64, N, H, W))] +
                               _dgrad_views(len(xs), 64, 3, N, H, W, needs[7:7 + len(xs)]))
            p2 = _pack_weights(w2, [(64, 64, 3, False, 0, 64, 0, (1, 64, N, H, W))] + _dgrad_views(1, 64, 3, N, H, W, [True]))
            h = _conv_launch(xs, p1[0], b1, None, N, H, W, 64, 64, 3, act1, False)
            y = _conv_launch([h], p2[0], b2, xs[0] if skip else None, N, H, W, 64, 64, 3, act2, False)
        ctx.packs = (p1[1:], p2[1])
        ctx.meta = (act1, act2, skip, len(xs), N, H, W)
        ctx.save_for_backward(w1, w2, h, *xs, *([y] if act2 != _lib.ACT_NONE else []))
        return y

    @staticmethod
    def backward(ctx, g):
        act1, act2, skip, nsrc, N, H, W = ctx.meta
        sv = ctx.saved_tensors
        w1, w2, h, xs = sv[0], sv[1], sv[2], sv[3:3 + nsrc]
        g = _check_c8(g, "conv_pair_c8 backward")
        needs = ctx.needs_input_grad
        L, dev = _lib.lib(), g.device
        with _OnDevice(dev):
            if act2 != _lib.ACT_NONE:
                g2 = torch.empty_like(g)
                _lib.check(L.rvsr_c8_act_bwd(_p(g), _p(sv[-1]), _p(g2), g.numel(), act2, _stream(dev)), "c8_act_bwd")
            else:
                g2 = g
            packs1, pack2 = iter(ctx.packs[0]), ctx.packs[1]
            gh = _dgrad(w2, g2, 0, 1, 64, 3, N, H, W, mask=h, slope=_SLOPE[act1], wp=pack2)  # gradient of conv1's pre-activation output
            gw1 = gb1 = gw2 = gb2 = None
            jobs = []
            if needs[2] or needs[3]:
                gw2, gb2 = _wgrad_alloc(1, 64, 3, needs[3], dev)
                jobs.append((h, g2, gw2, 64, 0, gb2))
            if needs[0] or needs[1]:
                gw1, gb1 = _wgrad_alloc(nsrc, 64, 3, needs[1], dev)
                jobs += [(x, gh, gw1, nsrc * 64, i * 64, gb1 if i == 0 else None) for i, x in enumerate(xs)]
            if jobs:  # both convolutions' weight gradients: one launch
                prm = ctx.params
                outs = [(prm[k], t) for k, t in enumerate((gw1, gb1, gw2, gb2)) if needs[k] and t is not None]
                if _wgrad_launch(jobs, N, H, W, 64, 3, prm, lambda: outs):
                    gw1 = gb1 = gw2 = gb2 = None
            gxs = [None] * nsrc
            for i in range(nsrc):
                if needs[7 + i]:
                    gxs[i] = _dgrad(w1, gh, i, nsrc, 64, 3, N, H, W, residual=g if (skip and i == 0) else None, wp=next(packs1))
        return (gw1 if needs[0] else None, gb1, gw2 if needs[2] else None, gb2, None, None, None, *gxs)


def conv_pair(xs, w1, b1, act1, w2, b2, act2=None, skip=False):
    """act2(conv2(act1(conv1(cat(xs))))) [+ xs]: see _ConvPairC8."""
    if isinstance(xs, torch.Tensor):
        xs = [xs]
    return _ConvPairC8.apply(w1, b1, w2, b2, ACT[act1], ACT[act2], bool(skip), *xs)


def conv(xs, weight, bias=None, act=None, residual=None, shuffle=False):
    """One nn.Conv2d site (3x3 / 1x1, stride 1) on C8 tensors; `xs` is a tensor or the list torch.cat would have joined."""
    if isinstance(xs, torch.Tensor):
        xs = [xs]
    return _ConvC8.apply(weight, bias, residual, ACT[act], bool(shuffle), *xs)


# ---------------------------------------------------------------- modulated deformable convolution pack
class _DcnPackC8(torch.autograd.Function):
    """y = act(dcn(x, offset, sigmoid(mask)) + bias) with offset / mask taken from `om`, the 256-channel C8 output of the
    conv_offset_mask convolution (ModulatedDeformConvPack.forward, deform_conv.py:274-292)."""

    @staticmethod
    def forward(ctx, x, om, weight, bias, act):
        x, om = _check_c8(x, "dcn_pack_c8 x"), _check_c8(om, "dcn_pack_c8 om")
        N, P, H, W, _ = x.shape
        if P != 8 or om.shape[1] != 32 or tuple(weight.shape) != (64, 64, 3, 3) or weight.dtype != torch.float32:
            raise NotImplementedError("dcn_pack_c8: built for 64 -> 64 channels, 3x3, 8 deformable groups, fp32 parameters")
        weight = weight.contiguous()
        y = torch.empty_like(x)
        L = _lib.lib()
        with _OnDevice(x.device):
            ws = torch.empty(L.rvsr_c8_mdcn_workspace_bytes(N, H, W, 0), dtype=torch.uint8, device=x.device)
            _lib.check(L.rvsr_c8_mdcn_fwd(_p(x), _p(om), _p(weight), _p(bias), _p(y), N, H, W, act, _p(ws), ws.numel(), _stream(x.device)),
                       "c8_mdcn_fwd")
        ctx.act, ctx.with_bias = act, bias is not None
        ctx.save_for_backward(x, om, weight, *([y] if act != _lib.ACT_NONE else []))
        return y

    @staticmethod
    def backward(ctx, g):
        x, om, weight = ctx.saved_tensors[:3]
        y = ctx.saved_tensors[3] if ctx.act != _lib.ACT_NONE else None
        g = _check_c8(g, "dcn_pack_c8 backward")
        N, _, H, W, _ = x.shape
        gx, gom = torch.empty_like(x), torch.empty_like(om)
        gw = torch.empty_like(weight)
        gb = torch.empty(64, dtype=torch.float32, device=x.device) if ctx.with_bias else None
        L = _lib.lib()
        with _OnDevice(x.device):
            ws = torch.empty(L.rvsr_c8_mdcn_workspace_bytes(N, H, W, 1), dtype=torch.uint8, device=x.device)
            _lib.check(L.rvsr_c8_mdcn_bwd(_p(x), _p(om), _p(weight), _p(g), _p(y), _p(gx), _p(gom), _p(gw), _p(gb), N, H, W, ctx.act,
                                          _p(ws), ws.numel(), _stream(x.device)), "c8_mdcn_bwd")
        return gx, gom, gw, gb, None


def dcn_pack(x, om, weight, bias=None, act=None):
    return _DcnPackC8.apply(x, om, weight, bias, ACT[act])


# ---------------------------------------------------------------- TSA temporal attention
class _TsaTemporalC8(torch.autograd.Function):
    """(aligned, emb: [B * N, 8, H, W, 8]; emb_ref: [B, 8, H, W, 8]) -> N tensors aligned[b, n] * sigmoid(<emb[b, n], emb_ref[b]>)
    (TSA_Fusion.forward, EDVR_arch.py:170-181): the N inputs of the 1x1 fusion convolutions, each [B, 8, H, W, 8]."""

    @staticmethod
    def forward(ctx, aligned, emb, emb_ref, N):
        aligned, emb, emb_ref = _check_c8(aligned, "tsa aligned"), _check_c8(emb, "tsa emb"), _check_c8(emb_ref, "tsa emb_ref")
        B, P, H, W, _ = emb_ref.shape
        if P != 8 or tuple(aligned.shape) != (B * N, 8, H, W, 8) or aligned.shape != emb.shape:
            raise NotImplementedError("tsa_temporal_c8: 64 channels, aligned / emb [B * N, 8, H, W, 8], emb_ref [B, 8, H, W, 8]")
        outs = [torch.empty_like(emb_ref) for _ in range(N)]
        prob = torch.empty((B, N, H, W), dtype=torch.float32, device=emb.device)
        ptrs = (ctypes.c_void_p * N)(*[o.data_ptr() for o in outs])
        with _OnDevice(emb.device):
            _lib.check(_lib.lib().rvsr_c8_tsa_temporal(_p(aligned), _p(emb), _p(emb_ref), ptrs, _p(prob), B, N, 64, H, W, _stream(emb.device)),
                       "c8_tsa_temporal")
        ctx.N = N
        ctx.save_for_backward(aligned, emb, emb_ref, prob)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        aligned, emb, emb_ref, prob = ctx.saved_tensors
        N = ctx.N
        B, _, H, W, _ = emb_ref.shape
        gouts = [None if g is None else _check_c8(g, "tsa backward") for g in gouts]
        ptrs = (ctypes.c_void_p * N)(*[0 if g is None else g.data_ptr() for g in gouts])
        g_aligned, g_emb, g_emb_ref = torch.empty_like(aligned), torch.empty_like(emb), torch.empty_like(emb_ref)
        with _OnDevice(emb.device):
            _lib.check(_lib.lib().rvsr_c8_tsa_temporal_bwd(ptrs, _p(aligned), _p(emb), _p(emb_ref), _p(prob), _p(g_aligned), _p(g_emb),
                                                           _p(g_emb_ref), B, N, 64, H, W, _stream(emb.device)), "c8_tsa_temporal_bwd")
        return g_aligned, g_emb, g_emb_ref, None


def tsa_temporal(aligned, emb, emb_ref, N):
    return list(_TsaTemporalC8.apply(aligned, emb, emb_ref, N))


class _PoolMaxAvgC8(torch.autograd.Function):
    """(MaxPool2d(3, 2, 1)(x), AvgPool2d(3, 2, 1)(x)) of a C8 tensor in one pass (TSA_Fusion, EDVR_arch.py:154-155, :187, :191)."""

    @staticmethod
    def forward(ctx, x):
        x = _check_c8(x, "pool_maxavg_c8")
        N, P, H, W, _ = x.shape
        shape = (N, P, (H - 1) // 2 + 1, (W - 1) // 2 + 1, 8)
        mx, av = torch.empty(shape, dtype=torch.bfloat16, device=x.device), torch.empty(shape, dtype=torch.bfloat16, device=x.device)
        with _OnDevice(x.device):
            _lib.check(_lib.lib().rvsr_c8_pool_maxavg(_p(x), _p(mx), _p(av), N * P, H, W, _stream(x.device)), "c8_pool_maxavg")
        ctx.save_for_backward(x)
        return mx, av

    @staticmethod
    def backward(ctx, g_max, g_avg):
        (x,) = ctx.saved_tensors
        N, P, H, W, _ = x.shape
        g_max = None if g_max is None else _check_c8(g_max, "pool backward")
        g_avg = None if g_avg is None else _check_c8(g_avg, "pool backward")
        gx = torch.empty_like(x)
        with _OnDevice(x.device):
            _lib.check(_lib.lib().rvsr_c8_pool_maxavg_bwd(_p(x), _p(g_max), _p(g_avg), _p(gx), N * P, H, W, _stream(x.device)),
                       "c8_pool_maxavg_bwd")
        return gx


def pool_maxavg(x):
    return list(_PoolMaxAvgC8.apply(x))


class _TsaFinalC8(torch.autograd.Function):
    """fea * sigmoid(att) * 2 + att_add (TSA_Fusion, EDVR_arch.py:206-207)."""

    @staticmethod
    def forward(ctx, fea, att, att_add):
        fea, att, att_add = _check_c8(fea, "tsa_final fea"), _check_c8(att, "tsa_final att"), _check_c8(att_add, "tsa_final att_add")
        if fea.shape != att.shape or fea.shape != att_add.shape:
            raise RuntimeError("tsa_final_c8: shapes differ")
        out = torch.empty_like(fea)
        with _OnDevice(fea.device):
            _lib.check(_lib.lib().rvsr_c8_tsa_final(_p(fea), _p(att), _p(att_add), _p(out), fea.numel(), _stream(fea.device)), "c8_tsa_final")
        ctx.save_for_backward(fea, att)
        return out

    @staticmethod
    def backward(ctx, g):
        fea, att = ctx.saved_tensors
        g = _check_c8(g, "tsa_final backward")
        g_fea, g_att = torch.empty_like(fea), torch.empty_like(att)
        with _OnDevice(fea.device):
            _lib.check(_lib.lib().rvsr_c8_tsa_final_bwd(_p(g), _p(fea), _p(att), _p(g_fea), _p(g_att), fea.numel(), _stream(fea.device)),
                       "c8_tsa_final_bwd")
        return g_fea, g_att, g


def tsa_final(fea, att, att_add):
    return _TsaFinalC8.apply(fea, att, att_add)


# ---------------------------------------------------------------- x2 bilinear upsample (optionally scaled)
class _Up2C8(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scale):
        x = _check_c8(x, "upsample2x_c8")
        N, P, H, W, _ = x.shape
        ctx.scale = float(scale)
        y = torch.empty((N, P, 2 * H, 2 * W, 8), dtype=torch.bfloat16, device=x.device)
        with _OnDevice(x.device):
            _lib.check(_lib.lib().rvsr_c8_upsample2x(_p(x), _p(y), N * P, H, W, ctx.scale, 0, _stream(x.device)), "c8_upsample2x")
        return y

    @staticmethod
    def backward(ctx, g):
        g = _check_c8(g, "upsample2x_c8 backward")
        N, P, H2, W2, _ = g.shape
        gx = torch.empty((N, P, H2 // 2, W2 // 2, 8), dtype=torch.bfloat16, device=g.device)
        with _OnDevice(g.device):
            _lib.check(_lib.lib().rvsr_c8_upsample2x(_p(g), _p(gx), N * P, H2 // 2, W2 // 2, ctx.scale, 1, _stream(g.device)),
                       "c8_upsample2x (adjoint)")
        return gx, None


def upsample2x(x, scale=1.0):
    """scale * F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False) on a C8 tensor."""
    return _Up2C8.apply(x, scale)


# ---------------------------------------------------------------- whole-step CUDA graph
class GraphedStep:
    """forward + loss + backward of `net` captured ONCE into a CUDA graph and replayed per step.

    A training step of this path is ~1000 small launches issued from Python; eager, the host cannot feed the GPU fast enough
    (the step is launch-bound).  Everything on the path is capture-safe: kernels go to torch's current stream, no host
    synchronisation, buffers come from torch's (graph-private) allocator, weights are re-packed INSIDE the graph so a replay
    sees the optimizer's latest values.  Usage:

        step = GraphedStep(net, torch.nn.functional.l1_loss, x_example, target_example)
        for x, target in loader:          # fixed shapes
            loss = step(x, target)        # p.grad of every parameter now holds this step's gradient (overwritten, not summed)
            optimizer.step()

    The usual whole-network capture rules apply (torch.cuda.graphs): static shapes, optimizer.zero_grad(set_to_none=True) must
    NOT be called between replays (the graph owns the .grad tensors), and no autograd graph of an earlier eager step over the
    same parameters may still be alive at construction (drop such a `loss` first: it pins the parameters' gradient accumulators
    to the stream that step ran on, which invalidates the capture)."""

    def __init__(self, net, loss_fn, x, target, amp_dtype=torch.bfloat16, warmup=3, overlap_wgrad=None, pack_ahead=None):
        if not x.is_cuda:
            raise NotImplementedError("GraphedStep: CUDA tensors only")
        self.net, self.loss_fn, self.amp_dtype = net, loss_fn, amp_dtype
        # weight-gradient launches as a parallel branch of the graph (see _Defer); RVSR_WGRAD_OVERLAP=0 keeps one stream
        self.overlap_wgrad = (os.environ.get("RVSR_WGRAD_OVERLAP", "1") != "0") if overlap_wgrad is None else bool(overlap_wgrad)
        self.overlap_pack = (os.environ.get("RVSR_PACK_AHEAD", "1") != "0") if pack_ahead is None else bool(pack_ahead)
        self.x, self.target = x.detach().clone(), target.detach().clone()
        side = torch.cuda.Stream(device=x.device)
        side.wait_stream(torch.cuda.current_stream(x.device))
        with torch.cuda.stream(side):   # warm-up on a side stream: lazy initialisation (cuDNN plans, smem attributes) happens here
            for _ in range(warmup):
                net.zero_grad(set_to_none=True)
                self._fwd_bwd()
        torch.cuda.current_stream(x.device).wait_stream(side)
        net.zero_grad(set_to_none=True)
        self.graph = torch.cuda.CUDAGraph()
        prio = int(os.environ.get("RVSR_GRAPH_PRIO", "1"))   # main branch at high priority: the side branches fill its gaps
        kw = dict(stream=torch.cuda.Stream(device=x.device, priority=-1)) if prio else {}
        _Defer.LAG = int(os.environ.get("RVSR_WGRAD_LAG", "3"))
        with torch.cuda.graph(self.graph, **kw):
            self.loss = self._fwd_bwd()

    def _fwd_bwd(self):
        _Defer.begin(self.x.device, self.overlap_wgrad, self.overlap_pack)
        try:
            with torch.autocast("cuda", dtype=self.amp_dtype, enabled=self.amp_dtype is not None):
                loss = self.loss_fn(self.net(self.x).float(), self.target)
            _Defer.begin_backward(self.x.device)
            loss.backward()
        finally:
            _Defer.finish(self.x.device)
        return loss.detach()

    def __call__(self, x, target):
        self.x.copy_(x)
        self.target.copy_(target)
        self.graph.replay()
        return self.loss

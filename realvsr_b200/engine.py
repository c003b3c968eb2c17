"""Python handle on the C++ EDVR inference engine (rvsr_engine_* in include/rvsr_b200.h).

PyTorch is used only for device memory and the current CUDA stream; all compute is the
library's own kernels.  One engine per (device, precision); weights are handed over as
fp32 device tensors under the reference's state_dict key names.
"""
import ctypes

import torch

from . import _lib

_PRECISION = {"fp32": _lib.F32, "fp16": _lib.F16}


def _dt(t):
    if t.dtype == torch.float32:
        return _lib.F32
    if t.dtype == torch.float16:
        return _lib.F16
    raise RuntimeError("realvsr_b200 engine: unsupported dtype %s" % t.dtype)


class EDVREngine:
    def __init__(self, nf=64, nc=3, nframes=5, groups=8, front_RBs=5, back_RBs=10, center=None, predeblur=False,
                 HR_in=False, w_TSA=True, upsample=True, precision="fp16", device=None):
        self.L = _lib.lib()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.cfg = _lib.EdvrConfig(nf, nc, nframes, groups, front_RBs, back_RBs, -1 if center is None else center,
                                   int(bool(predeblur)), int(bool(HR_in)), int(bool(w_TSA)), int(bool(upsample)),
                                   _PRECISION[precision])
        self.precision = precision
        self.scale = 4 if (upsample and not HR_in) else 1   # HR_in: frames arrive at the output resolution
        h = ctypes.c_void_p()
        _lib.check(self.L.rvsr_engine_create(ctypes.byref(self.cfg), ctypes.byref(h)), "engine_create")
        self.h = h
        self._ws = None
        self._staging = {}
        self.loaded = False

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.rvsr_engine_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    def weight_names(self):
        n = self.L.rvsr_engine_num_weights(self.h)
        return [self.L.rvsr_engine_weight_name(self.h, i).decode() for i in range(n)]

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def load_state_dict(self, sd, strict=True):
        """Strict like the reference's test scripts (test_RealVSR_wi_GT.py:77): unexpected or
        missing keys raise RuntimeError.  A leading 'module.' (DataParallel) is stripped like
        base_model.load_network does (base_model.py:104-114)."""
        known = set(self.weight_names())
        with torch.cuda.device(self.device):
            for k, v in sd.items():
                name = k[7:] if k.startswith("module.") else k
                if not strict and name not in known:
                    continue  # strict=False skips unknown keys only; size mismatches and CUDA errors still raise
                t = v.detach().to(device=self.device, dtype=torch.float32).contiguous()
                shape = (ctypes.c_int64 * t.dim())(*t.shape)
                _lib.check(self.L.rvsr_engine_set_weight(self.h, name.encode(), ctypes.c_void_p(t.data_ptr()),
                                                         shape, t.dim(), self._stream()), "load_state_dict")
            _lib.check(self.L.rvsr_engine_finalize(self.h, self._stream()), "load_state_dict")
            torch.cuda.current_stream(self.device).synchronize()  # source tensors may be temporaries
        self.loaded = True

    # ------------------------------------------------------------------ forward
    def _workspace(self, B, H, W):
        need = self.L.rvsr_engine_workspace_bytes(self.h, B, H, W)
        if need == 0:
            _lib.check(_lib.E_INVALID, "workspace_bytes")
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def forward(self, x, out_dtype=None, out=None):
        """x: [B, N, C, H, W] CUDA tensor (fp32 or fp16) -> [B, C, sH, sW] of out_dtype (default x.dtype);
        `out`: optional preallocated result tensor."""
        if not x.is_cuda:
            raise NotImplementedError("realvsr_b200 engine runs on CUDA tensors only (no CPU fallback)")
        if x.dim() != 5 or x.shape[1] != self.cfg.nframes or x.shape[2] != self.cfg.nc:
            raise RuntimeError("expected input [B, %d, %d, H, W], got %s" % (self.cfg.nframes, self.cfg.nc,
                                                                               tuple(x.shape)))
        x = x.contiguous()
        B, _, _, H, W = x.shape
        if out is None:
            out = torch.empty(B, self.cfg.nc, H * self.scale, W * self.scale, device=x.device,
                              dtype=out_dtype or x.dtype)
        elif tuple(out.shape) != (B, self.cfg.nc, H * self.scale, W * self.scale) or not out.is_cuda or not out.is_contiguous():
            raise RuntimeError("forward: `out` must be a contiguous CUDA tensor of shape %s" %
                               ((B, self.cfg.nc, H * self.scale, W * self.scale),))
        with torch.cuda.device(self.device):
            ws = self._workspace(B, H, W)
            _lib.check(self.L.rvsr_engine_forward(self.h, ctypes.c_void_p(x.data_ptr()), _dt(x),
                                                  ctypes.c_void_p(out.data_ptr()), _dt(out), B, H, W,
                                                  ctypes.c_void_p(ws.data_ptr()), ws.numel(), self._stream()),
                       "engine_forward")
        return out

    __call__ = forward

    def forward_host(self, x_host, out_host=None):
        """Host tensors in, host tensors out (pinned for async copies): H2D + forward + D2H on
        the current stream, then a stream synchronise -- what the reference's
        util.single_forward (utils/util.py:222-237) does around the model."""
        if x_host.is_cuda or x_host.dim() != 5 or not x_host.is_contiguous():
            raise RuntimeError("forward_host: x_host must be a contiguous host tensor [B, N, C, H, W]")
        B, _, _, H, W = x_host.shape
        if out_host is not None and (out_host.is_cuda or not out_host.is_contiguous() or out_host.dtype != x_host.dtype or
                                     tuple(out_host.shape) != (B, self.cfg.nc, H * self.scale, W * self.scale)):
            raise RuntimeError("forward_host: out_host must be a contiguous host tensor of x_host's dtype and shape %s" %
                               ((B, self.cfg.nc, H * self.scale, W * self.scale),))
        key = (tuple(x_host.shape), x_host.dtype)
        if key not in self._staging:
            self._staging.clear()
            self._staging[key] = (torch.empty(x_host.shape, dtype=x_host.dtype, device=self.device),
                                  torch.empty(B, self.cfg.nc, H * self.scale, W * self.scale,
                                              dtype=x_host.dtype, device=self.device))
        din, dout = self._staging[key]
        if out_host is None:
            out_host = torch.empty(dout.shape, dtype=dout.dtype, pin_memory=True)
        with torch.cuda.device(self.device):
            ws = self._workspace(B, H, W)
            _lib.check(self.L.rvsr_engine_forward_host(
                self.h, ctypes.c_void_p(x_host.data_ptr()), _dt(x_host), ctypes.c_void_p(out_host.data_ptr()),
                _dt(out_host), B, H, W, ctypes.c_void_p(din.data_ptr()), ctypes.c_void_p(dout.data_ptr()),
                ctypes.c_void_p(ws.data_ptr()), ws.numel(), self._stream()), "engine_forward_host")
            torch.cuda.current_stream(self.device).synchronize()
        return out_host

    def host_pipeline(self, depth=2):
        """Pipelined variant of forward_host for streams of windows: see HostPipeline."""
        return HostPipeline(self, depth)

    # ------------------------------------------------------------------ sliding-window feature cache
    def make_cache(self, n_slots, H, W):
        """Device buffer for the feature pyramids of n_slots frames (rvsr_engine_cache_bytes)."""
        nbytes = self.L.rvsr_engine_cache_bytes(self.h, n_slots, H, W)
        if nbytes == 0:
            _lib.check(_lib.E_INVALID, "cache_bytes (H and W must be multiples of 4)")
        return torch.empty(nbytes, dtype=torch.uint8, device=self.device)

    def extract_features(self, frames, cache, n_slots, slot0=0):
        """frames [F, C, H, W] (CUDA) -> cache slots [slot0, slot0 + F)."""
        frames = frames.contiguous()
        F_, _, H, W = frames.shape
        with torch.cuda.device(self.device):
            need = self.L.rvsr_engine_extract_workspace_bytes(self.h, F_, H, W)
            if self._ws is None or self._ws.numel() < need:
                self._ws = None
                self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            _lib.check(self.L.rvsr_engine_extract_features(
                self.h, ctypes.c_void_p(frames.data_ptr()), _dt(frames), F_, H, W, ctypes.c_void_p(cache.data_ptr()),
                n_slots, slot0, ctypes.c_void_p(self._ws.data_ptr()), self._ws.numel(), self._stream()),
                "extract_features")

    def forward_cached(self, cache, n_slots, window_slots, frames, out_dtype=None):
        """window_slots: B lists of nframes slot indices; frames: the LQ frames of all slots
        [n_slots, C, H, W] (CUDA).  -> [B, C, sH, sW]."""
        B = len(window_slots)
        flat = [int(v) for w in window_slots for v in w]
        if len(flat) != B * self.cfg.nframes:
            raise RuntimeError("forward_cached: every window needs %d slot indices" % self.cfg.nframes)
        frames = frames.contiguous()
        _, _, H, W = frames.shape
        out = torch.empty(B, self.cfg.nc, H * self.scale, W * self.scale, device=frames.device,
                          dtype=out_dtype or frames.dtype)
        arr = (ctypes.c_int * len(flat))(*flat)
        with torch.cuda.device(self.device):
            ws = self._workspace(B, H, W)
            _lib.check(self.L.rvsr_engine_forward_cached(
                self.h, ctypes.c_void_p(cache.data_ptr()), n_slots, arr, ctypes.c_void_p(frames.data_ptr()), _dt(frames),
                ctypes.c_void_p(out.data_ptr()), _dt(out), B, H, W, ctypes.c_void_p(ws.data_ptr()), ws.numel(),
                self._stream()), "forward_cached")
        return out

    def last_launch_count(self):
        return self.L.rvsr_engine_last_launch_count(self.h)

    def profile(self, x, steps=3):
        """Run `steps` forwards with per-launch CUDA-event timing; returns a list of dicts
        {label, ms, flops, bytes} (ms averaged over the steps), in launch order."""
        rows = None
        _lib.check(self.L.rvsr_engine_set_profiling(self.h, 1))
        try:
            for _ in range(steps):
                self.forward(x)
                n = self.L.rvsr_engine_profile_collect(self.h)
                if n < 0:
                    _lib.check(n, "profile_collect")
                cur = []
                buf = ctypes.create_string_buffer(128)
                ms, fl, by = ctypes.c_float(), ctypes.c_double(), ctypes.c_double()
                for i in range(n):
                    _lib.check(self.L.rvsr_engine_profile_entry(self.h, i, buf, 128, ctypes.byref(ms),
                                                                ctypes.byref(fl), ctypes.byref(by)))
                    cur.append(dict(label=buf.value.decode(), ms=ms.value, flops=fl.value, bytes=by.value))
                if rows is None:
                    rows = cur
                else:
                    for r, c in zip(rows, cur):
                        r["ms"] += c["ms"]
            for r in rows:
                r["ms"] /= steps
        finally:
            self.L.rvsr_engine_set_profiling(self.h, 0)
        return rows

    def read_tap(self, name, shape):
        dst = torch.empty(shape, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.L.rvsr_engine_read_tap(self.h, name.encode(), ctypes.c_void_p(dst.data_ptr()),
                                                   dst.numel(), self._stream()), "read_tap")
        return dst


class HostPipeline:
    """Host tensors in, host tensors out, for a STREAM of window batches: the H2D copy of batch i + 1 and the D2H
    copy of batch i - 1 run on their own CUDA streams while batch i is computed (the reference's test loop,
    test_RealVSR_wi_GT.py:114-119 + utils/util.py:222-237, copies, computes and reads back strictly in turn).

        pipe = engine.host_pipeline()
        for x_host, out_host in batches:      # pinned host tensors
            pipe.submit(x_host, out_host)     # returns a ticket at once; wait_input(t): x_host reusable, wait(t): out_host valid
        pipe.drain()

    `depth` device staging buffers per direction; submit() blocks the host only when all of them are in flight."""

    def __init__(self, engine, depth=2):
        self.e, self.depth, self.i = engine, max(2, int(depth)), 0
        dev = engine.device
        self.s_in, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        self.slots = None
        self.key = None

    def _alloc(self, x_host):
        e = self.e
        B, _, _, H, W = x_host.shape
        self.key = (tuple(x_host.shape), x_host.dtype)
        self.slots = [dict(din=torch.empty(x_host.shape, dtype=x_host.dtype, device=e.device),
                           dout=torch.empty(B, e.cfg.nc, H * e.scale, W * e.scale, dtype=x_host.dtype, device=e.device),
                           ev_in=torch.cuda.Event(), ev_done=torch.cuda.Event(), ev_out=torch.cuda.Event(), busy=False)
                      for _ in range(self.depth)]

    @staticmethod
    def _check_host(t, what, dtype=None, shape=None):
        if t.is_cuda or not t.is_contiguous() or not t.is_pinned():
            raise RuntimeError("HostPipeline: %s must be a contiguous pinned host tensor" % what)
        if dtype is not None and t.dtype != dtype:
            raise RuntimeError("HostPipeline: %s must be %s (got %s)" % (what, dtype, t.dtype))
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise RuntimeError("HostPipeline: %s must have shape %s (got %s)" % (what, tuple(shape), tuple(t.shape)))

    def submit(self, x_host, out_host):
        """Queue one batch.  Returns a ticket for wait_input(): x_host may be refilled once wait_input(ticket)
        has returned (its H2D copy has left the host buffer); out_host is valid after wait(ticket) / drain()."""
        e = self.e
        if x_host.dim() != 5:
            raise RuntimeError("HostPipeline: x_host must be [B, N, C, H, W]")
        B, _, _, H, W = x_host.shape
        self._check_host(x_host, "x_host")
        self._check_host(out_host, "out_host", x_host.dtype, (B, e.cfg.nc, H * e.scale, W * e.scale))
        if self.key != (tuple(x_host.shape), x_host.dtype):
            self.drain()
            self._alloc(x_host)
        sl = self.slots[self.i % self.depth]
        self.i += 1
        compute = torch.cuda.current_stream(self.e.device)
        if sl["busy"]:
            sl["ev_out"].synchronize()      # the slot's previous result has left the device
        with torch.cuda.stream(self.s_in):
            sl["din"].copy_(x_host, non_blocking=True)
            sl["ev_in"].record(self.s_in)
        compute.wait_event(sl["ev_in"])
        self.e.forward(sl["din"], out=sl["dout"])
        sl["ev_done"].record(compute)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(sl["ev_done"])
            out_host.copy_(sl["dout"], non_blocking=True)
            sl["ev_out"].record(self.s_out)
        sl["busy"] = True
        return self.i - 1

    def wait_input(self, ticket):
        """Block until the H2D copy of submit() number `ticket` has read its x_host (safe to refill it)."""
        if self.slots and self.i - ticket <= self.depth:
            self.slots[ticket % self.depth]["ev_in"].synchronize()

    def wait(self, ticket):
        """Block until the result of submit() number `ticket` has landed in its out_host."""
        if self.slots and self.i - ticket <= self.depth:
            self.slots[ticket % self.depth]["ev_out"].synchronize()

    def drain(self):
        if self.slots:
            for sl in self.slots:
                if sl["busy"]:
                    sl["ev_out"].synchronize()
                    sl["busy"] = False

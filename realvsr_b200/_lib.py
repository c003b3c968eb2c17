"""ctypes binding of the C ABI in include/rvsr_b200.h.

There is deliberately NO fallback: if the CUDA library is missing or fails to load, every
operator raises.  (The CPU oracle under oracle/ is test infrastructure and is never
imported from here.)
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librvsr_b200.so")

F32, F16, BF16 = 0, 1, 2
OK, E_INVALID, E_CUDA, E_WORKSPACE, E_STATE, E_UNSUPPORTED = 0, -1, -2, -3, -4, -5
ACT_NONE, ACT_LRELU, ACT_RELU = 0, 1, 2

c_int, c_size_t, c_void_p, c_char_p = ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_char_p


class EdvrConfig(ctypes.Structure):
    _fields_ = [(n, c_int) for n in ("nf", "nc", "nframes", "groups", "front_RBs", "back_RBs", "center",
                                     "predeblur", "HR_in", "w_TSA", "upsample", "precision")]


# name -> (restype, argtypes); exactly the symbols include/rvsr_b200.h declares
_I12 = [c_int] * 12
SIGNATURES = {
    "rvsr_version": (c_int, []),
    "rvsr_last_error": (c_char_p, []),
    "rvsr_device_ok": (c_int, []),
    "rvsr_mdcn_fwd_workspace_bytes": (c_size_t, _I12 + [c_int]),
    "rvsr_mdcn_fwd": (c_int, [c_void_p] * 6 + _I12 + [c_int, c_void_p, c_size_t, c_void_p]),
    "rvsr_mdcn_bwd_workspace_bytes": (c_size_t, _I12 + [c_int]),
    "rvsr_mdcn_bwd": (c_int, [c_void_p] * 10 + _I12 + [c_int, c_void_p, c_size_t, c_void_p]),
    "rvsr_mdcn_pack_fwd_workspace_bytes": (c_size_t, [c_int] * 7),
    "rvsr_mdcn_pack_fwd": (c_int, [c_void_p] * 7 + [c_int] * 8 + [c_void_p, c_size_t, c_void_p]),
    "rvsr_conv2d_fwd_workspace_bytes": (c_size_t, [c_int] * 8),
    "rvsr_conv2d_fwd": (c_int, [c_void_p] * 6 + [c_int] * 12 + [c_void_p, c_size_t, c_void_p]),
    "rvsr_c8_from_nchw": (c_int, [c_void_p, c_int, c_void_p] + [c_int] * 5 + [c_void_p]),
    "rvsr_c8_to_nchw": (c_int, [c_void_p, c_void_p, c_int] + [c_int] * 5 + [c_void_p]),
    "rvsr_c8_conv_weight_bytes": (c_size_t, [c_int] * 4),
    "rvsr_c8_conv_pack_weight": (c_int, [c_void_p, c_void_p] + [c_int] * 8 + [c_void_p]),
    "rvsr_c8_conv_pack_weights": (c_int, [c_void_p, c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_void_p), c_void_p]),
    "rvsr_c8_conv_layouts": (c_int, [c_int] * 8),
    "rvsr_c8_conv_fwd": (c_int, [ctypes.POINTER(c_void_p), ctypes.POINTER(ctypes.c_longlong), c_int, c_int, c_void_p,
                                 c_void_p, c_void_p, c_void_p] + [c_int] * 9 + [ctypes.c_float, c_void_p]),
    "rvsr_c8_conv_wgrad_workspace_bytes": (c_size_t, [c_int] * 5),
    "rvsr_c8_conv_wgrad": (c_int, [c_int, ctypes.POINTER(c_void_p), ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(c_void_p),
                                   ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p), ctypes.POINTER(c_int), ctypes.POINTER(c_int)] +
                           [c_int] * 6 + [c_void_p, c_size_t, c_void_p]),
    "rvsr_c8_act_bwd": (c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_longlong, c_int, c_void_p]),
    "rvsr_c8_unshuffle2_act_bwd": (c_int, [c_void_p, c_void_p, c_void_p] + [c_int] * 5 + [c_void_p]),
    "rvsr_c8_upsample2x": (c_int, [c_void_p, c_void_p, ctypes.c_longlong, c_int, c_int, ctypes.c_float, c_int, c_void_p]),
    "rvsr_c8_tsa_temporal": (c_int, [c_void_p, c_void_p, c_void_p, ctypes.POINTER(c_void_p), c_void_p] + [c_int] * 5 + [c_void_p]),
    "rvsr_c8_tsa_temporal_bwd": (c_int, [ctypes.POINTER(c_void_p)] + [c_void_p] * 7 + [c_int] * 5 + [c_void_p]),
    "rvsr_c8_pool_maxavg": (c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_longlong, c_int, c_int, c_void_p]),
    "rvsr_c8_pool_maxavg_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_longlong, c_int, c_int, c_void_p]),
    "rvsr_c8_tsa_final": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_longlong, c_void_p]),
    "rvsr_c8_tsa_final_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_longlong, c_void_p]),
    "rvsr_c8_mdcn_workspace_bytes": (c_size_t, [c_int] * 4),
    "rvsr_c8_mdcn_fwd": (c_int, [c_void_p] * 5 + [c_int] * 4 + [c_void_p, c_size_t, c_void_p]),
    "rvsr_c8_mdcn_bwd": (c_int, [c_void_p] * 9 + [c_int] * 4 + [c_void_p, c_size_t, c_void_p]),
    "rvsr_upsample2x_nchw": (c_int, [c_void_p, c_void_p, ctypes.c_longlong, c_int, c_int, ctypes.c_float, c_int, c_int, c_void_p]),
    "rvsr_frames_from_u8": (c_int, [c_void_p, c_void_p] + [c_int] * 6 + [c_void_p]),
    "rvsr_frames_to_u8": (c_int, [c_void_p, c_int, c_void_p] + [c_int] * 5 + [c_void_p]),
    "rvsr_engine_create": (c_int, [ctypes.POINTER(EdvrConfig), ctypes.POINTER(c_void_p)]),
    "rvsr_engine_destroy": (None, [c_void_p]),
    "rvsr_engine_set_weight": (c_int, [c_void_p, c_char_p, c_void_p, ctypes.POINTER(ctypes.c_int64), c_int,
                                       c_void_p]),
    "rvsr_engine_finalize": (c_int, [c_void_p, c_void_p]),
    "rvsr_engine_num_weights": (c_int, [c_void_p]),
    "rvsr_engine_weight_name": (c_char_p, [c_void_p, c_int]),
    "rvsr_engine_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int, c_int]),
    "rvsr_engine_forward": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                    c_size_t, c_void_p]),
    "rvsr_engine_forward_host": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int,
                                         c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "rvsr_engine_cache_bytes": (c_size_t, [c_void_p, c_int, c_int, c_int]),
    "rvsr_engine_extract_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int, c_int]),
    "rvsr_engine_extract_features": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int,
                                             c_void_p, c_size_t, c_void_p]),
    "rvsr_engine_forward_cached": (c_int, [c_void_p, c_void_p, c_int, ctypes.POINTER(c_int), c_void_p, c_int, c_void_p,
                                           c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "rvsr_engine_last_launch_count": (c_int, [c_void_p]),
    "rvsr_engine_set_profiling": (c_int, [c_void_p, c_int]),
    "rvsr_engine_profile_collect": (c_int, [c_void_p]),
    "rvsr_engine_profile_entry": (c_int, [c_void_p, c_int, c_char_p, c_int, ctypes.POINTER(ctypes.c_float),
                                          ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]),
    "rvsr_engine_read_tap": (c_int, [c_void_p, c_char_p, c_void_p, c_size_t, c_void_p]),
}

_lib = None


def lib():
    """The loaded library; raises if it has not been built (no silent fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "realvsr_b200: CUDA library %s is missing. Build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). "
                "There is no CPU or PyTorch fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError here = header/library drift
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def last_error():
    msg = lib().rvsr_last_error()
    return msg.decode(errors="replace") if msg else ""


def check(rc, what=""):
    """Map a negative return code to the exception the reference would have raised:
    TORCH_CHECK/AT_ERROR -> RuntimeError; an unbuilt feature -> NotImplementedError."""
    if rc == OK:
        return
    msg = "%s%s (rvsr rc=%d)" % (what + ": " if what else "", last_error(), rc)
    if rc == E_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)

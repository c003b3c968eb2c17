"""CPU tests of the boundary: the C-ABI library loads and exports exactly the symbols that
include/rvsr_b200.h declares; host-side argument validation works without a GPU."""
import ctypes
import os
import re

import pytest

from realvsr_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "rvsr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rvsr_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported_and_bound():
    names = _declared()
    assert len(names) >= 20
    L = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), "library does not export %s" % n
    assert sorted(_lib.SIGNATURES.keys()) == names, "python binding and header drifted"
    assert _lib.lib().rvsr_version() >= 100


def test_engine_config_validation_without_gpu():
    L = _lib.lib()
    h = ctypes.c_void_p()
    bad = _lib.EdvrConfig(60, 3, 5, 8, 5, 10, -1, 0, 0, 1, 1, _lib.F16)  # nf must be a multiple of 8
    assert L.rvsr_engine_create(ctypes.byref(bad), ctypes.byref(h)) == _lib.E_UNSUPPORTED
    assert b"nf" in L.rvsr_last_error()
    bad = _lib.EdvrConfig(16, 3, 5, 8, 5, 10, -1, 0, 0, 1, 0, _lib.F16)  # NoUp needs nf == 64
    assert L.rvsr_engine_create(ctypes.byref(bad), ctypes.byref(h)) == _lib.E_INVALID
    with pytest.raises(NotImplementedError):
        _lib.check(_lib.E_UNSUPPORTED)
    with pytest.raises(RuntimeError):
        _lib.check(_lib.E_INVALID)


def test_engine_state_dict_contract_names():
    from helpers import edvr_state_shapes
    L = _lib.lib()
    for up, kw in ((1, dict(nf=64, nframes=5, groups=8, front_RBs=5, back_RBs=10, w_TSA=True)),
                   (0, dict(nf=64, nframes=3, groups=8, front_RBs=5, back_RBs=10, w_TSA=False)),
                   (1, dict(nf=16, nframes=3, groups=4, front_RBs=1, back_RBs=1, w_TSA=True, predeblur=True, HR_in=True)),
                   (1, dict(nf=16, nframes=3, groups=4, front_RBs=1, back_RBs=1, w_TSA=True, predeblur=True)),
                   (1, dict(nf=16, nframes=3, groups=4, front_RBs=1, back_RBs=1, w_TSA=False, HR_in=True))):
        cfg = _lib.EdvrConfig(kw["nf"], 3, kw["nframes"], kw["groups"], kw["front_RBs"], kw["back_RBs"], -1,
                              int(kw.get("predeblur", False)), int(kw.get("HR_in", False)), int(kw["w_TSA"]), up, _lib.F32)
        h = ctypes.c_void_p()
        assert L.rvsr_engine_create(ctypes.byref(cfg), ctypes.byref(h)) == 0
        names = [L.rvsr_engine_weight_name(h, i).decode() for i in range(L.rvsr_engine_num_weights(h))]
        assert names == list(edvr_state_shapes("EDVR" if up else "EDVR_NoUp", **kw).keys())
        # workspace sizing is host-only arithmetic: 0 for bad dims, >0 and monotone otherwise
        assert L.rvsr_engine_workspace_bytes(h, 1, 30, 32) == 0
        if kw.get("HR_in"):
            assert L.rvsr_engine_workspace_bytes(h, 1, 36, 32) == 0     # HR_in: multiples of 16
        a, b = L.rvsr_engine_workspace_bytes(h, 1, 32, 32), L.rvsr_engine_workspace_bytes(h, 2, 32, 32)
        assert 0 < a < b
        L.rvsr_engine_destroy(h)


def test_dcn_argument_validation_without_gpu():
    L = _lib.lib()
    # channels not divisible by deformable groups -> INVALID (reference: AT_ERROR)
    rc = L.rvsr_mdcn_fwd(None, None, None, None, None, None, 1, 10, 8, 8, 8, 3, 3, 1, 1, 1, 1, 4, 0, None, 0, None)
    assert rc == _lib.E_INVALID
    assert L.rvsr_mdcn_fwd_workspace_bytes(1, 16, 8, 8, 8, 3, 3, 1, 1, 1, 1, 4, 0) > 0
    # empty batch is a no-op, like the reference
    assert L.rvsr_mdcn_fwd(None, None, None, None, None, None, 0, 16, 8, 8, 8, 3, 3, 1, 1, 1, 1, 4, 0, None, 0, None) == 0


def test_conv_argument_validation_without_gpu():
    L = _lib.lib()
    args = lambda B, ks: (None, None, None, None, None, None, B, 64, 0, 8, 8, 64, ks, 1, 0, 0, 1, 1, None, 0, None)  # noqa: E731
    assert L.rvsr_conv2d_fwd(*args(0, 3)) == 0            # empty batch: no-op
    assert L.rvsr_conv2d_fwd(*args(1, 5)) == _lib.E_INVALID  # 5x5 kernels are not built
    assert L.rvsr_conv2d_fwd_workspace_bytes(1, 64, 64, 16, 16, 64, 3, 1) > 0


def test_training_entry_points_validate_arguments_without_gpu():
    """rvsr_c8_* (the bf16 training path): bad arguments come back as error codes before anything touches the device."""
    L = _lib.lib()
    nul = ctypes.c_void_p(0)
    one = ctypes.c_void_p(256)   # never dereferenced: every call below fails its argument checks first
    ptrs, strides = (ctypes.c_void_p * 1)(256), (ctypes.c_longlong * 1)(64 * 16)
    # no sources / too many sources
    assert L.rvsr_c8_conv_fwd(ptrs, strides, 0, 64, one, nul, nul, one, 1, 4, 4, 64, 3, 1, 0, 0, 0, 0.0, nul) == _lib.E_INVALID
    assert b"sources" in L.rvsr_last_error()
    # pixel shuffle with a residual; mask mode without a mask tensor; 5x5 kernel
    assert L.rvsr_c8_conv_fwd(ptrs, strides, 1, 64, one, nul, one, one, 1, 4, 4, 256, 3, 1, 0, 1, 0, 0.0, nul) == _lib.E_INVALID
    assert L.rvsr_c8_conv_fwd(ptrs, strides, 1, 64, one, nul, nul, one, 1, 4, 4, 64, 3, 1, 0, 0, 2, 0.1, nul) == _lib.E_INVALID
    assert L.rvsr_c8_conv_fwd(ptrs, strides, 1, 64, one, nul, nul, one, 1, 4, 4, 64, 5, 1, 0, 0, 0, 0.0, nul) == _lib.E_INVALID
    # shapes the tcgen05 kernels do not cover are UNSUPPORTED (NotImplementedError on the Python side), never a silent fallback
    assert L.rvsr_c8_conv_weight_bytes(216, 64, 3, 0) == 0 and L.rvsr_c8_conv_weight_bytes(256, 64, 3, 0) > 0
    assert L.rvsr_c8_conv_pack_weight(one, one, 216, 64, 3, 0, 0, 64, 0, 3, nul) == _lib.E_UNSUPPORTED
    assert L.rvsr_c8_conv_pack_weight(one, ctypes.c_void_p(264), 64, 64, 3, 0, 0, 64, 0, 3, nul) == _lib.E_INVALID  # unaligned destination
    assert L.rvsr_c8_conv_pack_weight(one, one, 64, 64, 3, 0, 1, 128, 96, 3, nul) == _lib.E_INVALID               # slice past the row
    job = ((ctypes.c_void_p * 1)(256), (ctypes.c_longlong * 1)(64 * 16), (ctypes.c_void_p * 1)(256), (ctypes.c_void_p * 1)(256),
           (ctypes.c_void_p * 1)(0), (ctypes.c_int * 1)(64), (ctypes.c_int * 1)(0))
    assert L.rvsr_c8_conv_wgrad(1, *job, 1, 4, 4, 32, 64, 3, one, 1 << 20, nul) == _lib.E_UNSUPPORTED   # Cin != 64
    assert L.rvsr_c8_conv_wgrad(0, *job, 1, 4, 4, 64, 64, 3, one, 1 << 20, nul) == _lib.E_INVALID       # no jobs
    assert L.rvsr_c8_conv_wgrad_workspace_bytes(1, 80, 64, 64, 64) >= 148 * 6 * 128 * 64 * 4
    assert L.rvsr_c8_conv_wgrad_workspace_bytes(3, 80, 64, 64, 64) >= 147 * 6 * 128 * 64 * 4           # 3 jobs share the SMs
    assert L.rvsr_c8_act_bwd(one, one, one, 64, 0, nul) == _lib.E_INVALID                       # no activation to differentiate
    assert L.rvsr_c8_unshuffle2_act_bwd(one, one, one, 1, 48, 4, 4, 1, nul) == _lib.E_INVALID   # C % 32
    assert L.rvsr_c8_tsa_temporal(one, one, one, ptrs, one, 1, 5, 32, 4, 4, nul) == _lib.E_UNSUPPORTED
    assert L.rvsr_c8_from_nchw(one, _lib.BF16, one, 1, 20, 4, 4, 2, nul) == _lib.E_INVALID      # 2 channel blocks cannot hold 20 channels
    assert L.rvsr_c8_mdcn_fwd(one, one, one, nul, one, 1, 4, 4, 0, one, 16, nul) == _lib.E_INVALID  # workspace too small
    # empty batches are no-ops
    assert L.rvsr_c8_conv_fwd(ptrs, strides, 1, 64, one, nul, nul, one, 0, 4, 4, 64, 3, 1, 0, 0, 0, 0.0, nul) == _lib.OK
    assert L.rvsr_c8_upsample2x(nul, nul, 0, 4, 4, 1.0, 0, nul) == _lib.OK


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.lib()

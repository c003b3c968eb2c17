"""GPU parity tests of the EDVR engine (rvsr_engine_forward through realvsr_b200.archs).

Errors are measured on the network's own contribution  out - base  (base = bilinear x4 of the
centre frame, or the centre frame for EDVR_NoUp) so the comparison is not flattered by the
image-sized base term, as max|diff| / max|reference residual|.
  fp32 engine vs reference-generated golden : < 1e-3  (north_star tolerance)
  fp16 engine vs the same golden            : < 1e-2  (fp16 storage of ~100 chained layers;
                                               fp32 coordinates, blend and accumulation)
"""
import pytest
import torch
import torch.nn.functional as F

from helpers import EDVR_CASES, load_case, rel_err
from realvsr_b200.archs import EDVR_arch as E

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ENGINE_CASES = list(EDVR_CASES)   # incl. the predeblur / HR_in stems (EDVR_arch.py:15-59, :228-231)


def _base(c):
    xc = c["x"][:, c["kwargs"]["nframes"] // 2]
    up = c["cls"] == "EDVR" and not c["kwargs"].get("HR_in", False)   # HR_in: base = the centre frame itself (:315-316)
    return F.interpolate(xc, scale_factor=4, mode="bilinear", align_corners=False) if up else xc


def _net(c, path):
    net = getattr(E, c["cls"])(**c["kwargs"]).eval()
    net.load_state_dict(c["sd"], strict=True)
    net.exec_path = path
    return net.to(DEV)


@pytest.mark.parametrize("name", ENGINE_CASES)
def test_engine_fp32_matches_golden(name):
    c = load_case(name)
    net = _net(c, "engine")
    with torch.no_grad():
        y = net(c["x"].to(DEV))
    assert y.dtype == torch.float32 and y.shape == c["out"].shape
    base = _base(c)
    assert rel_err(y.cpu() - base, c["out"] - base) < 1e-3
    eng = net._get_engine(c["x"].to(DEV))
    al = eng.read_tap("aligned", (c["x"].shape[0] * c["x"].shape[1],) + tuple(c["aligned0"].shape[1:]))   # feature resolution
    B, N = c["x"].shape[:2]
    al0 = al.view(B, N, *al.shape[1:])[:, 0].cpu()
    assert rel_err(al0, c["aligned0"]) < 1e-3


@pytest.mark.parametrize("name", ENGINE_CASES)
def test_engine_fp16_matches_golden(name):
    c = load_case(name)
    net = _net(c, "engine").half()
    with torch.no_grad():
        y = net(c["x"].to(DEV).half())
    assert y.dtype == torch.float16
    base = _base(c)
    assert rel_err(y.float().cpu() - base, c["out"] - base) < 1e-2


@pytest.mark.parametrize("name", ["edvr_tiny", "edvr_noup_3f", "edvr_predeblur", "edvr_hr_in", "edvr_predeblur_hr_in"])
def test_module_path_matches_golden(name):
    """nn.Conv2d modules + our DCN operator (the autograd-capable path)."""
    c = load_case(name)
    net = _net(c, "module")
    with torch.no_grad():
        y = net(c["x"].to(DEV))
    base = _base(c)
    assert rel_err(y.cpu() - base, c["out"] - base) < 1e-3


def test_engine_batch_invariance_and_determinism():
    c = load_case("edvr_tiny")
    net = _net(c, "engine")
    x = c["x"].to(DEV)
    with torch.no_grad():
        y1 = net(x)
        y2 = net(torch.cat([x, x.flip(1), x], 0))
        y3 = net(x)
    assert torch.equal(y1, y3)                      # bit-exact rerun
    assert torch.equal(y2[0:1], y1) and torch.equal(y2[2:3], y1)   # batch position does not matter


def test_engine_nf128_batch_invariance_fp16():
    """BASELINE cfg4's architecture (nf = 128: 64-channel source split, output halves, two-pass DCN contraction) at batch 3:
    every window's frame equals the single-window result bit for bit, wherever it sits in the batch."""
    c = load_case("edvr_nf128_7f")
    net = _net(c, "engine").half()
    x = c["x"].to(DEV).half()
    with torch.no_grad():
        y1 = net(x)
        y3 = net(torch.cat([x, x.flip(1), x], 0))
    assert torch.equal(y3[0:1], y1) and torch.equal(y3[2:3], y1) and not torch.equal(y3[1:2], y1)


def test_engine_reloads_weights_when_parameters_change():
    c = load_case("edvr_tiny")
    net = _net(c, "engine")
    x = c["x"].to(DEV)
    with torch.no_grad():
        y1 = net(x)
        net.conv_last.bias.add_(0.25)
        y2 = net(x)
    assert rel_err(y2 - y1, torch.full_like(y1, 0.25)) < 1e-5


def test_engine_rejects_bad_sizes():
    c = load_case("edvr_tiny")
    net = _net(c, "engine")
    with pytest.raises(RuntimeError):
        with torch.no_grad():
            net(torch.zeros(1, 5, 3, 30, 32, device=DEV))   # H not a multiple of 4


def test_engine_host_buffers_roundtrip():
    c = load_case("edvr_tiny")
    net = _net(c, "engine")
    x = c["x"].to(DEV)
    with torch.no_grad():
        y = net(x)
    eng = net._get_engine(x)
    out = eng.forward_host(c["x"].pin_memory())
    assert torch.equal(out, y.cpu())


def test_engine_host_pipeline_matches_forward_host():
    """HostPipeline (H2D / forward / D2H on three streams, double-buffered) returns bit-identical frames to the
    one-at-a-time forward_host, for a stream of different inputs and with staging slots being reused."""
    c = load_case("edvr_tiny")
    net = _net(c, "engine")
    x = c["x"].to(DEV)
    with torch.no_grad():
        net(x)
    eng = net._get_engine(x)
    g = torch.Generator().manual_seed(5)
    xs = [c["x"].pin_memory()] + [torch.rand(c["x"].shape, generator=g).pin_memory() for _ in range(4)]
    want = [eng.forward_host(xh).clone() for xh in xs]
    outs = [torch.empty_like(want[0]).pin_memory() for _ in xs]
    pipe = eng.host_pipeline(depth=2)
    for xh, oh in zip(xs, outs):
        pipe.submit(xh, oh)
    pipe.drain()
    for w, o in zip(want, outs):
        assert torch.equal(w, o)


def test_overlapped_launches_match_serialized_launches(monkeypatch):
    """Programmatic dependent launch lets every kernel's prologue overlap the previous kernel's tail; a kernel that touched
    activations before its griddepcontrol.wait would read the PREVIOUS forward's data.  Alternating inputs makes that
    visible: frames with overlapped launches must equal the fully serialized ones (RVSR_PDL=0) bit for bit."""
    from helpers import edvr_state_shapes
    from synth import synth_input, synth_state_dict
    kw = dict(nf=64, nc=3, nframes=5, groups=8, front_RBs=5, back_RBs=10, w_TSA=True)
    net = E.EDVR(**kw).eval()
    net.load_state_dict(synth_state_dict(edvr_state_shapes("EDVR", **kw), 7), strict=True)
    net = net.to(DEV).half()
    net.exec_path = "engine"
    xs = [synth_input((1, 5, 3, 180, 320), 29 + i).to(DEV).half() for i in range(2)]
    with torch.no_grad():
        monkeypatch.setenv("RVSR_PDL", "0")
        refs = [net(x).clone() for x in xs]
        monkeypatch.delenv("RVSR_PDL", raising=False)
        for i in range(10):
            assert torch.equal(net(xs[i % 2]), refs[i % 2]), "overlapped launches differ from serialized ones in repeat %d" % i


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_two_devices_in_one_process():
    """The reference wraps the model in nn.DataParallel (VideoSR_AllPair_model_YCbCr_Split.py:33-36): one process, several
    devices.  Per-device state of the library (opt-in shared-memory attribute, SM count) must follow the current device."""
    c = load_case("edvr_nf64_crop")
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        net = getattr(E, c["cls"])(**c["kwargs"]).eval()
        net.load_state_dict(c["sd"], strict=True)
        net = net.to(dev).half()
        net.exec_path = "engine"
        with torch.no_grad():
            outs.append(net(c["x"].to(dev).half()).cpu())
    assert torch.equal(outs[0], outs[1])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_dataparallel_wrapper_train_and_eval():
    """The wrapper the reference's training models use (nn.DataParallel over gpu_ids, VideoSR_AllPair_model_YCbCr_Split.py:36;
    the shipped YAMLs select it).  Replicas have no parameters() / state_dict() (replicate() sets the broadcast weights as
    plain attributes): with autograd they must take the module path and deliver gradients to the wrapped module's
    parameters; under no_grad (model.test()) they run the engine with the replica's weights."""
    c = load_case("edvr_tiny_b2_g2")
    net = getattr(E, c["cls"])(**c["kwargs"])
    net.load_state_dict(c["sd"], strict=True)
    net = net.to(DEV)
    dp = torch.nn.DataParallel(net, device_ids=[0, 1])
    x = c["x"].to(DEV)
    base = _base(c)
    dp.eval()
    with torch.no_grad():
        y = dp(x)
    assert rel_err(y.cpu() - base, c["out"] - base) < 1e-3
    dp.train()
    out = dp(x)
    out.mean().backward()
    assert all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in net.parameters())
    assert rel_err(out.detach().cpu() - base, c["out"] - base) < 1e-3


def test_engine_sees_data_writes_after_invalidate():
    """In-place writes through p.data do not bump p._version (torch semantics): invalidate_engine() / the 'checksum' weight
    check make the engine re-read the weights."""
    c = load_case("edvr_tiny")
    net = _net(c, "engine")
    x = c["x"].to(DEV)
    with torch.no_grad():
        y1 = net(x)
        net.conv_last.bias.data.add_(0.25)
        net.invalidate_engine()
        y2 = net(x)
        assert rel_err(y2 - y1, torch.full_like(y1, 0.25)) < 1e-5
        net.engine_weight_check = "checksum"
        net(x)
        net.conv_last.bias.data.add_(0.25)
        y3 = net(x)
        assert rel_err(y3 - y1, torch.full_like(y1, 0.5)) < 1e-5
    import copy
    net2 = copy.deepcopy(net)                      # engine cache holds ctypes handles: the copy starts without one
    with torch.no_grad():
        assert torch.equal(net2(x), y3)


def test_cfg2_full_size_fp16_properties():
    """BASELINE cfg2 size (5x3x180x320 -> 720x1280, nf=64, TSA): too slow for the CPU oracle,
    so check size-independent properties: finite, deterministic, zero-initialised offset
    convs turn every DCN into 0.5 * conv (checked against the module path in fp32)."""
    from synth import synth_input, synth_state_dict
    from helpers import edvr_state_shapes
    kw = dict(nf=64, nc=3, nframes=5, groups=8, front_RBs=5, back_RBs=10, w_TSA=True)
    sd = synth_state_dict(edvr_state_shapes("EDVR", **kw), 7)
    net = E.EDVR(**kw).eval()
    net.load_state_dict(sd, strict=True)
    net = net.to(DEV).half()
    net.exec_path = "engine"
    x = synth_input((1, 5, 3, 180, 320), 8).to(DEV).half()
    with torch.no_grad():
        y = net(x)
        y2 = net(x)
    assert y.shape == (1, 3, 720, 1280) and bool(torch.isfinite(y).all()) and torch.equal(y, y2)
    # interior crop consistency with the fp32 module path on a 64x64 crop is covered by the golden
    # nf64 case; here compare fp16 engine vs fp32 engine at full size
    net32 = E.EDVR(**kw).eval()
    net32.load_state_dict(sd, strict=True)
    net32 = net32.to(DEV)
    net32.exec_path = "engine"
    with torch.no_grad():
        y32 = net32(x.float())
    base = F.interpolate(x[:, 2].float(), scale_factor=4, mode="bilinear", align_corners=False)
    assert rel_err(y.float() - base, y32 - base) < 1e-2


def test_shipped_realvsr_config_full_frame():
    """The configuration the reference's test scripts actually run (test_RealVSR_wi_GT.py:50-51):
    EDVR_NoUp(nf=64, nframes=3, groups=8, front_RBs=5, back_RBs=10, w_TSA=False) on whole 1024x512
    RealVSR frames (RealVSR_dataset.py:284), one window per call, fp32 in / fp32 out via single_forward.
    Full size is out of reach of the CPU oracle: check fp16-engine vs fp32-engine agreement,
    determinism, and that the strict-loaded weights are the ones used (bias probe)."""
    from helpers import edvr_state_shapes
    from realvsr_b200 import video as V
    from synth import synth_input, synth_state_dict
    kw = dict(nf=64, nc=3, nframes=3, groups=8, front_RBs=5, back_RBs=10, w_TSA=False)
    sd = synth_state_dict(edvr_state_shapes("EDVR_NoUp", **kw), 17)
    net = E.EDVR_NoUp(**kw).eval()
    net.load_state_dict(sd, strict=True)
    net = net.to(DEV)
    net.exec_path = "engine"
    x = synth_input((1, 3, 3, 512, 1024), 18).to(DEV)
    y32 = V.single_forward(net, x)                      # fp32 engine (CUDA-core kernels)
    assert y32.shape == (1, 3, 512, 1024) and y32.dtype == torch.float32 and bool(torch.isfinite(y32).all())
    net16 = E.EDVR_NoUp(**kw).eval()
    net16.load_state_dict(sd, strict=True)
    net16 = net16.to(DEV).half()
    net16.exec_path = "engine"
    y16 = V.single_forward(net16, x.half())
    base = x[:, 1].cpu()
    assert rel_err(y16 - base, y32 - base) < 1e-2
    assert torch.equal(y16, V.single_forward(net16, x.half()))

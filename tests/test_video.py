"""Video-level helpers: index_generation against vectors produced by the reference's
own function (integer work: bit-exact), the flip ensemble wiring on CPU, and -- on the GPU -- the
feature-cache path against the plain forward (bit-identical frames)."""
import json
import os

import pytest
import torch

from helpers import GOLDEN, load_case
from realvsr_b200 import video as V


def test_index_generation_matches_reference_vectors():
    cases = json.load(open(os.path.join(GOLDEN, "index_generation.json")))
    assert len(cases) > 200
    for c in cases:
        assert V.index_generation(c["crt"], c["max_n"], c["N"], c["padding"]) == c["out"], c
    with pytest.raises(ValueError):
        V.index_generation(0, 10, 5, "zeros")


def test_single_and_flipx4_forward_wiring():
    class M(torch.nn.Module):  # not flip-equivariant on purpose: a fixed left-to-right ramp is added
        def forward(self, x):
            return [x[:, 1] * 2 + torch.arange(x.shape[-1], dtype=x.dtype)]
    x = torch.rand(1, 3, 2, 4, 6)
    y = V.single_forward(M(), x)
    assert y.dtype == torch.float32 and y.device.type == "cpu" and y.shape == (1, 2, 4, 6)
    ramp = torch.arange(6.)
    expect = x[:, 1] * 2 + (ramp + ramp.flip(0)) / 2  # two of the four passes see the ramp mirrored
    assert torch.allclose(V.flipx4_forward(M(), x), expect, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("name,half", [("edvr_tiny", False), ("edvr_nf64_crop", True), ("edvr_noup_nf64_ship", True)])
def test_sr_sequence_cache_is_bit_identical_to_plain_forward(name, half):
    from realvsr_b200.archs import EDVR_arch as E
    from synth import synth_input
    c = load_case(name)
    net = getattr(E, c["cls"])(**c["kwargs"]).eval()
    net.load_state_dict(c["sd"], strict=True)
    net = net.to("cuda:0")
    net.exec_path = "engine"
    H, W = c["x"].shape[-2:]
    frames = synth_input((9, 3, H, W), 321).to("cuda:0")
    if half:
        net, frames = net.half(), frames.half()
    for padding in ("replicate", "reflection"):
        a = V.sr_sequence(net, frames, padding=padding, batch=4, cache=True)
        b = V.sr_sequence(net, frames, padding=padding, batch=2, cache=False)
        assert a.shape[0] == 9 and torch.equal(a, b), padding
    # one window against the plain model call, frame order as the reference loop builds it
    idx = V.index_generation(4, 9, c["kwargs"]["nframes"], "replicate")
    with torch.no_grad():
        y = net(frames[idx].unsqueeze(0))
    assert torch.equal(V.sr_sequence(net, frames, batch=3)[4:5], y)


@pytest.mark.gpu
def test_flipx4_on_the_engine_matches_reference_semantics():
    """utils/util.py:240-261 on the real model: four forwards on flipped inputs, flipped back and averaged -- against the
    CPU oracle doing the same, and the batch-4 variant against the four separate calls (bit-identical)."""
    from helpers import rel_err
    from oracle import edvr_oracle as O
    from realvsr_b200.archs import EDVR_arch as E
    c = load_case("edvr_tiny")
    kw = c["kwargs"]
    net = E.EDVR(**kw).eval()
    net.load_state_dict(c["sd"], strict=True)
    net = net.to("cuda:0")
    net.exec_path = "engine"
    x = c["x"]

    def oracle_model(t):
        return O.edvr_forward(c["sd"], t, groups=kw["groups"], w_TSA=kw["w_TSA"], upsample=True)

    ref = V.flipx4_forward(oracle_model, x)
    got = V.flipx4_forward(net, x.to("cuda:0"))
    assert rel_err(got, ref) < 1e-4
    assert torch.equal(V.flipx4_forward_batched(net, x.to("cuda:0")), got)


@pytest.mark.gpu
@pytest.mark.parametrize("half", [False, True])
def test_tiled_forward_matches_oracle_with_identical_tiling(half):
    """BASELINE cfg4's architecture (7 frames, nf = 128, 16 channels per deformable group) on a 2 x 2-tile crop.  The
    network's receptive field exceeds any practical halo, so tiles + halo != whole frame near the seams (SURVEY.md 7,
    "hard parts"): parity for tiled inference is therefore defined PER TILING -- the CPU oracle is run through the same
    video.tiled_forward (same tile origins, halos and crop) and must agree everywhere: fp32 engine < 1e-3, fp16 < 1e-2.
    One tile covering the frame is bit-identical to the plain forward."""
    import torch.nn.functional as F
    from helpers import rel_err
    from oracle import edvr_oracle as O
    from realvsr_b200.archs import EDVR_arch as E
    from synth import synth_input
    c = load_case("edvr_nf128_7f")
    kw = c["kwargs"]
    net = E.EDVR(**kw).eval()
    net.load_state_dict(c["sd"], strict=True)
    net = net.to("cuda:0")
    net.exec_path = "engine"
    x = synth_input((1, 7, 3, 40, 56), 118)
    xd = x.to("cuda:0")
    if half:
        net, xd = net.half(), xd.half()

    def oracle_model(t):
        with torch.no_grad():
            return O.edvr_forward(c["sd"], t, groups=kw["groups"], w_TSA=True, upsample=True)

    tile, halo = (20, 28), 8
    ref = V.tiled_forward(oracle_model, x, tile=tile, halo=halo)
    got = V.tiled_forward(net, xd, tile=tile, halo=halo).float().cpu()
    base = F.interpolate(x[:, 3], scale_factor=4, mode="bilinear", align_corners=False)
    err = rel_err(got - base, ref - base)
    print("tiled nf128/7f %s engine vs identically tiled oracle: %.2e" % ("fp16" if half else "fp32", err))
    assert got.shape == ref.shape == (1, 3, 160, 224)
    assert err < (1e-2 if half else 1e-3)
    with torch.no_grad():
        full = net(xd)
    assert torch.equal(V.tiled_forward(net, xd, tile=(40, 56), halo=8), full)
    with pytest.raises(RuntimeError):
        V.tiled_forward(net, xd[..., :38, :], tile=tile)

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
def test_tiled_forward_matches_whole_frame():
    """One tile covering the frame is exact; 2 x 2 tiles with a 16-pixel halo differ from the whole-frame forward only by
    the truncated receptive field at the seams (bounded), and every pixel farther than the halo from a seam... is still
    only approximately equal (the pyramid's receptive field exceeds the halo), so the bound is global."""
    import torch
    from helpers import edvr_state_shapes, rel_err
    from realvsr_b200 import video
    from realvsr_b200.archs import EDVR_arch as E
    from synth import synth_input, synth_state_dict
    kw = dict(nf=64, nc=3, nframes=3, groups=8, front_RBs=2, back_RBs=2, w_TSA=True)
    net = E.EDVR(**kw).eval()
    net.load_state_dict(synth_state_dict(edvr_state_shapes("EDVR", **kw), 17), strict=True)
    net = net.to("cuda:0")
    net.exec_path = "engine"
    x = synth_input((1, 3, 3, 96, 128), 18).to("cuda:0")
    with torch.no_grad():
        full = net(x)
    assert torch.equal(video.tiled_forward(net, x, tile=(96, 128), halo=16), full)
    tiled = video.tiled_forward(net, x, tile=(48, 64), halo=16)
    assert tiled.shape == full.shape
    base = torch.nn.functional.interpolate(x[:, 1], scale_factor=4, mode="bilinear", align_corners=False)
    assert rel_err((tiled - base).cpu(), (full - base).cpu()) < 0.2       # seams: truncated context
    inner = (slice(None), slice(None), slice(4 * 8, 4 * 40), slice(4 * 8, 4 * 56))   # interior of the first tile
    assert rel_err((tiled - base)[inner].cpu(), (full - base)[inner].cpu()) < 5e-2
    with pytest.raises(RuntimeError):
        video.tiled_forward(net, x[..., :94, :], tile=(48, 64))

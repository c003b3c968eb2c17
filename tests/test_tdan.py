"""TDAN `Align` (reference codes/models/archs/TDAN_arch.py:17-72) on this package's DCN operator -- SURVEY.md 8f rank 3.
Fixtures: tests/golden/make_golden_tdan.py ran the reference's own Align (DCN call routed to torchvision)."""
import ast
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, rel_err

CASES = ["tdan_align", "tdan_align_g8"]


def _load(name):
    from synth import synth_input, synth_state_dict
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    kw = ast.literal_eval(str(z["kwargs"]))
    shapes = {str(k): ast.literal_eval(str(s)) for k, s in zip(z["keys"], z["shapes"])}
    return dict(kwargs=kw, shapes=shapes, x=synth_input(tuple(z["shape"]), int(z["xseed"])),
                sd=synth_state_dict(shapes, int(z["wseed"])), out=torch.from_numpy(z["out"]))


@pytest.mark.parametrize("name", CASES)
def test_align_state_dict_contract(name):
    """Same parameter names, order and shapes as the reference module (strict loads of reference checkpoints)."""
    from realvsr_b200.archs.TDAN_arch import Align
    c = _load(name)
    net = Align(**c["kwargs"])
    got = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    assert list(got.keys()) == list(c["shapes"].keys())
    assert got == c["shapes"]
    net.load_state_dict(c["sd"], strict=True)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_align_fp32_matches_reference_golden(name):
    from realvsr_b200.archs.TDAN_arch import Align
    c = _load(name)
    net = Align(**c["kwargs"]).eval()
    net.load_state_dict(c["sd"], strict=True)
    net = net.to("cuda:0")
    x = c["x"].to("cuda:0")
    with torch.no_grad():
        y = net(x).cpu()                 # fused pack operator (rvsr_mdcn_pack_fwd)
    assert y.shape == c["out"].shape
    assert rel_err(y, c["out"]) < 1e-3   # north_star: 1e-3 relative (measured ~1e-6)
    y2 = net(x).detach().cpu()           # autograd graph: torch conv + rvsr_mdcn_fwd
    assert rel_err(y2, c["out"]) < 1e-3


@pytest.mark.gpu
def test_align_fp16_tcgen05_pack():
    """nf = 64, 8 groups: the packs run on the tcgen05 offset conv + gather kernels (fp16 storage)."""
    from realvsr_b200.archs.TDAN_arch import Align
    c = _load("tdan_align_g8")
    net = Align(**c["kwargs"]).eval()
    net.load_state_dict(c["sd"], strict=True)
    net = net.to("cuda:0").half()
    with torch.no_grad():
        y = net(c["x"].to("cuda:0").half()).float().cpu()
    assert rel_err(y, c["out"]) < 2e-2   # fp16 storage of every layer; fp32 coordinates / accumulation

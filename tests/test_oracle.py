"""CPU tests: pin the oracle (oracle/dcn_oracle.c + oracle/edvr_oracle.py) against the
golden fixtures that tests/golden/make_golden.py produced by running the reference."""
import os

import numpy as np
import pytest
import torch

from helpers import EDVR_CASES, GOLDEN, load_case, rel_err
from oracle import edvr_oracle as O
from synth import synth_normal


def _dcn_unit():
    z = np.load(os.path.join(GOLDEN, "dcn_unit.npz"))
    B, C, H, W, Cout, dg = [int(v) for v in z["dims"]]
    x = synth_normal((B, C, H, W), 71).double()
    off = synth_normal((B, dg * 18, H, W), 72, std=4.0).double()
    msk = torch.sigmoid(synth_normal((B, dg * 9, H, W), 73).double())
    w = synth_normal((Cout, C, 3, 3), 74, std=0.1).double()
    b = synth_normal((Cout,), 75).double()
    go = synth_normal((B, Cout, H, W), 76).double()
    return z, (x, off, msk, w, b, go, dg)


def test_dcn_oracle_forward_matches_golden_fp64():
    z, (x, off, msk, w, b, go, dg) = _dcn_unit()
    y = O.dcn_forward(x, off, msk, w, b, 1, 1, 1, 1, dg)
    assert rel_err(y, torch.from_numpy(z["out"])) < 1e-12


def test_dcn_oracle_backward_matches_golden_fp64():
    z, (x, off, msk, w, b, go, dg) = _dcn_unit()
    gx, goff, gm, gw, gb = O.dcn_backward(x, off, msk, w, go, 1, 1, 1, 1, dg)
    for got, key in ((gx, "gx"), (goff, "goff"), (gm, "gmask"), (gw, "gw"), (gb, "gb")):
        assert rel_err(got, torch.from_numpy(z[key])) < 1e-11, key


def test_dcn_oracle_fp32_edge_and_errors():
    z, (x, off, msk, w, b, go, dg) = _dcn_unit()
    y = O.dcn_forward(x.float(), off.float(), msk.float(), w.float(), b.float(), 1, 1, 1, 1, dg)
    assert rel_err(y, torch.from_numpy(z["out"])) < 1e-5
    with pytest.raises(RuntimeError):  # channels not divisible by deformable groups
        O.dcn_forward(x[:, :15], off, msk, w[:, :15], b, 1, 1, 1, 1, dg)


def test_dcn_oracle_zero_offset_is_half_conv():
    # reference init: offsets 0, mask 0.5  ->  0.5 * ordinary conv (SURVEY.md section 4 trap 1)
    x = synth_normal((1, 8, 9, 10), 1).double()
    w = synth_normal((8, 8, 3, 3), 2, std=0.2).double()
    y = O.dcn_forward(x, torch.zeros(1, 36, 9, 10).double(), torch.full((1, 18, 9, 10), 0.5).double(),
                      w, None, 1, 1, 1, 1, 2)
    assert rel_err(y, 0.5 * torch.nn.functional.conv2d(x, w, padding=1)) < 1e-12


def test_dcn_oracle_stride_groups_dilation_vs_torchvision():
    tvo = pytest.importorskip("torchvision.ops")
    x = synth_normal((2, 8, 12, 9), 3).double()
    w = synth_normal((6, 4, 3, 3), 4, std=0.2).double()  # groups=2
    b = synth_normal((6,), 5).double()
    for stride, pad, dil in ((2, 1, 1), (1, 2, 2), (2, 0, 1)):
        Ho = (12 + 2 * pad - (dil * 2 + 1)) // stride + 1
        Wo = (9 + 2 * pad - (dil * 2 + 1)) // stride + 1
        off = synth_normal((2, 4 * 18, Ho, Wo), 6, std=2.0).double()
        msk = torch.sigmoid(synth_normal((2, 4 * 9, Ho, Wo), 7).double())
        y = O.dcn_forward(x, off, msk, w, b, stride, pad, dil, 2, 4)
        ref = tvo.deform_conv2d(x, off, w, b, stride=stride, padding=pad, dilation=dil, mask=msk)
        assert rel_err(y, ref) < 1e-12


@pytest.mark.parametrize("name", EDVR_CASES)
def test_edvr_oracle_matches_reference_golden(name):
    c = load_case(name)
    kw = c["kwargs"]
    taps = {}
    with torch.no_grad():
        y = O.edvr_forward(c["sd"], c["x"], groups=kw["groups"], center=kw.get("center"),
                           w_TSA=kw.get("w_TSA", True), upsample=c["cls"] == "EDVR",
                           is_predeblur=kw.get("predeblur", False), HR_in=kw.get("HR_in", False),
                           taps=taps)
    assert y.shape == c["out"].shape
    assert rel_err(taps["aligned"][:, 0], c["aligned0"]) < 2e-5
    assert rel_err(y, c["out"]) < 2e-5


def test_pixel_shuffle_index_map_bit_exact():
    x = torch.arange(2 * 8 * 3 * 5, dtype=torch.int64).view(2, 8, 3, 5)
    y = torch.nn.functional.pixel_shuffle(x, 2)
    for c in range(2):
        for i in range(2):
            for j in range(2):
                ci, _, _ = O.pixel_shuffle2_index(c, 0, 0, i, j)
                assert torch.equal(y[:, c, i::2, j::2], x[:, ci])

"""GPU parity tests of the fused convolution operator (rvsr_conv2d_fwd), CUDA-core and
tcgen05 kernels, against a plain PyTorch fp32 reference of the same op (F.conv2d + the
surrounding cat / activation / residual / pixel_shuffle).

Tolerances: fp32 kernel 1e-4 of max|ref|; fp16-storage kernels 2e-3 against the fp32
reference evaluated on the same fp16-rounded inputs and weights (fp32 accumulate inside)."""
import pytest
import torch
import torch.nn.functional as F

from helpers import rel_err
from realvsr_b200 import ops
from synth import synth_normal

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ref(x, w, b, x2, res, stride, act, shuffle):
    xin = x if x2 is None else torch.cat([x, x2], 1)
    y = F.conv2d(xin, w, b, stride=stride, padding=w.shape[-1] // 2)
    y = F.leaky_relu(y, 0.1) if act == "lrelu" else (F.relu(y) if act == "relu" else y)
    if res is not None:
        y = y + res
    return F.pixel_shuffle(y, 2) if shuffle else y


def _mk(B, C1, H, W, Cout, ks=3, C2=0, res=False, stride=1, seed=300):
    x = synth_normal((B, C1, H, W), seed)
    x2 = synth_normal((B, C2, H, W), seed + 1) if C2 else None
    w = synth_normal((Cout, C1 + C2, ks, ks), seed + 2, std=(1.0 / ((C1 + C2) * ks * ks)) ** 0.5)
    b = synth_normal((Cout,), seed + 3, std=0.3)
    Ho, Wo = (H, W) if stride == 1 else ((H - 1) // 2 + 1, (W - 1) // 2 + 1)
    r = synth_normal((B, Cout, Ho, Wo), seed + 4) if res else None
    return x, w, b, x2, r


CASES = [  # (name, make-kwargs, call-kwargs)
    ("rb_conv1_relu", dict(B=2, C1=64, H=24, W=40, Cout=64), dict(act="relu")),
    ("rb_conv2_residual", dict(B=1, C1=64, H=12, W=68, Cout=64, res=True), dict()),
    ("cat2_lrelu_ragged", dict(B=2, C1=64, C2=64, H=10, W=34, Cout=64), dict(act="lrelu")),
    ("stride2_lrelu", dict(B=2, C1=64, H=24, W=64, Cout=64, stride=2), dict(act="lrelu", stride=2)),
    ("fusion_1x1", dict(B=1, C1=64, C2=64, H=16, W=36, Cout=64, ks=1), dict(act="lrelu")),
    ("upconv_shuffle", dict(B=1, C1=64, H=12, W=32, Cout=256), dict(act="lrelu", shuffle=True)),
    ("conv_last_cout3", dict(B=1, C1=64, H=20, W=36, Cout=3), dict()),
    ("conv_first_cin3", dict(B=2, C1=3, H=16, W=32, Cout=64), dict(act="lrelu")),
    ("big_180x320", dict(B=1, C1=64, H=180, W=320, Cout=64), dict(act="lrelu")),
]


@pytest.mark.parametrize("name,mk,kw", CASES, ids=[c[0] for c in CASES])
def test_conv_simt_fp32(name, mk, kw):
    x, w, b, x2, r = _mk(**mk)
    ref = _ref(x, w, b, x2, r, kw.get("stride", 1), kw.get("act"), kw.get("shuffle", False))
    d = lambda t: None if t is None else t.to(DEV)  # noqa: E731
    y = ops.conv2d_fused(d(x), d(w), d(b), x2=d(x2), residual=d(r), **kw)
    assert y.shape == ref.shape
    assert rel_err(y.cpu(), ref) < 1e-4


@pytest.mark.parametrize("use_tc", [False, True], ids=["simt", "tcgen05"])
@pytest.mark.parametrize("name,mk,kw", CASES, ids=[c[0] for c in CASES])
def test_conv_fp16(name, mk, kw, use_tc):
    x, w, b, x2, r = _mk(**mk)
    h = lambda t: None if t is None else t.half().float()  # noqa: E731
    ref = _ref(h(x), h(w), h(b), h(x2), h(r), kw.get("stride", 1), kw.get("act"), kw.get("shuffle", False))
    d = lambda t: None if t is None else t.to(DEV).half()  # noqa: E731
    y = ops.conv2d_fused(d(x), d(w), d(b), x2=d(x2), residual=d(r), use_tc=use_tc, **kw)
    assert y.dtype == torch.float16 and y.shape == ref.shape
    assert rel_err(y.float().cpu(), ref) < 2e-3


def test_conv_tc_matches_simt_bitwise_close_on_structured_input():
    """A delta image makes every tap/shift/tile-seam mistake visible as a misplaced weight."""
    x = torch.zeros(1, 64, 12, 70)
    x[0, 5, 6, 31] = 1.0   # on a tile seam (valid width 30)
    x[0, 63, 0, 0] = 2.0
    x[0, 17, 11, 69] = -1.0
    w = synth_normal((64, 64, 3, 3), 9, std=0.1)
    ref = F.conv2d(x.half().float(), w.half().float(), padding=1)
    y = ops.conv2d_fused(x.to(DEV).half(), w.to(DEV).half(), None, use_tc=True)
    assert rel_err(y.float().cpu(), ref) < 2e-3

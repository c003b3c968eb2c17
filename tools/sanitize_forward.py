"""One small EDVR forward per engine precision for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck  python tools/sanitize_forward.py            > profiles/r02_sanitizer_memcheck.txt
    compute-sanitizer --tool racecheck python tools/sanitize_forward.py            > profiles/r02_sanitizer_racecheck.txt

Covers the tcgen05 conv kernels (single CTA and CTA pair), the fused DCN pack, the taps-in-N conv_last, the glue kernels and
the CUDA-core fp32 engine, at a size where tiles straddle the image border (nf = 64 network, 2 windows of 5x3x36x68)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from helpers import edvr_state_shapes  # noqa: E402
from realvsr_b200.archs import EDVR_arch as E  # noqa: E402
from synth import synth_input, synth_state_dict  # noqa: E402

kw = dict(nf=64, nc=3, nframes=5, groups=8, front_RBs=2, back_RBs=2, w_TSA=True)
sd = synth_state_dict(edvr_state_shapes("EDVR", **kw), 7)
x = synth_input((2, 5, 3, 36, 68), 8).cuda()
for half in (True, False) if "--fp16-only" not in sys.argv else (True,):
    net = E.EDVR(**kw).eval()
    net.load_state_dict(sd, strict=True)
    net = net.cuda()
    net.exec_path = "engine"
    xi = x
    if half:
        net, xi = net.half(), x.half()
    with torch.no_grad():
        y = net(xi)
    torch.cuda.synchronize()
    print("forward", "fp16" if half else "fp32", tuple(y.shape), "finite", bool(torch.isfinite(y).all()),
          "launches", net._get_engine(xi).last_launch_count())

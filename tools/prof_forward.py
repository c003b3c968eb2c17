"""Run a couple of cfg2 forwards (fp16 engine) -- the target command for ncu captures.
    ncu ... python tools/prof_forward.py [n_forwards]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from helpers import edvr_state_shapes  # noqa: E402
from realvsr_b200.archs import EDVR_arch as E  # noqa: E402
from synth import synth_input, synth_state_dict  # noqa: E402

CFG = dict(nf=64, nc=3, nframes=5, groups=8, front_RBs=5, back_RBs=10, w_TSA=True)
net = E.EDVR(**CFG).eval()
net.load_state_dict(synth_state_dict(edvr_state_shapes("EDVR", **CFG), 7), strict=True)
net = net.to("cuda:0").half()
net.exec_path = "engine"
x = synth_input((1, 5, 3, 180, 320), 8).to("cuda:0").half()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
with torch.no_grad():
    for _ in range(n):
        y = net(x)
torch.cuda.synchronize()
print("ok", tuple(y.shape), float(y.float().abs().mean()))

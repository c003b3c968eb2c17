"""Target command for ncu captures: cfg2 forwards on the fp16 engine.
    ncu ... python tools/prof_forward.py [n_forwards] [batch]
Also writes the ordered launch labels of one forward to gpurun_out/launch_labels.json so the ncu
launch list (same order) can be joined with kernel classes."""
import json
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
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
x = synth_input((B, 5, 3, 180, 320), 8).to("cuda:0").half()
with torch.no_grad():
    for _ in range(n):
        y = net(x)
torch.cuda.synchronize()
if os.environ.get("RVSR_DUMP_LABELS"):
    with torch.no_grad():
        rows = net._get_engine(x).profile(x, steps=1)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump([dict(label=r["label"], flops=r["flops"], bytes=r["bytes"]) for r in rows],
              open(os.path.join(ROOT, "gpurun_out", "launch_labels.json"), "w"))
print("ok", tuple(y.shape), float(y.float().abs().mean()))

"""Stress the chained trunks for nondeterminism: the shipped RealVSR config on a 512x1024 frame, many repeats."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch
from helpers import edvr_state_shapes
from realvsr_b200.archs import EDVR_arch as E
from synth import synth_input, synth_state_dict
kw = dict(nf=64, nc=3, nframes=3, groups=8, front_RBs=5, back_RBs=10, w_TSA=False)
net = E.EDVR_NoUp(**kw).eval()
net.load_state_dict(synth_state_dict(edvr_state_shapes("EDVR_NoUp", **kw), 17), strict=True)
net = net.to("cuda:0").half(); net.exec_path = "engine"
x = synth_input((1, 3, 3, 512, 1024), 18).to("cuda:0").half()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
with torch.no_grad():
    os.environ["RVSR_CHAIN"] = "0"
    ref = net(x).clone()
    os.environ.pop("RVSR_CHAIN")
    bad = 0
    for i in range(n):
        y = net(x)
        if not torch.equal(y, ref):
            bad += 1
            d = (y.float() - ref.float()).abs()
            if bad <= 3:
                print("  mismatch in repeat %d: %d elements differ, max |d| %.3g" % (i, int((d > 0).sum()), float(d.max())))
print("chain debug=%s: %d of %d repeats differ from the per-layer result" % (os.environ.get("RVSR_CHAIN_DEBUG", "0"), bad, n))

"""Soak test: many cfg2 forwards on alternating inputs, every result compared bit for bit with the first one for that input
(a rare synchronisation bug in a persistent kernel shows up as a few wrong pixels once in a while)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch
from helpers import edvr_state_shapes
from realvsr_b200.archs import EDVR_arch as E
from synth import synth_input, synth_state_dict
CFG = dict(nf=64, nc=3, nframes=5, groups=8, front_RBs=5, back_RBs=10, w_TSA=True)
net = E.EDVR(**CFG).eval()
net.load_state_dict(synth_state_dict(edvr_state_shapes("EDVR", **CFG), 7), strict=True)
net = net.to("cuda:0").half(); net.exec_path = "engine"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
bad = 0
for B, H, W in ((4, 180, 320), (1, 180, 320), (2, 68, 100)):
    xs = [synth_input((B, 5, 3, H, W), 40 + i).to("cuda:0").half() for i in range(3)]
    with torch.no_grad():
        refs = [net(x).clone() for x in xs]
        for i in range(n):
            y = net(xs[i % 3])
            if not torch.equal(y, refs[i % 3]):
                bad += 1
                print("MISMATCH B=%d %dx%d iteration %d: %d elements differ" % (B, H, W, i, int((y != refs[i % 3]).sum())))
    torch.cuda.synchronize()
    print("B=%d %dx%d: %d forwards compared" % (B, H, W, n))
print("soak: %d mismatches" % bad)
sys.exit(1 if bad else 0)

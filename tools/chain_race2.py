"""First forward of a fresh engine vs the second one (the failing assertion of test_shipped_realvsr_config_full_frame)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch
from helpers import edvr_state_shapes
from realvsr_b200.archs import EDVR_arch as E
from synth import synth_input, synth_state_dict
kw = dict(nf=64, nc=3, nframes=3, groups=8, front_RBs=5, back_RBs=10, w_TSA=False)
sd = synth_state_dict(edvr_state_shapes("EDVR_NoUp", **kw), 17)
x = synth_input((1, 3, 3, 512, 1024), 18).to("cuda:0").half()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
bad = 0
for i in range(n):
    net = E.EDVR_NoUp(**kw).eval()
    net.load_state_dict(sd, strict=True)
    net = net.to("cuda:0").half(); net.exec_path = "engine"
    with torch.no_grad():
        a = net(x).clone()
        b = net(x).clone()
        c = net(x).clone()
    if not (torch.equal(a, b) and torch.equal(b, c)):
        bad += 1
        d1 = (a.float() - b.float()).abs(); d2 = (b.float() - c.float()).abs()
        print("  iter %d: first vs second differ in %d elements (max %.3g); second vs third in %d" % (i, int((d1 > 0).sum()), float(d1.max()), int((d2 > 0).sum())))
    del net
    torch.cuda.empty_cache()
print("chain=%s debug=%s: %d of %d fresh engines nondeterministic" % (os.environ.get("RVSR_CHAIN", "1"), os.environ.get("RVSR_CHAIN_DEBUG", "0"), bad, n))

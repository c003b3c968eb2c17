"""Per-launch event timings of the chain launches (and totals) of one cfg2 forward, B from env."""
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
B = int(os.environ.get("B", "4"))
x = synth_input((B, 5, 3, 180, 320), 8).to("cuda:0").half()
with torch.no_grad():
    for _ in range(3): net(x)
    eng = net._get_engine(x)
    rows = eng.profile(x, steps=5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): net(x)
    e1.record(); torch.cuda.synchronize()
tot = sum(r["ms"] for r in rows)
print("chain=%s dbg=%s B=%d step %.3f ms profile-sum %.3f ms" % (os.environ.get("RVSR_CHAIN", "1"), os.environ.get("RVSR_CHAIN_DEBUG", "0"), B,
                                                          e0.elapsed_time(e1) / 20, tot), end=" | ")
for r in rows:
    if "chain" in r["label"] or r["label"].endswith("feature_extraction.0.conv1") or r["label"].endswith("recon_trunk.0.conv1"):
        print("%s %.1fus" % (r["label"].split(":")[-1][-28:], r["ms"] * 1e3), end="  ")
print()

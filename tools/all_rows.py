"""Per-launch table of one cfg2 forward (fp16 engine, B windows): label, us, TFLOP/s, GB/s -- from the engine's own
per-launch CUDA events (rvsr_engine_profile_*).   B=4 python tools/all_rows.py"""
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
H, W = int(os.environ.get("H", "180")), int(os.environ.get("W", "320"))
x = synth_input((B, 5, 3, H, W), 8).to("cuda:0").half()
with torch.no_grad():
    for _ in range(3): net(x)
    rows = net._get_engine(x).profile(x, steps=5)
print("workspace %.3f GB (RVSR_ARENA_REUSE=%s)" % (net._get_engine(x)._ws.numel() / 1e9, os.environ.get("RVSR_ARENA_REUSE", "1")))
tot = sum(r["ms"] for r in rows)
print("B=%d %dx%d: %d launches, sum %.3f ms" % (B, H, W, len(rows), tot))
for r in rows:
    ms = max(r["ms"], 1e-6)
    print("%-58s %8.1f us %8.1f TF/s %8.1f GB/s" % (r["label"][:58], ms * 1e3, r["flops"] / ms / 1e9, r["bytes"] / ms / 1e6))

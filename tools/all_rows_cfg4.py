"""Per-launch table of one cfg4 tile forward (nf = 128, 7 frames, 212x352 LQ = 180x320 tile + 16 halo), fp16 engine."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch
from helpers import edvr_state_shapes
from realvsr_b200.archs import EDVR_arch as E
from synth import synth_input, synth_state_dict
kw = dict(nf=128, nc=3, nframes=7, groups=8, front_RBs=5, back_RBs=10, w_TSA=True)
net = E.EDVR(**kw).eval()
net.load_state_dict(synth_state_dict(edvr_state_shapes("EDVR", **kw), 7), strict=True)
net = net.to("cuda:0").half(); net.exec_path = "engine"
x = synth_input((1, 7, 3, 212, 352), 8).to("cuda:0").half()
with torch.no_grad():
    for _ in range(3): net(x)
    rows = net._get_engine(x).profile(x, steps=5)
tot = sum(r["ms"] for r in rows)
print("cfg4 tile 7x3x212x352: %d launches, sum %.3f ms" % (len(rows), tot))
agg = {}
for r in rows:
    k = ":".join(r["label"].split(":")[:2])
    a = agg.setdefault(k, [0.0, 0, 0.0]); a[0] += r["ms"]; a[1] += 1; a[2] += r["flops"]
for k, (ms, n, fl) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-34s %8.1f us  n=%3d  %7.1f TF/s" % (k, ms * 1e3, n, fl / max(ms, 1e-9) / 1e9))
for r in sorted(rows, key=lambda r: -r["ms"])[:14]:
    print("   %-58s %8.1f us %8.1f TF/s" % (r["label"][:58], r["ms"] * 1e3, r["flops"] / max(r["ms"], 1e-9) / 1e9))

"""Print per-launch event timings of one cfg2 forward (fp16 engine) for selected labels."""
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
B = int(os.environ.get("B", "1"))
x = synth_input((B, 5, 3, 180, 320), 8).to("cuda:0").half()
with torch.no_grad():
    for _ in range(3): net(x)
    eng = net._get_engine(x)
    rows = eng.profile(x, steps=5)
want = sys.argv[1:] or ["feature_extraction.0.conv1", "pcd_align.L1_offset_conv1", "L1_dcnpack.conv_offset_mask", "pcd_align.L1_dcnpack", "recon_trunk.0.conv1", "HRconv", "upconv2", "conv_last", "tsa_fusion.fea_fusion"]
tot = sum(r["ms"] for r in rows)
print("dbg=%s B=%d total %.3f ms" % (os.environ.get("RVSR_TC_DEBUG", "0"), B, tot), end=" | ")
for w in want:
    for r in rows:
        if r["label"].endswith(w):
            print("%s %.1fus" % (w[-24:], r["ms"] * 1e3), end="  ")
            break
print()

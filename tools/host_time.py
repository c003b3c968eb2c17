"""Host-side enqueue time of one engine forward (must stay well below the GPU time, or the GPU waits for launches)."""
import os, sys, time
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
for B in (4, 1):
    x = synth_input((B, 5, 3, 180, 320), 8).to("cuda:0").half()
    with torch.no_grad():
        for _ in range(5): net(x)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            torch.cuda.synchronize()
            t0 = time.perf_counter(); net(x); ts.append(time.perf_counter() - t0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20): net(x)
        t_enq = time.perf_counter() - t0
        torch.cuda.synchronize()
        t_all = time.perf_counter() - t0
    print("B=%d: host enqueue of one forward %.3f ms (median), 20 forwards enqueued in %.2f ms, finished in %.2f ms" % (
        B, sorted(ts)[5] * 1e3, t_enq * 1e3, t_all * 1e3))

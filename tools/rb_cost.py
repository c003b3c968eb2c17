"""Marginal cost of one 64->64 3x3 conv launch inside the real (un-instrumented) forward:
step time with extra residual blocks minus the base step, per added conv.  B=4 cfg2 windows."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch
from helpers import edvr_state_shapes
from realvsr_b200.archs import EDVR_arch as E
from synth import synth_input, synth_state_dict

def step_ms(front, back, B=4, iters=20):
    cfg = dict(nf=64, nc=3, nframes=5, groups=8, front_RBs=front, back_RBs=back, w_TSA=True)
    net = E.EDVR(**cfg).eval()
    net.load_state_dict(synth_state_dict(edvr_state_shapes("EDVR", **cfg), 7), strict=True)
    net = net.to("cuda:0").half(); net.exec_path = "engine"
    xs = [synth_input((B, 5, 3, 180, 320), 8 + i).to("cuda:0").half() for i in range(4)]
    with torch.no_grad():
        for i in range(3): net(xs[i % 4])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters): net(xs[i % 4])
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

B = int(os.environ.get("B", "4"))
base = step_ms(5, 10, B)
back = step_ms(5, 30, B)
front = step_ms(15, 10, B)
print("PDL=%s B=%d base %.3f ms | recon conv (%d img) %.2f us each | front conv (%d img) %.2f us each" % (
    os.environ.get("RVSR_PDL", "1"), B, base, B, (back - base) * 1e3 / 40, 5 * B, (front - base) * 1e3 / 20))

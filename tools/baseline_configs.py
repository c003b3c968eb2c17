"""Measured numbers for the BASELINE.json configs that bench.py does not time (cfg4, cfg5); cfg1/2/3 are the parity
gate, the bench line and the multi-GPU bench.  Prints one line per config."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch
import torch.nn.functional as F
from helpers import edvr_state_shapes
from realvsr_b200 import video
from realvsr_b200.archs import EDVR_arch as E
from synth import synth_input, synth_state_dict

dev = "cuda:0"

def timed(fn, n):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n

which = sys.argv[1:] or ["cfg4", "cfg5"]
if "cfg4" in which:
    # 7-frame window, 128-channel variant, 540x960 -> 2160x3840, tiled (nf = 128 on the tcgen05 kernels: 64-channel source split, output halves)
    kw = dict(nf=128, nc=3, nframes=7, groups=8, front_RBs=5, back_RBs=10, w_TSA=True)
    net = E.EDVR(**kw).eval()
    net.load_state_dict(synth_state_dict(edvr_state_shapes("EDVR", **kw), 7), strict=True)
    net = net.to(dev).half(); net.exec_path = "engine"
    x = synth_input((1, 7, 3, 540, 960), 8).to(dev).half()
    y = video.tiled_forward(net, x, tile=(180, 320), halo=16)
    t = timed(lambda: video.tiled_forward(net, x, tile=(180, 320), halo=16), 2)
    print("cfg4: 7x3x540x960 -> %s, nf=128, 9 tiles of 180x320 + 16 halo, fp16 engine: %.1f ms per 4K frame (%.2f frames/s), finite=%s" % (
        tuple(y.shape), t * 1e3, 1 / t, bool(torch.isfinite(y).all())))
    del net, x, y
    torch.cuda.empty_cache()
if "cfg5" in which:
    # training step: batch 16, 5x3x64x64 patches, forward + backward + L1 loss in fp32 (module path, fp32 DCN backward kernel); the bf16
    # step on the train_c8 path (BASELINE cfg5 proper) is bench.py's `cfg5` key / tools/cfg5_profile.py
    kw = dict(nf=64, nc=3, nframes=5, groups=8, front_RBs=5, back_RBs=10, w_TSA=True)
    net = E.EDVR(**kw)
    net.load_state_dict(synth_state_dict(edvr_state_shapes("EDVR", **kw), 7), strict=True)
    net = net.to(dev).train()
    x = synth_input((16, 5, 3, 64, 64), 9).to(dev)
    gt = synth_input((16, 3, 256, 256), 10).to(dev)
    def step():
        net.zero_grad(set_to_none=True)
        loss = F.l1_loss(net(x), gt)
        loss.backward()
        return loss
    t = timed(step, 3)
    gn = sum(float(p.grad.abs().sum()) for p in net.parameters() if p.grad is not None)
    print("cfg5: training step B=16 5x3x64x64 fwd+bwd+L1, fp32 module path (torch convs + rvsr_mdcn_fwd/bwd): %.1f ms/step, grad-sum finite=%s" % (
        t * 1e3, gn == gn and gn != float("inf")))

"""Training soak on the train_c8 path: cfg5's network and batch, N optimizer steps (Adam) replayed from one CUDA graph on a
fixed synthetic batch (the loss must fall), twice from the same initial weights (run-to-run deviation: only the fp32 atomics
inside dcn_bwd_tc_kernel are order-dependent), plus the same number of eager steps on the cuDNN autocast path for reference."""
import copy, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch, torch.nn.functional as F
from helpers import edvr_state_shapes
from realvsr_b200 import train_c8
from realvsr_b200.archs import EDVR_arch as E
from synth import synth_input, synth_state_dict
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
kw = dict(nf=64, nc=3, nframes=5, groups=8, front_RBs=5, back_RBs=10, w_TSA=True)
net0 = E.EDVR(**kw)
net0.load_state_dict(synth_state_dict(edvr_state_shapes("EDVR", **kw), 7), strict=True)
net0 = net0.to("cuda:0").train()
x = synth_input((16, 5, 3, 64, 64), 9).to("cuda:0")
gt = F.interpolate(x[:, 2], scale_factor=4, mode="bicubic", align_corners=False).clamp(0, 1)   # a learnable target
def run_graph():
    net = copy.deepcopy(net0)
    step = train_c8.GraphedStep(net, F.l1_loss, x, gt)
    opt = torch.optim.Adam(net.parameters(), lr=1e-4)
    losses = torch.zeros(steps, device="cuda:0")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(steps):
        losses[i] = step(x, gt)
        opt.step()
    torch.cuda.synchronize()
    return losses.cpu(), (time.perf_counter() - t0) / steps * 1e3
def run_eager_cudnn():
    net = copy.deepcopy(net0); net.exec_path = "module"
    opt = torch.optim.Adam(net.parameters(), lr=1e-4)
    losses = torch.zeros(steps, device="cuda:0")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(steps):
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = F.l1_loss(net(x).float(), gt)
        loss.backward(); opt.step(); losses[i] = loss.detach()
    torch.cuda.synchronize()
    return losses.cpu(), (time.perf_counter() - t0) / steps * 1e3
a, ta = run_graph(); b, tb = run_graph(); c, tc = run_eager_cudnn()
fmt = lambda l: " ".join("%.5f" % float(l[i]) for i in (0, steps // 4, steps // 2, 3 * steps // 4, steps - 1))
print("train_c8 graph  run 1: %.2f ms/step incl. Adam, loss at steps 0, 1/4, 1/2, 3/4, end: %s  finite %s" % (ta, fmt(a), bool(torch.isfinite(a).all())))
print("train_c8 graph  run 2: %.2f ms/step incl. Adam, loss: %s  max |run 1 - run 2| over all steps %.2e" % (tb, fmt(b), float((a - b).abs().max())))
print("cuDNN autocast eager : %.2f ms/step incl. Adam, loss: %s  max |train_c8 - cuDNN| over all steps %.2e" % (tc, fmt(c), float((a - c).abs().max())))

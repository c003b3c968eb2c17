"""Per-parameter gradient agreement of one training step three ways (EDVR nf = 64 crop): fp32 module path, torch autocast(bf16) on
the module path (cuDNN), and the train_c8 path.  Prints cosine / norm ratio against fp32 for every parameter."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch, torch.nn.functional as F
from helpers import load_case
from synth import synth_normal
from realvsr_b200.archs import EDVR_arch as E
torch.backends.cudnn.allow_tf32 = False
c = load_case("edvr_nf64_crop")
net = E.EDVR(**c["kwargs"]).train()
net.load_state_dict(c["sd"], strict=True)
net = net.to("cuda")
NB = int(os.environ.get("NB", "2"))
x = torch.cat([c["x"], c["x"].flip(3), c["x"].flip(4), c["x"].flip(3).flip(4)][:NB], 0).to("cuda")
gt = synth_normal((NB,) + tuple(c["out"].shape[1:]), 55, std=0.3).to("cuda") + 0.5
def grads(path, amp):
    net.exec_path = path
    net.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
        out = net(x).float()
        loss = F.l1_loss(out, gt)
    loss.backward()
    return out.detach(), float(loss.detach()), {n: p.grad.detach().float().clone() for n, p in net.named_parameters()}
o32, l32, g32 = grads("module", False)
oac, lac, gac = grads("module", True)
oc8, lc8, gc8 = grads("train_c8", False)
print("loss", l32, lac, lc8, "out err ac %.3e c8 %.3e" % (float((oac - o32).abs().max()), float((oc8 - o32).abs().max())))
cos = lambda u, v: float(torch.dot(u.flatten(), v.flatten()) / (u.norm() * v.norm()).clamp_min(1e-30))
for n in g32:
    print("%-45s cos c8 %.4f ac %.4f  norm ratio c8 %.3f ac %.3f" % (n, cos(gc8[n], g32[n]), cos(gac[n], g32[n]), float(gc8[n].norm() / g32[n].norm()), float(gac[n].norm() / g32[n].norm())))

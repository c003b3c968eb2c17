"""One small bf16 training step (forward + L1 + backward) on the train_c8 path for compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_train.py
    compute-sanitizer --tool racecheck python tools/sanitize_train.py
    compute-sanitizer --tool synccheck python tools/sanitize_train.py

Covers the bf16 mode of the tcgen05 conv kernels (forward, data gradient, mask / residual epilogues, pixel shuffle, 16-wide
tile), conv_wgrad_tc_kernel + its reduce, the DCN pack forward / backward, the TSA kernels and the layout kernels, at a size
where tiles straddle the image border (nf = 64 network, 2 windows of 3x3x20x36)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from helpers import edvr_state_shapes  # noqa: E402
from realvsr_b200.archs import EDVR_arch as E  # noqa: E402
from synth import synth_input, synth_state_dict  # noqa: E402

kw = dict(nf=64, nc=3, nframes=3, groups=8, front_RBs=1, back_RBs=1, w_TSA=True)
net = E.EDVR(**kw).train()
net.load_state_dict(synth_state_dict(edvr_state_shapes("EDVR", **kw), 7), strict=True)
net = net.cuda()
net.exec_path = "train_c8"
x = synth_input((2, 3, 3, 20, 36), 8).cuda()
gt = synth_input((2, 3, 80, 144), 9).cuda()
loss = F.l1_loss(net(x).float(), gt)
loss.backward()
torch.cuda.synchronize()
gn = sum(float(p.grad.abs().sum()) for p in net.parameters())
print("training step: loss %.5f, |grad| sum %.4f, finite %s, parameters with a gradient %d / %d" % (
    float(loss), gn, gn == gn and gn != float("inf"), sum(p.grad is not None for p in net.parameters()), len(list(net.parameters()))))

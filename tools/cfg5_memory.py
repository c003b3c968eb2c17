"""Peak device memory of one cfg5 training step (B=16, 5x3x64x64) per execution path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch, torch.nn.functional as F
from helpers import edvr_state_shapes
from realvsr_b200.archs import EDVR_arch as E
from synth import synth_input, synth_state_dict
kw = dict(nf=64, nc=3, nframes=5, groups=8, front_RBs=5, back_RBs=10, w_TSA=True)
net = E.EDVR(**kw)
net.load_state_dict(synth_state_dict(edvr_state_shapes("EDVR", **kw), 7), strict=True)
net = net.to("cuda:0").train()
x = synth_input((16, 5, 3, 64, 64), 9).to("cuda:0"); gt = synth_input((16, 3, 256, 256), 10).to("cuda:0")
for name, path, amp in (("train_c8 (bf16, own kernels)", "auto", True), ("module path, bf16 autocast (cuDNN)", "module", True), ("module path, fp32", "module", False)):
    net.exec_path = path
    for i in range(2):
        net.zero_grad(set_to_none=True)
        torch.cuda.synchronize(); torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
            loss = F.l1_loss(net(x).float(), gt)
        loss.backward()
        torch.cuda.synchronize()
    print("%-40s peak allocated %.2f GB" % (name, torch.cuda.max_memory_allocated() / 2**30))

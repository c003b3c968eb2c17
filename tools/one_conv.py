import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from realvsr_b200 import ops
mode = sys.argv[1] if len(sys.argv) > 1 else "c64"
if mode == "c64":
    x = torch.randn(5, 64, 180, 320, device="cuda").half(); w = (torch.randn(64, 64, 3, 3, device="cuda") * 0.05).half()
    kw = dict(act="relu")
elif mode == "c64b":
    x = torch.randn(20, 64, 180, 320, device="cuda").half(); w = (torch.randn(64, 64, 3, 3, device="cuda") * 0.05).half()
    kw = dict(act="relu")
elif mode == "up":
    x = torch.randn(4, 64, 360, 640, device="cuda").half(); w = (torch.randn(256, 64, 3, 3, device="cuda") * 0.05).half()
    kw = dict(act="lrelu", shuffle=True)
if mode == "pack":
    x = torch.randn(20, 64, 180, 320, device="cuda").half(); f = torch.randn(20, 64, 180, 320, device="cuda").half()
    wom = (torch.randn(216, 64, 3, 3, device="cuda") * 0.02).half(); bom = torch.zeros(216, device="cuda").half()
    w = (torch.randn(64, 64, 3, 3, device="cuda") * 0.05).half(); b = torch.zeros(64, device="cuda").half()
    for _ in range(3):
        y = ops.mdcn_pack(x, f, wom, bom, w, b, 8, act="lrelu")
    torch.cuda.synchronize()
    sys.exit(0)
b = torch.zeros(w.shape[0], device="cuda").half()
for _ in range(3):
    y = ops.conv2d_fused(x, w, b, use_tc=True, **kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

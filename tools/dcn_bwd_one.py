"""One ModulatedDeformConvPack backward on the C8 path (dcn_bwd_tc_kernel) at cfg5's L1 size, for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from realvsr_b200 import train_c8 as T
N, H, W = [int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (80, 64, 64))]
g = torch.Generator(device="cuda").manual_seed(1)
x = T.to_c8(torch.randn(N, 64, H, W, device="cuda", generator=g)).requires_grad_()
om = torch.zeros(N, 256, H, W, device="cuda")
# offsets as a network produces them: a smooth field (per-channel constant + low-amplitude pixel noise); SMOOTH=0: i.i.d. per pixel
if os.environ.get("SMOOTH", "1") == "1":
    om[:, :144] = torch.randn(1, 144, 1, 1, device="cuda", generator=g) * 1.5 + torch.randn(N, 144, H, W, device="cuda", generator=g) * 0.1
else:
    om[:, :144] = torch.randn(N, 144, H, W, device="cuda", generator=g) * 1.5
om[:, 144:216] = torch.randn(N, 72, H, W, device="cuda", generator=g)
om = T.to_c8(om).requires_grad_()
w = (torch.randn(64, 64, 3, 3, device="cuda", generator=g) * 0.05).requires_grad_()
b = torch.zeros(64, device="cuda").requires_grad_()
for _ in range(2):
    y = T.dcn_pack(x, om, w, b, "lrelu")
    y.backward(torch.ones_like(y))
torch.cuda.synchronize()

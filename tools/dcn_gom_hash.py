import os, sys, hashlib
sys.path.insert(0, '/root/repo')
import torch
from realvsr_b200 import train_c8 as T
g = torch.Generator(device="cuda").manual_seed(1)
N,H,W = 6, 40, 50
x = T.to_c8(torch.randn(N, 64, H, W, device="cuda", generator=g)).requires_grad_()
om = torch.zeros(N, 256, H, W, device="cuda")
om[:, :144] = torch.randn(1, 144, 1, 1, device="cuda", generator=g) * 2.5 + torch.randn(N, 144, H, W, device="cuda", generator=g) * 0.3
om[:, 144:216] = torch.randn(N, 72, H, W, device="cuda", generator=g)
om = T.to_c8(om).requires_grad_()
w = (torch.randn(64, 64, 3, 3, device="cuda", generator=g) * 0.05).requires_grad_()
b = torch.zeros(64, device="cuda").requires_grad_()
y = T.dcn_pack(x, om, w, b, "lrelu")
y.backward(torch.randn(y.shape, device="cuda", generator=g).to(y.dtype))
torch.cuda.synchronize()
print(hashlib.md5(om.grad.cpu().view(torch.int16).numpy().tobytes()).hexdigest(), float(om.grad.float().abs().sum()), float(om.grad[:, 27:].float().abs().max()))

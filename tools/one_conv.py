import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from realvsr_b200 import ops
x = torch.randn(5, 64, 180, 320, device="cuda").half(); w = (torch.randn(64, 64, 3, 3, device="cuda") * 0.05).half()
b = torch.zeros(64, device="cuda").half()
for _ in range(2):
    y = ops.conv2d_fused(x, w, b, act="relu", use_tc=True)
torch.cuda.synchronize()

"""One conv_wgrad_tc_kernel launch shape, a few launches (for ncu)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from realvsr_b200 import _lib
L = _lib.lib()
N, H, W, Cout = [int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (80, 64, 64, 64))]
x = torch.randn(N, 8, H, W, 8, device="cuda").bfloat16()
g = torch.randn(N, Cout // 8, H, W, 8, device="cuda").bfloat16()
dw = torch.zeros(Cout, 64, 3, 3, device="cuda"); db = torch.zeros(Cout, device="cuda")
ws = torch.empty(L.rvsr_c8_conv_wgrad_workspace_bytes(N, H, W, Cout), dtype=torch.uint8, device="cuda")
s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(3):
    _lib.check(L.rvsr_c8_conv_wgrad(x.data_ptr(), x.stride(0), g.data_ptr(), dw.data_ptr(), db.data_ptr(), N, H, W, 64, Cout, 3, 64, 0, ws.data_ptr(), ws.numel(), s))
torch.cuda.synchronize()

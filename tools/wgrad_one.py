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
ws = torch.empty(L.rvsr_c8_conv_wgrad_workspace_bytes(1, N, H, W, Cout), dtype=torch.uint8, device="cuda")
JOB = ((ctypes.c_void_p * 1)(x.data_ptr()), (ctypes.c_longlong * 1)(x.stride(0)), (ctypes.c_void_p * 1)(g.data_ptr()), (ctypes.c_void_p * 1)(dw.data_ptr()),
       (ctypes.c_void_p * 1)(db.data_ptr()), (ctypes.c_int * 1)(64), (ctypes.c_int * 1)(0))
s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(3):
    _lib.check(L.rvsr_c8_conv_wgrad(1, *JOB, N, H, W, 64, Cout, 3, ws.data_ptr(), ws.numel(), s))
torch.cuda.synchronize()

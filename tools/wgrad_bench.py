"""conv_wgrad_tc_kernel / bf16 conv timing per launch (CUDA events, 20 launches) over image counts and sizes."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from realvsr_b200 import _lib, train_c8 as T
L = _lib.lib()
dev = "cuda:0"
def timeit(fn, n=20):
    """n launches captured into one CUDA graph (no host launch overhead between them), replayed 3 times"""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3): fn(side)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn(torch.cuda.current_stream())
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * n) * 1e3
S = lambda st: ctypes.c_void_p(st.cuda_stream)
for (N, H, W, Cout) in [(80, 64, 64, 64), (40, 64, 64, 64), (16, 64, 64, 64), (4, 64, 64, 64), (80, 32, 32, 64), (80, 16, 16, 64),
                        (16, 256, 256, 64), (16, 128, 128, 256), (80, 64, 64, 256)]:
    x = torch.randn(N, 8, H, W, 8, device=dev).bfloat16()
    g = torch.randn(N, Cout // 8, H, W, 8, device=dev).bfloat16()
    dw = torch.zeros(Cout, 64, 3, 3, device=dev)
    ws = torch.empty(L.rvsr_c8_conv_wgrad_workspace_bytes(1, N, H, W, Cout), dtype=torch.uint8, device=dev)
    db = torch.zeros(Cout, device=dev)
    JOB = ((ctypes.c_void_p * 1)(x.data_ptr()), (ctypes.c_longlong * 1)(x.stride(0)), (ctypes.c_void_p * 1)(g.data_ptr()), (ctypes.c_void_p * 1)(dw.data_ptr()),
           (ctypes.c_void_p * 1)(db.data_ptr()), (ctypes.c_int * 1)(64), (ctypes.c_int * 1)(0))
    w = torch.randn(Cout, 64, 3, 3, device=dev) * 0.05
    t_w = timeit(lambda st: _lib.check(L.rvsr_c8_conv_wgrad(1, *JOB, N, H, W, 64, Cout, 3, ws.data_ptr(), ws.numel(), S(st))))
    wp = T._pack_weight(w, Cout, 64, 3, False, 0, 64, 0, (1, 64, N, H, W))
    t_f = timeit(lambda st: T._conv_launch([x], wp, None, None, N, H, W, 64, Cout, 3, 1, False))
    fl = 2.0 * N * H * W * 64 * Cout * 9
    print("N=%3d %3dx%3d Cout=%3d: wgrad %7.1f us (%5.2f PFLOP/s)   fwd conv %7.1f us (%5.2f PFLOP/s)" % (N, H, W, Cout, t_w, fl / t_w * 1e-9, t_f, fl / t_f * 1e-9))

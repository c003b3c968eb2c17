"""Image I/O around the model, on the GPU.

The reference's test loop converts on the host: ``read_img_seq`` (data/util.py:104-122) turns uint8 files into a
float [T, C, H, W] tensor, and after the model ``tensor2img`` (utils/util.py:151-181) + ``ycbcr2bgr``
(data/util.py:397-416) produce the uint8 BGR image that ``cv2.imwrite`` stores (test_RealVSR_wi_GT.py:98,:121-128).
Here both directions are one CUDA kernel each, so only uint8 crosses PCIe (1 byte per sample instead of 4)."""
import ctypes

import torch

from . import _lib


def _dt(t):
    return _lib.F16 if t.dtype == torch.float16 else _lib.F32


def frames_from_uint8(u8, reverse_channels=True, dtype=torch.float32):
    """u8 [T, H, W, C] uint8 CUDA tensor (as cv2.imread returns the files) -> [T, C, H, W] in [0, 1].
    reverse_channels=True reproduces read_img_seq's ``imgs[:, :, :, [2, 1, 0]]``."""
    if not u8.is_cuda:
        raise NotImplementedError("realvsr_b200.color runs on CUDA tensors only (no CPU fallback)")
    if u8.dtype != torch.uint8 or u8.dim() != 4:
        raise RuntimeError("frames_from_uint8: expected a [T, H, W, C] uint8 tensor, got %s %s" % (u8.dtype, tuple(u8.shape)))
    if dtype not in (torch.float32, torch.float16):
        raise RuntimeError("frames_from_uint8: dtype must be float32 or float16")
    u8 = u8.contiguous()
    T, H, W, C = u8.shape
    out = torch.empty(T, C, H, W, dtype=dtype, device=u8.device)
    with torch.cuda.device(u8.device):
        _lib.check(_lib.lib().rvsr_frames_from_u8(ctypes.c_void_p(u8.data_ptr()), ctypes.c_void_p(out.data_ptr()), T, C, H, W,
                                                  1 if reverse_channels else 0, _dt(out),
                                                  ctypes.c_void_p(torch.cuda.current_stream(u8.device).cuda_stream)),
                   "frames_from_u8")
    return out


def frames_to_bgr_uint8(x, color="YCbCr"):
    """Network output [B, 3, H, W] (fp32 / fp16, CUDA) -> uint8 BGR images [B, H, W, 3].
    color="YCbCr": tensor2img(float32, reverse_channel=False) + ycbcr2bgr + clip/round (test_RealVSR_wi_GT.py:122-123);
    color="RGB":   tensor2img(uint8, reverse_channel=True) (:128)."""
    if not x.is_cuda:
        raise NotImplementedError("realvsr_b200.color runs on CUDA tensors only (no CPU fallback)")
    if x.dim() != 4 or x.shape[1] != 3 or x.dtype not in (torch.float32, torch.float16):
        raise RuntimeError("frames_to_bgr_uint8: expected a [B, 3, H, W] float tensor, got %s %s" % (x.dtype, tuple(x.shape)))
    if color not in ("YCbCr", "RGB"):
        raise RuntimeError("frames_to_bgr_uint8: color must be 'YCbCr' or 'RGB'")
    x = x.contiguous()
    B, C, H, W = x.shape
    out = torch.empty(B, H, W, 3, dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().rvsr_frames_to_u8(ctypes.c_void_p(x.data_ptr()), _dt(x), ctypes.c_void_p(out.data_ptr()), B, C, H, W,
                                                1 if color == "YCbCr" else 0,
                                                ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)),
                   "frames_to_u8")
    return out

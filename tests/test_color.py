"""Image I/O around the model: the numpy oracle and the CUDA kernels against vectors produced by
the reference's own read_img_seq / tensor2img / ycbcr2bgr (tests/golden/make_golden_color.py).  Integer outputs:
bit-exact."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN


def _gold():
    return np.load(os.path.join(GOLDEN, "color_io.npz"))


def test_oracle_matches_reference_vectors():
    from oracle import color_oracle as C
    z = _gold()
    assert np.array_equal(C.frames_from_uint8(z["u8"]), z["frames"])
    assert np.array_equal(C.frames_to_bgr_uint8(z["out"], "YCbCr"), z["bgr_from_ycbcr"])
    assert np.array_equal(C.frames_to_bgr_uint8(z["out"], "RGB"), z["bgr_from_rgb"])


def test_color_rejects_cpu_tensors():
    from realvsr_b200 import color
    with pytest.raises(NotImplementedError):
        color.frames_from_uint8(torch.zeros(1, 4, 4, 3, dtype=torch.uint8))
    with pytest.raises(NotImplementedError):
        color.frames_to_bgr_uint8(torch.zeros(1, 3, 4, 4))


@pytest.mark.gpu
def test_frames_from_uint8_bit_exact():
    from realvsr_b200 import color
    z = _gold()
    got = color.frames_from_uint8(torch.from_numpy(z["u8"]).cuda()).cpu().numpy()
    assert np.array_equal(got, z["frames"])
    half = color.frames_from_uint8(torch.from_numpy(z["u8"]).cuda(), dtype=torch.float16).cpu().numpy()
    assert np.array_equal(half, z["frames"].astype(np.float16))
    keep = color.frames_from_uint8(torch.from_numpy(z["u8"]).cuda(), reverse_channels=False).cpu().numpy()
    assert np.array_equal(keep, z["frames"][:, ::-1])


@pytest.mark.gpu
@pytest.mark.parametrize("mode,key", [("YCbCr", "bgr_from_ycbcr"), ("RGB", "bgr_from_rgb")])
def test_frames_to_bgr_uint8_bit_exact(mode, key):
    from realvsr_b200 import color
    z = _gold()
    got = color.frames_to_bgr_uint8(torch.from_numpy(z["out"]).cuda(), mode).cpu().numpy()
    assert got.shape == z[key].shape
    assert np.array_equal(got, z[key])


@pytest.mark.gpu
def test_frames_to_bgr_uint8_full_frame_against_oracle():
    """720p frame, fp32 and fp16 inputs, against the numpy oracle (bit-exact)."""
    from oracle import color_oracle as C
    from realvsr_b200 import color
    g = torch.Generator().manual_seed(3)
    x = torch.rand(2, 3, 720, 1280, generator=g) * 1.2 - 0.1
    for mode in ("YCbCr", "RGB"):
        assert np.array_equal(color.frames_to_bgr_uint8(x.cuda(), mode).cpu().numpy(), C.frames_to_bgr_uint8(x.numpy(), mode))
    xh = x.half()
    assert np.array_equal(color.frames_to_bgr_uint8(xh.cuda(), "YCbCr").cpu().numpy(),
                          C.frames_to_bgr_uint8(xh.float().numpy(), "YCbCr"))


@pytest.mark.gpu
def test_color_error_behaviour():
    from realvsr_b200 import color
    with pytest.raises(RuntimeError):
        color.frames_from_uint8(torch.zeros(1, 4, 4, 3, device="cuda"))          # not uint8
    with pytest.raises(RuntimeError):
        color.frames_to_bgr_uint8(torch.zeros(1, 1, 4, 4, device="cuda"))        # one channel
    with pytest.raises(RuntimeError):
        color.frames_to_bgr_uint8(torch.zeros(1, 3, 4, 4, device="cuda"), "HSV")

"""oracle/color_oracle.py -- TEST INFRASTRUCTURE (imported by tests/ only).

numpy restatement of the reference's host-side image I/O around the model:
  frames_from_uint8   data/util.py::read_img (:87-101) + ::read_img_seq (:104-122)
  frames_to_bgr_uint8 utils/util.py::tensor2img (:151-181), data/util.py::ycbcr2bgr (:397-416) and the clip / round of
                      test_RealVSR_wi_GT.py:122-123, :128
Pinned against tests/golden/color_io.npz, which tests/golden/make_golden_color.py produced by calling the
reference's own functions."""
import numpy as np


def frames_from_uint8(u8_thwc, reverse_channels=True):
    img = u8_thwc.astype(np.float32) / 255.                 # read_img :97
    if reverse_channels and img.shape[-1] == 3:
        img = img[:, :, :, [2, 1, 0]]                        # read_img_seq :120
    return np.ascontiguousarray(np.transpose(img, (0, 3, 1, 2)))  # :121


def frames_to_bgr_uint8(x_bchw, color="YCbCr"):
    out = []
    for t in x_bchw.astype(np.float32):
        t = np.clip(t, 0, 1)                                 # tensor2img :157-158 (min_max = (0, 1))
        if color == "RGB":
            img = np.transpose(t[[2, 1, 0], :, :], (1, 2, 0))  # :169 CHW -> HWC, RGB -> BGR
            out.append((img * 255.0).round().astype(np.uint8))  # :178-181
        else:
            img = np.transpose(t, (1, 2, 0)).copy()          # :171 (reverse_channel=False), out_type float32
            img *= 255.                                      # ycbcr2bgr :405-406 (float input)
            rlt = np.matmul(img, [[0.00456621, 0.00456621, 0.00456621], [0.00791071, -0.00153632, 0],
                                  [0, -0.00318811, 0.00625893]]) * 255.0 + [-276.836, 135.576, -222.921]  # :408-410
            rlt /= 255.                                      # :414
            rlt = rlt.astype(np.float32)                     # :415
            out.append((np.clip(rlt, 0, 1) * 255.).round().astype(np.uint8))  # test_RealVSR_wi_GT.py:123
    return np.stack(out, 0)

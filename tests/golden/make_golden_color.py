"""Golden vectors for the image I/O around the model (tests/golden/color_io.npz), produced by the REFERENCE's own
functions in the build container:

    python tests/golden/make_golden_color.py

data/util.py imports cleanly here (cv2, PIL present); utils/util.py imports ``ffmpeg`` at :9, which is absent and
unused by tensor2img -- it is stubbed with an empty module.  read_img_seq reads files: the uint8 frames are written
as PNGs into a temporary folder first (lossless), so the reference's cv2.imread path is exercised as well."""
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("RVSR_REFERENCE", "/root/reference")


def main():
    sys.path.insert(0, os.path.join(REF, "codes"))
    for missing in ("ffmpeg", "lmdb", "kornia"):
        try:
            __import__(missing)
        except ImportError:
            sys.modules[missing] = types.ModuleType(missing)
    import cv2
    import data.util as data_util
    import utils.util as util

    rng = np.random.RandomState(91)
    T, H, W = 3, 20, 28
    u8 = rng.randint(0, 256, size=(T, H, W, 3)).astype(np.uint8)
    u8[0, :2] = 0
    u8[0, 2:4] = 255                                         # extremes
    with tempfile.TemporaryDirectory() as d:
        for t in range(T):
            cv2.imwrite(os.path.join(d, "%03d.png" % t), u8[t])
        frames = data_util.read_img_seq(d).numpy()           # [T, C, H, W] float32, channels reversed
    # model outputs: beyond [0, 1] on both sides, exact ties of x * 255 at .5, fp16-representable values
    out = (rng.standard_normal((4, 3, H, W)) * 0.35 + 0.5).astype(np.float32)
    out[0, :, 0, :8] = (np.arange(8, dtype=np.float32) + 0.5) / 255.0
    out[1] = out[1].astype(np.float16).astype(np.float32)
    ycc, rgb = [], []
    for b in range(out.shape[0]):
        o = util.tensor2img(torch.from_numpy(out[b]), out_type=np.float32, reverse_channel=False)   # test_RealVSR_wi_GT.py:122
        ycc.append((np.clip(data_util.ycbcr2bgr(o), 0, 1) * 255.).round().astype(np.uint8))         # :123
        rgb.append(util.tensor2img(torch.from_numpy(out[b]), out_type=np.uint8, reverse_channel=True))  # :128
    np.savez_compressed(os.path.join(HERE, "color_io.npz"), u8=u8, frames=frames, out=out, bgr_from_ycbcr=np.stack(ycc),
                        bgr_from_rgb=np.stack(rgb))
    print("color_io", frames.shape, np.stack(ycc).shape, float(frames.mean()))


if __name__ == "__main__":
    main()

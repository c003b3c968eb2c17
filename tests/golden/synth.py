"""Deterministic synthetic weights / inputs shared by the golden generator, the
parity tests, smoke() and bench.py.

Everything is drawn from numpy's legacy ``RandomState`` (MT19937), whose stream is
frozen across numpy versions, so fixtures only need to store a seed -- not the
3.3 M parameters of the full-size network.

Why not the reference's own init: ``ModulatedDeformConvPack`` zero-initialises
``conv_offset_mask`` (reference dcn/deform_conv.py:270-272), which makes every
offset 0 and every mask 0.5 -- the bilinear gather and its border rules would
never be exercised (trap 1).
"""
import numpy as np
import torch


def synth_state_dict(shapes, seed, offset_std=0.03, offset_bias_std=0.6):
    """shapes: ordered {name: shape} as produced by ``module.state_dict()``.

    * conv weights: N(0, (gain/sqrt(fan_in))^2), gain 1.0 (0.6 inside residual blocks so
      15 stacked blocks stay bounded)
    * biases: N(0, 0.05^2)  (non-zero so the bias path is tested)
    * ``*.conv_offset_mask.weight``: N(0, offset_std^2); its bias N(0, offset_bias_std^2)
      -> offsets of a pixel or two, some leaving the image at the borders.
    """
    rng = np.random.RandomState(seed)
    out = {}
    for name, shape in shapes.items():
        shape = tuple(int(s) for s in shape)
        if name.endswith("conv_offset_mask.weight"):
            a = rng.standard_normal(shape) * offset_std
        elif name.endswith("conv_offset_mask.bias"):
            a = rng.standard_normal(shape) * offset_bias_std
        elif name.endswith(".weight") and len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            gain = 0.6 if (".conv1." in name or ".conv2." in name) else 1.0
            a = rng.standard_normal(shape) * (gain / np.sqrt(fan_in))
        else:
            a = rng.standard_normal(shape) * 0.05
        out[name] = torch.from_numpy(a.astype(np.float32))
    return out


def synth_input(shape, seed):
    """LQ clip in [0, 1), like image data."""
    rng = np.random.RandomState(seed)
    return torch.from_numpy(rng.random_sample(tuple(shape)).astype(np.float32))


def synth_normal(shape, seed, std=1.0):
    rng = np.random.RandomState(seed)
    return torch.from_numpy((rng.standard_normal(tuple(shape)) * std).astype(np.float32))

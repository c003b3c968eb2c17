"""Shared helpers for the parity tests (fixtures, tolerances)."""
import ast
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EDVR_CASES = ["edvr_tiny", "edvr_tiny_b2_g2", "edvr_noup_3f", "edvr_nf64_crop",
              "edvr_noup_nf64_ship", "edvr_predeblur", "edvr_nf128_7f", "edvr_hr_in", "edvr_predeblur_hr_in"]


def load_case(name):
    """-> dict(cls, kwargs, x, sd, out, aligned0) for a golden EDVR case; weights and
    input are regenerated from the stored seeds (tests/golden/synth.py)."""
    from synth import synth_input, synth_state_dict
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    kwargs = ast.literal_eval(str(z["kwargs"]))
    cls = str(z["cls"])
    shapes = edvr_state_shapes(cls, **kwargs)
    assert list(shapes.keys()) == [str(k) for k in z["keys"]], "state_dict key order drifted"
    return dict(cls=cls, kwargs=kwargs, x=synth_input(tuple(z["shape"]), int(z["xseed"])),
                sd=synth_state_dict(shapes, int(z["wseed"])), out=torch.from_numpy(z["out"]),
                aligned0=torch.from_numpy(z["aligned0"]))


def edvr_state_shapes(cls, nf=64, nc=3, nframes=5, groups=8, front_RBs=5, back_RBs=10, center=None,
                      predeblur=False, HR_in=False, w_TSA=True):
    """The reference's state_dict contract: ordered {name: shape}, built
    without instantiating any module so tests can check the product's modules against it."""
    up = cls == "EDVR"
    s = {}

    def conv(name, co, ci, k):
        s[name + ".weight"] = (co, ci, k, k)
        s[name + ".bias"] = (co,)

    def rb(name, n=nf):
        conv(name + ".conv1", n, n, 3)
        conv(name + ".conv2", n, n, 3)

    if up and predeblur:
        p = "pre_deblur"
        if HR_in:
            conv(p + ".conv_first_1", nf, 3, 3); conv(p + ".conv_first_2", nf, nf, 3)
            conv(p + ".conv_first_3", nf, nf, 3)
        else:
            conv(p + ".conv_first", nf, 3, 3)
        for n in ("RB_L1_1", "RB_L1_2", "RB_L1_3", "RB_L1_4", "RB_L1_5", "RB_L2_1", "RB_L2_2",
                  "RB_L3_1"):
            rb(p + "." + n)
        conv(p + ".deblur_L2_conv", nf, nf, 3); conv(p + ".deblur_L3_conv", nf, nf, 3)
        conv("conv_1x1", nf, nf, 1)
    elif up and HR_in:
        conv("conv_first_1", nf, nc, 3); conv("conv_first_2", nf, nf, 3); conv("conv_first_3", nf, nf, 3)
    else:
        conv("conv_first", nf, nc, 3)
    for i in range(front_RBs):
        rb("feature_extraction.%d" % i)
    for n in ("fea_L2_conv1", "fea_L2_conv2", "fea_L3_conv1", "fea_L3_conv2"):
        conv(n, nf, nf, 3)
    p = "pcd_align."

    def dcn(name):
        s[p + name + ".weight"] = (nf, nf, 3, 3)
        s[p + name + ".bias"] = (nf,)
        conv(p + name + ".conv_offset_mask", groups * 27, nf, 3)

    conv(p + "L3_offset_conv1", nf, 2 * nf, 3); conv(p + "L3_offset_conv2", nf, nf, 3); dcn("L3_dcnpack")
    conv(p + "L2_offset_conv1", nf, 2 * nf, 3); conv(p + "L2_offset_conv2", nf, 2 * nf, 3)
    conv(p + "L2_offset_conv3", nf, nf, 3); dcn("L2_dcnpack"); conv(p + "L2_fea_conv", nf, 2 * nf, 3)
    conv(p + "L1_offset_conv1", nf, 2 * nf, 3); conv(p + "L1_offset_conv2", nf, 2 * nf, 3)
    conv(p + "L1_offset_conv3", nf, nf, 3); dcn("L1_dcnpack"); conv(p + "L1_fea_conv", nf, 2 * nf, 3)
    conv(p + "cas_offset_conv1", nf, 2 * nf, 3); conv(p + "cas_offset_conv2", nf, nf, 3); dcn("cas_dcnpack")
    if w_TSA:
        t = "tsa_fusion."
        conv(t + "tAtt_1", nf, nf, 3); conv(t + "tAtt_2", nf, nf, 3)
        conv(t + "fea_fusion", nf, nframes * nf, 1); conv(t + "sAtt_1", nf, nframes * nf, 1)
        conv(t + "sAtt_2", nf, 2 * nf, 1); conv(t + "sAtt_3", nf, nf, 3); conv(t + "sAtt_4", nf, nf, 1)
        conv(t + "sAtt_5", nf, nf, 3); conv(t + "sAtt_L1", nf, nf, 1); conv(t + "sAtt_L2", nf, 2 * nf, 3)
        conv(t + "sAtt_L3", nf, nf, 3); conv(t + "sAtt_add_1", nf, nf, 1); conv(t + "sAtt_add_2", nf, nf, 1)
    else:
        conv("tsa_fusion", nf, nframes * nf, 1)
    for i in range(back_RBs):
        rb("recon_trunk.%d" % i)
    if up:
        conv("upconv1", nf * 4, nf, 3); conv("upconv2", 256, nf, 3)
    conv("HRconv", 64, 64, 3); conv("conv_last", nc, 64, 3)
    return s


def rel_err(a, b):
    """max |a-b| / max |b|  -- the 'relative fp32 tolerance' used throughout."""
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))

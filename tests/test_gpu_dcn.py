"""GPU parity tests of the DCNv2 operator through the C ABI (rvsr_mdcn_fwd via the
autograd Function in realvsr_b200/archs/dcn/deform_conv.py).

Tolerances (north_star: 1e-3 relative, fp32): fp32 path asserts 1e-4 of max|ref| (only the
summation order differs from the oracle); fp16 I/O asserts 2e-3 against the fp32 oracle
evaluated on the same fp16-rounded inputs (fp32 coordinates / blend / accumulate inside)."""
import os
import sys

import numpy as np
import pytest
import torch

import realvsr_b200.archs.dcn.deform_conv  # noqa: F401
from helpers import GOLDEN, rel_err
from oracle import edvr_oracle as O
from synth import synth_normal

D = sys.modules['realvsr_b200.archs.dcn.deform_conv']
pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _case(B, C, H, W, Cout, dg, groups=1, stride=1, pad=1, dil=1, k=3, off_std=3.0, seed=100):
    Ho = (H + 2 * pad - (dil * (k - 1) + 1)) // stride + 1
    Wo = (W + 2 * pad - (dil * (k - 1) + 1)) // stride + 1
    x = synth_normal((B, C, H, W), seed)
    off = synth_normal((B, dg * 2 * k * k, Ho, Wo), seed + 1, std=off_std)
    msk = torch.sigmoid(synth_normal((B, dg * k * k, Ho, Wo), seed + 2))
    w = synth_normal((Cout, C // groups, k, k), seed + 3, std=0.1)
    b = synth_normal((Cout,), seed + 4)
    return x, off, msk, w, b, (stride, pad, dil, groups, dg)


CASES = [
    dict(B=2, C=16, H=11, W=13, Cout=12, dg=4),                       # ragged tile edges, offsets leave the image
    dict(B=1, C=64, H=24, W=40, Cout=64, dg=8),                       # EDVR shape class (8 ch / group)
    dict(B=1, C=8, H=16, W=16, Cout=8, dg=8),                         # 1 channel per group (tiny arch)
    dict(B=2, C=8, H=12, W=9, Cout=6, dg=4, groups=2, stride=2),      # conv groups + stride 2
    dict(B=1, C=16, H=10, W=10, Cout=16, dg=2, pad=2, dil=2),         # dilation
    dict(B=1, C=24, H=9, W=7, Cout=70, dg=3),                         # Cout > one tile, C not /16
    dict(B=1, C=16, H=8, W=8, Cout=8, dg=2, k=1, pad=0),              # 1x1 kernel
    dict(B=1, C=16, H=6, W=6, Cout=8, dg=4, off_std=40.0),            # everything out of range -> bias only
]


@pytest.mark.parametrize("cfg", CASES)
def test_mdcn_fwd_fp32_vs_oracle(cfg):
    x, off, msk, w, b, (s, p, d, g, dg) = _case(**cfg)
    ref = O.dcn_forward(x, off, msk, w, b, s, p, d, g, dg)
    y = D.modulated_deform_conv(x.to(DEV), off.to(DEV), msk.to(DEV), w.to(DEV), b.to(DEV), s, p, d, g, dg)
    assert y.shape == ref.shape and y.dtype == torch.float32
    assert rel_err(y.cpu(), ref) < 1e-4


@pytest.mark.parametrize("cfg", CASES[:4])
def test_mdcn_fwd_fp16_vs_oracle(cfg):
    x, off, msk, w, b, (s, p, d, g, dg) = _case(**cfg)
    h = lambda t: t.half().float()  # noqa: E731
    ref = O.dcn_forward(h(x), h(off), h(msk), h(w), h(b), s, p, d, g, dg)
    y = D.modulated_deform_conv(x.to(DEV).half(), off.to(DEV).half(), msk.to(DEV).half(), w.to(DEV).half(),
                                b.to(DEV).half(), s, p, d, g, dg)
    assert y.dtype == torch.float16
    # EDVR's 64-channel / 8-group shape runs the tcgen05 gather -> UMMA kernel (fp16x2 bilinear blend: up to three more fp16
    # roundings per sample, same bound as the fused pack); the other shapes the CUDA-core kernel with an fp32 blend
    assert rel_err(y.float().cpu(), ref) < (4e-3 if cfg["C"] == 64 else 2e-3)


def test_mdcn_fwd_matches_reference_generated_golden():
    z = np.load(os.path.join(GOLDEN, "dcn_unit.npz"))
    B, C, H, W, Cout, dg = [int(v) for v in z["dims"]]
    x = synth_normal((B, C, H, W), 71); off = synth_normal((B, dg * 18, H, W), 72, std=4.0)
    msk = torch.sigmoid(synth_normal((B, dg * 9, H, W), 73).double()).float()
    w = synth_normal((Cout, C, 3, 3), 74, std=0.1); b = synth_normal((Cout,), 75)
    y = D.modulated_deform_conv(x.to(DEV), off.to(DEV), msk.to(DEV), w.to(DEV), b.to(DEV), 1, 1, 1, 1, dg)
    assert rel_err(y.cpu(), torch.from_numpy(z["out"]).float()) < 1e-4


def test_mdcn_zero_offset_is_half_conv_and_no_bias():
    x = synth_normal((2, 16, 20, 20), 5).to(DEV)
    w = synth_normal((16, 16, 3, 3), 6, std=0.1).to(DEV)
    y = D.modulated_deform_conv(x, torch.zeros(2, 2 * 18, 20, 20, device=DEV),
                                torch.full((2, 2 * 9, 20, 20), 0.5, device=DEV), w, None, 1, 1, 1, 1, 2)
    assert rel_err(y, 0.5 * torch.nn.functional.conv2d(x, w, padding=1)) < 1e-4


def test_mdcn_error_behaviour():
    x, off, msk, w, b, (s, p, d, g, dg) = _case(B=1, C=16, H=8, W=8, Cout=8, dg=4)
    with pytest.raises(NotImplementedError):            # CPU tensors, like the reference
        D.modulated_deform_conv(x, off, msk, w, b, s, p, d, g, dg)
    with pytest.raises(RuntimeError):                   # non-contiguous input (reference TORCH_CHECK)
        D.modulated_deform_conv(x.to(DEV).transpose(2, 3), off.to(DEV), msk.to(DEV), w.to(DEV), b.to(DEV), s, p, d,
                                g, dg)
    with pytest.raises(RuntimeError):                   # channels % deformable groups
        D.modulated_deform_conv(x.to(DEV)[:, :15].contiguous(), off.to(DEV), msk.to(DEV), w.to(DEV)[:, :15].contiguous(),
                                b.to(DEV), s, p, d, g, dg)
    y = D.modulated_deform_conv(x.to(DEV)[:0], off.to(DEV)[:0], msk.to(DEV)[:0], w.to(DEV), b.to(DEV), s, p, d, g, dg)
    assert y.shape[0] == 0                              # empty batch


def test_mdcn_fwd_vs_reference_cuda_extension():
    """Second oracle: the reference's own deform_conv_cuda extension, compiled unmodified
    from /root/reference into oracle/_ref (oracle/build_ref.py)."""
    from oracle.build_ref import load_ref
    ref_ext = load_ref()
    if ref_ext is None:
        pytest.skip("oracle/_ref not built")
    x, off, msk, w, b, (s, p, d, g, dg) = _case(B=2, C=64, H=20, W=28, Cout=64, dg=8, off_std=2.0)
    xd, od, md, wd, bd = [t.to(DEV) for t in (x, off, msk, w, b)]
    out_ref = xd.new_empty(2, 64, 20, 28)
    ref_ext.modulated_deform_conv_cuda_forward(xd, wd, bd, xd.new_empty(0), od, md, out_ref, xd.new_empty(0), 3, 3,
                                               s, s, p, p, d, d, g, dg, True)
    y = D.modulated_deform_conv(xd, od, md, wd, bd, s, p, d, g, dg)
    # the reference GEMM (cuBLAS addmm_) may run in TF32; both sides only agree to ~1e-3
    assert rel_err(y, out_ref) < 1e-3
    assert rel_err(out_ref.cpu(), O.dcn_forward(x, off, msk, w, b, s, p, d, g, dg)) < 1e-3


# ---------------------------------------------------------------- fused pack operator (rvsr_mdcn_pack_fwd)
def _pack_case(B, C, H, W, Cout, dg, seed=500, off_w=0.05, off_b=1.0):
    x = synth_normal((B, C, H, W), seed)
    feat = synth_normal((B, C, H, W), seed + 1)
    wom = synth_normal((27 * dg, C, 3, 3), seed + 2, std=off_w)
    bom = synth_normal((27 * dg,), seed + 3, std=off_b)
    w = synth_normal((Cout, C, 3, 3), seed + 4, std=(1.0 / (C * 9)) ** 0.5)
    b = synth_normal((Cout,), seed + 5, std=0.3)
    return x, feat, wom, bom, w, b


def _pack_ref(x, feat, wom, bom, w, b, dg, act):
    om = torch.nn.functional.conv2d(feat, wom, bom, padding=1)
    off, msk = om[:, :18 * dg].contiguous(), torch.sigmoid(om[:, 18 * dg:]).contiguous()
    y = O.dcn_forward(x, off, msk, w, b, 1, 1, 1, 1, dg)
    return torch.nn.functional.leaky_relu(y, 0.1) if act == "lrelu" else y


PACK_CASES = [dict(B=2, C=64, H=20, W=36, Cout=64, dg=8), dict(B=1, C=64, H=9, W=70, Cout=64, dg=8, off_b=3.0),
              dict(B=1, C=64, H=16, W=32, Cout=64, dg=4), dict(B=1, C=16, H=12, W=12, Cout=16, dg=4),
              dict(B=1, C=64, H=12, W=16, Cout=64, dg=1)]


@pytest.mark.parametrize("cfg", PACK_CASES)
@pytest.mark.parametrize("act", [None, "lrelu"])
def test_mdcn_pack_fp32_vs_oracle(cfg, act):
    from realvsr_b200 import ops
    dg = cfg["dg"]
    t = _pack_case(**cfg)
    ref = _pack_ref(*t, dg, act)
    y = ops.mdcn_pack(*[v.to(DEV) for v in t], dg, act=act)
    assert rel_err(y.cpu(), ref) < 1e-4


@pytest.mark.parametrize("cfg", PACK_CASES)
def test_mdcn_pack_fp16_vs_oracle(cfg):
    """64-channel cases run the tcgen05 pair (offset conv -> OUT_OM24 -> gather + UMMA).
    Tolerance 4e-3: fp16 storage of x / weights / output, masks stored as fp16; offsets stay fp32."""
    from realvsr_b200 import ops
    dg = cfg["dg"]
    t = [v.half().float() for v in _pack_case(**cfg)]
    ref = _pack_ref(*t, dg, "lrelu")
    y = ops.mdcn_pack(*[v.to(DEV).half() for v in t], dg, act="lrelu")
    assert y.dtype == torch.float16
    assert rel_err(y.float().cpu(), ref) < 4e-3


def test_dcn_pack_module_inference_uses_fused_operator():
    m = D.ModulatedDeformConvPack(64, 64, 3, stride=1, padding=1, dilation=1, deformable_groups=8,
                                  extra_offset_mask=True).to(DEV)
    torch.nn.init.normal_(m.conv_offset_mask.weight, std=0.05)
    torch.nn.init.normal_(m.conv_offset_mask.bias, std=1.0)
    x, feat = synth_normal((1, 64, 16, 24), 900).to(DEV), synth_normal((1, 64, 16, 24), 901).to(DEV)
    with torch.no_grad():
        y_fused = m([x, feat])
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False  # cuDNN's default TF32 convs are only ~1e-3 accurate
    try:
        with torch.enable_grad():  # autograd on -> unfused graph (torch conv + our DCN op)
            y_graph = m([x, feat]).detach()
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    assert rel_err(y_fused, y_graph) < 1e-4


# ---------------------------------------------------------------- backward (rvsr_mdcn_bwd)
BWD_CASES = [dict(B=2, C=16, H=11, W=13, Cout=12, dg=4), dict(B=1, C=64, H=12, W=20, Cout=64, dg=8),
             dict(B=2, C=8, H=12, W=9, Cout=6, dg=4, groups=2, stride=2), dict(B=1, C=16, H=10, W=10, Cout=16, dg=2, pad=2, dil=2),
             dict(B=1, C=8, H=8, W=8, Cout=70, dg=8), dict(B=1, C=32, H=9, W=9, Cout=8, dg=2)]


@pytest.mark.parametrize("cfg", BWD_CASES)
def test_mdcn_bwd_fp32_vs_oracle(cfg):
    """All five gradients through the autograd Function against oracle/dcn_oracle.c (double inside).
    Tolerance 2e-4 of max|ref| per gradient: fp32 atomics in arbitrary order (the reference's col2im
    scatter is order-nondeterministic too, deform_conv_cuda_kernel.cu:688)."""
    x, off, msk, w, b, (s, p, d, g, dg) = _case(**cfg)
    go = synth_normal(tuple(O.dcn_forward(x, off, msk, w, b, s, p, d, g, dg).shape), 77)
    ref = O.dcn_backward(x, off, msk, w, go, s, p, d, g, dg)
    leaves = [t.to(DEV).requires_grad_() for t in (x, off, msk, w, b)]
    y = D.modulated_deform_conv(*leaves, s, p, d, g, dg)
    y.backward(go.to(DEV))
    for t, r, name in zip(leaves, ref, ("input", "offset", "mask", "weight", "bias")):
        assert t.grad is not None and t.grad.shape == r.shape, name
        assert rel_err(t.grad.cpu(), r) < 2e-4, name


def test_mdcn_bwd_matches_reference_generated_golden():
    z = np.load(os.path.join(GOLDEN, "dcn_unit.npz"))
    B, C, H, W, Cout, dg = [int(v) for v in z["dims"]]
    x = synth_normal((B, C, H, W), 71); off = synth_normal((B, dg * 18, H, W), 72, std=4.0)
    msk = torch.sigmoid(synth_normal((B, dg * 9, H, W), 73).double()).float()
    w = synth_normal((Cout, C, 3, 3), 74, std=0.1); b = synth_normal((Cout,), 75); go = synth_normal((B, Cout, H, W), 76)
    leaves = [t.to(DEV).requires_grad_() for t in (x, off, msk, w, b)]
    D.modulated_deform_conv(*leaves, 1, 1, 1, 1, dg).backward(go.to(DEV))
    for t, key in zip(leaves, ("gx", "goff", "gmask", "gw", "gb")):
        assert rel_err(t.grad.cpu(), torch.from_numpy(z[key]).float()) < 2e-4, key


def test_mdcn_bwd_no_bias_and_fp16_inputs():
    x, off, msk, w, b, (s, p, d, g, dg) = _case(B=1, C=16, H=8, W=8, Cout=8, dg=4)
    h = lambda t: t.half().float()  # noqa: E731
    go = synth_normal((1, 8, 8, 8), 78)
    ref = O.dcn_backward(h(x), h(off), h(msk), h(w), h(go), s, p, d, g, dg, with_bias=False)
    leaves = [t.to(DEV).half().requires_grad_() for t in (x, off, msk, w)]
    y = D.modulated_deform_conv(*leaves, None, s, p, d, g, dg)
    y.backward(go.to(DEV).half())
    for t, r in zip(leaves, ref[:4]):
        assert t.grad.dtype == torch.float16 and rel_err(t.grad.float().cpu(), r) < 3e-3


def test_mdcn_bf16_forward_and_backward_vs_oracle():
    """RVSR_BF16 (BASELINE cfg5's dtype): bfloat16 tensors through the ABI, fp32 arithmetic inside.  Against the fp32 C oracle
    on the same bf16-rounded inputs; the only extra error is the final rounding of each result to bf16 (2^-9 relative)."""
    x, off, msk, w, b, (s, p, d, g, dg) = _case(B=2, C=16, H=12, W=10, Cout=16, dg=4)
    r = lambda t: t.bfloat16().float()  # noqa: E731
    go = synth_normal((2, 16, 12, 10), 79)
    ref = O.dcn_forward(r(x), r(off), r(msk), r(w), r(b), s, p, d, g, dg)
    gref = O.dcn_backward(r(x), r(off), r(msk), r(w), r(go), s, p, d, g, dg, with_bias=True)
    leaves = [t.to(DEV).bfloat16().requires_grad_() for t in (x, off, msk, w, b)]
    y = D.modulated_deform_conv(*leaves, s, p, d, g, dg)
    assert y.dtype == torch.bfloat16 and rel_err(y.float().cpu(), ref) < 6e-3
    y.backward(go.to(DEV).bfloat16())
    for t, gr, name in zip(leaves, gref, ("x", "offset", "mask", "weight", "bias")):
        assert t.grad.dtype == torch.bfloat16 and rel_err(t.grad.float().cpu(), gr) < 8e-3, name


@pytest.mark.parametrize("shape", [(2, 20, 36), (1, 64, 64), (3, 7, 33)])
def test_mdcn_bf16_backward_on_tensor_cores_vs_oracle(shape, monkeypatch):
    """EDVR's DCN shape class (64 -> 64, 3x3, 8 deformable groups) in bf16: rvsr_mdcn_bwd runs dcn_bwd_tc_kernel (grad_col and
    grad_weight contractions on tcgen05, bf16 operands, fp32 accumulate).  Against the fp32 C oracle on the same
    bf16-rounded inputs, and against the CUDA-core bf16 path (RVSR_DCN_BWD_TC=0 is read once per process, so the comparison
    is with the oracle only).  Ragged tiles (W = 36, 33; H = 20, 7), offsets that leave the image."""
    B, H, W = shape
    x, off, msk, w, b, (s, p, d, g, dg) = _case(B=B, C=64, H=H, W=W, Cout=64, dg=8, off_std=2.5, seed=700 + H)
    w = w * 0.3
    r = lambda t: t.bfloat16().float()  # noqa: E731
    go = synth_normal((B, 64, H, W), 779 + W) * 0.01           # gradient-sized values (bf16 keeps their range)
    gref = O.dcn_backward(r(x), r(off), r(msk), r(w), r(go), s, p, d, g, dg, with_bias=True)
    leaves = [t.to(DEV).bfloat16().requires_grad_() for t in (x, off, msk, w, b)]
    y = D.modulated_deform_conv(*leaves, s, p, d, g, dg)
    ef = rel_err(y.float().cpu(), O.dcn_forward(r(x), r(off), r(msk), r(w), r(b), s, p, d, g, dg))
    print("bf16 DCN forward on tcgen05 %s: rel err %.2e" % (shape, ef))
    assert y.dtype == torch.bfloat16 and ef < 8e-3          # forward: dcn_tc_kernel, fp16 operands inside, bf16 result
    y.backward(go.to(DEV).bfloat16())
    for t, gr, name in zip(leaves, gref, ("x", "offset", "mask", "weight", "bias")):
        e = rel_err(t.grad.float().cpu(), gr)
        print("bf16 tensor-core DCN backward %s: grad_%s rel err %.2e" % (shape, name, e))
        assert t.grad.dtype == torch.bfloat16 and e < 1e-2, name


def test_training_step_bf16_autocast():
    """cfg5 as BASELINE states it: the training step under torch.autocast(bfloat16) -- torch's convolutions in bf16, the DCN
    through RVSR_BF16.  Loss close to the fp32 step's, finite gradients on every parameter."""
    from helpers import load_case
    from realvsr_b200.archs import EDVR_arch as E
    c = load_case("edvr_tiny")
    net = E.EDVR(**c["kwargs"]).train()
    net.load_state_dict(c["sd"], strict=True)
    net = net.to(DEV)
    x = c["x"].to(DEV)
    gt = synth_normal(tuple(c["out"].shape), 5, std=0.3).to(DEV)
    l32 = float(torch.nn.functional.l1_loss(net(x), gt).detach())
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = torch.nn.functional.l1_loss(net(x).float(), gt)
    loss.backward()
    assert abs(float(loss.detach()) - l32) < 2e-2 * max(l32, 1e-3)
    assert all(p_.grad is not None and bool(torch.isfinite(p_.grad).all()) for p_ in net.parameters())


def test_bf16_training_gradients_agree_with_fp32_on_the_nf64_network():
    """The tensor-core DCN kernels inside a real network: EDVR nf = 64 (the shape class dcn_tc_kernel / dcn_bwd_tc_kernel
    cover) on a small crop, one training step under torch.autocast(bfloat16) against the same step in fp32 (CUDA-core DCN
    kernels, cuDNN TF32 off).  bf16 storage puts ~1e-2 of noise on every activation, so the check is directional: the
    gradient of every DCN-related parameter points the same way (cosine > 0.97) and has the same size (within 10 %)."""
    from helpers import load_case
    from realvsr_b200.archs import EDVR_arch as E
    c = load_case("edvr_nf64_crop")
    net = E.EDVR(**c["kwargs"]).train()
    net.load_state_dict(c["sd"], strict=True)
    net = net.to(DEV)
    x = torch.cat([c["x"], c["x"].flip(3)], 0).to(DEV)
    gt = synth_normal((2,) + tuple(c["out"].shape[1:]), 55, std=0.3).to(DEV) + 0.5
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        torch.nn.functional.l1_loss(net(x), gt).backward()
        g32 = {n: p_.grad.clone() for n, p_ in net.named_parameters()}
        net.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = torch.nn.functional.l1_loss(net(x).float(), gt)
        loss.backward()
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    worst = (1.0, "")
    for n, p_ in net.named_parameters():
        if "dcnpack" not in n and "offset_conv" not in n:
            continue
        a, b = p_.grad.float().flatten(), g32[n].float().flatten()
        cos = float(torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-30))
        ratio = float(a.norm() / b.norm().clamp_min(1e-30))
        if cos < worst[0]:
            worst = (cos, n)
        assert 0.9 < ratio < 1.1, (n, ratio)
    print("bf16 (tcgen05 DCN fwd + bwd) vs fp32 gradients: worst cosine %.4f (%s)" % worst)
    assert worst[0] > 0.97, worst


def test_training_step_through_module_path():
    """cfg5-shaped smoke (tiny): EDVR module path forward + L1 loss + backward through our DCN fwd/bwd;
    gradients reach every parameter and match a finite-difference probe on one DCN weight."""
    from helpers import load_case
    from realvsr_b200.archs import EDVR_arch as E
    c = load_case("edvr_tiny")
    net = E.EDVR(**c["kwargs"]).train()
    net.load_state_dict(c["sd"], strict=True)
    net = net.to(DEV)
    x = c["x"].to(DEV)
    gt = synth_normal(tuple(c["out"].shape), 5, std=0.3).to(DEV)
    loss = torch.nn.functional.l1_loss(net(x), gt)
    loss.backward()
    missing = [n for n, p_ in net.named_parameters() if p_.grad is None]
    assert not missing, missing
    assert all(bool(torch.isfinite(p_.grad).all()) for p_ in net.parameters())
    wparam = net.pcd_align.cas_dcnpack.weight
    idx = (3, 2, 1, 1)
    g = float(wparam.grad[idx])
    eps = 1e-2
    with torch.no_grad():
        wparam[idx] += eps
        lp = float(torch.nn.functional.l1_loss(net(x), gt))
        wparam[idx] -= 2 * eps
        lm = float(torch.nn.functional.l1_loss(net(x), gt))
        wparam[idx] += eps
    fd = (lp - lm) / (2 * eps)
    assert abs(fd - g) < 0.15 * max(abs(g), 1e-4) + 2e-5, (fd, g)


def test_whole_network_gradients_match_reference_golden():
    """BASELINE cfg5's step (forward + L1 loss + backward) on a tiny architecture: EVERY parameter's gradient against
    the golden produced by the reference's own EDVR_arch.py in float64 (tests/golden/make_golden.py, edvr_tiny_grads:
    torchvision's deform_conv2d supplies the DCN autograd there).  Module path here: torch convs + rvsr_mdcn_fwd /
    rvsr_mdcn_bwd.  Tolerance 1e-3 of each gradient tensor's max (fp32 vs float64, atomics-free summation order)."""
    import ast
    from helpers import edvr_state_shapes
    from realvsr_b200.archs import EDVR_arch as E
    from synth import synth_input, synth_state_dict
    z = np.load(os.path.join(GOLDEN, "edvr_tiny_grads.npz"))
    kw = ast.literal_eval(str(z["kwargs"]))
    net = E.EDVR(**kw).train()
    net.load_state_dict(synth_state_dict(edvr_state_shapes("EDVR", **kw), int(z["wseed"])), strict=True)
    net = net.to(DEV)
    x = synth_input(tuple(z["shape"]), int(z["xseed"])).to(DEV)
    gt = (synth_normal((2, 3, 96, 96), int(z["gtseed"]), std=0.3) + 0.5).to(DEV)
    # the plain convolutions of the module path are cuDNN: TF32 (torch's default for convolutions, 10-bit mantissa) would
    # put ~1e-2 on the small gradients of the offset branches -- strict fp32 for a 1e-3 comparison
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        loss = torch.nn.functional.l1_loss(net(x), gt)
        loss.backward()
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    assert abs(float(loss.detach()) - float(z["loss"])) < 1e-5
    worst = ("", 0.0)
    for name, p_ in net.named_parameters():
        e = rel_err(p_.grad.cpu(), torch.from_numpy(z["g:" + name]))
        if e > worst[1]:
            worst = (name, e)
    print("whole-network gradients: worst %s %.2e" % worst)
    assert worst[1] < 1e-3, worst

"""bf16 training path (realvsr_b200/train_c8.py) against torch autograd in fp32 on the same bf16-rounded operands.

Reference semantics: nn.Conv2d autograd (EDVR_arch.py:71-91, :229-253, arch_util.py:121-139), F.interpolate x2 bilinear
(:109-121), PixelShuffle(2) + lrelu (:313-314).  Tolerances: outputs and data gradients are stored in bf16 (8 mantissa
bits): max |err| <= 1e-2 of the tensor's max magnitude.  Weight / bias gradients are fp32 sums of bf16 products:
<= 3e-3 of the max magnitude.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _r(t):  # round to bf16, keep fp32
    return t.to(torch.bfloat16).float()


def _rel(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-20))


@pytest.fixture(autouse=True)
def _no_tf32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_layout_roundtrip_and_gradient():
    from realvsr_b200 import train_c8 as T
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(3, 19, 7, 11, device="cuda", generator=g, requires_grad=True)
    c = T.to_c8(x)
    assert tuple(c.shape) == (3, 3, 7, 11, 8) and c.dtype == torch.bfloat16
    assert float(c.detach()[:, 2, :, :, 3:].abs().max()) == 0.0  # channels 19..23 are zero padding
    y = T.from_c8(c, 19, torch.float32)
    assert torch.equal(y, _r(x.detach()))
    w = torch.randn_like(y)
    (y * w).sum().backward()
    assert torch.equal(x.grad, _r(w))


@pytest.mark.parametrize("case", ["lrelu", "relu", "none_residual", "cat2", "shuffle", "shuffle_none", "k1_lrelu", "k1_cat5", "cout16"])
@pytest.mark.parametrize("shape", [(2, 16, 32), (3, 18, 40), (1, 64, 64)])
def test_conv_forward_backward(case, shape):
    from realvsr_b200 import train_c8 as T
    N, H, W = shape
    g = torch.Generator(device="cuda").manual_seed(len(case) * 131 + N * 17 + H)
    nsrc = {"cat2": 2, "k1_cat5": 5}.get(case, 1)
    Cout = 256 if case.startswith("shuffle") else (16 if case == "cout16" else 64)
    ks = 1 if case.startswith("k1") else 3
    act = {"lrelu": "lrelu", "relu": "relu", "none_residual": None, "cat2": "lrelu", "shuffle": "lrelu", "shuffle_none": None,
           "k1_lrelu": "lrelu", "k1_cat5": "lrelu", "cout16": None}[case]
    xs = [_r(torch.randn(N, 64, H, W, device="cuda", generator=g)).requires_grad_() for _ in range(nsrc)]
    w = _r(torch.randn(Cout, 64 * nsrc, ks, ks, device="cuda", generator=g) * 0.05).requires_grad_()  # a bf16-exact leaf: no cast in the graph
    b = (torch.randn(Cout, device="cuda", generator=g) * 0.1).requires_grad_()
    res = _r(torch.randn(N, 64, H, W, device="cuda", generator=g)).requires_grad_() if case == "none_residual" else None
    # product
    xs2 = [x.detach().clone().requires_grad_() for x in xs]
    w2, b2 = w.detach().clone().requires_grad_(), b.detach().clone().requires_grad_()
    res2 = res.detach().clone().requires_grad_() if res is not None else None
    y = T.conv([T.to_c8(x) for x in xs2], w2, b2, act=act, residual=None if res2 is None else T.to_c8(res2),
               shuffle=case.startswith("shuffle"))
    Cy = Cout // 4 if case.startswith("shuffle") else Cout
    y_nchw = T.from_c8(y, Cy, torch.float32)
    # reference: fp32 math on the same bf16-exact operands.  The branch of the activation is taken from the PRODUCT's output
    # sign: two fp32 summation orders disagree on it for the ~1e-6 of outputs whose pre-activation is ~1e-7 (forward
    # difference 1e-7, but a different gradient at that element), which is no property of either implementation.
    pre = F.conv2d(torch.cat(xs, 1), w, b, padding=ks // 2)
    pos = (F.pixel_unshuffle(y_nchw.detach(), 2) if case.startswith("shuffle") else y_nchw.detach()) > 0
    assert act is None or int((pos != (pre.detach() > 0)).sum()) <= 4
    if act == "lrelu":
        y_ref = pre * torch.where(pos, 1.0, 0.1)
    elif act == "relu":
        y_ref = pre * pos
    else:
        y_ref = pre
    if res is not None:
        y_ref = y_ref + res
    if case.startswith("shuffle"):
        y_ref = F.pixel_shuffle(y_ref, 2)
    gy = _r(torch.randn(y_ref.shape, device="cuda", generator=g))
    leaves = xs + [w, b] + ([res] if res is not None else [])
    ref_grads = torch.autograd.grad(y_ref, leaves, gy)
    assert y_nchw.shape == y_ref.shape
    assert _rel(y_nchw, y_ref.detach()) < 1e-2
    leaves2 = xs2 + [w2, b2] + ([res2] if res2 is not None else [])
    grads = torch.autograd.grad(y_nchw, leaves2, gy)
    names = ["dx%d" % i for i in range(nsrc)] + ["dw", "db"] + (["dres"] if res is not None else [])
    for name, a, r in zip(names, grads, ref_grads):
        tol = 3e-3 if name in ("dw", "db") else 1e-2
        assert a.shape == r.shape, name
        assert _rel(a, r) < tol, (name, _rel(a, r))


@pytest.mark.parametrize("shape", [(2, 8, 5, 7), (3, 16, 16, 16)])
def test_upsample2x(shape):
    from realvsr_b200 import train_c8 as T
    N, C, H, W = shape
    g = torch.Generator(device="cuda").manual_seed(3)
    x = _r(torch.randn(N, C, H, W, device="cuda", generator=g)).requires_grad_()
    y_ref = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False) * 2
    gy = _r(torch.randn(y_ref.shape, device="cuda", generator=g))
    (gx_ref,) = torch.autograd.grad(y_ref, [x], gy)
    x2 = x.detach().clone().requires_grad_()
    y = T.from_c8(T.upsample2x(T.to_c8(x2), 2.0), C, torch.float32)
    assert _rel(y, y_ref) < 1e-2
    (gx,) = torch.autograd.grad(y, [x2], gy)
    assert _rel(gx, gx_ref) < 1e-2


def test_network_gradients_c8_path_vs_fp32_and_vs_autocast():
    """EDVR nf = 64 on a small crop, one training step three ways: fp32 module path (cuDNN TF32 off; its gradients are pinned to
    the float64 oracle by tests/test_gpu_dcn.py), torch.autocast(bfloat16) on the module path (cuDNN convolutions), and the
    train_c8 path (this library's convolution kernels).  bf16 storage puts ~1e-2 of noise on every activation, so the check is
    directional, for EVERY parameter: cosine with the fp32 gradient > 0.95 and not worse than autocast's by more than 0.03,
    norm within 10 %."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from helpers import load_case
    from synth import synth_normal
    from realvsr_b200.archs import EDVR_arch as E
    c = load_case("edvr_nf64_crop")
    net = E.EDVR(**c["kwargs"]).train()
    net.load_state_dict(c["sd"], strict=True)
    net = net.to("cuda")
    x = torch.cat([c["x"], c["x"].flip(3)], 0).to("cuda")
    gt = synth_normal((2,) + tuple(c["out"].shape[1:]), 55, std=0.3).to("cuda") + 0.5

    def grads(path, amp):
        net.exec_path = path
        net.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
            loss = F.l1_loss(net(x).float(), gt)
        loss.backward()
        return float(loss.detach()), {n: p.grad.detach().float().clone() for n, p in net.named_parameters()}

    l32, g32 = grads("module", False)
    lac, gac = grads("module", True)
    lc8, gc8 = grads("train_c8", False)
    assert abs(lc8 - l32) < 2e-2 * l32 and abs(lc8 - lac) < 2e-2 * l32
    worst = (1.0, "", 1.0)
    for n in g32:
        a, b, r = gc8[n].flatten(), gac[n].flatten(), g32[n].flatten()
        cos = lambda u, v: float(torch.dot(u, v) / (u.norm() * v.norm()).clamp_min(1e-30))  # noqa: E731
        c8, ac = cos(a, r), cos(b, r)
        if c8 < worst[0]:
            worst = (c8, n, ac)
        assert c8 > 0.95 and c8 > ac - 0.03, (n, c8, ac)
        assert 0.9 < float(a.norm() / r.norm().clamp_min(1e-30)) < 1.1, n
    print("train_c8 vs fp32 gradients: worst cosine %.4f (%s; torch autocast there: %.4f)" % worst)


def test_layout_with_spare_channel_blocks():
    """from_c8 of the first C channels of a wider tensor (the 216 offset / mask channels of a 256-channel convolution output)
    and its gradient (zeros in the dropped blocks), batch > 1."""
    from realvsr_b200 import train_c8 as T
    g = torch.Generator(device="cuda").manual_seed(5)
    x = _r(torch.randn(3, 32, 6, 10, device="cuda", generator=g))
    c = T.to_c8(x).requires_grad_()
    y = T.from_c8(c, 20, torch.float32)
    assert torch.equal(y, x[:, :20])
    w = _r(torch.randn_like(y))
    (y * w).sum().backward()
    back = T.from_c8(c.grad, 32, torch.float32)
    assert torch.equal(back[:, :20], w) and float(back[:, 20:].abs().max()) == 0.0


@pytest.mark.parametrize("act", [None, "lrelu"])
def test_dcn_pack_c8_against_the_nchw_operator(act):
    """train_c8.dcn_pack (C8 tensors, sigmoid / chunk inside) against ModulatedDeformConvFunction on NCHW bf16 tensors -- the
    operator tests/test_gpu_dcn.py pins to the C oracle and the float64 gradients -- forward and all five gradients."""
    from realvsr_b200 import train_c8 as T
    from realvsr_b200.archs.dcn.deform_conv import modulated_deform_conv
    g = torch.Generator(device="cuda").manual_seed(11)
    N, H, W = 3, 20, 40
    x = _r(torch.randn(N, 64, H, W, device="cuda", generator=g)).requires_grad_()
    om = torch.zeros(N, 256, H, W, device="cuda")
    om[:, :144] = torch.randn(N, 144, H, W, device="cuda", generator=g) * 2.0
    om[:, 144:216] = torch.randn(N, 72, H, W, device="cuda", generator=g)
    om = _r(om).requires_grad_()
    w = _r(torch.randn(64, 64, 3, 3, device="cuda", generator=g) * 0.05).requires_grad_()
    b = _r(torch.randn(64, device="cuda", generator=g) * 0.1).requires_grad_()  # bf16-exact: the NCHW operator rounds its bias to bf16
    gy = _r(torch.randn(N, 64, H, W, device="cuda", generator=g))
    # reference: the NCHW operator in bf16 (tensor-core kernels, fp32 master weights as under autocast)
    y_ref = modulated_deform_conv(x.bfloat16(), om[:, :144].bfloat16(), torch.sigmoid(om[:, 144:216]).bfloat16(), w, b, 1, 1, 1, 1, 8).float()
    x2, om2, w2, b2 = [t.detach().clone().requires_grad_() for t in (x, om, w, b)]
    y = T.from_c8(T.dcn_pack(T.to_c8(x2), T.to_c8(om2), w2, b2, act), 64, torch.float32)
    if act:
        # the NCHW operator receives the mask rounded to bf16, the C8 operator computes the sigmoid itself: outputs differ by
        # ~1e-3, which flips the LeakyReLU branch of the ~1e-3 of outputs that close to zero.  Same branch for both (see
        # test_conv_forward_backward).
        pos = y.detach() > 0
        assert float((pos != (y_ref.detach() > 0)).float().mean()) < 5e-3
        y_ref = y_ref * torch.where(pos, 1.0, 0.1)
    ref = torch.autograd.grad(y_ref, [x, om, w, b], gy)
    assert _rel(y, y_ref.detach()) < 1e-2
    got = torch.autograd.grad(y, [x2, om2, w2, b2], gy)
    for name, a, r in zip(("dx", "dom", "dw", "db"), got, ref):
        assert a.shape == r.shape
        assert _rel(a, r) < 2e-2, (name, _rel(a, r))
    assert float(got[1][:, 216:].abs().max()) == 0.0


def test_conv_first_c8():
    """3 -> 64 channels from an NCHW fp32 image: forward, weight and bias gradients against torch autograd."""
    from realvsr_b200 import train_c8 as T
    g = torch.Generator(device="cuda").manual_seed(21)
    N, H, W = 5, 20, 40
    x = _r(torch.rand(N, 3, H, W, device="cuda", generator=g))
    w = _r(torch.randn(64, 3, 3, 3, device="cuda", generator=g) * 0.2).requires_grad_()
    b = (torch.randn(64, device="cuda", generator=g) * 0.1).requires_grad_()
    w2, b2 = w.detach().clone().requires_grad_(), b.detach().clone().requires_grad_()
    y = T.from_c8(T.conv_first(x, w2, b2, "lrelu"), 64, torch.float32)
    pre = F.conv2d(x, w, b, padding=1)
    y_ref = pre * torch.where(y.detach() > 0, 1.0, 0.1)
    assert _rel(y, y_ref.detach()) < 1e-2
    gy = _r(torch.randn(y.shape, device="cuda", generator=g))
    ref = torch.autograd.grad(y_ref, [w, b], gy)
    got = torch.autograd.grad(y, [w2, b2], gy)
    for name, a, r in zip(("dw", "db"), got, ref):
        assert a.shape == r.shape and _rel(a, r) < 3e-3, (name, _rel(a, r))


def test_tsa_temporal_c8():
    """Temporal attention (EDVR_arch.py:170-181) forward and the three gradients against torch autograd in fp32."""
    from realvsr_b200 import train_c8 as T
    g = torch.Generator(device="cuda").manual_seed(31)
    B, N, H, W = 2, 5, 12, 20
    al = _r(torch.randn(B * N, 64, H, W, device="cuda", generator=g)).requires_grad_()
    em = _r(torch.randn(B * N, 64, H, W, device="cuda", generator=g) * 0.3).requires_grad_()
    er = _r(torch.randn(B, 64, H, W, device="cuda", generator=g) * 0.3).requires_grad_()
    prob = torch.sigmoid((em.view(B, N, 64, H, W) * er.unsqueeze(1)).sum(2, keepdim=True))
    ref = (al.view(B, N, 64, H, W) * prob)
    gy = _r(torch.randn(ref.shape, device="cuda", generator=g))
    gy[:, 3] = 0
    gref = torch.autograd.grad(ref, [al, em, er], gy)
    al2, em2, er2 = [t.detach().clone().requires_grad_() for t in (al, em, er)]
    outs = T.tsa_temporal(T.to_c8(al2), T.to_c8(em2), T.to_c8(er2), N)
    y = torch.stack([T.from_c8(o, 64, torch.float32) for o in outs], 1)
    assert _rel(y, ref.detach()) < 1e-2
    got = torch.autograd.grad(y, [al2, em2, er2], gy)
    for name, a, r in zip(("d_aligned", "d_emb", "d_emb_ref"), got, gref):
        assert a.shape == r.shape and _rel(a, r) < 1e-2, (name, _rel(a, r))


@pytest.mark.parametrize("cls,kw", [("EDVR_NoUp", dict(nf=64, nframes=3, groups=8, front_RBs=2, back_RBs=2, w_TSA=False)),
                                    ("EDVR", dict(nf=64, nframes=7, groups=8, front_RBs=1, back_RBs=1, w_TSA=True)),
                                    ("EDVR", dict(nf=64, nframes=3, groups=8, front_RBs=1, back_RBs=1, w_TSA=False))])
def test_other_architectures_on_the_c8_path(cls, kw):
    """The shipped RealVSR configuration (EDVR_NoUp, 3 frames, no TSA: train_EDVR_woTSA_RealVSR_YCbCr_Split.yml:38-50) and other
    frame counts: train_c8 gradients against the fp32 module path, every parameter (same criterion as the cfg5 network test).
    Sizes that are not tile multiples (36 x 44)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import edvr_state_shapes
    from synth import synth_input, synth_state_dict
    from realvsr_b200.archs import EDVR_arch as E
    net = getattr(E, cls)(**kw).train()
    net.load_state_dict(synth_state_dict(edvr_state_shapes(cls, **kw), 3), strict=True)
    net = net.to("cuda")
    x = synth_input((2, kw["nframes"], 3, 36, 44), 4).to("cuda")
    s = 4 if cls == "EDVR" else 1
    gt = synth_input((2, 3, 36 * s, 44 * s), 5).to("cuda")

    def grads(path):
        net.exec_path = path
        net.zero_grad(set_to_none=True)
        loss = F.l1_loss(net(x).float(), gt)
        loss.backward()
        return float(loss.detach()), {n: p.grad.detach().float().clone() for n, p in net.named_parameters()}

    l32, g32 = grads("module")
    lc8, gc8 = grads("train_c8")
    assert abs(lc8 - l32) < 2e-2 * l32
    for n in g32:
        a, r = gc8[n].flatten(), g32[n].flatten()
        cos = float(torch.dot(a, r) / (a.norm() * r.norm()).clamp_min(1e-30))
        assert cos > 0.95, (n, cos)
        assert 0.9 < float(a.norm() / r.norm().clamp_min(1e-30)) < 1.1, n


@pytest.mark.parametrize("kind", ["resblock", "pcd_pair"])
def test_conv_pair_c8(kind):
    """The fused two-convolution Function (activation gradient in the data-gradient epilogue, skip gradient as its residual)
    against the same two layers as separate train_c8.conv calls -- which test_conv_forward_backward pins to torch autograd."""
    from realvsr_b200 import train_c8 as T
    g = torch.Generator(device="cuda").manual_seed(41)
    N, H, W = 3, 20, 44
    nsrc = 1 if kind == "resblock" else 2
    xs = [_r(torch.randn(N, 64, H, W, device="cuda", generator=g)) for _ in range(nsrc)]
    w1 = _r(torch.randn(64, 64 * nsrc, 3, 3, device="cuda", generator=g) * 0.05)
    w2 = _r(torch.randn(64, 64, 3, 3, device="cuda", generator=g) * 0.05)
    b1, b2 = torch.randn(64, device="cuda", generator=g) * 0.1, torch.randn(64, device="cuda", generator=g) * 0.1
    gy = _r(torch.randn(N, 64, H, W, device="cuda", generator=g))

    def run(fused):
        leaves = [t.detach().clone().requires_grad_() for t in xs + [w1, b1, w2, b2]]
        cx, (lw1, lb1, lw2, lb2) = [T.to_c8(t) for t in leaves[:nsrc]], leaves[nsrc:]
        if kind == "resblock":
            y = (T.conv_pair(cx, lw1, lb1, "relu", lw2, lb2, None, skip=True) if fused
                 else T.conv(T.conv(cx, lw1, lb1, act="relu"), lw2, lb2, residual=cx[0]))
        else:
            y = (T.conv_pair(cx, lw1, lb1, "lrelu", lw2, lb2, "lrelu") if fused
                 else T.conv(T.conv(cx, lw1, lb1, act="lrelu"), lw2, lb2, act="lrelu"))
        y = T.from_c8(y, 64, torch.float32)
        return y.detach(), torch.autograd.grad(y, leaves, gy)

    y0, g0 = run(False)
    y1, g1 = run(True)
    assert torch.equal(y0, y1)                       # identical forward kernels
    for a, r in zip(g1, g0):
        assert a.shape == r.shape and _rel(a, r) < 1e-2   # one bf16 rounding less on the fused path


def test_pool_and_tsa_final_c8():
    """MaxPool2d(3,2,1) + AvgPool2d(3,2,1) (with TIES: bf16 feature maps have them, the gradient must go to the first maximum
    like torch's) and fea * sigmoid(att) * 2 + att_add, forward and gradients, against torch."""
    from realvsr_b200 import train_c8 as T
    g = torch.Generator(device="cuda").manual_seed(51)
    N, C, H, W = 3, 16, 11, 14
    x = (torch.randint(0, 6, (N, C, H, W), device="cuda", generator=g).float() * 0.25).requires_grad_()   # few distinct values -> ties
    mx_ref, av_ref = F.max_pool2d(x, 3, 2, 1), F.avg_pool2d(x, 3, 2, 1)
    g1, g2 = _r(torch.randn(mx_ref.shape, device="cuda", generator=g)), _r(torch.randn(mx_ref.shape, device="cuda", generator=g))
    (gx_ref,) = torch.autograd.grad([mx_ref, av_ref], [x], [g1, g2])
    x2 = x.detach().clone().requires_grad_()
    mx, av = T.pool_maxavg(T.to_c8(x2))
    mx, av = T.from_c8(mx, C, torch.float32), T.from_c8(av, C, torch.float32)
    assert torch.equal(mx, mx_ref.detach()) and _rel(av, av_ref.detach()) < 1e-2
    (gx,) = torch.autograd.grad([mx, av], [x2], [g1, g2])
    assert _rel(gx, gx_ref) < 1e-2
    fea, att, add = [_r(torch.randn(N, C, H, W, device="cuda", generator=g)).requires_grad_() for _ in range(3)]
    ref = fea * torch.sigmoid(att) * 2 + add
    gy = _r(torch.randn(ref.shape, device="cuda", generator=g))
    gref = torch.autograd.grad(ref, [fea, att, add], gy)
    leaves = [t.detach().clone().requires_grad_() for t in (fea, att, add)]
    out = T.from_c8(T.tsa_final(*[T.to_c8(t) for t in leaves]), C, torch.float32)
    assert _rel(out, ref.detach()) < 1e-2
    for a, r in zip(torch.autograd.grad(out, leaves, gy), gref):
        assert _rel(a, r) < 1e-2


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_dataparallel_training_step_on_the_c8_path():
    """nn.DataParallel (the wrapper the reference's training models use, VideoSR_AllPair_model_YCbCr_Split.py:36) around the
    module, bf16 autocast step: every replica runs the train_c8 path on its device with the broadcast weights, the gradients
    arrive on the wrapped module's parameters and agree with the single-device step on the whole batch."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import load_case
    from synth import synth_normal
    from realvsr_b200.archs import EDVR_arch as E
    c = load_case("edvr_nf64_crop")
    net = E.EDVR(**c["kwargs"]).train()
    net.load_state_dict(c["sd"], strict=True)
    net = net.to("cuda:0")
    x = torch.cat([c["x"], c["x"].flip(3), c["x"].flip(4), c["x"].flip(3).flip(4)], 0).to("cuda:0")
    gt = synth_normal((4,) + tuple(c["out"].shape[1:]), 56, std=0.3).to("cuda:0") + 0.5

    # torch's parallel_apply re-enters autocast in the replica threads WITHOUT the dtype (they see the thread default, float16),
    # so "bf16 autocast around a DataParallel model" is not bf16 inside the replicas -- the explicit switch is the way here
    net.exec_path = "train_c8"

    def grads(model):
        net.zero_grad(set_to_none=True)
        loss = F.l1_loss(model(x).float(), gt)
        loss.backward()
        return float(loss.detach()), [p.grad.detach().clone() for p in net.parameters()]

    l1, g1 = grads(net)
    l2, g2 = grads(torch.nn.DataParallel(net, device_ids=[0, 1]))
    diffs = [float((a - b).abs().max() / b.abs().max().clamp_min(1e-12)) for a, b in zip(g2, g1)]
    worst = max(diffs)
    if worst >= 2e-2:
        for (n_, _), d_ in zip(net.named_parameters(), diffs):
            if d_ > 5e-3:
                print("   %-50s %.3e" % (n_, d_))
    print("DataParallel vs single device: loss %.6f vs %.6f, worst relative gradient difference %.2e" % (l2, l1, worst))
    assert abs(l1 - l2) < 1e-3 * l1
    # two devices sum a weight gradient as (batch half 0) + (batch half 1), one device in tile order: fp32 sums of the same bf16
    # products in another order, on top of the atomics inside dcn_bwd_tc_kernel
    assert worst < 2e-2


def test_graphed_step_follows_the_optimizer():
    """train_c8.GraphedStep: the captured step re-packs the weights INSIDE the graph, so every replay sees the optimizer's last
    update.  Three SGD steps replayed from the graph against the same three steps run eagerly: same losses, same final weights."""
    import copy
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import load_case
    from synth import synth_normal
    from realvsr_b200 import train_c8 as T
    from realvsr_b200.archs import EDVR_arch as E
    c = load_case("edvr_nf64_crop")
    net = E.EDVR(**c["kwargs"]).train()
    net.load_state_dict(c["sd"], strict=True)
    net = net.to("cuda")
    ref = copy.deepcopy(net)
    x = torch.cat([c["x"], c["x"].flip(3)], 0).to("cuda")
    gt = synth_normal((2,) + tuple(c["out"].shape[1:]), 57, std=0.3).to("cuda") + 0.5
    lr = 2e-3
    losses_eager = []
    opt = torch.optim.SGD(ref.parameters(), lr=lr)
    for _ in range(3):
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = F.l1_loss(ref(x).float(), gt)
        loss.backward()
        opt.step()
        losses_eager.append(float(loss.detach()))
    step = T.GraphedStep(net, F.l1_loss, x, gt)          # warm-up runs forward / backward only: the weights are untouched
    opt = torch.optim.SGD(net.parameters(), lr=lr)
    losses_graph = []
    for _ in range(3):
        losses_graph.append(float(step(x, gt)))
        opt.step()
    assert losses_eager[2] < losses_eager[0]             # the steps do something
    for a, b in zip(losses_graph, losses_eager):
        assert abs(a - b) < 2e-3 * b, (losses_graph, losses_eager)
    for (n, p), q in zip(net.named_parameters(), ref.parameters()):
        assert float((p.detach() - q.detach()).abs().max()) <= 2e-2 * lr * 50 + 1e-3 * float(q.detach().abs().max()), n


@pytest.mark.parametrize("overlap", [False, True])
def test_graphed_step_gradients_equal_the_eager_step(overlap):
    """Every parameter's gradient after one replay of GraphedStep equals the eager step's -- with the weight-gradient launches on
    the main stream and as a parallel branch of the graph (train_c8._Defer: gradients delivered to .grad behind autograd's back;
    derived weights such as the permuted offset / mask convolution must still take autograd's route).  The convolution weight
    gradients are summed in a fixed order (bit-identical); the DCN backward's atomics reorder sums, hence the small tolerance."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import load_case
    from synth import synth_normal
    from realvsr_b200 import train_c8 as T
    from realvsr_b200.archs import EDVR_arch as E
    c = load_case("edvr_nf64_crop")
    net = E.EDVR(**c["kwargs"]).train()
    net.load_state_dict(c["sd"], strict=True)
    net = net.to("cuda")
    x = torch.cat([c["x"], c["x"].flip(3)], 0).to("cuda")
    gt = synth_normal((2,) + tuple(c["out"].shape[1:]), 58, std=0.3).to("cuda") + 0.5
    net.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = F.l1_loss(net(x).float(), gt)
    loss.backward()
    eager = {n: p.grad.detach().clone() for n, p in net.named_parameters()}
    loss_eager = float(loss.detach())
    del loss   # a live autograd graph of the same parameters pins their AccumulateGrad nodes to the stream it ran on
    step = T.GraphedStep(net, F.l1_loss, x, gt, overlap_wgrad=overlap, pack_ahead=overlap)
    for _ in range(2):   # a replay overwrites, never accumulates
        lg = step(x, gt)
    assert abs(float(lg) - loss_eager) < 1e-4 * loss_eager
    assert not T._Defer.active and not T._Defer.packing and not T._Defer.pending and not T._Defer.keep
    for n, p in net.named_parameters():
        assert p.grad is not None, n
        a, b = p.grad.float(), eager[n].float()
        assert float((a - b).abs().max()) <= 2e-2 * float(b.abs().max()) + 1e-7, n
    # downstream of the last DCN nothing is summed by atomics: those gradients are bit-identical from replay to replay and to
    # the eager step (a side-stream launch reading a buffer the allocator had already recycled would show up here)
    first = {n: p.grad.clone() for n, p in net.named_parameters()}
    for _ in range(3):
        step(x, gt)
    for n, p in net.named_parameters():
        if n.startswith(("recon_trunk", "upconv", "HRconv", "conv_last")):
            assert torch.equal(p.grad, first[n]) and torch.equal(p.grad, eager[n]), n


def test_ft_tsa_only_freezing_on_the_c8_path():
    """The reference's `ft_tsa_only` option freezes every parameter whose name lacks 'tsa_fusion'
    (VideoSR_AllPair_model_YCbCr_Split.py:94-99).  On the train_c8 path frozen layers must skip their weight gradients but
    still pass the data gradient back to the TSA module; the TSA gradients equal those of the unfrozen step."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import load_case
    from synth import synth_normal
    from realvsr_b200.archs import EDVR_arch as E
    c = load_case("edvr_nf64_crop")
    net = E.EDVR(**c["kwargs"]).train()
    net.load_state_dict(c["sd"], strict=True)
    net = net.to("cuda")
    net.exec_path = "train_c8"
    x = c["x"].to("cuda")
    gt = synth_normal(tuple(c["out"].shape), 58, std=0.3).to("cuda") + 0.5

    def step():
        net.zero_grad(set_to_none=True)
        F.l1_loss(net(x).float(), gt).backward()
        return {n: (None if p.grad is None else p.grad.detach().clone()) for n, p in net.named_parameters()}

    full = step()
    for n, p in net.named_parameters():
        p.requires_grad = "tsa_fusion" in n
    part = step()
    for n in full:
        if "tsa_fusion" in n:
            # identical kernels on identical data; only the fp32 atomics inside dcn_bwd_tc_kernel (not on this path: the DCN
            # sits before the TSA module) could differ -- so bit-identical
            assert torch.equal(part[n], full[n]), n
        else:
            assert part[n] is None, n
    # the same frozen network replayed from a graph with its parallel branches (train_c8._Defer): frozen layers launch no weight
    # gradient on the side stream and get no .grad; the derived offset / mask weights of frozen packs are packed on the main stream
    from realvsr_b200 import train_c8 as T
    net.zero_grad(set_to_none=True)
    gs = T.GraphedStep(net, F.l1_loss, x, gt, amp_dtype=None)
    gs(x, gt)
    for n, p in net.named_parameters():
        if "tsa_fusion" in n:
            assert torch.equal(p.grad, full[n]), n
        else:
            assert p.grad is None, n


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.float16, 2e-3), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("shape", [(2, 5, 7, 9), (3, 8, 16, 16), (1, 3, 1, 5)])
def test_upsample2x_nchw_module_path(dtype, tol, shape):
    """ops.upsample2x (the module path's F.interpolate x2 replacement on CUDA NCHW tensors) forward and adjoint against torch."""
    from realvsr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(61)
    x = torch.randn(shape, device="cuda", generator=g).to(dtype).requires_grad_()
    ref = F.interpolate(x.float(), scale_factor=2, mode="bilinear", align_corners=False)
    gy = torch.randn(ref.shape, device="cuda", generator=g).to(dtype)
    (gx_ref,) = torch.autograd.grad(ref, [x], gy.float())
    x2 = x.detach().clone().requires_grad_()
    y = ops.upsample2x(x2)
    assert y.dtype == dtype and y.shape == ref.shape
    assert _rel(y, ref.detach()) < tol
    (gx,) = torch.autograd.grad(y, [x2], gy)
    assert _rel(gx, gx_ref) < tol

"""CPU tests of the host-side mirror of the reference API (realvsr_b200/archs): state_dict
contract, constructor surface, and the module-path graph wiring.  The DCN operator is
CUDA-only (as in the reference), so here -- and only here, in a test -- its call is routed
to the CPU oracle to check the graph around it against the reference-generated goldens."""
import pytest
import torch

from helpers import EDVR_CASES, edvr_state_shapes, load_case, rel_err
from oracle import edvr_oracle as O
from realvsr_b200.archs import EDVR_arch as E
from realvsr_b200.archs import arch_util
import sys

import realvsr_b200.archs.dcn.deform_conv  # noqa: F401  (the package also exports a *function* of this name)
D = sys.modules['realvsr_b200.archs.dcn.deform_conv']


@pytest.mark.parametrize("cls,kw", [
    ("EDVR", {}), ("EDVR_NoUp", dict(nframes=3, w_TSA=False)), ("EDVR", dict(nf=16, predeblur=True, HR_in=True)),
    ("EDVR", dict(nf=8, groups=2, nframes=7, front_RBs=1, back_RBs=1, center=1)),
])
def test_state_dict_contract(cls, kw):
    net = getattr(E, cls)(**kw)
    sd = net.state_dict()
    exp = edvr_state_shapes(cls, **kw)
    assert list(sd.keys()) == list(exp.keys())
    assert all(tuple(sd[k].shape) == tuple(exp[k]) for k in sd)
    if not kw:
        assert len(sd) == 144 and sum(v.numel() for v in sd.values()) == 3300131
    assert any('tsa_fusion' in k for k, _ in net.named_parameters())  # ft_tsa_only relies on this substring


def test_dcn_pack_init_and_cpu_behaviour():
    m = D.ModulatedDeformConvPack(16, 16, 3, stride=1, padding=1, dilation=1, deformable_groups=4,
                                  extra_offset_mask=True)
    assert m.conv_offset_mask.weight.shape == (4 * 27, 16, 3, 3)
    assert float(m.conv_offset_mask.weight.abs().sum()) == 0 and float(m.conv_offset_mask.bias.abs().sum()) == 0
    assert float(m.bias.abs().sum()) == 0 and float(m.weight.abs().max()) <= 1 / (16 * 9) ** 0.5
    with pytest.raises(NotImplementedError):  # CUDA only, like the reference (deform_conv.py:109-110)
        m([torch.zeros(1, 16, 8, 8), torch.zeros(1, 16, 8, 8)])
    with pytest.raises(NotImplementedError):
        D.DeformConv(4, 4, 3)


def test_resblock_init_scale():
    torch.manual_seed(0)
    rb = arch_util.ResidualBlock_noBN(64)
    std = float(rb.conv1.weight.std())
    assert abs(std - 0.1 * (2.0 / (64 * 9)) ** 0.5) < 2e-4 and float(rb.conv1.bias.abs().sum()) == 0
    assert len(arch_util.make_layer(arch_util.ResidualBlock_noBN, 3, nf=8)) == 3


@pytest.mark.parametrize("name", ["edvr_tiny", "edvr_tiny_b2_g2", "edvr_noup_3f", "edvr_predeblur", "edvr_hr_in",
                                  "edvr_predeblur_hr_in"])
def test_module_path_graph_matches_reference_golden(name, monkeypatch):
    c = load_case(name)
    net = getattr(E, c["cls"])(**c["kwargs"]).eval()
    net.load_state_dict(c["sd"], strict=True)

    def cpu_dcn(x, offset, mask, weight, bias, stride, padding, dilation, groups, dg):
        return O.dcn_forward(x, offset.contiguous(), mask.contiguous(), weight, bias, stride, padding, dilation,
                             groups, dg)

    monkeypatch.setattr(D, "modulated_deform_conv", cpu_dcn)
    with torch.no_grad():
        y = net._forward_modules(c["x"])
    assert rel_err(y, c["out"]) < 2e-5


def test_engine_path_refuses_cpu_tensors():
    net = E.EDVR(nf=8, groups=8, front_RBs=1, back_RBs=1).eval()
    net.exec_path = "engine"
    with pytest.raises(RuntimeError):
        with torch.no_grad():
            net(torch.zeros(1, 5, 3, 16, 16))


def test_tiled_forward_stitches_exactly_for_a_local_model():
    """video.tiled_forward on the CPU with a stand-in model whose receptive field is smaller than the halo (x4 bilinear
    of the centre frame + a 3x3 box filter): the stitched result must equal the whole-frame result everywhere, for tile
    grids that divide the frame and for ragged last tiles."""
    import torch
    import torch.nn.functional as F
    from realvsr_b200 import video

    def model(x):                                   # [B, N, C, H, W] -> [B, C, 4H, 4W]
        c = x[:, x.shape[1] // 2]
        c = F.avg_pool2d(F.pad(c, (1, 1, 1, 1), mode="replicate"), 3, 1)
        return F.interpolate(c, scale_factor=4, mode="bilinear", align_corners=False)

    g = torch.Generator().manual_seed(0)
    x = torch.rand(2, 3, 3, 40, 56, generator=g)
    full = model(x)
    for tile in ((40, 56), (20, 28), (16, 24), (12, 20)):
        out = video.tiled_forward(model, x, tile=tile, halo=8)
        assert out.shape == full.shape
        # replicate padding at the true image border is the only place a tile can differ: tiles see the same border
        assert torch.allclose(out, full, atol=1e-6), tile


def test_bench_reference_arm_prints_exactly_one_json_line():
    """The driver parses ONE JSON line from bench.py's stdout; native libraries may print to fd 1 (NCCL banner), so
    bench.py keeps a private copy of stdout for the result.  The reference arm runs on the CPU (the oracle port)."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


# ---------------------------------------------------------------- engine-cache bookkeeping (no GPU needed)
def _fake_replica(net):
    """What torch.nn.parallel.replicate() produces (replicate.py:167-197): no _parameters, the broadcast copies as
    plain tensor attributes + _former_parameters, sub-modules replaced by their replicas, __dict__ shallow-copied."""
    memo = {}
    for m in net.modules():
        r = m._replicate_for_data_parallel()
        r._former_parameters = {}
        memo[m] = r
    for m, r in memo.items():
        for k, child in m._modules.items():
            if child is not None:
                r._modules[k] = memo[child]
        for k, p in m._parameters.items():
            if p is not None:
                cp = p.detach().clone().requires_grad_(p.requires_grad) * 1.0  # non-leaf, like Broadcast.apply's output
                setattr(r, k, cp)
                r._former_parameters[k] = cp
    return memo[net]


def test_named_weights_on_dataparallel_replica():
    net = E.EDVR(nf=8, groups=2, front_RBs=1, back_RBs=1)
    rep = _fake_replica(net)
    assert len(list(rep.parameters())) == 0 and len(rep.state_dict()) == 0   # why state_dict() cannot feed the engine
    names = [k for k, _ in rep._named_weights()]
    assert names == list(net.state_dict().keys())
    assert rep._on_replica() and not net._on_replica()
    # replica weights require grad (non-leaf copies): with autograd on, the engine must not be chosen
    x = torch.zeros(1, 5, 3, 8, 8)
    rep.exec_path = "auto"
    assert rep._engine_ok(x) is False
    assert rep._engines is net._engines      # shared on purpose: replica d reuses device d's engine across iterations
    assert rep._weights_changed([None, ("anything", None)], "anything") is True   # replicas always re-hand weights


def test_engine_cache_survives_deepcopy_and_pickle():
    import copy
    import ctypes
    import pickle
    net = E.EDVR(nf=8, groups=2, front_RBs=1, back_RBs=1)
    net._engines[(0, "fp16")] = [ctypes.c_void_p(1234), None]   # stands for an EDVREngine (ctypes handles inside)
    c = copy.deepcopy(net)
    assert c._engines == {} and len(net._engines) == 1
    assert list(c.state_dict().keys()) == list(net.state_dict().keys())
    c2 = pickle.loads(pickle.dumps(net))
    assert c2._engines == {}


def test_weight_stamp_and_invalidate():
    net = E.EDVR(nf=8, groups=2, front_RBs=1, back_RBs=1)
    slot = [None, None]
    st = net._weight_stamp()
    assert net._weights_changed(slot, st)
    slot[1] = (st, None)
    assert not net._weights_changed(slot, net._weight_stamp())
    with torch.no_grad():
        net.conv_last.bias.add_(1.0)                       # in-place on the parameter: version bump -> seen
    assert net._weights_changed(slot, net._weight_stamp())
    slot[1] = (net._weight_stamp(), None)
    net.conv_last.bias.data.add_(1.0)                      # through .data: NOT seen by the stamp (torch semantics) ...
    assert not net._weights_changed(slot, net._weight_stamp())
    net._engines[(0, "fp32")] = slot
    net.invalidate_engine()                                # ... hence the explicit call
    assert net._weights_changed(slot, net._weight_stamp())
    net.engine_weight_check = "checksum"                   # ... or the checksum mode
    slot[1] = (net._weight_stamp(), net._weight_checksum())
    assert not net._weights_changed(slot, net._weight_stamp())
    net.conv_last.bias.data.add_(1.0)
    assert net._weights_changed(slot, net._weight_stamp())
    net.load_state_dict(net.state_dict())                  # hooks: load_state_dict / .to() invalidate
    assert net._engines[(0, "fp32")][1] is None

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


@pytest.mark.parametrize("name", ["edvr_tiny", "edvr_tiny_b2_g2", "edvr_noup_3f", "edvr_predeblur"])
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

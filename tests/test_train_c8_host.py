"""Host-side logic of the bf16 training path (realvsr_b200/train_c8.py) that needs no GPU."""
import pytest
import torch


def test_cpu_tensors_are_rejected_not_routed_elsewhere():
    from realvsr_b200 import train_c8 as T
    with pytest.raises(NotImplementedError):
        T.to_c8(torch.zeros(1, 8, 4, 4))
    with pytest.raises(NotImplementedError):
        T.GraphedStep(torch.nn.Identity(), torch.nn.functional.l1_loss, torch.zeros(1, 1), torch.zeros(1, 1))
    with pytest.raises(RuntimeError):
        T.conv(torch.zeros(1, 8, 4, 4, 8, dtype=torch.bfloat16), torch.zeros(64, 64, 3, 3))


def test_weight_layout_decision_is_the_librarys():
    """rvsr_c8_conv_layouts: which packed operand layout a launch reads -- the CTA-pair kernel's from 4 tiles on (4 rows x 30
    valid columns per tile for 3x3), the single-CTA kernels' below that and for 1x1 / 16-wide outputs.  The Python side packs
    exactly what this returns, so the decision exists once (tc_kernels.cu)."""
    from realvsr_b200 import _lib
    L = _lib.lib()
    assert L.rvsr_c8_conv_layouts(1, 64, 1, 4, 30, 64, 3, 0) == 1      # 1 tile
    assert L.rvsr_c8_conv_layouts(1, 64, 1, 8, 60, 64, 3, 0) == 2      # 4 tiles
    assert L.rvsr_c8_conv_layouts(2, 64, 80, 64, 64, 64, 3, 0) == 2    # torch.cat of two sources
    assert L.rvsr_c8_conv_layouts(4, 64, 16, 64, 64, 64, 3, 0) == 2    # data gradient of a 256-channel output: pair kernel only
    assert L.rvsr_c8_conv_layouts(1, 64, 16, 64, 64, 256, 3, 1) == 2   # pixel-shuffle convolution
    assert L.rvsr_c8_conv_layouts(5, 64, 16, 64, 64, 64, 1, 0) == 1    # 1x1
    assert L.rvsr_c8_conv_layouts(1, 64, 16, 256, 256, 16, 3, 0) == 1  # conv_last on the 16-wide tile
    assert L.rvsr_c8_conv_layouts(1, 64, 16, 64, 64, 216, 3, 0) == 0   # not a whole number of 64-wide tiles
    assert L.rvsr_c8_conv_layouts(1, 24, 16, 64, 64, 64, 3, 0) == 0    # source channels not a multiple of 16


def test_training_path_routing_rules():
    """EDVR._train_c8_ok: only under bf16 autocast (or explicit exec_path), nf == 64, standard stem, CUDA input."""
    from realvsr_b200.archs import EDVR_arch as E
    net = E.EDVR(nf=64, nframes=3, front_RBs=1, back_RBs=1)
    x = torch.zeros(1, 3, 3, 8, 8)
    assert not net._train_c8_ok(x)                      # CPU tensor, no autocast
    with torch.autocast("cuda", dtype=torch.bfloat16):
        assert not net._train_c8_ok(x)                  # CPU tensor
    net.exec_path = "train_c8"
    with pytest.raises(RuntimeError):
        net._train_c8_ok(x)                             # forced path, unsupported input: loud, no fallback
    small = E.EDVR(nf=8, nframes=3, groups=1, front_RBs=1, back_RBs=1)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        assert not small._train_c8_ok(x)

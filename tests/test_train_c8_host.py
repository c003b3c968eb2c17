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


def test_tile_count_matches_the_kernel_geometry():
    """_tiles decides which weight layout a launch will read (CTA-pair kernel from 4 tiles on): 4 rows x 30 (3x3) / 32 (1x1)
    valid columns per tile, as tc_kernels.cu's launch_conv_tc computes them."""
    from realvsr_b200.train_c8 import _tiles
    assert _tiles(1, 4, 30, 3) == 1 and _tiles(1, 4, 31, 3) == 2 and _tiles(1, 5, 30, 3) == 2
    assert _tiles(80, 64, 64, 3) == 80 * 16 * 3
    assert _tiles(2, 8, 32, 1) == 2 * 2 * 1 and _tiles(2, 8, 33, 1) == 2 * 2 * 2


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

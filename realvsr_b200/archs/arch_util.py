"""Host-side mirror of the hot-path part of the reference's ``codes/models/archs/arch_util.py``:
``initialize_weights`` (:8-25), ``make_layer`` (:28-39), ``ResidualBlock_noBN`` (:121-139).
The rest of that file (flow_warp, ResBlock, Upsampler, ...) serves other archs and is out of scope."""
import torch.nn as nn
import torch.nn.functional as F
import torch.nn.init as init


def initialize_weights(net_l, scale=1):
    """Kaiming-normal (fan_in) conv/linear weights times ``scale``, zero biases; BN -> (1, 0)."""
    for net in (net_l if isinstance(net_l, list) else [net_l]):
        for m in net.modules():
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                init.kaiming_normal_(m.weight, a=0, mode='fan_in')
                m.weight.data *= scale
                if m.bias is not None:
                    m.bias.data.zero_()
            elif isinstance(m, nn.BatchNorm2d):
                init.constant_(m.weight, 1)
                init.constant_(m.bias.data, 0.0)


def make_layer(basic_block, num_basic_block, **kwarg):
    """``num_basic_block`` instances of ``basic_block(**kwarg)`` in an nn.Sequential."""
    return nn.Sequential(*[basic_block(**kwarg) for _ in range(num_basic_block)])


class ResidualBlock_noBN(nn.Module):
    """x + conv2(relu(conv1(x))), both 3x3 nf->nf, initialised at 0.1x Kaiming."""

    def __init__(self, nf=64):
        super(ResidualBlock_noBN, self).__init__()
        self.conv1 = nn.Conv2d(nf, nf, 3, 1, 1, bias=True)
        self.conv2 = nn.Conv2d(nf, nf, 3, 1, 1, bias=True)
        initialize_weights([self.conv1, self.conv2], 0.1)

    def forward(self, x):
        return x + self.conv2(F.relu(self.conv1(x)))

"""Parameter containers with the reference's state_dict key names (factory/Norm.py:4-50).

These modules only HOLD parameters (so ``load_state_dict`` of a reference checkpoint works and the
initialisation statistics match); the arithmetic runs in libavc_b200.so, never through these modules."""
import torch.nn as nn


class ConvNorm(nn.Module):
    """Conv1d weights under ``.conv`` with xavier-uniform init (factory/Norm.py:4-37)."""

    def __init__(self, in_channels, out_channels, kernel_size=5, stride=1, padding=None, dilation=1, bias=True,
                 w_init_gain="linear"):
        super().__init__()
        if padding is None:
            assert kernel_size % 2 == 1
            padding = dilation * (kernel_size - 1) // 2
        assert stride == 1, "only stride-1 convolutions are on the conversion path"
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                              dilation=dilation, bias=bias)
        nn.init.xavier_uniform_(self.conv.weight, gain=nn.init.calculate_gain(w_init_gain))


class LinearNorm(nn.Module):
    """Linear weights under ``.linear_layer`` with xavier-uniform init (factory/Norm.py:40-50)."""

    def __init__(self, in_dim, out_dim, bias=True, w_init_gain="linear"):
        super().__init__()
        self.linear_layer = nn.Linear(in_dim, out_dim, bias=bias)
        nn.init.xavier_uniform_(self.linear_layer.weight, gain=nn.init.calculate_gain(w_init_gain))

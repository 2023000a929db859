"""CPU oracle for the MelGAN generator ``melgan/modules.py:72-130`` and
``MelVocoder.inverse`` (``melgan/interface.py:43-53``).  Test infrastructure.

The state_dict holds only ``weight_g`` / ``weight_v`` / ``bias`` (old-style
``torch.nn.utils.weight_norm``, dim=0): W = g * v / ||v||, the norm taken over all axes
but 0 (axis 0 is C_out for Conv1d and C_in for ConvTranspose1d).
"""
import torch
import torch.nn.functional as F

from .layers import cast_state_dict

RATIOS = (8, 8, 2, 2)          # melgan/modules.py:91


def wn_weight(sd, prefix):
    """Fold weight norm (melgan/modules.py:18-23)."""
    v = sd[prefix + ".weight_v"]
    g = sd[prefix + ".weight_g"]
    norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape([-1] + [1] * (v.dim() - 1))
    return g * v / norm


def lrelu(x):
    return F.leaky_relu(x, 0.2)


def resnet_block(sd, prefix, x, dilation):
    """``ResnetBlock.forward`` (melgan/modules.py:72-85): shortcut sees the raw x."""
    y = lrelu(x)
    y = F.pad(y, (dilation, dilation), mode="reflect")
    y = F.conv1d(y, wn_weight(sd, prefix + ".block.2"), sd[prefix + ".block.2.bias"], dilation=dilation)
    y = lrelu(y)
    y = F.conv1d(y, wn_weight(sd, prefix + ".block.4"), sd[prefix + ".block.4.bias"])
    s = F.conv1d(x, wn_weight(sd, prefix + ".shortcut"), sd[prefix + ".shortcut.bias"])
    return s + y


@torch.no_grad()
def melgan_forward(sd, mel, dtype=torch.float32, taps=None):
    """``Generator(80, 32, 3).forward`` (melgan/modules.py:88-130).  mel (B,80,T) -> (B,1,256*T)."""
    sd = cast_state_dict(sd, dtype)
    x = mel.to(dtype)
    x = F.pad(x, (3, 3), mode="reflect")                                   # model.0
    x = F.conv1d(x, wn_weight(sd, "model.1"), sd["model.1.bias"])          # model.1
    if taps is not None:
        taps["stem"] = x
    idx = 2
    for si, r in enumerate(RATIOS):
        x = lrelu(x)                                                       # model.{idx}
        x = F.conv_transpose1d(x, wn_weight(sd, f"model.{idx + 1}"), sd[f"model.{idx + 1}.bias"],
                               stride=r, padding=r // 2 + r % 2, output_padding=r % 2)
        if taps is not None:
            taps[f"up{si}"] = x
        for j in range(3):
            x = resnet_block(sd, f"model.{idx + 2 + j}", x, 3 ** j)
        if taps is not None:
            taps[f"stage{si}"] = x
        idx += 5
    x = lrelu(x)                                                           # model.22
    x = F.pad(x, (3, 3), mode="reflect")                                   # model.23
    x = F.conv1d(x, wn_weight(sd, "model.24"), sd["model.24.bias"])        # model.24
    return torch.tanh(x)                                                   # model.25


def vocoder_inverse(sd, mel, dtype=torch.float32):
    """``MelVocoder.inverse`` (melgan/interface.py:43-53): (B,80,T) -> (B,256T)."""
    return melgan_forward(sd, mel, dtype).squeeze(1)

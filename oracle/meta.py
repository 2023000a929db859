"""CPU oracle for ``factory/MetaPool.py`` / ``factory/MetaConv.py`` / ``factory/MLPMixer.py`` (eval-mode forward).
Test infrastructure.  Tensors are channels-first (B, C, L) exactly as in the reference; every function cites the
reference lines it restates.  ``kind`` is "pool" (MetaPool) or "conv" (MetaConv)."""
import torch
import torch.nn.functional as F

from .autovc import postnet_residual, upsample_codes
from .layers import batchnorm_eval, cast_state_dict


def group_norm1(sd, prefix, x, eps=1e-5):
    """``GroupNorm(1, C)`` (factory/Norm.py:53-60): per sample mean / biased variance over all C x L elements."""
    mean = x.mean(dim=(1, 2), keepdim=True)
    var = x.var(dim=(1, 2), keepdim=True, unbiased=False)
    w = sd[prefix + ".weight"].view(1, -1, 1)
    b = sd[prefix + ".bias"].view(1, -1, 1)
    return (x - mean) / torch.sqrt(var + eps) * w + b


def conv_bn_relu(sd, prefix, x):
    """``nn.Sequential(ConvNorm(k5,p2), BatchNorm1d, ReLU)`` (factory/MetaPool.py:37-43,53-63,92-96)."""
    y = F.conv1d(x, sd[prefix + ".0.conv.weight"], sd[prefix + ".0.conv.bias"], padding=2)
    return torch.relu(batchnorm_eval(sd, prefix + ".1", y))


def layer_norm(sd, prefix, x, eps=1e-5):
    mean = x.mean(dim=-1, keepdim=True)
    var = x.var(dim=-1, keepdim=True, unbiased=False)
    return (x - mean) / torch.sqrt(var + eps) * sd[prefix + ".weight"] + sd[prefix + ".bias"]


def gelu(x):
    return 0.5 * x * (1.0 + torch.erf(x / 2 ** 0.5))       # nn.GELU() default: exact erf


def mlp_mixer(sd, prefix, img, patch, taps=None):
    """``MLPMixer(image_size=S, channels=1, patch_size=p, dim=S, depth=1, out_dim=...)`` (factory/MLPMixer.py:58-92).

    img (B, S, S): rows = channel index of the preceding conv, columns = its length axis.  Returns (B, out, S)."""
    B, S, _ = img.shape
    n = S // patch
    # Rearrange "b c (h p1) (w p2) -> b (h w) (p1 p2 c)" with c = 1
    tok = img.reshape(B, n, patch, n, patch).permute(0, 1, 3, 2, 4).reshape(B, n * n, patch * patch)
    z = tok @ sd[prefix + ".1.weight"].t() + sd[prefix + ".1.bias"]                      # Linear(p^2 -> dim)
    # token mixing: PreNormResidual(dim, FeedForward(num_patches, dense=Conv1d k1))   (:80-84)
    y = layer_norm(sd, prefix + ".2.0.norm", z)
    u = F.conv1d(y, sd[prefix + ".2.0.fn.0.weight"], sd[prefix + ".2.0.fn.0.bias"])
    u = F.conv1d(gelu(u), sd[prefix + ".2.0.fn.3.weight"], sd[prefix + ".2.0.fn.3.bias"])
    z = u + z
    # channel mixing: PreNormResidual(dim, FeedForward(dim, dense=Linear))             (:85-87)
    y = layer_norm(sd, prefix + ".2.1.norm", z)
    v = gelu(y @ sd[prefix + ".2.1.fn.0.weight"].t() + sd[prefix + ".2.1.fn.0.bias"])
    z = v @ sd[prefix + ".2.1.fn.3.weight"].t() + sd[prefix + ".2.1.fn.3.bias"] + z
    if taps is not None:
        taps[prefix + ".z"] = z
    return F.conv1d(z, sd[prefix + ".3.weight"], sd[prefix + ".3.bias"], padding=2)       # Conv1d(np -> out, k5)


def meta_block(sd, prefix, x, kind, patch=8, taps=None):
    """``MetaBlock.forward`` (factory/MetaPool.py:66-77; factory/MetaConv.py:65-76)."""
    n1 = group_norm1(sd, prefix + ".norm1", x)
    if kind == "pool":        # Pooling: AvgPool1d(3,1,1,count_include_pad=False)(y) - y   (MetaPool.py:7-15)
        tm = F.avg_pool1d(n1, 3, stride=1, padding=1, count_include_pad=False) - n1
    else:                     # Conv5 + BN + ReLU                                           (MetaConv.py:24-34)
        tm = conv_bn_relu(sd, prefix + ".token_mixer", n1)
    x = x + tm
    a = conv_bn_relu(sd, prefix + ".conv_1", x)
    m = mlp_mixer(sd, prefix + ".mlp", group_norm1(sd, prefix + ".norm2", a), patch, taps)
    out = x + conv_bn_relu(sd, prefix + ".conv_2", m)
    if taps is not None:
        taps[prefix] = out
    return out


def meta_encoder(sd, x, c_org, dim_neck, freq, kind, taps=None):
    """``Encoder.forward`` (factory/MetaPool.py:109-133).  Returns codes (B, T/freq, 2*dim_neck)."""
    if x.dim() == 4:
        x = x.squeeze(1)
    B, T, _ = x.shape
    h = torch.cat((x.transpose(2, 1), c_org.unsqueeze(-1).expand(-1, -1, T)), dim=1)
    h = F.conv1d(h, sd["encoder.embding.proj.weight"], sd["encoder.embding.proj.bias"], padding=2)   # PatchEmbed
    if taps is not None:
        taps["enc_embed"] = h
    for i in range(3):
        h = meta_block(sd, f"encoder.metablock.{i}", h, kind, 8, taps)
    a = conv_bn_relu(sd, "encoder.output_conv", h)
    out = mlp_mixer(sd, "encoder.mlp", a, 16, taps).transpose(1, 2)          # (B, T, 2*dim_neck)
    if taps is not None:
        taps["enc_out"] = out
    return torch.cat((out[:, freq - 1::freq, :dim_neck], out[:, ::freq, dim_neck:]), dim=-1)


def meta_decoder(sd, dec_in, kind, taps=None):
    """``Decoder.forward`` (factory/MetaPool.py:160-182).  dec_in (B, T=176, 344) is consumed as
    (channels = T, length = 344); returns (B, 176, 80)."""
    h = F.conv1d(dec_in, sd["decoder.embding.proj.weight"], sd["decoder.embding.proj.bias"], padding=2)
    h = meta_block(sd, "decoder.metablock.0", h, kind, 8, taps)
    a = conv_bn_relu(sd, "decoder.output_conv_1", h)
    m = mlp_mixer(sd, "decoder.mlp", a, 8, taps)                              # (B, 88, 344)
    h = conv_bn_relu(sd, "decoder.output_conv_2", m.transpose(2, 1))          # (B, 176, 88)
    if taps is not None:
        taps["dec_conv2"] = h
    w = sd["decoder.linear_projection.linear_layer.weight"]
    b = sd["decoder.linear_projection.linear_layer.bias"]
    return h @ w.t() + b


@torch.no_grad()
def meta_forward(sd, x, c_org, c_trg, dim_neck, freq, kind, dtype=torch.float32, taps=None):
    """``MetaPool.forward`` / ``MetaConv.forward`` (factory/MetaPool.py:256-276)."""
    sd = cast_state_dict(sd, dtype)
    x, c_org = x.to(dtype), c_org.to(dtype)
    codes = meta_encoder(sd, x, c_org, dim_neck, freq, kind, taps)
    flat = codes.reshape(codes.shape[0], -1)
    if taps is not None:
        taps["codes"] = flat
    if c_trg is None:
        return flat
    T = x.shape[-2]
    dec_in = upsample_codes(codes, T, c_trg.to(dtype))
    mel = meta_decoder(sd, dec_in, kind, taps)
    post = mel + postnet_residual(sd, mel, taps)
    if taps is not None:
        taps["mel"] = mel
        taps["mel_postnet"] = post
    return mel.unsqueeze(1), post.unsqueeze(1), flat

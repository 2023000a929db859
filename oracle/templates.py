"""state_dict templates (key -> zero tensor of the reference's shape/dtype) written out by hand,
so tests and the bench can build seeded weights on a box without /root/reference.
Test infrastructure.  ``tests/test_oracle_golden.py::test_templates_match_reference`` compares them
with the reference's own ``state_dict()`` whenever the reference tree is present.
"""
import torch


def _z(*shape):
    return torch.zeros(shape, dtype=torch.float32)


def _conv_bn(sd, prefix, c_out, c_in, k=5):
    """nn.Sequential(ConvNorm, BatchNorm1d) keys (factory/AutoVC.py:26-39)."""
    sd[f"{prefix}.0.conv.weight"] = _z(c_out, c_in, k)
    sd[f"{prefix}.0.conv.bias"] = _z(c_out)
    _bn(sd, f"{prefix}.1", c_out)


def _bn(sd, prefix, c):
    sd[f"{prefix}.weight"] = _z(c)
    sd[f"{prefix}.bias"] = _z(c)
    sd[f"{prefix}.running_mean"] = _z(c)
    sd[f"{prefix}.running_var"] = _z(c)
    sd[f"{prefix}.num_batches_tracked"] = torch.zeros((), dtype=torch.int64)


def _lstm(sd, prefix, in_dim, hidden, layers, bidirectional=False):
    dirs = ["", "_reverse"] if bidirectional else [""]
    for l in range(layers):
        i = in_dim if l == 0 else hidden * len(dirs)
        for d in dirs:
            sd[f"{prefix}.weight_ih_l{l}{d}"] = _z(4 * hidden, i)
            sd[f"{prefix}.weight_hh_l{l}{d}"] = _z(4 * hidden, hidden)
            sd[f"{prefix}.bias_ih_l{l}{d}"] = _z(4 * hidden)
            sd[f"{prefix}.bias_hh_l{l}{d}"] = _z(4 * hidden)


def _postnet(sd, prefix="postnet"):
    """factory/AutoVC.py:117-171."""
    _conv_bn(sd, f"{prefix}.convolutions.0", 512, 80)
    for i in (1, 2, 3):
        _conv_bn(sd, f"{prefix}.convolutions.{i}", 512, 512)
    _conv_bn(sd, f"{prefix}.convolutions.4", 80, 512)


def autovc_template(dim_neck, dim_emb, dim_pre, freq):
    """factory/AutoVC.py:182-189."""
    sd = {}
    for i in range(3):
        _conv_bn(sd, f"encoder.convolutions.{i}", 512, 80 + dim_emb if i == 0 else 512)
    _lstm(sd, "encoder.lstm", 512, dim_neck, 2, bidirectional=True)
    _lstm(sd, "decoder.lstm1", dim_neck * 2 + dim_emb, dim_pre, 1)
    for i in range(3):
        _conv_bn(sd, f"decoder.convolutions.{i}", dim_pre, dim_pre)
    _lstm(sd, "decoder.lstm2", dim_pre, 1024, 2)
    sd["decoder.linear_projection.linear_layer.weight"] = _z(80, 1024)
    sd["decoder.linear_projection.linear_layer.bias"] = _z(80)
    _postnet(sd)
    return sd


def lstmdv_template():
    """factory/LstmDV.py:8-17."""
    sd = {}
    _lstm(sd, "lstm", 80, 768, 3)
    sd["embedding.weight"] = _z(256, 768)
    sd["embedding.bias"] = _z(256)
    return sd


def lstmdv_twin_template(num_classes=256):
    """make_data/factory/LstmDV.py:8-16: the embedder plus the ``output`` classifier head."""
    sd = lstmdv_template()
    sd["output.weight"] = _z(num_classes, 256)
    sd["output.bias"] = _z(num_classes)
    return sd


def _wn(sd, prefix, w_shape, bias_len):
    sd[f"{prefix}.bias"] = _z(bias_len)
    sd[f"{prefix}.weight_g"] = _z(w_shape[0], 1, 1)
    sd[f"{prefix}.weight_v"] = _z(*w_shape)


def melgan_template(input_size=80, ngf=32, n_residual_layers=3):
    """melgan/modules.py:88-127 with ratios [8, 8, 2, 2]."""
    sd = {}
    ratios = (8, 8, 2, 2)
    mult = 2 ** len(ratios)
    _wn(sd, "model.1", (mult * ngf, input_size, 7), mult * ngf)
    idx = 2
    for r in ratios:
        c_in, c_out = mult * ngf, mult * ngf // 2
        _wn(sd, f"model.{idx + 1}", (c_in, c_out, 2 * r), c_out)       # ConvTranspose1d: (C_in, C_out, K)
        for j in range(n_residual_layers):
            p = f"model.{idx + 2 + j}"
            _wn(sd, f"{p}.block.2", (c_out, c_out, 3), c_out)
            _wn(sd, f"{p}.block.4", (c_out, c_out, 1), c_out)
            _wn(sd, f"{p}.shortcut", (c_out, c_out, 1), c_out)
        idx += 2 + n_residual_layers
        mult //= 2
    _wn(sd, f"model.{idx + 2}", (1, ngf, 7), 1)
    return sd


def _mixer(sd, prefix, image, patch, out_dim):
    """factory/MLPMixer.py:58-92 with channels=1, dim=image, depth=1."""
    np_ = (image // patch) ** 2
    dim = image
    sd[f"{prefix}.1.weight"] = _z(dim, patch * patch)
    sd[f"{prefix}.1.bias"] = _z(dim)
    sd[f"{prefix}.2.0.fn.0.weight"] = _z(4 * np_, np_, 1)
    sd[f"{prefix}.2.0.fn.0.bias"] = _z(4 * np_)
    sd[f"{prefix}.2.0.fn.3.weight"] = _z(np_, 4 * np_, 1)
    sd[f"{prefix}.2.0.fn.3.bias"] = _z(np_)
    sd[f"{prefix}.2.0.norm.weight"] = _z(dim)
    sd[f"{prefix}.2.0.norm.bias"] = _z(dim)
    sd[f"{prefix}.2.1.fn.0.weight"] = _z(4 * dim, dim)
    sd[f"{prefix}.2.1.fn.0.bias"] = _z(4 * dim)
    sd[f"{prefix}.2.1.fn.3.weight"] = _z(dim, 4 * dim)
    sd[f"{prefix}.2.1.fn.3.bias"] = _z(dim)
    sd[f"{prefix}.2.1.norm.weight"] = _z(dim)
    sd[f"{prefix}.2.1.norm.bias"] = _z(dim)
    sd[f"{prefix}.3.weight"] = _z(out_dim, np_, 5)
    sd[f"{prefix}.3.bias"] = _z(out_dim)


def _conv_bn_seq(sd, prefix, c_out, c_in):
    sd[f"{prefix}.0.conv.weight"] = _z(c_out, c_in, 5)
    sd[f"{prefix}.0.conv.bias"] = _z(c_out)
    _bn(sd, f"{prefix}.1", c_out)


def _meta_block(sd, prefix, kind, dim, crop, out_neck=88, patch=8):
    """factory/MetaPool.py:18-64 / factory/MetaConv.py:8-63."""
    sd[f"{prefix}.norm1.weight"] = _z(dim)
    sd[f"{prefix}.norm1.bias"] = _z(dim)
    if kind == "conv":
        _conv_bn_seq(sd, f"{prefix}.token_mixer", 512, 512)
    sd[f"{prefix}.norm2.weight"] = _z(crop)
    sd[f"{prefix}.norm2.bias"] = _z(crop)
    _conv_bn_seq(sd, f"{prefix}.conv_1", crop, 512)
    _mixer(sd, f"{prefix}.mlp", crop, patch, out_neck)
    _conv_bn_seq(sd, f"{prefix}.conv_2", 512, out_neck)


def meta_template(kind, dim_neck, dim, dim_pre, freq):
    """MetaPool / MetaConv (factory/MetaPool.py:249-254): kind = "pool" | "conv"."""
    sd = {}
    sd["encoder.embding.proj.weight"] = _z(512, 336, 5)
    sd["encoder.embding.proj.bias"] = _z(512)
    for i in range(3):
        _meta_block(sd, f"encoder.metablock.{i}", kind, dim_pre, 176)
    _conv_bn_seq(sd, "encoder.output_conv", 176, 512)
    _mixer(sd, "encoder.mlp", 176, 16, 2 * dim_neck)
    sd["decoder.embding.proj.weight"] = _z(512, 176, 5)
    sd["decoder.embding.proj.bias"] = _z(512)
    _meta_block(sd, "decoder.metablock.0", kind, dim_pre, 344)
    _conv_bn_seq(sd, "decoder.output_conv_1", 344, 512)
    _mixer(sd, "decoder.mlp", 344, 8, 88)
    _conv_bn_seq(sd, "decoder.output_conv_2", 176, 344)
    sd["decoder.linear_projection.linear_layer.weight"] = _z(80, 88)
    sd["decoder.linear_projection.linear_layer.bias"] = _z(80)
    _postnet(sd)
    return sd


def _adjust(sd, dim_emb, prefix="adjust", dim_cell=768):
    """factory/Adjust.py:8-27."""
    for i in range(3):
        _conv_bn(sd, f"{prefix}.convolutions.{i}", 512, 80 + dim_emb if i == 0 else 512)
    _lstm(sd, f"{prefix}.lstm", 512, dim_cell, 3)
    sd[f"{prefix}.embedding.linear_layer.weight"] = _z(256, dim_cell)
    sd[f"{prefix}.embedding.linear_layer.bias"] = _z(256)


def autovc_adjust_template(dim_neck, dim_emb, dim_pre, freq):
    """factory/AutoVC_Adjust.py:169-175: AutoVC + ``adjust``."""
    sd = autovc_template(dim_neck, dim_emb, dim_pre, freq)
    _adjust(sd, dim_emb)
    return sd


def meta_adjust_template(kind, dim_neck, dim_emb, dim_pre, freq):
    """factory/MetaPool_Adjust.py:250-256 / MetaConv_Adjust.py: Meta model + ``adjust``."""
    sd = meta_template(kind, dim_neck, dim_emb, dim_pre, freq)
    _adjust(sd, dim_emb)
    return sd


def _adain_parts(sd):
    """factory/AutoVC2.py:19-33 (encoder.feature_pre_extract) and :176-190 (postnet.feature_last_combine)."""
    for i in range(3):
        _conv_bn(sd, f"encoder.feature_pre_extract.{i}", 80, 80)
        sd[f"postnet.feature_last_combine.{i}.0.conv.weight"] = _z(80, 80, 5)
        sd[f"postnet.feature_last_combine.{i}.0.conv.bias"] = _z(80)


def autovc2_template(dim_neck, dim_emb, dim_pre, freq):
    sd = autovc_template(dim_neck, dim_emb, dim_pre, freq)
    _adain_parts(sd)
    return sd


def meta2_template(kind, dim_neck, dim_emb, dim_pre, freq):
    sd = meta_template(kind, dim_neck, dim_emb, dim_pre, freq)
    _adain_parts(sd)
    return sd

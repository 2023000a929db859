"""The AdaIN "2" variants (factory/AutoVC2.py, MetaPool2.py, MetaConv2.py) as a mixin over the drop-in
models.

Encoder side (AutoVC2.py:54-60): three [Conv1d(80 -> 80, k5) + BatchNorm] layers run on the mel BEFORE the speaker
concat; after each one the scalar ``x.mean()`` and ``x.std()`` over the whole (B, 80, T) tensor are recorded as the
utterance batch's "features", and the encoder continues from the third layer's output.  Postnet side
(AutoVC2.py:192-203): after the five postnet convolutions, three times ``x = combine_i(AdaIN(x, mu_i, std_i))`` with
``AdaIN(x, mu, std) = (x - x.mean()) / x.std() * std + mu`` (factory/Norm.py:86-94) and combine_i = Conv1d(80 -> 80, k5).
``forward(x, c_org, c_trg, target_feature=None)`` returns ``(codes, features)`` when both c_trg and target_feature are
None, else the usual 3-tuple, styled with ``target_feature`` if given (util/evaluate.py:81-82) or with the source's own
features.

The statistics are batch-global scalars, exactly as in the reference: a call's result depends on which utterances share
the batch, so these variants are NOT utterance-shardable; per call they match the reference.
Kernels: ``avc_conv_gemm`` (the six small convolutions), ``avc_global_stats``, ``avc_adain``.
"""
import torch
import torch.nn as nn

from .. import layers, ops, packing
from .Norm import ConvNorm


class AdaIN(nn.Module):
    """Parameter-free placeholder so the module tree matches the reference (factory/Norm.py:86-94)."""


class AdaINPlan:
    def __init__(self, sd, precision):
        self.pre = [layers.conv_bn_layer(sd, f"encoder.feature_pre_extract.{i}", precision, "none") for i in range(3)]
        self.combine = [ops.ConvGemm(*packing.pack_conv(sd[f"postnet.feature_last_combine.{i}.0.conv.weight"],
                                                        sd[f"postnet.feature_last_combine.{i}.0.conv.bias"], precision),
                                     act="none", tag="conv") for i in range(3)]


class AdaINMixin:
    """Mix in BEFORE the base model class: ``class AutoVC2(AdaINMixin, AutoVC)``."""

    def _init_adain(self):
        self.encoder.feature_pre_extract = nn.ModuleList([
            nn.Sequential(ConvNorm(80, 80, kernel_size=5, stride=1, padding=2, dilation=1, w_init_gain="linear"),
                          nn.BatchNorm1d(80)) for _ in range(3)])
        self.postnet.adain = AdaIN()
        self.postnet.feature_last_combine = nn.ModuleList([
            nn.Sequential(ConvNorm(80, 80, kernel_size=5, stride=1, padding=2, dilation=1, w_init_gain="linear"))
            for _ in range(3)])
        self._adain_cache = layers.PlanCache()
        self._features = None
        self._target_stats = None

    def _adain_plan(self):
        def build():
            sd = layers.state_for_packing(self)
            return AdaINPlan(sd, self.precision)
        return self._adain_cache.get(self, (self.precision,), build)

    # ---- hooks called by the base forward
    def _encoder_features(self, x, plan, B, T):
        ap = self._adain_plan()
        prec = self.precision
        stats = torch.empty(3, 2, dtype=torch.float32, device=x.device)
        h = ops.to_act(x, prec)
        for i, conv in enumerate(ap.pre):                                  # AutoVC2.py:57-60
            f32 = torch.empty(B, T, 80, dtype=torch.float32, device=x.device)
            o = ops.alloc_act(B, T, 80, prec, x.device)
            conv(h, B, T, out=o, out2=f32.view(B * T, 80))
            ops.global_stats(f32, out=stats[i])
            h, x = o, f32
        self._features = stats
        return x                                                           # the encoder continues from the features

    def _run_postnet(self, plan, mel_op, mel, B, T, taps):
        ap = self._adain_plan()
        prec = self.precision
        style = self._target_stats if self._target_stats is not None else self._features
        h = plan.postnet.hidden(mel_op, B, T, taps)                        # AutoVC2.py:193-194
        cur = torch.empty(B, T, 80, dtype=torch.float32, device=mel.device)
        plan.postnet.convs[4](h, B, T, out2=cur.view(B * T, 80))           # :196
        for i, comb in enumerate(ap.combine):                              # :197-199
            a_op, _ = ops.adain(cur, ops.global_stats(cur), style[i], prec)
            nxt = torch.empty(B, T, 80, dtype=torch.float32, device=mel.device)
            last = i == len(ap.combine) - 1                                # the residual mel + postnet rides on the last
            comb(a_op, B, T, out2=nxt.view(B * T, 80), residual=mel.view(B * T, 80) if last else None)
            cur = nxt
        return cur

    # ---- public surface
    def features(self):
        """[[mean, std]] x 3 as 0-dim tensors, the structure Encoder.forward returns (AutoVC2.py:60)."""
        st = self._features
        return [[st[i, 0], st[i, 1]] for i in range(3)]

    @staticmethod
    def _as_stats(target_feature, device):
        rows = [torch.stack([torch.as_tensor(m, dtype=torch.float32, device=device).reshape(()),
                             torch.as_tensor(s, dtype=torch.float32, device=device).reshape(())])
                for m, s in target_feature]
        return torch.stack(rows).contiguous()

    @ops.on_device_of_input
    @torch.no_grad()
    def forward(self, x, c_org, c_trg, target_feature=None):
        ops._require_cuda(x)
        self._target_stats = self._as_stats(target_feature, x.device) if target_feature is not None else None
        try:
            if c_trg is None:
                if target_feature is not None:
                    raise AttributeError("'NoneType' object has no attribute 'unsqueeze' (c_trg is required when a "
                                         "target_feature is given, as in the reference)")
                codes = super().forward(x, c_org, None)
                return codes, self.features()                              # AutoVC2.py:219-220
            return super().forward(x, c_org, c_trg)
        finally:
            self._target_stats = None

"""Batched conversion pipeline: speaker embedding -> AutoVC conversion -> MelGAN vocoding (BASELINE config 4).

Follows the reference's own recipe for utterances whose length is not a multiple of ``freq``: zero-pad the mel at
the end, convert, trim the padding off again (util/evaluate.py:36-43 ``crop_mel`` pads with constant 0;
:85-92 trims ``mel_trans`` back), then vocode ``mel_trans.transpose(2, 1)`` (conversion.ipynb cell 14,
util/evaluate.py:96-98).  Utterances of one call share the same T (callers bucket by exact length: zero padding
changes the result)."""
import torch


def pad_to_multiple(mel, multiple):
    """(B, T, 80) -> ((B, T', 80) zero-padded at the end, pad_size)   (util/evaluate.py:36-43)."""
    T = mel.shape[1]
    pad = (-T) % multiple
    if pad:
        mel = torch.nn.functional.pad(mel, (0, 0, 0, pad))
    return mel, pad


@torch.no_grad()
def convert(model, mel_src, emb_org, emb_trg):
    """AutoVC conversion with pad -> convert -> trim.  Returns the converted mel (B, T, 80)."""
    T = mel_src.shape[1]
    x, pad = pad_to_multiple(mel_src, model.freq)
    _, mel_trans, _ = model(x, emb_org, emb_trg)                 # util/evaluate.py:83
    mel_trans = mel_trans.squeeze(1)                             # :85
    return mel_trans[:, :T, :] if pad else mel_trans             # :87-90


@torch.no_grad()
def convert_and_vocode(embedder, model, generator, mel_src, mel_trg_ref):
    """mel_src (B, T, 80): utterances to convert; mel_trg_ref (B, T2, 80): utterances of the target speakers.

    Returns (converted mel (B, T, 80), waveform (B, 256 T), emb_org, emb_trg)."""
    emb_org = embedder(mel_src)                                  # LstmDV d-vectors (factory/LstmDV.py:19-24)
    emb_trg = embedder(mel_trg_ref)
    mel_trans = convert(model, mel_src, emb_org, emb_trg)
    wav = generator(mel_trans.transpose(2, 1).contiguous()).squeeze(1)   # MelVocoder.inverse (interface.py:43-53)
    return mel_trans, wav, emb_org, emb_trg

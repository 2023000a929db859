"""``Evaluator`` with the reference's surface (util/evaluate.py:8-218) on top of the B200 conversion path.

The reference's notebooks drive conversion through this object (conversion.ipynb cell 8:
``E.get_trans_mel(model, source_id, target_id, sound_id, isAdjust, isAdain, isPlay)``, cell 14:
``E.get_wavs(mel_trans.transpose(2, 1))``), so it is the caller contract of the drop-in.  What matters numerically is
the padding recipe: ``crop_mel`` pads an utterance shorter than ``len_crop`` with zeros AT THE END UP TO ``len_crop``
(176) -- not to the next multiple of ``freq`` -- crops a longer one at a random offset, and ``get_trans_mel`` trims the
padded frames off again only when ``isPlay`` (:36-50, :85-92).  The backward encoder LSTM and the edge convolutions
see the padding, so a different pad length gives a different mel.

Config attributes are the reference's (:11-21): ``root, num_speaker, batch_size, max_uttr_idx, erroment_num,
len_crop, device, all_speaker, embedder, metadata``; one optional extra, ``vocoder`` (a ``MelVocoder`` or a MelGAN
``state_dict``), replaces the hard-wired ``model/static/multi_speaker.pt`` load when present."""
import random

import numpy as np
import torch
import torch.nn.functional as F

from ..melgan.interface import MelVocoder
from ..pipeline import crop_mel as _crop_mel


class Evaluator:
    def __init__(self, config):
        self.root = config.root
        self.num_speaker = config.num_speaker
        self.batch_size = config.batch_size
        self.max_uttr_idx = config.max_uttr_idx
        self.erroment_num = config.erroment_num
        self.len_crop = config.len_crop
        self.device = config.device
        self.all_speaker = config.all_speaker
        self.embedder = config.embedder
        self.enrroment_idx = []
        self.remain_idx = np.arange(2, config.max_uttr_idx)
        self.metadata = self.build_metadata(config.metadata)
        voc = getattr(config, "vocoder", None)
        if isinstance(voc, MelVocoder):
            self.vocoder = voc
        elif voc is not None:
            self.vocoder = MelVocoder(device=self.device, state_dict=voc)
        else:
            self.vocoder = MelVocoder(device=self.device, model_name="model/static/multi_speaker")   # :23
        if self.embedder is not None:
            self.all_dv = self.generate_real_dv()                                                    # :24-26

    # ------------------------------------------------------------------ data access (:28-54)
    def build_metadata(self, metadata):
        return sorted(entry for entry in metadata if str(entry[0]) in self.all_speaker)

    def crop_mel(self, tmp):
        """numpy (T, 80) -> (tensor (1, len_crop, 80) on the device, pad_size)."""
        mel, pad = _crop_mel(tmp, self.len_crop)
        return mel.to(self.device), pad

    def get_mel(self, speaker_id, sound_id):
        path = self.metadata[speaker_id][sound_id].replace("\\", "/")
        return self.crop_mel(np.load(f"{self.root}/{path}"))

    # ------------------------------------------------------------------ conversion call site (:56-98)
    def get_trans_mel(self, model, source_id, target_id, sound_id, isAdjust, isAdain, isPlay=False):
        mel_source, pad_source = self.get_mel(source_id, sound_id)
        mel_target, pad_target = self.get_mel(target_id, sound_id)
        emb_org = torch.from_numpy(self.metadata[source_id][1]).unsqueeze(0).to(self.device)
        emb_trg = torch.from_numpy(self.metadata[target_id][1]).unsqueeze(0).to(self.device)
        if isAdjust:
            _, _, mel_trans, _ = model(mel_source, emb_org, emb_trg, True, mel_target)
        elif isAdain:
            _, feature = model(mel_source, emb_org, None, None)
            _, mel_trans, _ = model(mel_source, emb_org, emb_trg, feature)
        else:
            _, mel_trans, _ = model(mel_source, emb_org, emb_trg)
        mel_trans = mel_trans.squeeze(1)
        if isPlay:
            if pad_source > 0:
                keep = self.len_crop - pad_source
                mel_source, mel_trans = mel_source[:, :keep, :], mel_trans[:, :keep, :]
            if pad_target > 0:
                mel_target = mel_target[:, :self.len_crop - pad_target, :]
        return mel_source, mel_target, mel_trans

    def get_wavs(self, mel):
        """mel (1, 80, crop_len) -> waveform (1, 256 crop_len)."""
        return self.vocoder.inverse(mel)

    # ------------------------------------------------------------------ d-vector enrolment (:100-121)
    def get_dv(self, speaker_id):
        dv = torch.zeros((1, 256))
        for idx in self.enrroment_idx:
            mel, _ = self.get_mel(speaker_id, idx)
            dv += self.embedder(mel)[1].detach().cpu()         # the make_data twin's (predictions, d_vec)
        return (dv / self.erroment_num).to(self.device)

    def generate_real_dv(self):
        for _ in range(self.erroment_num):
            pick = random.choice(self.remain_idx)
            self.remain_idx = np.delete(self.remain_idx, np.where(self.remain_idx == pick))
            self.enrroment_idx.append(pick)
        return [self.get_dv(i) for i, _ in enumerate(self.all_speaker)]

    # ------------------------------------------------------------------ cosine-similarity tables (:123-218)
    @staticmethod
    def _cos(a, b):
        return float(torch.clamp(F.cosine_similarity(a.detach().cpu(), b.detach().cpu(), dim=1, eps=1e-8), min=0.0)[0])

    def get_real_data_cos(self):
        n = self.num_speaker
        mean, std = np.zeros((n, n)), np.zeros((n, n))
        for i, _ in enumerate(self.all_speaker):
            styles = [self.embedder(self.get_mel(i, s)[0])[1] for s in self.remain_idx]
            for j, dv in enumerate(self.all_dv):
                c = np.array([self._cos(dv, st) for st in styles], dtype=np.float32)
                mean[i][j], std[i][j] = c.mean(), c.std()
        return mean, std

    def get_cos_rc(self, cos_res):
        n = len(cos_res)
        same = sum(cos_res[i][i][i] for i in range(n))
        other = sum((sum(cos_res[i][i]) - cos_res[i][i][i]) / (n - 1) for i in range(n))
        return same / n, other / n

    def get_cos_trans(self, cos_res):
        n = len(cos_res)
        to_target = [(np.sum(np.diagonal(cos_res[i])) - cos_res[i][i][i]) / (n - 1) for i in range(n)]
        to_other = [(np.sum(cos_res[i]) - np.sum(np.diagonal(cos_res[i]))) / ((n - 1) * n) for i in range(n)]
        return sum(to_target) / n, sum(to_other) / n

    def generate_result(self, models, sound_id=2, isAdjust=False, isAdain=False):
        n = self.num_speaker
        out = [np.zeros((n, n, n)) for _ in models]
        for s, _ in enumerate(self.metadata):
            for t, _ in enumerate(self.metadata):
                styles = [self.embedder(self.get_trans_mel(m, s, t, sound_id, isAdjust, isAdain)[2])[1] for m in models]
                for k, emb in enumerate(self.all_dv):
                    for mi, st in enumerate(styles):
                        out[mi][s][t][k] = self._cos(st, emb)
        return out

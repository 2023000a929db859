"""``MelVocoder`` with the reference's surface (melgan/interface.py:23-53), running on the GPU kernels.

``inverse(mel)`` is the conversion path's vocoder call (conversion.ipynb cell 14: ``E.get_wavs(mel.transpose(2, 1))``).
``__call__`` (audio -> mel, ``Audio2Mel``) needs librosa's mel filter bank and is outside this path."""
import os

import torch

from .modules import Generator


def get_default_device():
    return "cuda"


class MelVocoder:
    def __init__(self, device=get_default_device(), model_name="multi_speaker", state_dict=None):
        netG = Generator(80, 32, 3)
        if state_dict is None:
            path = f"{model_name}.pt"
            if not os.path.exists(path):
                raise FileNotFoundError(f"{path} not found (the reference loads it the same way, interface.py:29)")
            state_dict = torch.load(path, map_location="cpu")
        netG.load_state_dict(state_dict)
        self.mel2wav = netG.to(device).eval()
        self.device = device

    def __call__(self, audio):
        raise NotImplementedError("Audio2Mel (wav -> mel) is outside the accelerated conversion path")

    def inverse(self, mel):
        """mel (B, 80, T) -> waveform (B, 256 T)."""
        with torch.no_grad():
            return self.mel2wav(mel.to(self.device)).squeeze(1)

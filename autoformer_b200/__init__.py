"""autoformer_b200 -- B200-native (sm_100a) batched voice-conversion forward path.

Drop-in replacements for the reference's model classes on the conversion path
(``factory.AutoVC.AutoVC``, ``factory.LstmDV.LstmDV``, ``melgan`` generator / vocoder)
whose ``forward`` runs hand-written CUDA kernels from ``libavc_b200.so`` through a C ABI
(``include/avc_b200.h``).  There is no CPU or PyTorch-eager fallback: importing the
models works anywhere, calling them without the built library or without a GPU raises.
"""
__version__ = "0.1.0"

"""MelGAN vocoder on libavc_b200.so: ``modules.Generator`` and ``interface.MelVocoder`` mirror melgan/ of the reference."""

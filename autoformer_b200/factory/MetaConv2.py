"""Drop-in ``factory.MetaConv2.MetaConv2`` (AdaIN-styled MetaConv, factory/MetaConv2.py); see ``_adain.py``."""
from ._adain import AdaINMixin
from ._meta import MetaBase


class MetaConv2(AdaINMixin, MetaBase):
    KIND = "conv"

    def __init__(self, dim_neck, dim_emb, dim_pre, freq):
        MetaBase.__init__(self, dim_neck, dim_emb, dim_pre, freq)
        self._init_adain()

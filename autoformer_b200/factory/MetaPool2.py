"""Drop-in ``factory.MetaPool2.MetaPool2`` (AdaIN-styled MetaPool, factory/MetaPool2.py:296-330); see ``_adain.py``."""
from ._adain import AdaINMixin
from ._meta import MetaBase


class MetaPool2(AdaINMixin, MetaBase):
    KIND = "pool"

    def __init__(self, dim_neck, dim_emb, dim_pre, freq):
        MetaBase.__init__(self, dim_neck, dim_emb, dim_pre, freq)
        self._init_adain()

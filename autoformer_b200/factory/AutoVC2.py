"""Drop-in ``factory.AutoVC2.AutoVC2`` (AdaIN-styled AutoVC, factory/AutoVC2.py:204-243); see ``_adain.py``."""
from ._adain import AdaINMixin
from .AutoVC import AutoVC


class AutoVC2(AdaINMixin, AutoVC):
    def __init__(self, dim_neck, dim_emb, dim_pre, freq):
        AutoVC.__init__(self, dim_neck, dim_emb, dim_pre, freq)
        self._init_adain()

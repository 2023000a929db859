"""Drop-in ``factory.MetaConv.MetaConv(dim_neck, dim, dim_pre, freq)`` on libavc_b200.so (factory/MetaConv.py:247-274)."""
from ._meta import MetaBase


class MetaConv(MetaBase):
    KIND = "conv"

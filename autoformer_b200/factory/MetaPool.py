"""Drop-in ``factory.MetaPool.MetaPool(dim_neck, dim, dim_pre, freq)`` on libavc_b200.so (factory/MetaPool.py:249-276)."""
from ._meta import MetaBase


class MetaPool(MetaBase):
    KIND = "pool"

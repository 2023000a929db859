"""Drop-in ``factory.MetaPool_Adjust`` (factory/MetaPool_Adjust.py:250-283).  In the reference this file's class is
still called ``MetaPool``; unlike the other two ``*_Adjust`` models its encoder sees the UN-adjusted source code -- only
the target code passes through ``Adjust`` -- and ``c_org`` is returned unchanged.  Both names are exported."""
from .MetaConv_Adjust import MetaConv_Adjust


class MetaPool(MetaConv_Adjust):
    KIND = "pool"
    ADJUST_SOURCE = False


MetaPool_Adjust = MetaPool

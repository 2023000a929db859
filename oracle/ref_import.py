"""Import the UNMODIFIED reference from /root/reference.  Build-container only.

Nothing under ``tests/`` marked gpu, ``smoke()`` or ``bench.py`` may call this at run
time: ``/root/reference`` does not exist on the GPU box.  It is used by
``oracle/make_golden.py`` (fixture generation) and by CPU tests that skip when the
reference tree is absent.
"""
import os
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get("AUTOFORMER_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "factory", "AutoVC.py"))


def _stub_librosa():
    # melgan/modules.py:4 imports librosa.filters.mel at module import; only Audio2Mel
    # (out of scope) ever calls it.  librosa is not installed and there is no network.
    if "librosa" in sys.modules:
        return
    lib = types.ModuleType("librosa")
    filt = types.ModuleType("librosa.filters")

    def _mel(*a, **k):
        raise RuntimeError("librosa is stubbed; Audio2Mel is out of scope")
    filt.mel = _mel
    lib.filters = filt
    sys.modules["librosa"] = lib
    sys.modules["librosa.filters"] = filt


def load():
    """Return a namespace with the reference classes on the hot path."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _stub_librosa()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    warnings.filterwarnings("ignore", category=FutureWarning)
    import importlib
    ns = types.SimpleNamespace()
    ns.AutoVC = importlib.import_module("factory.AutoVC").AutoVC
    ns.MetaPool = importlib.import_module("factory.MetaPool").MetaPool
    ns.MetaConv = importlib.import_module("factory.MetaConv").MetaConv
    ns.LstmDV = importlib.import_module("factory.LstmDV").LstmDV
    ns.Generator = importlib.import_module("melgan.modules").Generator
    ns.Adjust = importlib.import_module("factory.Adjust").Adjust
    ns.AutoVC_Adjust = importlib.import_module("factory.AutoVC_Adjust").AutoVC_Adjust
    ns.MetaPool_Adjust = importlib.import_module("factory.MetaPool_Adjust").MetaPool     # (sic) the file's class name
    ns.MetaConv_Adjust = importlib.import_module("factory.MetaConv_Adjust").MetaConv_Adjust
    ns.AutoVC2 = importlib.import_module("factory.AutoVC2").AutoVC2
    ns.MetaPool2 = importlib.import_module("factory.MetaPool2").MetaPool2
    ns.MetaConv2 = importlib.import_module("factory.MetaConv2").MetaConv2
    return ns

"""Install the UNMODIFIED reference into baseline/_ref (git-ignored; travels to the GPU box like the built .so).

    python baseline/install_reference.py            # build container only: /root/reference must exist

The reference ships no setup.py / pyproject.toml, so the contract's
``pip install --target baseline/_ref /root/reference`` has nothing to build.  This script makes the smallest possible
package of it: a copy of the tree under /tmp (``/root/reference`` is read-only) plus a generated three-line setup.py that
lists the reference's own packages (``factory``, ``melgan``, ``util``, ``make_data.factory``) -- no source file is
edited -- and then runs exactly the prescribed pip command on that copy (``--no-deps``: wandb / librosa / soundfile are
neither needed by the forward path nor installable offline).  ``bench.py --impl reference`` and the ``cpu_baseline`` leg
import the classes from baseline/_ref; nothing on the product path does."""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TARGET = os.path.join(ROOT, "baseline", "_ref")
REFERENCE = os.environ.get("AUTOFORMER_REFERENCE", "/root/reference")

SETUP = '''from setuptools import setup
setup(name="autoformer-reference", version="0", packages=["factory", "melgan", "util", "make_data", "make_data.factory"])
'''


def installed():
    return os.path.isfile(os.path.join(TARGET, "factory", "AutoVC.py"))


def install(force=False):
    if installed() and not force:
        return TARGET
    if not os.path.isfile(os.path.join(REFERENCE, "factory", "AutoVC.py")):
        raise RuntimeError(f"reference tree not found at {REFERENCE}")
    tmp = tempfile.mkdtemp(prefix="autoformer_ref_")
    try:
        src = os.path.join(tmp, "src")
        shutil.copytree(REFERENCE, src, ignore=shutil.ignore_patterns(".git", "*.ipynb", "__pycache__"))
        with open(os.path.join(src, "setup.py"), "w") as f:
            f.write(SETUP)
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", TARGET, src]
        out = subprocess.run(cmd, capture_output=True, text=True)
        if out.returncode != 0:
            raise RuntimeError("pip install of the reference failed:\n" + out.stdout[-2000:] + out.stderr[-2000:])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    assert installed(), "pip reported success but baseline/_ref/factory/AutoVC.py is missing"
    return TARGET


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))

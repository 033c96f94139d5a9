"""TEST INFRASTRUCTURE ONLY -- import the UNMODIFIED live reference from /root/reference.

Works only where /root/reference exists (the build container).  ``import fasta`` there fails on
``fasta/__init__.py:26 -> plots.py:4`` because matplotlib is not installed, so a two-file stub
``matplotlib`` package is put on ``sys.path`` first (written to a temp dir, never into the repo's
product tree).  Must run in a process that has NOT imported this repo's own ``fasta`` package
(same top-level name): use it from ``oracle/make_golden.py`` or a subprocess.
"""

import importlib
import os
import sys
import tempfile

REFERENCE_ROOT = "/root/reference"


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "fasta"))


def load():
    """Return the live reference ``fasta`` module (with .linalg/.proximal/.stopping)."""
    if not available():
        raise RuntimeError("live reference not present (only exists in the build container)")
    if "fasta" in sys.modules and not getattr(sys.modules["fasta"], "__file__", "").startswith(REFERENCE_ROOT):
        raise RuntimeError("another 'fasta' package is already imported in this process")
    stub = tempfile.mkdtemp(prefix="mpl_stub_")
    os.makedirs(os.path.join(stub, "matplotlib"))
    for name in ("__init__.py", "pyplot.py"):
        with open(os.path.join(stub, "matplotlib", name), "w") as fh:
            fh.write("# stub so that the reference's plots.py imports; never called by the oracle\n")
    sys.path[:0] = [stub, REFERENCE_ROOT]
    mod = importlib.import_module("fasta")
    assert mod.__file__.startswith(REFERENCE_ROOT), mod.__file__
    return mod

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
# `pip install --target baseline/_ref /root/reference` (git-ignored, travels to the GPU box): the same files, unmodified
INSTALLED_ROOT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


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


def _roots():
    return [r for r in (INSTALLED_ROOT, REFERENCE_ROOT) if os.path.isfile(os.path.join(r, "fasta", "__init__.py"))]


def _stub_dir():
    stub = tempfile.mkdtemp(prefix="mpl_stub_")
    os.makedirs(os.path.join(stub, "matplotlib"))
    for name in ("__init__.py", "pyplot.py"):
        with open(os.path.join(stub, "matplotlib", name), "w") as fh:
            fh.write("# stub so that the reference's plots.py imports; never called\n")
    return stub


_isolated = None


def load_isolated():
    """The UNMODIFIED reference package as a module object, importable even in a process that already holds this
    repo's own ``fasta`` package (same top-level name): ``sys.modules`` entries named ``fasta*`` are set aside while
    the reference is imported from ``baseline/_ref`` (or ``/root/reference``) and put back afterwards; the reference's
    modules keep referring to each other through their own globals.  Returns ``(module, root)`` or ``(None, None)``
    where neither copy exists (then bench.py falls back to the oracle port and says so)."""
    global _isolated
    if _isolated is not None:
        return _isolated
    roots = _roots()
    if not roots:
        _isolated = (None, None)
        return _isolated
    root = roots[0]
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "fasta" or k.startswith("fasta.")}
    had_mpl = "matplotlib" in sys.modules
    try:
        import matplotlib  # noqa: F401  (present on some systems: then no stub is needed)
        extra = []
    except Exception:
        extra = [_stub_dir()]
    old_path = list(sys.path)
    try:
        sys.path[:0] = extra + [root]
        mod = importlib.import_module("fasta")
        importlib.import_module("fasta.linalg")
        assert os.path.realpath(mod.__file__).startswith(os.path.realpath(root)), mod.__file__
    finally:
        sys.path[:] = old_path
        for k in [k for k in sys.modules if k == "fasta" or k.startswith("fasta.")]:
            sys.modules.pop(k)
        if extra and not had_mpl:
            for k in [k for k in sys.modules if k == "matplotlib" or k.startswith("matplotlib.")]:
                sys.modules.pop(k)
        sys.modules.update(saved)
    _isolated = (mod, root)
    return _isolated

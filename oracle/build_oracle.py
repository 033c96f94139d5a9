"""TEST INFRASTRUCTURE ONLY -- build the plain-C parts of the oracle with gcc into oracle/_build/ (git-ignored).

    python oracle/build_oracle.py

Called by __graft_entry__.build() and lazily by the tests.  Building the checker is not using it: nothing under
fasta-python_b200/ loads this library.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libfasta_oracle_c.so")
SOURCES = [os.path.join(HERE, "np_legacy_rng.c")]
INCLUDES = [os.path.join(ROOT, "fasta-python_b200", "csrc")]


def _has_fma():
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("flags"):
                    return " fma " in line + " "
    except OSError:
        pass
    return False


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    deps = SOURCES + [os.path.join(INCLUDES[0], f) for f in ("glibc_log.h", "glibc_log_data.h")] + [os.path.abspath(__file__)]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", LIB] + SOURCES + ["-I" + i for i in INCLUDES] + ["-lm"]
    if _has_fma():
        cmd.insert(1, "-mfma")         # fma() inlines to vfmadd; without it libm's (equally exact) fma is called
    subprocess.run(cmd, check=True)
    return LIB


def load():
    import ctypes
    lib = ctypes.CDLL(build())
    lib.fb200_ref_randn.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                    ctypes.POINTER(ctypes.c_double), ctypes.c_int64, ctypes.c_void_p]
    lib.fb200_ref_randn.restype = None
    lib.fb200_ref_log_mismatches.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_double)]
    lib.fb200_ref_log_mismatches.restype = ctypes.c_int64
    lib.fb200_ref_log.argtypes = [ctypes.c_double]
    lib.fb200_ref_log.restype = ctypes.c_double
    return lib


def randn(state, n):
    """n draws continuing the numpy legacy state tuple `state` (np.random.get_state()); returns (values, new_state)."""
    import ctypes
    import numpy as np
    lib = load()
    name, key, pos, has_gauss, gauss = state
    key = np.array(key, dtype=np.uint32)
    cpos, chas, cg = ctypes.c_int(int(pos)), ctypes.c_int(int(has_gauss)), ctypes.c_double(float(gauss))
    out = np.empty(n, dtype=np.float64)
    lib.fb200_ref_randn(key.ctypes.data, ctypes.byref(cpos), ctypes.byref(chas), ctypes.byref(cg), n, out.ctypes.data)
    return out, (name, key, cpos.value, chas.value, cg.value)


if __name__ == "__main__":
    print(build(force=True))

"""Build libfasta_b200.so for sm_100a with nvcc (cross-compiles without a GPU).

    python fasta-python_b200/build.py [--force] [--verbose]

The library is built IN-TREE at fasta-python_b200/lib/libfasta_b200.so (git-ignored, but shipped
to the GPU box by gpurun).  Elementwise / stencil translation units are compiled with
-fmad=false so that every expression rounds exactly once per numpy operation of the reference
line it replaces; the dense streaming kernels keep FMA contraction.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libfasta_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", INCLUDE]
COMMON += os.environ.get("FB200_NVCC_DEFS", "").split()        # developer experiments: extra -D switches

# (source, extra flags)
UNITS = [
    ("api.cu", []),
    ("vector_kernels.cu", ["-fmad=false"]),
    ("tv_stencil.cu", ["-fmad=false"]),
    ("dense_stream.cu", []),
    ("dense_sweep.cu", []),
    ("dense_gsweep.cu", []),
    ("batched_gemm.cu", []),
    ("ozaki_gemm.cu", []),
    ("batched_vector.cu", ["-fmad=false"]),
    ("resident_loop.cu", ["-fmad=false"]),
    ("jacobi_svd.cu", []),
    ("legacy_rng.cu", ["-fmad=false"]),
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(INCLUDE, "fasta_b200.h"))
    headers.append(os.path.abspath(__file__))
    nvcc = _nvcc()
    objs = []
    jobs = []
    for src, extra in UNITS:
        s = os.path.join(CSRC, src)
        o = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            jobs.append([nvcc, "-c", s, "-o", o] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []))
    rebuilt = bool(jobs)
    if jobs:                                    # translation units are independent: compile them side by side
        from concurrent.futures import ThreadPoolExecutor

        def run(cmd):
            if verbose:
                print(" ".join(cmd))
            return subprocess.run(cmd, check=True)
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1, 8)) as pool:
            list(pool.map(run, jobs))
    if rebuilt or force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ARCH + ["-Xcompiler", "-fPIC", "-cudart", "static"]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

"""Developer check of the tcgen05 (int8 digit-plane) fp64 GEMM: plane exactness, GEMM error against an
extended-precision reference, timing against the DMMA kernel and cuBLAS.  Usage: ozaki_check.py [small|big|all]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fasta-python_b200")]
import numpy as np
import torch
from fasta import _cabi, _device
lib = _cabi.load()
dev = "cuda"
pad = lambda n, t: int(lib.fb200_ozaki_pad(n, t))
st = _device.stream_ptr


def slice_rows(P):
    R, C = P.shape
    S = torch.empty(8, pad(R, 128), pad(C, 128), dtype=torch.int8, device=dev)
    sc = torch.empty(pad(R, 128), dtype=torch.float64, device=dev)
    _cabi.check(lib.fb200_ozaki_slice_rows(P.data_ptr(), P.stride(0), R, C, S.data_ptr(), sc.data_ptr(), st()), "slice_rows")
    return S, sc


def slice_cols(P, tile, colmap=None):
    R = P.shape[0]
    n = P.shape[1] if colmap is None else len(colmap)
    S = torch.empty(8, pad(n, tile), pad(R, 128), dtype=torch.int8, device=dev)
    sc = torch.empty(pad(n, tile), dtype=torch.float64, device=dev)
    scratch = torch.empty(n, dtype=torch.int64, device=dev)
    cm = None if colmap is None else torch.as_tensor(colmap, dtype=torch.int32, device=dev)
    _cabi.check(lib.fb200_ozaki_slice_cols(P.data_ptr(), P.stride(0), R, 0 if cm is None else cm.data_ptr(), n, tile, S.data_ptr(),
                                           sc.data_ptr(), scratch.data_ptr(), st()), "slice_cols")
    return S, sc


def rebuild(S, sc):
    """sum_s plane_s 2^(-7s) * scale (row-wise), in fp64 -- exact up to the dropped 2^-56 tail"""
    acc = torch.zeros(S.shape[1:], dtype=torch.float64, device=dev)
    for s in range(7, -1, -1):
        acc = acc / 128.0 + S[s].double()
    return acc * sc[:, None]


def gemm(LS, lsc, Mg, RS, rsc, Ng, K, out_cols=None, colmap=None):
    S = int(lib.fb200_ozaki_splits(Mg, Ng, K))
    nc = Ng if out_cols is None else out_cols
    C = torch.zeros(S, Mg, nc, dtype=torch.float64, device=dev)
    cm = None if colmap is None else torch.as_tensor(colmap, dtype=torch.int32, device=dev)
    _cabi.check(lib.fb200_ozaki_gemm(LS.data_ptr(), lsc.data_ptr(), Mg, RS.data_ptr(), rsc.data_ptr(), Ng, K, C.data_ptr(), nc,
                                     0 if cm is None else cm.data_ptr(), S, Mg * nc, st()), "ozaki_gemm")
    return C.sum(0) if S > 1 else C[0]


def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def ref_ld(A, X):
    return (A.cpu().numpy().astype(np.longdouble) @ X.cpu().numpy().astype(np.longdouble)).astype(np.float64)


def small():
    g = torch.Generator(device="cpu").manual_seed(0)
    for (M, N, B) in [(128, 128, 64), (200, 1000, 8), (130, 258, 6), (333, 1414, 70), (1000, 2000, 64)]:
        A = torch.randn(M, N, generator=g, dtype=torch.float64).to(dev) * torch.logspace(-3, 3, M, dtype=torch.float64, device=dev)[:, None]
        X = torch.randn(N, B, generator=g, dtype=torch.float64).to(dev)
        X[torch.rand(N, B, generator=g).to(dev) < 0.7] = 0.0           # sparse iterates
        X[:, 0] = 0.0                                                   # an all-zero column
        Rm = torch.randn(M, B, generator=g, dtype=torch.float64).to(dev) * 1e-5
        rec = dict(M=M, N=N, B=B)
        AF, af = slice_rows(A)
        rec["planes_fwd_err"] = float(((rebuild(AF, af)[:M, :N] - A).abs() / af[:M, None]).max())     # in units of the row scale: <= 2^-56 * 64
        AT, at = slice_cols(A, 128)
        rec["planes_adj_err"] = float(((rebuild(AT, at)[:N, :M] - A.t()).abs() / at[:N, None].clamp_min(1e-300)).max())
        XS, xs = slice_cols(X, 64)
        rec["planes_x_err"] = float(((rebuild(XS, xs)[:B, :N] - X.t()).abs() / xs[:B, None].clamp_min(1e-300)).max())
        assert int(AF.abs().max()) <= 64 and int(XS.abs().max()) <= 64
        Z = gemm(AF, af, M, XS, xs, B, N)
        Zr = torch.from_numpy(ref_ld(A, X)).to(dev)
        rec["fwd_err"] = float((Z - Zr).norm() / Zr.norm())
        rec["fwd_err_cublas"] = float((A @ X - Zr).norm() / Zr.norm())
        RS, rs = slice_cols(Rm, 64)
        G = gemm(AT, at, N, RS, rs, B, M)
        Gr = torch.from_numpy(ref_ld(A.t(), Rm)).to(dev)
        rec["adj_err"] = float((G - Gr).norm() / Gr.norm())
        rec["adj_err_cublas"] = float((A.t() @ Rm - Gr).norm() / Gr.norm())
        # compacted columns scattered into a wider output
        cols = [c for c in range(B) if c % 3 != 1]
        XSc, xsc = slice_cols(X, 64, cols)
        Zc = gemm(AF, af, M, XSc, xsc, len(cols), N, out_cols=B, colmap=cols)
        rec["compact_err"] = float((Zc[:, cols] - Zr[:, cols]).norm() / Zr[:, cols].norm())
        rec["compact_untouched"] = float(Zc[:, [c for c in range(B) if c % 3 == 1]].abs().max()) if B > 1 else 0.0
        print(json.dumps(rec), flush=True)


def big():
    from tools_gemm import dmma_gemm
    for (M, N, B) in [(4000, 10000, 256), (20000, 50000, 256)]:
        A = torch.randn(M, N, dtype=torch.float64, device=dev)
        X = torch.randn(N, B, dtype=torch.float64, device=dev)
        Rm = torch.randn(M, B, dtype=torch.float64, device=dev)
        rec = dict(M=M, N=N, B=B)
        rec["slice_rows_ms"] = timeit(lambda: slice_rows(A), 1)
        AF, af = slice_rows(A)
        rec["slice_colsA_ms"] = timeit(lambda: slice_cols(A, 128), 1)
        AT, at = slice_cols(A, 128)
        XS, xs = slice_cols(X, 64)
        RS, rs = slice_cols(Rm, 64)
        rec["slice_x_ms"] = timeit(lambda: slice_cols(X, 64))
        rec["slice_r_ms"] = timeit(lambda: slice_cols(Rm, 64))
        Z = gemm(AF, af, M, XS, xs, B, N)
        G = gemm(AT, at, N, RS, rs, B, M)
        Zr, Gr = A @ X, A.t() @ Rm
        rec["fwd_vs_cublas"] = float((Z - Zr).norm() / Zr.norm())
        rec["adj_vs_cublas"] = float((G - Gr).norm() / Gr.norm())
        fl = 2.0 * M * N * B
        t = timeit(lambda: gemm(AF, af, M, XS, xs, B, N)); rec.update(fwd_ms=t, fwd_tflops_fp64_equiv=fl / t / 1e9, fwd_int8_tops=36 * fl / t / 1e9)
        t = timeit(lambda: gemm(AT, at, N, RS, rs, B, M)); rec.update(adj_ms=t, adj_tflops_fp64_equiv=fl / t / 1e9, adj_int8_tops=36 * fl / t / 1e9)
        t = timeit(lambda: dmma_gemm(0, A, X)); rec.update(dmma_fwd_ms=t)
        t = timeit(lambda: dmma_gemm(1, A, Rm)); rec.update(dmma_adj_ms=t)
        t = timeit(lambda: torch.matmul(A, X)); rec.update(cublas_fwd_ms=t)
        t = timeit(lambda: torch.matmul(A.t(), Rm)); rec.update(cublas_adj_ms=t)
        rec["splits"] = [int(lib.fb200_ozaki_splits(M, B, N)), int(lib.fb200_ozaki_splits(N, B, M))]
        print(json.dumps(rec), flush=True)
        del A, X, Rm, AF, AT


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("small", "all"):
        small()
    if what in ("big", "all"):
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        big()
    if what == "prof":          # one forward + one adjoint product at config 5 size (for ncu)
        M, N, B = 20000, 50000, 256
        A = torch.randn(M, N, dtype=torch.float64, device=dev)
        X = torch.randn(N, B, dtype=torch.float64, device=dev)
        AF, af = slice_rows(A)
        XS, xs = slice_cols(X, 64)
        for _ in range(2):
            Z = gemm(AF, af, M, XS, xs, B, N)
        torch.cuda.synchronize()
        print("ok", float(Z.norm()))

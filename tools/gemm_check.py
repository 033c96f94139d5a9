"""Developer check of the fp64 batched GEMM kernel against torch.matmul (cuBLAS) incl. timing."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fasta-python_b200")]
import torch
from fasta import _cabi, _device
lib = _cabi.load()

def gemm(adj, A, Bm):
    K = A.shape[0] if adj else A.shape[1]
    Mg = A.shape[1] if adj else A.shape[0]
    Ng = Bm.shape[1]
    S = lib.fb200_gemm_splits(Mg, Ng, K)
    C = torch.empty(S, Mg, Ng, dtype=torch.float64, device="cuda")
    _cabi.check(lib.fb200_gemm_f64(adj, A.data_ptr(), A.stride(0), Bm.data_ptr(), Bm.stride(0), C.data_ptr(), C.stride(1), Mg, Ng, K, S, Mg * Ng, _device.stream_ptr()))
    return C.sum(0) if S > 1 else C[0]

def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for (M, N, B) in [(200, 1000, 8), (130, 258, 6), (1000, 2000, 64), (4000, 10000, 256), (20000, 50000, 256)]:
    A = torch.randn(M, N, dtype=torch.float64, device="cuda")
    X = torch.randn(N, B, dtype=torch.float64, device="cuda")
    R = torch.randn(M, B, dtype=torch.float64, device="cuda")
    Z, G = gemm(0, A, X), gemm(1, A, R)
    Zr, Gr = A @ X, A.T @ R
    rec = dict(M=M, N=N, B=B, fwd_err=float((Z - Zr).norm() / Zr.norm()), adj_err=float((G - Gr).norm() / Gr.norm()))
    if M >= 4000:
        fl = 2.0 * M * N * B
        t = timeit(lambda: gemm(0, A, X)); rec.update(fwd_ms=t, fwd_tflops=fl / t / 1e9)
        t = timeit(lambda: gemm(1, A, R)); rec.update(adj_ms=t, adj_tflops=fl / t / 1e9)
        t = timeit(lambda: torch.matmul(A, X)); rec.update(cublas_fwd_ms=t, cublas_fwd_tflops=fl / t / 1e9)
        t = timeit(lambda: torch.matmul(A.T, R)); rec.update(cublas_adj_ms=t, cublas_adj_tflops=fl / t / 1e9)
    print(json.dumps(rec), flush=True)

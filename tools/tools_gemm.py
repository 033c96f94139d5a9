"""Shared helper of the developer tools: call the fp64 DMMA GEMM (fb200_gemm_f64) on torch tensors."""
import torch
from fasta import _cabi, _device


def dmma_gemm(adj, A, Bm):
    lib = _cabi.load()
    K = A.shape[0] if adj else A.shape[1]
    Mg = A.shape[1] if adj else A.shape[0]
    Ng = Bm.shape[1]
    S = lib.fb200_gemm_splits(Mg, Ng, K)
    C = torch.empty(S, Mg, Ng, dtype=torch.float64, device="cuda")
    _cabi.check(lib.fb200_gemm_f64(adj, A.data_ptr(), A.stride(0), Bm.data_ptr(), Bm.stride(0), C.data_ptr(), C.stride(1), Mg, Ng, K, S,
                                   Mg * Ng, _device.stream_ptr()))
    return C.sum(0) if S > 1 else C[0]

"""SUPERSEDED by `bench.py --workload ...` (round 2), which measures the same configs with the contract keys, the
unmodified reference as CPU arm and full-size parity; kept as a developer tool.

Measurement of BASELINE.json config 5: batched lasso regularisation path, 256 lambdas x M=20000 N=50000,
the two contractions per iteration as tensor-core GEMMs (tcgen05 int8 digit planes, fp64 accuracy).

    python tools/bench_batched.py [--M 20000 --N 50000 --B 256] [--gemm ozaki|dmma] [--max-iters K] [--check-cols 2]

One JSON line: column-iterations/s, lock-step iterations, time, the GEMM share, fp64-equivalent TFLOP/s and
int8 tensor-pipe utilisation; parity of sampled columns against the single-problem CUDA path (fasta.fasta,
itself parity-tested against the reference) and a CPU arm (numpy oracle on one column, bounded).
Not the driver's contract bench (bench.py, config 2); results are committed under profiles/.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fasta-python_b200")]

import numpy as np
import torch

INT8_DENSE_PEAK_TOPS = 4500.0        # B200 nominal dense int8 (no measured figure in MEASURED_PEAKS.json)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--M", type=int, default=20000)
    ap.add_argument("--N", type=int, default=50000)
    ap.add_argument("--B", type=int, default=256)
    ap.add_argument("--gemm", default="ozaki")
    ap.add_argument("--max-iters", type=int, default=1000)
    ap.add_argument("--check-cols", type=int, default=2)
    ap.add_argument("--cpu-iters", type=int, default=3)
    a = ap.parse_args()
    os.environ["FASTA_B200_GEMM"] = a.gemm
    import fasta
    from fasta import batched
    from oracle import fasta_oracle

    M, N, B = a.M, a.N, a.B
    K = N // 20
    gen = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn(M, N, dtype=torch.float64, device="cuda", generator=gen)
    A /= (np.sqrt(M) + np.sqrt(N))                              # SURVEY 8d scaling recipe (config 2 / 5)
    xt = torch.zeros(N, dtype=torch.float64, device="cuda")
    xt[torch.randperm(N, generator=torch.Generator().manual_seed(5))[:K].cuda()] = 1.0
    b = torch.mv(A, xt) + 0.01 * torch.randn(M, dtype=torch.float64, device="cuda", generator=gen)
    lam_max = float(torch.mv(A.t(), b).abs().max())
    mus = lam_max * np.logspace(-3, 0, B)
    op = fasta.linalg.LinearMap.from_matrix(A)
    opts = dict(adaptive=True, verbose=False, tolerance=1e-5, max_iters=a.max_iters, evaluate_objective=True)

    def run():
        np.random.seed(0)
        return batched.lasso_path(op, b, mus, **opts)

    batched.GEMM_EVENTS = None
    torch.cuda.synchronize()
    t0 = time.time()
    run()                                                        # warm-up (also pages in the kernels)
    torch.cuda.synchronize()
    warm_s = time.time() - t0
    batched.GEMM_EVENTS = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ev = batched.GEMM_EVENTS
    batched.GEMM_EVENTS = None
    gemm_ms = sum(x.elapsed_time(y) for x, y, _, _ in ev)
    pad64 = lambda n: (n + 63) // 64 * 64
    flops = sum(2.0 * M * N * n for _, _, _, n in ev)                        # useful fp64-equivalent flops (active columns)
    int8_ops = sum(36 * 2.0 * M * N * pad64(n) for _, _, _, n in ev)          # what the tensor pipe executed (padded tiles, 36 digit pairs)
    full = [(x.elapsed_time(y), adj) for x, y, adj, n in ev if n == B]
    col_iters = int(sum(r.iteration_count for r in out))
    meta = out[0].batch
    line = dict(metric="batched_lasso_path_column_iterations_per_sec", value=col_iters / (ms / 1e3), unit="column-iterations/s",
                n_gpus=1, dtype="f64", data="synthetic",
                config=dict(workload=f"config 5: lasso regularisation path, {B} lambdas log-spaced in [1e-3,1]*|A^T b|_inf, M={M} N={N} "
                                     f"fp64 ({M * N * 8 / 1e9:.1f} GB A), adaptive, tol 1e-5, lock-step batched solve"),
                gemm=meta["gemm"], ms_total=ms, warmup_s=warm_s, lockstep_iterations=meta["iterations_lockstep"],
                column_iterations=col_iters, iterations_min_max=[int(min(r.iteration_count for r in out)), int(max(r.iteration_count for r in out))],
                backtracks_total=int(sum(r.backtracks for r in out)), gemm_calls=len(ev), gemm_ms=gemm_ms, gemm_share=gemm_ms / ms,
                gemm_fp64_equiv_tflops=flops / gemm_ms / 1e9,
                full_width_gemm_ms=dict(forward=float(np.mean([t for t, adj in full if not adj])) if any(not adj for _, adj in full) else None,
                                        adjoint=float(np.mean([t for t, adj in full if adj])) if any(adj for _, adj in full) else None),
                gpu_launches=meta["kernel_launches"])
    if meta["gemm"].startswith("tcgen05"):
        line["roofline"] = dict(bound="tensor", achieved=int8_ops / gemm_ms / 1e9, peak=INT8_DENSE_PEAK_TOPS, unit="TOP/s (int8)",
                                frac=int8_ops / gemm_ms / 1e9 / INT8_DENSE_PEAK_TOPS, peak_source="nominal B200 dense int8",
                                note="36 int8 digit-pair GEMMs per fp64 product, slicing kernels included in the timed GEMM calls")
    else:
        line["roofline"] = dict(bound="tensor", achieved=flops / gemm_ms / 1e9, peak=40.0, unit="TFLOP/s (fp64 DMMA)",
                                frac=flops / gemm_ms / 1e9 / 40.0, peak_source="nominal B200 fp64 tensor")
    # parity of sampled columns against the single-problem CUDA path
    par = []
    for j in np.linspace(B - 1, B // 3, a.check_cols).astype(int):
        loss, pen = fasta.losses.LeastSquares(b), fasta.proximal.L1Norm(float(mus[j]))
        np.random.seed(0)
        one = fasta.fasta(op, loss.f, loss.gradf, pen.g, pen.prox, torch.zeros(N, dtype=torch.float64, device="cuda"),
                          **dict(opts, accelerate=False))
        r = out[j]
        n = one.iteration_count
        so, sb = one.solution.cpu().numpy() if hasattr(one.solution, "cpu") else one.solution, r.solution.cpu().numpy() if hasattr(r.solution, "cpu") else r.solution
        par.append(dict(column=int(j), mu_over_lam_max=float(mus[j] / lam_max), iterations=[int(r.iteration_count), int(n)],
                        backtracks=[int(r.backtracks), int(one.backtracks)],
                        solution_rel_err=float(np.linalg.norm(sb - so) / max(np.linalg.norm(so), 1e-300)),
                        objective_rel_err=float(np.max(np.abs(r.objectives[:n + 1] - one.objectives[:n + 1]) / np.abs(one.objectives[:n + 1])))
                        if r.iteration_count == n else None, nnz=int(np.count_nonzero(sb))))
    line["parity_vs_single_problem_cuda_path"] = par
    if a.cpu_iters > 0:
        An, bn = A.cpu().numpy(), b.cpu().numpy()
        mu = float(mus[B // 2])
        f = lambda z: .5 * np.linalg.norm((z - bn).ravel()) ** 2
        gradf = lambda z: z - bn
        g = lambda x: mu * np.linalg.norm(x.ravel(), 1)
        proxg = lambda x, t: fasta_oracle.shrink(x, t * mu)
        np.random.seed(0)
        ref = fasta_oracle.solve(lambda x: An @ x, lambda y: An.T @ y, f, gradf, g, proxg, np.zeros(N),
                                 **dict(opts, accelerate=False, max_iters=a.cpu_iters))
        loop = ref.times[ref.iteration_count] - ref.times[0]
        line["cpu_baseline"] = dict(value=ref.iteration_count / loop, unit="column-iterations/s", cores=len(os.sched_getaffinity(0)), kind="port",
                                    sample=f"oracle/fasta_oracle.py (numpy {np.__version__}) ONE column (mu index {B // 2}) of the same problem, "
                                           f"first {ref.iteration_count} iterations, in-loop {loop:.2f} s; the reference solves columns one at a time")
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()

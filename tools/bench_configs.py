"""SUPERSEDED by `bench.py --workload ...` (round 2), which measures the same configs with the contract keys, the
unmodified reference as CPU arm and full-size parity; kept as a developer tool.

Measurements of the other BASELINE.json configs (3: sparse logistic, 4: TV denoising; 1: the
reference's own small lasso) in the style of bench.py: one JSON line per config with iterations/s,
time-to-tolerance, achieved algorithmic GB/s and the CPU arm (numpy oracle) beside it.

    python tools/bench_configs.py [logistic] [tv] [small] [--steps K]

Not the driver's contract bench (that is bench.py, config 2); results are committed under profiles/.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fasta-python_b200")]

import numpy as np
import torch

import fasta
from oracle import fasta_oracle, problems

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timed_solves(solve, steps, warmup=3):
    for _ in range(warmup):
        solve()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    outs = [solve() for _ in range(steps)]
    e1.record()
    torch.cuda.synchronize()
    return outs, e0.elapsed_time(e1)


def summarise(name, workload, outs, ms, steps, bytes_iter, bytes_bt, cpu, extra=None):
    iters = sum(r.iteration_count for r in outs)
    bts = sum(r.backtracks for r in outs)
    loop = sum(r.times[r.iteration_count] - r.times[0] for r in outs)
    alg = iters * bytes_iter + bts * bytes_bt
    line = dict(metric=f"{name}_fbs_iterations_per_sec", value=iters / (ms / 1e3), unit="iterations/s", n_gpus=1,
                steps=steps, ms_per_step=ms / steps, dtype="f64", data="synthetic", config=dict(workload=workload),
                iterations_per_solve=iters / steps, backtracks_per_solve=bts / steps,
                time_to_tol_ms=1e3 * loop / steps, iters_per_sec_in_loop=iters / loop,
                roofline=dict(bound="hbm", achieved=alg / loop / 1e9, peak=PEAK, unit="GB/s", frac=alg / loop / 1e9 / PEAK,
                              frac_of_nominal_8TBs=alg / loop / 1e9 / 8000.0,
                              algorithmic_bytes_per_iteration=bytes_iter, note="whole in-loop time, all kernels + host syncs"),
                cpu_baseline=cpu, backend=outs[-1].backend, single_pass=outs[-1].single_pass,
                gpu_launches=sum(r.kernel_launches for r in outs),
                speculation=getattr(outs[-1], "speculation", None),
                resident=bool(getattr(outs[-1], "resident", False)), resident_cluster=bool(getattr(outs[-1], "resident_cluster", False)))
    if extra:
        line.update(extra)
    print(json.dumps(line), flush=True)


def cpu_sample(op, adj, f, gradf, g, proxg, x0, iters, opts, what):
    np.random.seed(0)
    o = dict(opts)
    o["max_iters"] = iters
    t0 = time.time()
    res = fasta_oracle.solve(op, adj, f, gradf, g, proxg, x0, **o)
    wall = time.time() - t0
    n = res.iteration_count
    loop = res.times[n] - res.times[0]
    return dict(value=n / loop, unit="iterations/s", cores=len(os.sched_getaffinity(0)), kind="port",
                sample=f"oracle/fasta_oracle.py (numpy {np.__version__}) {what}: first {n} iterations, in-loop {loop:.2f} s, "
                       f"whole call {wall:.2f} s")


def bench_logistic(steps):
    """Config 3: M=100000, N=20000, K=50, mu = 40*sqrt(M/1000) = 400, adaptive BB."""
    M, N, K, mu = 100000, 20000, 50, 400.0
    gen = torch.Generator(device="cuda").manual_seed(3)
    A = torch.randn(M, N, dtype=torch.float64, device="cuda", generator=gen)
    xt = torch.zeros(N, dtype=torch.float64, device="cuda")
    xt[torch.randperm(N, generator=torch.Generator().manual_seed(3))[:K].cuda()] = 1.0
    p = torch.sigmoid(torch.mv(A, xt))
    b = 2.0 * (torch.rand(M, dtype=torch.float64, device="cuda", generator=gen) < p).double() - 1.0
    x0 = torch.zeros(N, dtype=torch.float64, device="cuda")
    op, loss, pen = fasta.linalg.LinearMap.from_matrix(A), fasta.losses.Logistic(b), fasta.proximal.L1Norm(mu)
    opts = dict(adaptive=True, accelerate=False, verbose=False, tolerance=1e-5, evaluate_objective=True)

    def solve():
        np.random.seed(0)
        return fasta.fasta(op, loss.f, loss.gradf, pen.g, pen.prox, x0, **opts)

    outs, ms = timed_solves(solve, steps)
    An, bn = A.cpu().numpy(), b.cpu().numpy()
    f = lambda z: np.sum(np.log(1 + np.exp(z)) - (bn == 1) * z)
    gradf = lambda z: -bn / (1 + np.exp(bn * z))
    g = lambda x: mu * np.linalg.norm(x.ravel(), 1)
    proxg = lambda x, t: fasta_oracle.shrink(x, t * mu)
    np.random.seed(0)
    ref = fasta_oracle.solve(lambda x: An @ x, lambda y: An.T @ y, f, gradf, g, proxg, np.zeros(N), **opts)
    n = ref.iteration_count
    sol = outs[-1].solution.cpu().numpy()
    parity = dict(ref_iterations=n, ref_backtracks=ref.backtracks, iterations=outs[-1].iteration_count,
                  backtracks=outs[-1].backtracks,
                  solution_rel_err=float(np.linalg.norm(sol - ref.solution) / max(np.linalg.norm(ref.solution), 1e-300)),
                  objective_rel_err=float(np.max(np.abs(outs[-1].objectives[:n + 1] - ref.objectives[:n + 1]) /
                                                 np.abs(ref.objectives[:n + 1]))) if outs[-1].iteration_count == n else None,
                  nnz=int(np.count_nonzero(sol)))
    loop = ref.times[n] - ref.times[0]
    cpu = dict(value=n / loop, unit="iterations/s", cores=len(os.sched_getaffinity(0)), kind="port",
               sample=f"oracle/fasta_oracle.py (numpy {np.__version__}) full solve of the same problem: {n} iterations, in-loop {loop:.2f} s")
    summarise("logistic", f"sparse logistic regression M={M} N={N} K={K} mu={mu} fp64 (16 GB A), adaptive, tol 1e-5",
              outs, ms, steps, 2 * M * N * 8, M * N * 8, cpu, dict(parity_vs_oracle_full_size=parity))


def bench_tv(steps, n=4096):
    """Config 4: n=4096 checkerboard + 0.1*randn, mu=0.1, matrix-free div/grad stencils."""
    rng = np.random.RandomState(0)
    img = problems.checkerboard(n, 64)
    img /= img.max()
    img += 0.1 * rng.randn(n, n)
    mu = 0.1
    b = torch.from_numpy(img / mu).cuda()
    Y0 = torch.zeros(n, n, 2, dtype=torch.float64, device="cuda")
    op, loss, pen = fasta.tv.divergence_map((n, n)), fasta.losses.LeastSquares(b), fasta.proximal.TVBall()
    U = n * n * 8
    for mode, mopts in (("accelerated", dict(adaptive=False, accelerate=True)), ("adaptive", dict(adaptive=True, accelerate=False))):
        opts = dict(verbose=False, tolerance=1e-5, max_iters=300, **mopts)

        def solve():
            np.random.seed(0)
            return fasta.fasta(op, loss.f, loss.gradf, pen.g, pen.prox, Y0, **opts)

        outs, ms = timed_solves(solve, steps, warmup=2)
        bn = img / mu
        f = lambda Z: .5 * np.linalg.norm((Z - bn).ravel()) ** 2
        gradf = lambda Z: Z - bn
        cpu = cpu_sample(problems.tv_div, problems.tv_grad, f, gradf, lambda Y: 0, problems._tv_ball, np.zeros((n, n, 2)), 3,
                         dict(opts), f"TV {n}x{n} {mode}")
        # per iteration: read x0(2U)+g0(2U), write xhat(2U)+x1(2U)+dx(2U) | read x1(2U)+b(U), write z(U)+r(U) |
        # read r(U), write g(2U) [+ BB: read xhat,x0,dx 6U]; SURVEY 8d counts the 11U a minimal fusion needs
        summarise("tv_" + mode, f"TV denoising {n}x{n} (dual, matrix-free div/grad stencils) fp64, {mode}, tol 1e-5, max 300 iterations",
                  outs, ms, steps, 11 * U, 7 * U, cpu)


def bench_small(steps):
    """Config 1: the reference's own CPU-sized lasso (M=200, N=1000): launch-latency bound on a GPU."""
    p = problems.build("lasso_200x1000_k10", 0)
    op = fasta.linalg.LinearMap.from_matrix(p.A)
    loss, pen = fasta.losses.LeastSquares(p.b), fasta.proximal.L1Norm(p.mu)
    x0 = torch.zeros(1000, dtype=torch.float64, device="cuda")
    for mode in problems.MODES:
        opts = dict(problems.HARNESS_OPTS, **problems.MODES[mode])

        def solve():
            np.random.seed(0)
            return fasta.fasta(op, loss.f, loss.gradf, pen.g, pen.prox, x0, **opts)

        outs, ms = timed_solves(solve, steps)
        f, gradf, g, proxg = problems.numpy_callables(p)
        o, a, _, _ = problems.numpy_operator(p)
        cpu = cpu_sample(o, a, f, gradf, g, proxg, p.x0, 1000, opts, f"config 1 {mode}")
        summarise("small_lasso_" + mode, f"config 1 lasso M=200 N=1000 K=10 {mode}", outs, ms, steps, 2 * 200 * 1000 * 8,
                  200 * 1000 * 8, cpu)


def bench_batched(steps, M=20000, N=50000, B=256):
    """Config 5: 256 lambdas x (M=20000, N=50000): lock-step FBS with fp64 GEMM contractions."""
    gen = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn(M, N, dtype=torch.float64, device="cuda", generator=gen)
    A /= (np.sqrt(M) + np.sqrt(N))
    xt = torch.zeros(N, dtype=torch.float64, device="cuda")
    xt[torch.randperm(N, generator=torch.Generator().manual_seed(5))[:N // 20].cuda()] = 1.0
    b = torch.mv(A, xt) + 0.01 * torch.randn(M, dtype=torch.float64, device="cuda", generator=gen)
    lam_max = float(torch.mv(A.t(), b).abs().max())
    mus = lam_max * np.logspace(-3, 0, B)
    op = fasta.linalg.LinearMap.from_matrix(A)
    opts = dict(adaptive=True, tolerance=1e-5, evaluate_objective=True, max_iters=200)

    def solve():
        np.random.seed(0)
        return fasta.batched.lasso_path(op, b, mus, x0=torch.zeros(N, B, dtype=torch.float64, device="cuda"), **opts)

    solve()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    outs = [solve() for _ in range(steps)]
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    res = outs[-1]
    lock = res[0].batch["iterations_lockstep"]
    col_iters = sum(r.iteration_count for r in res)
    bts = sum(r.backtracks for r in res)
    # GEMM throughput alone, and the fp64 library GEMM beside it
    X = torch.randn(N, B, dtype=torch.float64, device="cuda")
    from fasta import _cabi, _device
    lib = _cabi.load()
    S = lib.fb200_gemm_splits(M, B, N)
    C = torch.empty(S, M, B, dtype=torch.float64, device="cuda")

    def tm(fn, n=5):
        fn(); torch.cuda.synchronize()
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        c.record(); torch.cuda.synchronize()
        return a.elapsed_time(c) / n

    t_ours = tm(lambda: lib.fb200_gemm_f64(0, A.data_ptr(), N, X.data_ptr(), B, C.data_ptr(), B, M, B, N, S, M * B, _device.stream_ptr()))
    t_lib = tm(lambda: torch.matmul(A, X))
    flop = 2.0 * M * N * B
    # CPU arm: the numpy oracle on ONE column of the same path (columns are independent runs)
    An, bn = A.cpu().numpy(), b.cpu().numpy()
    mu = float(mus[B // 2])
    f = lambda z: .5 * np.linalg.norm((z - bn).ravel()) ** 2
    gradf = lambda z: z - bn
    g = lambda x: mu * np.linalg.norm(x.ravel(), 1)
    proxg = lambda x, t: fasta_oracle.shrink(x, t * mu)
    cpu = cpu_sample(lambda x: An @ x, lambda y: An.T @ y, f, gradf, g, proxg, np.zeros(N), 5, dict(opts), "one column of the path")
    mid = res[B // 2]
    line = dict(metric="batched_lasso_path_column_iterations_per_sec", value=col_iters * steps / (ms / 1e3) / steps,
                unit="column-iterations/s", n_gpus=1, steps=steps, ms_per_step=ms / steps, dtype="f64", data="synthetic",
                config=dict(workload=f"lasso regularisation path: {B} lambdas x (M={M}, N={N}) fp64, adaptive, tol 1e-5, lock-step GEMM iterations"),
                lockstep_iterations=lock, column_iterations=col_iters, backtracks=bts,
                iterations_per_column=dict(min=min(r.iteration_count for r in res), max=max(r.iteration_count for r in res)),
                roofline=dict(bound="tensor", unit="TFLOP/s", achieved=flop / t_ours / 1e9, peak=flop / t_lib / 1e9,
                              frac=t_lib / t_ours, peak_source="fp64 cuBLAS DGEMM (torch.matmul) of the same shape, measured in this run; "
                              "tcgen05 has no f64 kind, the kernel is DMMA mma.sync.m8n8k4", gemm_ms=t_ours, library_gemm_ms=t_lib,
                              whole_loop_TFLOPs=(2 * lock + 0) * flop / (ms / steps / 1e3) / 1e12),
                cpu_baseline=dict(cpu, note="per-column rate of ONE independent reference-style run; the path has %d columns" % B),
                kernel_launches=res[0].batch["kernel_launches"])
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 3
    which = args or ["small", "logistic", "tv"]
    if "small" in which:
        bench_small(max(steps, 5))
    if "logistic" in which:
        bench_logistic(steps)
    if "tv" in which:
        bench_tv(steps)
    if "batched" in which:
        bench_batched(max(1, min(steps, 2)))

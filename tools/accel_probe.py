"""Developer probe: accelerated (FISTA) mode on the config-2 lasso, fused single-pass sweep vs separate kernels."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fasta-python_b200")]
import numpy as np, torch
import fasta, bench
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "lasso_40000x100000"]
dev = torch.device("cuda", 0)
A, b, _ = bench.make_local_problem(w, 0, 1, dev)
x0 = torch.zeros(w["N"], dtype=torch.float64, device=dev)
op, loss, pen = fasta.linalg.LinearMap.from_matrix(A), fasta.losses.LeastSquares(b), fasta.proximal.L1Norm(w["mu"])
opts = dict(bench.SOLVER_OPTS, adaptive=False, accelerate=True)
for fused in ("1", "0"):
    os.environ["FASTA_B200_SWEEP_ACCEL"] = fused
    for rep in range(3):
        np.random.seed(0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fasta.fasta(op, loss.f, loss.gradf, pen.g, pen.prox, x0, **opts)
        e1.record(); torch.cuda.synchronize()
    loop = r.times[r.iteration_count] - r.times[0]
    print(json.dumps(dict(fused_fista_sweep=fused == "1", single_pass=r.single_pass, iterations=r.iteration_count, backtracks=r.backtracks,
                          ms_total=e0.elapsed_time(e1), loop_ms=1e3 * loop, iters_per_sec_in_loop=r.iteration_count / loop,
                          ms_per_iteration=1e3 * loop / r.iteration_count, final_residual=float(r.residuals[r.iteration_count - 1]))), flush=True)

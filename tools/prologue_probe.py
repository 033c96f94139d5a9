"""Developer probe: where the non-loop time of one config-2 solve goes (host RNG, Lipschitz sweeps, start,
teardown), with and without the nvidia-smi clock sampler of bench.py running beside it."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fasta-python_b200")]
import numpy as np, torch
import fasta, bench
from fasta import _backends
w = bench.WORKLOADS["lasso_40000x100000"]
dev = torch.device("cuda", 0)
A, b, _ = bench.make_local_problem(w, 0, 1, dev)
x0 = torch.zeros(w["N"], dtype=torch.float64, device=dev)
op, loss, pen = fasta.linalg.LinearMap.from_matrix(A), fasta.losses.LeastSquares(b), fasta.proximal.L1Norm(w["mu"])
marks = []
def wrap(cls, name):
    orig = getattr(cls, name)
    def f(self, *a, **k):
        t0 = time.perf_counter()
        out = orig(self, *a, **k)
        marks.append((name, time.perf_counter() - t0))
        return out
    setattr(cls, name, f)
for n in ("load", "lipschitz", "start", "solution", "close", "__init__"):
    wrap(_backends.FusedBackend, n)
orig_randn = np.random.randn
def randn(*a):
    t0 = time.perf_counter(); out = orig_randn(*a); marks.append(("randn", time.perf_counter() - t0)); return out
np.random.randn = randn
def solve():
    np.random.seed(0)
    return fasta.fasta(op, loss.f, loss.gradf, pen.g, pen.prox, x0, **bench.SOLVER_OPTS)
for _ in range(3): solve()
for sampler_on in (False, True, False):
    p = f = None
    if sampler_on: p, f = bench.clocks_sampler()
    torch.cuda.synchronize()
    for rep in range(4):
        marks.clear()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        r = solve()
        e1.record(); torch.cuda.synchronize(); tot = time.perf_counter() - t0
        loop = r.times[r.iteration_count] - r.times[0]
        print(f"sampler={sampler_on} rep {rep}: wall {tot*1e3:.1f} ms, events {e0.elapsed_time(e1):.1f} ms, loop {loop*1e3:.1f} ms ({r.iteration_count} it), other {1e3*(tot-loop):.1f} ms :: " +
              ", ".join(f"{n} {1e3*t:.2f}" for n, t in marks), flush=True)
    if p is not None: p.terminate()

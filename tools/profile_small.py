"""Developer tool: cProfile of repeated config-1 solves (where does the host time of a 0.66 ms solve go?)."""
import cProfile
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fasta-python_b200")]
import numpy as np
import torch
import fasta

np.random.seed(0)
M, N, K = 200, 1000, 10
x = np.zeros(N); x[np.random.permutation(N)[:K]] = 1
A = np.random.randn(M, N); A /= np.linalg.norm(A, 2)
b = A @ x + 0.01 * np.random.randn(M)
Ad, bd = torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda()
x0 = torch.zeros(N, dtype=torch.float64, device="cuda")
op, loss, pen = fasta.linalg.LinearMap.from_matrix(Ad), fasta.losses.LeastSquares(bd), fasta.proximal.L1Norm(0.02)
opts = dict(adaptive=True, accelerate=False, verbose=False, tolerance=1e-5, max_iters=1000, evaluate_objective=True)


def solve():
    np.random.seed(0)
    return fasta.fasta(op, loss.f, loss.gradf, pen.g, pen.prox, x0, **opts)


for _ in range(50):
    solve()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(300):
    solve()
torch.cuda.synchronize()
print(f"{(time.perf_counter() - t0) / 300 * 1e3:.3f} ms per solve")
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    solve()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)

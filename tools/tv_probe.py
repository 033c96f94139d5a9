import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fasta-python_b200")]
import numpy as np, torch, fasta
from oracle import problems
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 12
rng = np.random.RandomState(0)
img = problems.checkerboard(n, 64); img /= img.max(); img += 0.1 * rng.randn(n, n)
b = torch.from_numpy(img / 0.1).cuda()
Y0 = torch.zeros(n, n, 2, dtype=torch.float64, device="cuda")
op, loss, pen = fasta.tv.divergence_map((n, n)), fasta.losses.LeastSquares(b), fasta.proximal.TVBall()
for mode in (dict(adaptive=True), dict(adaptive=False, accelerate=True)):
    r = fasta.fasta(op, loss.f, loss.gradf, pen.g, pen.prox, Y0, verbose=False, max_iters=iters, L=1.0, tau0=0.02, **mode)
    torch.cuda.synchronize()
    print(mode, r.iteration_count, r.backtracks, r.tv_fused, (r.times[r.iteration_count] - r.times[0]) / r.iteration_count * 1e3, "ms/iter")

import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fasta-python_b200")]
import numpy as np
import fasta
from oracle import fasta_oracle, problems
p = problems.build("lasso_200x1000_k50", 0)
lam_max = np.max(np.abs(p.A.T @ p.b))
mus = lam_max * np.logspace(-1.5, -0.3, 8)
for mode in ("adaptive", "plain"):
    opts = dict(problems.HARNESS_OPTS, **problems.MODES[mode]); opts.pop("accelerate")
    np.random.seed(11)
    out = fasta.batched.lasso_path(p.A, p.b, mus, **opts)
    A = fasta.linalg.LinearMap.from_matrix(p.A)
    for j, mu in enumerate(mus):
        f = lambda z: .5 * np.linalg.norm((z - p.b).ravel()) ** 2
        gradf = lambda z: z - p.b
        g = lambda x: mu * np.linalg.norm(x.ravel(), 1)
        proxg = lambda x, t: fasta_oracle.shrink(x, t * mu)
        np.random.seed(11)
        ref = fasta_oracle.solve(lambda x: p.A @ x, lambda y: p.A.T @ y, f, gradf, g, proxg, p.x0, **opts)
        loss, pen = fasta.losses.LeastSquares(p.b), fasta.proximal.L1Norm(mu)
        np.random.seed(11)
        one = fasta.fasta(A, loss.f, loss.gradf, pen.g, pen.prox, p.x0, **opts)
        n = min(ref.iteration_count, out[j].iteration_count)
        dtau = np.max(np.abs(out[j].stepsizes[:n] - ref.stepsizes[:n]) / ref.stepsizes[:n])
        first = int(np.argmax(np.abs(out[j].stepsizes[:n] - ref.stepsizes[:n]) / ref.stepsizes[:n] > 1e-6)) if dtau > 1e-6 else -1
        print(mode, j, f"mu/lam={mu/lam_max:.4f}", "oracle", ref.iteration_count, ref.backtracks, "| single", one.iteration_count, one.backtracks,
              "| batched", out[j].iteration_count, out[j].backtracks, f"| tau0 {ref.stepsizes[0]:.6e} {out[j].stepsizes[0]:.6e} dtau {dtau:.2e} first {first}")

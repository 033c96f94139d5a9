"""GPU: BASELINE.json's full-size configurations through checks that do not need a full CPU solve.

Config 4 (TV denoising 4096 x 4096): the first iterations against the numpy oracle on the same inputs (the oracle
needs ~1.5 s per iteration at this size), adjointness of the stencil pair, feasibility of every iterate.
Config 2 (dense lasso 40000 x 100000, 32 GB) is cross-checked inside ``bench.py``'s CPU-baseline leg
(``cpu_baseline.parity_full_size``); config 3 inside ``tools/bench_configs.py``; config 5 in test_gpu_batched.py.
"""
import numpy as np
import pytest

from oracle import fasta_oracle, problems

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["adaptive", "accelerated"])
def test_tv_4096_first_iterations_match_oracle(mode):
    import fasta
    np.random.seed(0)
    p = problems.tv_denoising(n=4096, cell=64)
    opts = dict(problems.MODES[mode], verbose=False, max_iters=3, L=8.0, tau0=0.02, evaluate_objective=True)
    f, gradf, g, proxg = problems.numpy_callables(p)
    ref = fasta_oracle.solve(problems.tv_div, problems.tv_grad, f, gradf, g, proxg, p.x0, **opts)
    A = fasta.tv.divergence_map(p.x0.shape[:2])
    loss, pen = fasta.losses.LeastSquares(p.b), fasta.proximal.TVBall()
    res = fasta.fasta(A, loss.f, loss.gradf, pen.g, pen.prox, p.x0, **opts)
    assert res.tv_fused
    n = ref.iteration_count
    assert (res.iteration_count, res.backtracks) == (n, ref.backtracks)
    assert np.linalg.norm(res.solution - ref.solution) <= 1e-9 * np.linalg.norm(ref.solution)
    assert np.max(np.abs(res.objectives[:n + 1] - ref.objectives[:n + 1]) / np.abs(ref.objectives[:n + 1])) <= 1e-10
    assert np.allclose(res.stepsizes[:n], ref.stepsizes[:n], rtol=1e-9, atol=0)
    if mode != "accelerated":      # a prox output is feasible, |Y_ij|_2 <= 1 (tv_denoising.py:89-96); a FISTA iterate is
        assert np.max(np.linalg.norm(res.solution, axis=-1)) <= 1.0 + 1e-15      # the extrapolated point and need not be


def test_tv_4096_adjointness_and_linearity():
    import fasta
    import torch
    n = 4096
    g = torch.Generator(device="cuda").manual_seed(1)
    X = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g)
    Y = torch.randn(n, n, 2, dtype=torch.float64, device="cuda", generator=g)
    lhs = torch.sum(fasta.tv.div(Y) * X).item()
    rhs = torch.sum(Y * fasta.tv.grad(X)).item()
    assert abs(lhs - rhs) <= 1e-11 * max(abs(lhs), 1.0)
    # div(grad(const)) = 0 and linearity of the stencil at full size
    assert torch.count_nonzero(fasta.tv.grad(torch.full((n, n), 3.25, dtype=torch.float64, device="cuda"))).item() == 0
    Y2 = torch.randn(n, n, 2, dtype=torch.float64, device="cuda", generator=g)
    d = fasta.tv.div(Y + Y2) - (fasta.tv.div(Y) + fasta.tv.div(Y2))
    assert d.abs().max().item() <= 1e-13 * (Y.abs().max().item() + Y2.abs().max().item())

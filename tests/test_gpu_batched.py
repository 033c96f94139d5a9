"""GPU: the batched (regularisation-path / multi-RHS) loop: every column must reproduce an
independent reference run on that column (oracle run side by side, same seed before each)."""
import numpy as np
import pytest

from oracle import fasta_oracle, problems

pytestmark = pytest.mark.gpu


def _oracle_column(p, mu, b, opts, seed):
    f = lambda z: .5 * np.linalg.norm((z - b).ravel()) ** 2
    gradf = lambda z: z - b
    g = lambda x: mu * np.linalg.norm(x.ravel(), 1)
    proxg = lambda x, t: fasta_oracle.shrink(x, t * mu)
    np.random.seed(seed)
    return fasta_oracle.solve(lambda x: p.A @ x, lambda y: p.A.T @ y, f, gradf, g, proxg, p.x0, **opts)


def _check(res, ref, label, sol_tol=1e-9, obj_tol=1e-10):
    n = ref.iteration_count
    assert res.iteration_count == n, f"{label}: iterations {res.iteration_count} != {n}"
    assert res.backtracks == ref.backtracks, f"{label}: backtracks {res.backtracks} != {ref.backtracks}"
    err = np.linalg.norm(res.solution - ref.solution) / max(np.linalg.norm(ref.solution), 1e-300)
    assert err <= sol_tol, f"{label}: solution rel err {err:.2e}"
    scale = np.maximum(np.abs(ref.objectives[:n + 1]), 1e-3 * abs(ref.objectives[0]))
    oerr = np.max(np.abs(res.objectives[:n + 1] - ref.objectives[:n + 1]) / scale)
    assert oerr <= obj_tol, f"{label}: objective rel err {oerr:.2e}"
    assert np.all(res.residuals[n:] == 0)


@pytest.mark.parametrize("mode", ["adaptive", "plain"])
def test_lasso_path_columns_match_independent_runs(mode):
    import fasta
    p = problems.build("lasso_200x1000_k50", 0)
    lam_max = np.max(np.abs(p.A.T @ p.b))
    # benign range: for mu < 0.05*lam_max the adaptive runs (100+ iterations, 15+ backtracks) are chaotic --
    # oracle (CPU), single-problem GPU path and batched path then all differ from one another in counts
    # (tools/diag_batched.py), exactly like TV+adaptive
    mus = lam_max * np.logspace(-1.2, -0.3, 8)
    opts = dict(problems.HARNESS_OPTS, **problems.MODES[mode])
    opts.pop("accelerate")
    np.random.seed(11)
    out = fasta.batched.lasso_path(p.A, p.b, mus, **opts)
    assert len(out) == 8
    counts = set()
    for j, mu in enumerate(mus):
        ref = _oracle_column(p, mu, p.b, dict(opts, accelerate=False), 11)
        _check(out[j], ref, f"{mode}/mu[{j}]")
        counts.add(ref.iteration_count)
    assert len(counts) > 1          # columns really stop at different iterations
    assert out[0].batch["kernel_launches"] > 0


def _oracle_column_reordered(p, mu, b, opts, seed):
    """The same oracle run with the two contractions summed in a different (permuted) order: the
    distance between this run and the plain one is the problem's own sensitivity to last-bit
    reduction-order noise, i.e. the best agreement ANY other implementation of the sums can have."""
    perm = np.random.RandomState(101).permutation(p.A.shape[1])
    permr = np.random.RandomState(102).permutation(p.A.shape[0])
    Ap, ATp = np.ascontiguousarray(p.A[:, perm]), np.ascontiguousarray(p.A[permr].T)
    f = lambda z: .5 * np.linalg.norm((z - b).ravel()) ** 2
    gradf = lambda z: z - b
    g = lambda x: mu * np.linalg.norm(x.ravel(), 1)
    proxg = lambda x, t: fasta_oracle.shrink(x, t * mu)
    np.random.seed(seed)
    return fasta_oracle.solve(lambda x: Ap @ x[perm], lambda y: ATp @ y[permr], f, gradf, g, proxg, p.x0, **opts)


def test_multi_rhs_batch_with_backtracking_columns():
    """Different right-hand sides per column; K=50 problems backtrack in adaptive mode (config 1b).

    Adaptive runs with many backtracks amplify 1e-16 reduction-order noise (the larger right-hand sides
    here run 130+ iterations with 8+ backtracks; the oracle perturbed by a permuted summation order
    moves by up to 5e-8 from itself, and changes COUNTS on still larger ones).  So every column is held to
    the north-star bar (1e-9 / 1e-10) where the problem is well conditioned, and to 20x the oracle's own
    self-divergence where it is not; at least three columns must be held to the strict bar."""
    import fasta
    p = problems.build("lasso_200x1000_k50", 0)
    rng = np.random.RandomState(5)
    Bn = 6
    bs = np.stack([p.b * (1 + 0.2 * k) + 0.01 * rng.randn(p.b.size) for k in range(Bn)], axis=1)
    opts = dict(problems.HARNESS_OPTS, adaptive=True)
    np.random.seed(3)
    out = fasta.batched.fasta_batched(p.A, fasta.losses.LeastSquares(bs), fasta.proximal.L1Norm(p.mu),
                                      np.zeros((p.A.shape[1], Bn)), **opts)
    total_bt, strict = 0, 0
    for j in range(Bn):
        o = dict(opts, accelerate=False)
        ref = _oracle_column(p, p.mu, bs[:, j], o, 3)
        alt = _oracle_column_reordered(p, p.mu, bs[:, j], o, 3)
        if (alt.iteration_count, alt.backtracks) != (ref.iteration_count, ref.backtracks):
            continue                      # the reference itself is not reproducible on this column
        n = ref.iteration_count
        s_sol = np.linalg.norm(alt.solution - ref.solution) / np.linalg.norm(ref.solution)
        scale = np.maximum(np.abs(ref.objectives[:n + 1]), 1e-3 * abs(ref.objectives[0]))
        s_obj = np.max(np.abs(alt.objectives[:n + 1] - ref.objectives[:n + 1]) / scale)
        sol_tol, obj_tol = max(1e-9, 20 * s_sol), max(1e-10, 20 * s_obj)
        strict += (sol_tol == 1e-9 and obj_tol == 1e-10)
        _check(out[j], ref, f"rhs[{j}]", sol_tol=sol_tol, obj_tol=obj_tol)
        total_bt += ref.backtracks
    assert total_bt > 0 and strict >= 3


def test_batched_gemm_matches_numpy():
    import torch
    from fasta import _cabi, _device
    lib = _cabi.load()
    rng = np.random.RandomState(2)
    for (M, N, B) in [(130, 258, 6), (200, 1000, 8), (1000, 2048, 64), (33 * 2, 4098, 258)]:
        A, X, R = rng.randn(M, N), rng.randn(N, B), rng.randn(M, B)
        Ad, Xd, Rd = (torch.from_numpy(v).cuda() for v in (A, X, R))
        for adj, (rows, K, rhs, want) in enumerate([(M, N, Xd, A @ X), (N, M, Rd, A.T @ R)]):
            S = lib.fb200_gemm_splits(rows, B, K)
            C = torch.zeros(S, rows, B, dtype=torch.float64, device="cuda")
            _cabi.check(lib.fb200_gemm_f64(adj, Ad.data_ptr(), N, rhs.data_ptr(), B, C.data_ptr(), B, rows, B, K, S, rows * B,
                                           _device.stream_ptr()))
            got = C.sum(0).cpu().numpy()
            assert np.linalg.norm(got - want) <= 1e-13 * np.linalg.norm(want), (M, N, B, adj)

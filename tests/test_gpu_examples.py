"""GPU: the reference's remaining example problems (mmv, max_norm, democratic_representation, svm,
nn_factorization, logistic_matrix_completion) through the public API, written the way
the examples call it -- legacy 7-argument form with arrays or ``None, None`` operators, user callables on CUDA
tensors, the prox bodies from ``fasta.proximal`` -- against live-reference golden trajectories."""
import numpy as np
import pytest

from helpers import assert_trajectory, load_golden
from oracle import examples_extra, problems

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    return torch


def _dev(a):
    return _torch().from_numpy(np.ascontiguousarray(a)).cuda()


def _solve(case, opts):
    import fasta
    t = _torch()
    e = examples_extra.build(case, 0)
    d, x0 = e.data, _dev(e.x0)
    norm = t.linalg.norm
    if e.name == "mmv":                                             # mmv.py:49-65
        B, pen = _dev(d["B"]), fasta.proximal.RowGroupL2(d["mu"])
        f = lambda Z: .5 * norm((Z - B).ravel()) ** 2
        gradf = lambda Z: Z - B
        return fasta.fasta(d["A"], d["A"].T, f, gradf, pen.g, pen.prox, x0, **opts)
    if e.name == "max_norm":                                        # max_norm.py:49-61
        S, pen = _dev(d["S"]), fasta.proximal.RowNormBall(d["mu"])
        f = lambda X: t.sum(S * (X @ X.T))
        gradf = lambda X: (S + S.T) @ X
        return fasta.fasta(None, None, f, gradf, pen.g, pen.prox, x0, **opts)
    if e.name == "democratic":                                      # democratic_representation.py:41-46
        b, pen = _dev(d["b"]), fasta.proximal.LinfNorm(d["mu"])
        A = fasta.linalg.LinearMap.from_matrix(d["A"])              # dense matrix of mask * DCT
        f = lambda z: .5 * norm((z - b).ravel()) ** 2
        gradf = lambda z: z - b
        return fasta.fasta(A, f, gradf, pen.g, pen.prox, x0, **opts)
    if e.name == "svm":                                             # svm.py:68-74
        D, l, pen = _dev(d["D"]), _dev(d["l"]), fasta.proximal.Box(0.0, d["C"])
        f = lambda y: .5 * norm((D.T @ (l * y)).ravel()) ** 2 - t.sum(y)
        gradf = lambda y: l * (D @ (D.T @ (l * y))) - 1
        return fasta.fasta(None, None, f, gradf, pen.g, pen.prox, x0, **opts)
    if e.name == "nn_factorization":                                # nn_factorization.py:48-63
        S, mu, n = _dev(d["S"]), d["mu"], d["n"]
        f = lambda Z: .5 * norm((S - Z[:n] @ Z[n:].T).ravel()) ** 2

        def gradf(Z):
            X, Y = Z[:n], Z[n:]
            dd = X @ Y.T - S
            return t.cat((dd @ Y, dd.T @ X))

        g = lambda Z: mu * Z[:n].abs().sum()
        proxg = lambda Z, tt: t.cat((fasta.proximal.shrink(Z[:n], tt * mu), t.clamp(Z[n:], 0, 1)))
        return fasta.fasta(None, None, f, gradf, g, proxg, x0, **opts)
    if e.name == "logistic_matrix_completion":                      # logistic_matrix_completion.py:41-46
        B, mu = _dev(d["B"]), d["mu"]
        f = lambda Z: t.sum(t.log(1 + t.exp(Z)) - (B == 1) * Z)
        gradf = lambda Z: -B / (1 + t.exp(B * Z))
        # la.norm(np.diag(s), 1) of the reference is the MATRIX 1-norm of diag(s), i.e. the largest singular value
        g = lambda X: mu * fasta.proximal.singular_values(X).max()
        proxg = lambda X, tt: fasta.proximal.project_Lnuc_ball(X, tt * mu)
        return fasta.fasta(None, None, f, gradf, g, proxg, x0, **opts)
    raise ValueError(e.name)


@pytest.mark.parametrize("mode", list(problems.MODES))
@pytest.mark.parametrize("case", list(examples_extra.CASES))
def test_example_problem_matches_reference(case, mode):
    gold = load_golden(case, mode)
    res = _solve(case, gold["opts"])
    assert res.backend == "GenericBackend" and res.kernel_launches > 0
    res.solution = res.solution.cpu().numpy()
    assert_trajectory(res, gold, label=f"{case}/{mode}")


def test_row_prox_known_answers(golden_dir):
    import fasta
    with np.load(f"{golden_dir}/kat_row_prox.npz") as z:
        for i in range(int(z["count"])):
            X, t = z[f"X{i}"], float(z[f"t{i}"])
            for got, want in ((fasta.proximal.shrink_rows(X, t), z[f"mmv{i}"]),
                              (fasta.proximal.project_rows_L2_ball(X, t), z[f"ball{i}"]),
                              (fasta.proximal.row_norms(X), z[f"norms{i}"])):
                assert isinstance(got, np.ndarray) and got.shape == want.shape
                assert np.max(np.abs(got - want)) <= 4e-16 * max(1.0, np.max(np.abs(want)))
            # all-zero rows stay zero (the (norms == 0) guard of the reference)
            zero_rows = np.where(~X.any(axis=1))[0]
            assert np.all(fasta.proximal.shrink_rows(X, t)[zero_rows] == 0)


@pytest.mark.parametrize("M,N,L", [(20, 30, 10), (61, 90, 7), (33, 129, 1), (128, 256, 64)])
def test_matrix_iterate_dense_map(M, N, L):
    """A @ X and A.T @ Z for matrix unknowns (mmv.py:65) incl. odd sizes that take the padded path."""
    import fasta
    rng = np.random.RandomState(M + N + L)
    A, X, Z = rng.randn(M, N), rng.randn(N, L), rng.randn(M, L)
    op = fasta.linalg.LinearMap.from_matrix(A).with_columns(L)
    assert op.Vshape == (N, L) and op.Wshape == (M, L) and op.H.Vshape == (M, L)
    got, want = op(X), A @ X
    assert isinstance(got, np.ndarray) and np.linalg.norm(got - want) <= 1e-14 * np.linalg.norm(want)
    got, want = op.H(Z), A.T @ Z
    assert np.linalg.norm(got - want) <= 1e-14 * np.linalg.norm(want)
    with pytest.raises(AssertionError):
        op(X[:, :-1] if L > 1 else np.zeros((N, 2)))

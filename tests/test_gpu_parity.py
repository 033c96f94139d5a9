"""GPU: trajectory parity of the CUDA path (public fasta.fasta API -> C ABI) with the reference.

The bar (BASELINE.json north_star): identical iteration and backtrack counts, final iterate within
1e-9 relative, objective history within 1e-10 relative -- against the committed live-reference
golden trajectories AND against the numpy oracle run on the same seeded inputs.
"""
import numpy as np
import pytest

from helpers import assert_trajectory, golden_cases, load_golden
from oracle import fasta_oracle, problems

pytestmark = pytest.mark.gpu


def tagged(p, sharded=False):
    import fasta
    if p.kind == "dense":
        A = fasta.linalg.LinearMap.from_matrix(p.A)
    else:
        A = fasta.tv.divergence_map(p.x0.shape[:2])
    loss = {"least_squares": fasta.losses.LeastSquares, "logistic": fasta.losses.Logistic}[p.loss](p.b)
    pen = {"l1": lambda: fasta.proximal.L1Norm(p.mu), "l1ball": lambda: fasta.proximal.L1Ball(p.mu),
           "nonneg": fasta.proximal.NonNegative, "tv_ball": fasta.proximal.TVBall}[p.penalty]()
    return A, loss, pen


@pytest.mark.parametrize("case,mode", golden_cases())
def test_fused_path_matches_golden(case, mode):
    import fasta
    gold = load_golden(case, mode)
    p = problems.build(case, int(gold["seed"]))
    A, loss, pen = tagged(p)
    res = fasta.fasta(A, loss.f, loss.gradf, pen.g, pen.prox, p.x0, **gold["opts"])
    assert res.backend == "FusedBackend" and res.kernel_launches > 0
    assert isinstance(res.solution, np.ndarray)
    assert_trajectory(res, gold, label=f"{case}/{mode}")


SMALL_DENSE = ("lasso_200x1000_k10", "lasso_200x1000_k50", "lasso_333x1414_k40", "nnls_200x1000", "logistic_1000x2000")


@pytest.mark.parametrize("case,mode", golden_cases(prefixes=SMALL_DENSE))
def test_device_resident_loop_matches_golden(case, mode, monkeypatch):
    """Small dense problems run the whole loop in one cooperative kernel (csrc/resident_loop.cu)."""
    import fasta
    gold = load_golden(case, mode)
    p = problems.build(case, int(gold["seed"]))
    A, loss, pen = tagged(p)
    res = fasta.fasta(A, loss.f, loss.gradf, pen.g, pen.prox, p.x0, **gold["opts"])
    assert res.resident and res.backend == "FusedBackend"
    assert_trajectory(res, gold, label=f"resident/{case}/{mode}")
    n = res.iteration_count
    assert np.all(np.diff(res.times[:n + 1]) >= 0) and res.times[n] > res.times[0]
    # 200 x 1000 fits one cluster's shared memory twice: single-cluster variant; the larger cases take the grid variant
    assert res.resident_cluster == (p.A.shape == (200, 1000))
    if res.resident_cluster:
        monkeypatch.setenv("FASTA_B200_RESIDENT_CLUSTER", "0")
        p = problems.build(case, int(gold["seed"]))
        A, loss, pen = tagged(p)
        ref = fasta.fasta(A, loss.f, loss.gradf, pen.g, pen.prox, p.x0, **gold["opts"])
        assert ref.resident and not ref.resident_cluster
        assert_trajectory(ref, gold, label=f"resident-grid/{case}/{mode}")


@pytest.mark.parametrize("case,mode", golden_cases(prefixes=SMALL_DENSE))
def test_host_driven_loop_matches_golden_on_small_problems(case, mode, monkeypatch):
    """The same cases with the device-resident loop disabled (host loop + single-pass sweep)."""
    import fasta
    monkeypatch.setenv("FASTA_B200_RESIDENT", "0")
    gold = load_golden(case, mode)
    p = problems.build(case, int(gold["seed"]))
    A, loss, pen = tagged(p)
    res = fasta.fasta(A, loss.f, loss.gradf, pen.g, pen.prox, p.x0, **gold["opts"])
    assert not res.resident and res.single_pass
    assert_trajectory(res, gold, label=f"host-loop/{case}/{mode}")
    assert res.speculation is not None and res.speculation["speculated"] > 0 and res.speculation["mismatched"] == 0


@pytest.mark.parametrize("case,mode", [cm for cm in golden_cases(prefixes=("lasso_200x1000_k50", "logistic", "lasso_4000", "tv_64", "nnls"))
                                       if cm[1] != "accelerated"])
def test_without_speculative_run_ahead_matches_golden(case, mode, monkeypatch):
    """Host loop with the speculation disabled (the next trial is queued only after this one's sums were read)."""
    import fasta
    monkeypatch.setenv("FASTA_B200_RESIDENT", "0")
    gold = load_golden(case, mode)
    p = problems.build(case, int(gold["seed"]))
    A, loss, pen = tagged(p)
    state = np.random.get_state()                        # both runs must draw the same Lipschitz probes
    res = fasta.fasta(A, loss.f, loss.gradf, pen.g, pen.prox, p.x0, **gold["opts"])
    assert res.speculative and not res.resident          # the default
    assert res.speculation["mismatched"] == 0 and res.speculation["speculated"] >= res.iteration_count - 1
    assert res.speculation["dropped"] <= res.backtracks and (res.backtracks == 0 or res.speculation["dropped"] > 0)
    assert_trajectory(res, gold, label=f"speculative/{case}/{mode}")
    monkeypatch.setenv("FASTA_B200_SPECULATE", "0")
    np.random.set_state(state)
    ref = fasta.fasta(A, loss.f, loss.gradf, pen.g, pen.prox, p.x0, **gold["opts"])
    assert not ref.speculative
    assert_trajectory(ref, gold, label=f"non-speculative/{case}/{mode}")
    n = ref.iteration_count
    # the device-side step-size algebra is the host's up to the last bit of numpy's scalar power
    k = min(n, 20)
    assert np.allclose(res.stepsizes[:k], ref.stepsizes[:k], rtol=1e-12, atol=0)


@pytest.mark.parametrize("case,mode", golden_cases(prefixes=("lasso_200x1000_k50", "lasso_333", "tv_64", "lasso_4000")))
def test_two_probe_lipschitz_estimate_matches_golden(case, mode, monkeypatch):
    """Least-squares problems estimate L from ONE contraction pair on v1 - v2 (the gradient is affine); with
    FASTA_B200_AFFINE_PROBE=0 the two terms of reference :106-110 are evaluated separately.  Same trajectory, and the
    two estimates agree to rounding."""
    import fasta
    monkeypatch.setenv("FASTA_B200_RESIDENT", "0")
    gold = load_golden(case, mode)
    p = problems.build(case, int(gold["seed"]))
    A, loss, pen = tagged(p)
    state = np.random.get_state()
    one = fasta.fasta(A, loss.f, loss.gradf, pen.g, pen.prox, p.x0, **gold["opts"])
    monkeypatch.setenv("FASTA_B200_AFFINE_PROBE", "0")
    np.random.set_state(state)
    two = fasta.fasta(A, loss.f, loss.gradf, pen.g, pen.prox, p.x0, **gold["opts"])
    assert_trajectory(two, gold, label=f"two-probe/{case}/{mode}")
    assert abs(one.stepsizes[0] - two.stepsizes[0]) <= 1e-13 * two.stepsizes[0]
    assert one.kernel_launches < two.kernel_launches


def test_device_resident_loop_options(capsys):
    """Other stop rules, no backtracking, no objective, verbose lines, user-supplied L / tau0."""
    import fasta
    from oracle import fasta_oracle
    p = problems.build("lasso_200x1000_k50", 0)
    A, loss, pen = tagged(p)
    f, gradf, g, proxg = problems.numpy_callables(p)
    for opts in (dict(stop_rule_name="residual", tolerance=1e-3), dict(stop_rule_name="norm_residual", tolerance=1e-4),
                 dict(stop_rule_name="ratio_residual", tolerance=1e-4), dict(backtrack=False, adaptive=False, max_iters=50),
                 dict(L=1.3, tau0=0.11, max_iters=40, evaluate_objective=False), dict(window=3, stepsize_shrink=0.5, max_iters=60),
                 dict(window=70, max_iters=90)):
        o = dict(verbose=False, evaluate_objective=True)
        o.update(opts)
        name = o.pop("stop_rule_name", None)
        o_ref = dict(o)
        if name:
            o["stop_rule"] = getattr(fasta.stopping, name)
            o_ref["stop_rule"] = getattr(fasta_oracle, "stop_" + name)
        np.random.seed(5)
        res = fasta.fasta(A, loss.f, loss.gradf, pen.g, pen.prox, p.x0, **o)
        np.random.seed(5)
        ref = fasta_oracle.solve(lambda x: p.A @ x, lambda y: p.A.T @ y, f, gradf, g, proxg, p.x0, **o_ref)
        assert res.resident
        assert (res.iteration_count, res.backtracks) == (ref.iteration_count, ref.backtracks), opts
        n = ref.iteration_count
        assert np.linalg.norm(res.solution - ref.solution) <= 1e-9 * np.linalg.norm(ref.solution)
        assert np.allclose(res.stepsizes[:n], ref.stepsizes[:n], rtol=1e-6, atol=0)
        assert np.all(res.residuals[n:] == 0)
        if o["evaluate_objective"]:
            assert np.max(np.abs(res.objectives[:n + 1] - ref.objectives[:n + 1]) / np.abs(ref.objectives[:n + 1])) <= 1e-10
        else:
            assert res.objectives is None
    capsys.readouterr()
    fasta.fasta(A, loss.f, loss.gradf, pen.g, pen.prox, p.x0, max_iters=3, evaluate_objective=True)      # verbose default
    outp = capsys.readouterr().out
    assert outp.startswith("Initializing FASTA...\n\nIteration #\tResidual") and outp.count("\n[") == 3


@pytest.mark.parametrize("case,mode", golden_cases(prefixes=("lasso_200x1000_k50", "logistic", "lasso_333", "l1ball")))
def test_two_pass_path_matches_golden(case, mode, monkeypatch):
    """Same parity bar with the single-pass sweep disabled (separate A x and A^T r kernels)."""
    import fasta
    monkeypatch.setenv("FASTA_B200_SWEEP", "0")
    gold = load_golden(case, mode)
    p = problems.build(case, int(gold["seed"]))
    A, loss, pen = tagged(p)
    res = fasta.fasta(A, loss.f, loss.gradf, pen.g, pen.prox, p.x0, **gold["opts"])
    assert res.backend == "FusedBackend" and not res.single_pass
    assert_trajectory(res, gold, label=f"two-pass/{case}/{mode}")


@pytest.mark.parametrize("case,mode", golden_cases(prefixes=("tv_",)))
def test_tv_unfused_kernels_match_golden(case, mode, monkeypatch):
    """TV with the fused iteration kernels disabled (separate step / div / grad kernels)."""
    import fasta
    monkeypatch.setattr(fasta._backends.TVDriver, "fused_step_ok", False)
    gold = load_golden(case, mode)
    p = problems.build(case, int(gold["seed"]))
    A, loss, pen = tagged(p)
    res = fasta.fasta(A, loss.f, loss.gradf, pen.g, pen.prox, p.x0, **gold["opts"])
    assert not res.tv_fused
    assert_trajectory(res, gold, label=f"tv-unfused/{case}/{mode}")


@pytest.mark.parametrize("case,mode", [cm for cm in golden_cases(prefixes=("tv_",)) if cm[1] != "accelerated"])
def test_tv_two_kernel_iteration_matches_golden(case, mode, monkeypatch):
    """TV with the whole-iteration kernel disabled (step+div+loss and grad+BB as two fused kernels)."""
    import fasta
    monkeypatch.setattr(fasta._backends.TVDriver, "iter_fused_ok", False)
    gold = load_golden(case, mode)
    p = problems.build(case, int(gold["seed"]))
    A, loss, pen = tagged(p)
    res = fasta.fasta(A, loss.f, loss.gradf, pen.g, pen.prox, p.x0, **gold["opts"])
    assert res.tv_fused
    assert_trajectory(res, gold, label=f"tv-two-kernel/{case}/{mode}")


def test_tv_fused_is_default(monkeypatch):
    import fasta
    p = problems.build("tv_64", 0)
    A, loss, pen = tagged(p)
    assert fasta.fasta(A, loss.f, loss.gradf, pen.g, pen.prox, p.x0, verbose=False, max_iters=3).tv_fused
    res = fasta.fasta(A, loss.f, loss.gradf, pen.g, pen.prox, p.x0, verbose=False, max_iters=3, accelerate=True)
    assert res.tv_fused
    monkeypatch.setattr(fasta._backends.TVDriver, "fista_fused_ok", False)
    ref = fasta.fasta(A, loss.f, loss.gradf, pen.g, pen.prox, p.x0, verbose=False, max_iters=3, accelerate=True)
    assert not ref.tv_fused and ref.kernel_launches >= res.kernel_launches + 3 * 3     # one kernel per trial instead of four


@pytest.mark.parametrize("case", ["tv_64", "tv_128"])
def test_tv_accelerated_without_the_fista_kernel_matches_golden(case, monkeypatch):
    """Accelerated TV with the separate step / div / extrapolate / grad kernels (the fused FISTA kernel is the default)."""
    import fasta
    monkeypatch.setattr(fasta._backends.TVDriver, "fista_fused_ok", False)
    gold = load_golden(case, "accelerated")
    p = problems.build(case, int(gold["seed"]))
    A, loss, pen = tagged(p)
    res = fasta.fasta(A, loss.f, loss.gradf, pen.g, pen.prox, p.x0, **gold["opts"])
    assert not res.tv_fused
    assert_trajectory(res, gold, label=f"tv-accel-unfused/{case}")


def test_single_pass_is_default_for_dense_solves(monkeypatch):
    import fasta
    p = problems.build("lasso_200x1000_k10", 0)
    A, loss, pen = tagged(p)
    res = fasta.fasta(A, loss.f, loss.gradf, pen.g, pen.prox, p.x0, verbose=False, max_iters=3)
    assert res.single_pass
    res = fasta.fasta(A, loss.f, loss.gradf, pen.g, pen.prox, p.x0, verbose=False, max_iters=3, accelerate=True)
    assert res.single_pass                                # FISTA mode of the sweep (extrapolated z formed per row)
    monkeypatch.setenv("FASTA_B200_SWEEP_ACCEL", "0")
    res = fasta.fasta(A, loss.f, loss.gradf, pen.g, pen.prox, p.x0, verbose=False, max_iters=3, accelerate=True)
    assert not res.single_pass


@pytest.mark.parametrize("case", ["lasso_200x1000_k50", "logistic_1000x2000", "l1ball_200x1000", "lasso_333x1414_k40"])
def test_accelerated_without_the_fused_fista_sweep_matches_golden(case, monkeypatch):
    """Accelerated mode with the separate forward / extrapolate / adjoint kernels (the fused FISTA sweep is the default)."""
    import fasta
    monkeypatch.setenv("FASTA_B200_SWEEP_ACCEL", "0")
    monkeypatch.setenv("FASTA_B200_RESIDENT", "0")
    gold = load_golden(case, "accelerated")
    p = problems.build(case, int(gold["seed"]))
    A, loss, pen = tagged(p)
    res = fasta.fasta(A, loss.f, loss.gradf, pen.g, pen.prox, p.x0, **gold["opts"])
    assert not res.single_pass and not res.resident
    assert_trajectory(res, gold, label=f"accel-unfused/{case}")


def test_device_resident_loop_fista_options(capsys, monkeypatch):
    """FISTA in the device-resident loop: restart on/off, together with the adaptive step size, verbose lines."""
    import fasta
    p = problems.build("lasso_200x1000_k50", 0)
    A, loss, pen = tagged(p)
    f, gradf, g, proxg = problems.numpy_callables(p)
    for opts in (dict(restart=False, max_iters=120), dict(adaptive=True, max_iters=80), dict(backtrack=False, max_iters=60),
                 dict(evaluate_objective=False, max_iters=90)):
        o = dict(verbose=False, evaluate_objective=True, accelerate=True, adaptive=False)
        o.update(opts)
        np.random.seed(7)
        res = fasta.fasta(A, loss.f, loss.gradf, pen.g, pen.prox, p.x0, **o)
        np.random.seed(7)
        ref = fasta_oracle.solve(lambda x: p.A @ x, lambda y: p.A.T @ y, f, gradf, g, proxg, p.x0, **o)
        assert res.resident
        assert (res.iteration_count, res.backtracks) == (ref.iteration_count, ref.backtracks), opts
        n = ref.iteration_count
        assert np.linalg.norm(res.solution - ref.solution) <= 1e-9 * np.linalg.norm(ref.solution), opts
        assert np.allclose(res.stepsizes[:n], ref.stepsizes[:n], rtol=1e-6, atol=0)
        assert np.allclose(res.residuals[:n], ref.residuals[:n], rtol=1e-6, atol=0)
        if o["evaluate_objective"]:
            assert np.max(np.abs(res.objectives[:n + 1] - ref.objectives[:n + 1]) / np.abs(ref.objectives[:n + 1])) <= 1e-10
    # verbose output is the host loop's (= the reference's format), line for line: alpha column, "Restarted acceleration."
    o = dict(accelerate=True, adaptive=False, max_iters=200, evaluate_objective=True)
    capsys.readouterr()
    np.random.seed(7)
    assert fasta.fasta(A, loss.f, loss.gradf, pen.g, pen.prox, p.x0, **o).resident
    got = capsys.readouterr().out
    monkeypatch.setenv("FASTA_B200_RESIDENT", "0")
    np.random.seed(7)
    assert not fasta.fasta(A, loss.f, loss.gradf, pen.g, pen.prox, p.x0, **o).resident
    want = capsys.readouterr().out
    assert "Restarted acceleration." in want
    gl, wl = got.splitlines(), want.splitlines()
    assert len(gl) == len(wl)
    for a, b in zip(gl, wl):
        if a.startswith("["):
            fa, fb = a.split("\t"), b.split("\t")
            assert fa[0] == fb[0] and fa[4] == fb[4]
            assert np.allclose([float(v) for v in fa[1:4] + fa[5:]], [float(v) for v in fb[1:4] + fb[5:]], rtol=1e-5)
        else:
            assert a == b


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("mode", list(problems.MODES))
def test_fused_path_matches_oracle_other_seeds(seed, mode):
    """Fresh seeded inputs (not in the fixtures): CUDA path vs the numpy oracle run side by side."""
    import fasta
    opts = dict(problems.HARNESS_OPTS, **problems.MODES[mode])
    for case in ("lasso_200x1000_k50", "logistic_1000x2000", "l1ball_200x1000"):
        p = problems.build(case, seed)
        state = np.random.get_state()
        ref = fasta_oracle.solve_problem(p, **opts)
        np.random.set_state(state)          # the CUDA run must see the same tau0 draws
        A, loss, pen = tagged(p)
        res = fasta.fasta(A, loss.f, loss.gradf, pen.g, pen.prox, p.x0, **opts)
        n = ref.iteration_count
        gold = dict(iteration_count=n, backtracks=ref.backtracks, solution=ref.solution,
                    objectives=ref.objectives[:n + 1], residuals=ref.residuals[:n],
                    stepsizes=ref.stepsizes[:n], norm_residuals=ref.norm_residuals[:n])
        assert_trajectory(res, gold, label=f"{case}/{mode}/seed{seed}")


def test_legacy_seven_argument_form_with_arrays():
    """fasta(A, At, f, gradf, g, proxg, x0) with ndarrays (sparse_least_squares.py:46,76)."""
    import fasta
    gold = load_golden("lasso_200x1000_k10", "adaptive")
    p = problems.build("lasso_200x1000_k10", 0)
    loss, pen = fasta.losses.LeastSquares(p.b), fasta.proximal.L1Norm(p.mu)
    res = fasta.fasta(p.A, p.A.T, loss.f, loss.gradf, pen.g, pen.prox, p.x0, **gold["opts"])
    assert res.backend == "FusedBackend"
    assert_trajectory(res, gold, label="legacy/arrays")


def test_legacy_form_with_tv_callables():
    """fasta(div, grad, ...) as tv_denoising.py:99 calls it."""
    import fasta
    gold = load_golden("tv_64", "accelerated")
    p = problems.build("tv_64", 0)
    loss, pen = fasta.losses.LeastSquares(p.b), fasta.proximal.TVBall()
    res = fasta.fasta(fasta.tv.div, fasta.tv.grad, loss.f, loss.gradf, pen.g, pen.prox, p.x0, **gold["opts"])
    assert res.backend == "FusedBackend"
    assert_trajectory(res, gold, label="legacy/tv")


@pytest.mark.parametrize("mode", list(problems.MODES))
def test_generic_path_with_user_callables(mode):
    """Untagged torch lambdas (the reference's own formulas) run through GenericBackend on the GPU."""
    import fasta
    import torch
    gold = load_golden("lasso_200x1000_k50", mode)
    p = problems.build("lasso_200x1000_k50", 0)
    b = torch.from_numpy(p.b).cuda()
    mu = p.mu
    f = lambda z: .5 * torch.linalg.norm((z - b).ravel()) ** 2
    gradf = lambda z: z - b
    g = lambda x: mu * torch.linalg.norm(x.ravel(), 1)
    proxg = lambda x, t: fasta.proximal.shrink(x, t * mu)
    A = fasta.linalg.LinearMap.from_matrix(p.A)
    res = fasta.fasta(A, f, gradf, g, proxg, torch.from_numpy(p.x0).cuda(), **gold["opts"])
    assert res.backend == "GenericBackend"
    assert isinstance(res.solution, torch.Tensor) and res.solution.is_cuda
    res.solution = res.solution.cpu().numpy()
    assert_trajectory(res, gold, label=f"generic/{mode}")


def test_gradient_descent_when_g_is_none_and_hooks():
    import fasta
    p = problems.build("lasso_200x1000_k10", 0)
    f, gradf, _, _ = problems.numpy_callables(p)
    op, adj, _, _ = problems.numpy_operator(p)
    opts = dict(verbose=False, max_iters=25, evaluate_objective=True, record_iterates=True)
    np.random.seed(3)
    ref = fasta_oracle.solve(op, adj, f, gradf, None, None, p.x0, func=lambda x: np.abs(x).max(), **opts)
    np.random.seed(3)
    loss = fasta.losses.LeastSquares(p.b)
    res = fasta.fasta(fasta.linalg.LinearMap.from_matrix(p.A), loss.f, loss.gradf, None, None, p.x0,
                      func=lambda x: np.abs(x).max(), **opts)
    assert res.backend == "FusedBackend"
    assert res.iteration_count == ref.iteration_count and res.backtracks == ref.backtracks
    n = ref.iteration_count
    # the unregularised objective decays towards 0: compare on the scale of the initial objective
    scale = np.maximum(np.abs(ref.objectives[:n + 1]), 1e-3 * abs(ref.objectives[0]))
    assert np.max(np.abs(res.objectives[:n + 1] - ref.objectives[:n + 1]) / scale) <= 1e-10
    np.testing.assert_allclose(res.iterates[:n + 1], ref.iterates[:n + 1], rtol=0, atol=1e-9 * np.abs(ref.iterates).max())
    np.testing.assert_allclose(res.function_hist[:n + 1], ref.function_hist[:n + 1], rtol=1e-9)
    assert np.array_equal(res.iterates[0], p.x0)


def test_x0_not_mutated_and_torch_in_torch_out():
    import fasta
    import torch
    p = problems.build("nnls_200x1000", 0)
    x0 = torch.from_numpy(p.x0.copy()).cuda() + 0.25
    keep = x0.clone()
    A = fasta.linalg.LinearMap.from_matrix(torch.from_numpy(p.A).cuda())
    loss = fasta.losses.LeastSquares(torch.from_numpy(p.b).cuda())
    pen = fasta.proximal.NonNegative()
    res = fasta.fasta(A, loss.f, loss.gradf, pen.g, pen.prox, x0, verbose=False, max_iters=5)
    assert torch.equal(x0, keep)
    assert isinstance(res.solution, torch.Tensor) and res.solution.is_cuda and res.solution.shape == x0.shape
    assert res.solution.data_ptr() != x0.data_ptr()


def _verbose_sets():
    import os
    from conftest import GOLDEN
    with np.load(os.path.join(GOLDEN, "kat_verbose.npz")) as z:
        return [(str(z[f"case{k}"]), str(z[f"mode{k}"]), eval(str(z[f"extra{k}"])), str(z[f"text{k}"])) for k in range(int(z["count"]))]


@pytest.mark.parametrize("k", range(7))
@pytest.mark.parametrize("resident", ["1", "0"])
def test_verbose_text_is_the_reference_stdout(k, resident, capsys, monkeypatch):
    """F-12 on the GPU: the text fasta() prints (host-driven loop and device-resident loop) against the live
    reference's captured stdout -- same lines, header, restart notices, indices and backtrack counts; numeric columns to
    the printed precision."""
    import fasta
    from helpers import assert_verbose_text
    case, mode, extra, text = _verbose_sets()[k]
    monkeypatch.setenv("FASTA_B200_RESIDENT", resident)
    p = problems.build(case, 0)
    A, loss, pen = tagged(p)
    opts = dict(problems.HARNESS_OPTS, **problems.MODES[mode])
    opts.update(extra)
    opts["verbose"] = True
    capsys.readouterr()
    fasta.fasta(A, loss.f, loss.gradf, pen.g, pen.prox, p.x0, **opts)
    assert_verbose_text(capsys.readouterr().out, text, rtol=5e-6, label=f"{case}/{mode}/resident={resident}")

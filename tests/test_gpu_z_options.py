"""GPU: option semantics of fasta() through the public API against the LIVE reference's answers
(tests/golden/kat_options.npz, oracle/make_golden.py OPTION_SETS): stop rules, backtrack off, user L / tau0, window /
shrink overrides, restart off, adaptive + accelerated together, hooks, max_iters = 1, immediate stop -- on the
device-resident loop (default for this 200 x 1000 problem) and on the host-driven loop."""
import ast

import numpy as np
import pytest

from oracle import problems

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("resident", ["1", "0"])
def test_option_sets_match_live_reference(resident, golden_dir, monkeypatch):
    import fasta
    monkeypatch.setenv("FASTA_B200_RESIDENT", resident)
    with np.load(f"{golden_dir}/kat_options.npz", allow_pickle=False) as z:
        g = {k: z[k] for k in z.files}
    for k in range(int(g["count"])):
        o = ast.literal_eval(str(g[f"opts{k}"]))
        p = problems.build(str(g["case"]), 0)
        A = fasta.linalg.LinearMap.from_matrix(p.A)
        loss, pen = fasta.losses.LeastSquares(p.b), fasta.proximal.L1Norm(p.mu)
        opts = dict(verbose=False, evaluate_objective=True)
        opts.update(o)
        if "stop_rule" in opts:
            opts["stop_rule"] = getattr(fasta.stopping, opts["stop_rule"])
        if opts.get("func") == "max_abs":
            opts["func"] = lambda x: np.abs(x).max()
        np.random.seed(int(g["seed"]))
        res = fasta.fasta(A, loss.f, loss.gradf, pen.g, pen.prox, p.x0, **opts)
        n = int(g[f"n{k}"])
        assert (res.iteration_count, res.backtracks) == (n, int(g[f"bt{k}"])), (o, resident)
        assert np.linalg.norm(res.solution - g[f"solution{k}"]) <= 1e-9 * np.linalg.norm(g[f"solution{k}"]), o
        assert res.residuals.shape == g[f"residuals{k}"].shape and np.all(res.residuals[n:] == 0)
        np.testing.assert_allclose(res.stepsizes[:n], g[f"stepsizes{k}"][:n], rtol=1e-6)
        if f"objectives{k}" in g:
            np.testing.assert_allclose(res.objectives[:n + 1], g[f"objectives{k}"][:n + 1], rtol=1e-10)
        else:
            assert res.objectives is None
        if f"iterates{k}" in g:
            np.testing.assert_allclose(res.iterates, g[f"iterates{k}"], rtol=0, atol=1e-9 * np.abs(g[f"iterates{k}"]).max())
            np.testing.assert_allclose(res.function_hist, g[f"function_hist{k}"], rtol=1e-9)

"""CPU: the reference arm of bench.py (`--impl reference`, the one bench leg that runs without a GPU) prints ONE JSON
line with the keys the driver's contract names, for every workload family; the parity block shared with the GPU arm
reports a wrong solver instead of hiding it."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_line(workload, extra=(), env=None):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload,
                          "--steps", "1", "--warmup", "0", *extra], capture_output=True, text=True, timeout=600, cwd=ROOT,
                         env=dict(os.environ, CUDA_VISIBLE_DEVICES="", **(env or {})))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


@pytest.mark.parametrize("workload,metric,unit", [
    ("lasso_200x1000", "lasso_fbs_iterations_per_sec", "iterations/s"),
    ("logistic_10000x2000", "logistic_fbs_iterations_per_sec", "iterations/s"),
    ("tv_512", "tv_fbs_iterations_per_sec", "iterations/s"),
    ("batched_32x2000x5000", "batched_lasso_path_column_iterations_per_sec", "column-iterations/s"),
])
def test_reference_arm_prints_the_contract_line(workload, metric, unit):
    # OMP_NUM_THREADS=1 is what torchrun exports to its workers: the arm must undo it (it ran 12x slow)
    d = _reference_line(workload, env=dict(OMP_NUM_THREADS="1"))
    assert d["impl"] == "reference" and d["metric"] == metric and d["unit"] == unit
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["value"] > 0 and abs(d["ms_per_step"] - 1e3 / d["value"]) < 1e-6 * d["ms_per_step"]
    cores = len(os.sched_getaffinity(0))
    assert d["cpu_baseline"]["cores"] == cores and d["num_threads"] == cores
    have_ref = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "fasta")) or os.path.isdir("/root/reference/fasta")
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert d["e2e"] == dict(value=d["value"], unit=unit, h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_is_silent_on_other_ranks():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT,
                         env=dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", CUDA_VISIBLE_DEVICES=""))
    assert out.returncode == 0 and not [l for l in out.stdout.splitlines() if l.startswith("{")]


def _bench_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(bench)
    finally:
        sys.argv = argv
    return bench


def test_parity_block_reports_a_wrong_solver():
    """`parity_block` is the north_star bar (counts, solution 1e-9, objective history 1e-10): green for the same run,
    red -- not hidden -- for a solver that differs, absolute where the reference value is exactly zero."""
    bench = _bench_module()
    from oracle import fasta_oracle
    rng = np.random.RandomState(0)
    M, N, mu = 60, 200, 0.02
    A = rng.randn(M, N) / (np.sqrt(M) + np.sqrt(N))
    b = A @ (rng.rand(N) < 0.05) + 0.01 * rng.randn(M)

    def solver(scale=1.0, **o):
        np.random.seed(1)
        return fasta_oracle.solve(lambda v: A @ v, lambda y: A.T @ y, lambda z: .5 * np.linalg.norm(z - b) ** 2,
                                  lambda z: z - b, lambda v: mu * np.abs(v).sum(),
                                  lambda v, t: fasta_oracle.shrink(v, t * mu * scale), np.zeros(N),
                                  evaluate_objective=True, **o)

    want = solver()
    good = bench.parity_block(solver(), want)
    assert good["ok"] and good["solution_rel_err"] == 0.0 and good["objective_history_rel_err"] == 0.0
    assert good["iterations"] == [want.iteration_count] * 2
    bad = bench.parity_block(solver(1.5), want)
    assert not bad["ok"] and bad["solution_rel_err"] > 1e-6
    assert bench.rel([0.0, 1.0], [0.0, 1.0]) == 0.0 and bench.rel([1e-20, 1.0], [0.0, 1.0]) == 1e-20


def test_reference_loader_runs_the_unmodified_reference_next_to_this_package():
    """oracle/ref_loader.load_isolated: the reference package (same top-level name `fasta`) imported in a process that
    already holds this repo's package, without disturbing it."""
    from oracle import problems, ref_loader
    import fasta as ours
    ref, root = ref_loader.load_isolated()
    if ref is None:
        pytest.skip("neither baseline/_ref nor /root/reference present")
    assert sys.modules["fasta"] is ours and ref is not ours and os.path.realpath(ref.__file__).startswith(os.path.realpath(root))
    p = problems.build("lasso_200x1000_k10", 0)
    f, gradf, g, proxg = problems.numpy_callables(p)
    np.random.seed(3)
    res = ref.fasta(ref.linalg.LinearMap.from_matrix(p.A), f, gradf, g, proxg, p.x0, verbose=False, max_iters=5)
    assert res.iteration_count == 5 and np.all(np.isfinite(res.solution))

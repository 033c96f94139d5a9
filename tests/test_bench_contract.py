"""CPU: the reference arm of bench.py (`--impl reference`, the one bench leg that runs without a GPU) prints ONE JSON
line with the keys the driver's contract names; the pieces shared with the GPU arm (`config_of`, the CPU-baseline
leg incl. its full-size parity cross-check hook) are exercised on the small workload."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                          "lasso_8000x20000", "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "lasso_fbs_iterations_per_sec" and d["unit"] == "iterations/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["value"] > 0 and abs(d["ms_per_step"] - 1e3 / d["value"]) < 1e-6 * d["ms_per_step"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "first 3 iterations" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == dict(value=d["value"], unit="iterations/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_is_silent_on_other_ranks():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT,
                         env=dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", CUDA_VISIBLE_DEVICES=""))
    assert out.returncode == 0 and not [l for l in out.stdout.splitlines() if l.startswith("{")]


def test_cpu_leg_parity_hook():
    """`cpu_arm(..., gpu_solve=...)` compares the solver it is handed with the oracle run on the same seed."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(bench)
    finally:
        sys.argv = argv
    from oracle import fasta_oracle
    rng = np.random.RandomState(0)
    M, N, mu = 60, 200, 0.02
    A = rng.randn(M, N) / (np.sqrt(M) + np.sqrt(N))
    b = A @ (rng.rand(N) < 0.05) + 0.01 * rng.randn(M)

    def solver(o, scale=1.0):
        return fasta_oracle.solve(lambda v: A @ v, lambda y: A.T @ y, lambda z: .5 * np.linalg.norm(z - b) ** 2,
                                  lambda z: z - b, lambda v: mu * np.abs(v).sum(),
                                  lambda v, t: fasta_oracle.shrink(v, t * mu * scale), np.zeros(N), **o)

    good = bench.cpu_arm(dict(M=M, N=N, mu=mu), A, b, 3, gpu_solve=solver)["parity_full_size"]
    assert good["iterations"] == [3, 3] and good["iterate_rel_err"] == 0.0 and good["stepsizes_rel_err"] == 0.0
    bad = bench.cpu_arm(dict(M=M, N=N, mu=mu), A, b, 3, gpu_solve=lambda o: solver(o, 1.5))["parity_full_size"]
    assert bad["iterate_rel_err"] > 1e-6                                  # a wrong solver is reported, not hidden
    err = bench.cpu_arm(dict(M=M, N=N, mu=mu), A, b, 3, gpu_solve=lambda o: 1 / 0)["parity_full_size"]
    assert "error" in err

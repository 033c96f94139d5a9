"""TEST INFRASTRUCTURE ONLY -- run in a fresh process where /root/reference exists (the build container):

    python oracle/live_check.py [count] [seed]

Draws `count` random small problems (lasso / non-negative least squares / L1-ball / logistic; random shapes, sparsity,
modes and solver options) and compares the numpy oracle with the UNMODIFIED live reference on each, bit for bit
(same process, same numpy / BLAS), every third one also on the verbose text fasta() prints.  Prints one line per problem and `live_check ok <count>` at the end; exit code 1
on the first difference.  Used by tests/test_oracle_live_reference.py (skipped where the reference is absent) --
a randomized complement to the committed fixtures of tests/golden/.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_loader  # noqa: E402
from oracle import fasta_oracle, problems  # noqa: E402


def main(count=12, seed=2024):
    ref = ref_loader.load()
    rng = np.random.RandomState(seed)
    gens = [problems.sparse_least_squares, problems.nonneg_least_squares, problems.l1ball_lasso, problems.sparse_logistic]
    for k in range(count):
        gen = gens[rng.randint(len(gens))]
        M, N = int(rng.randint(5, 120)), int(rng.randint(5, 300))
        K = int(rng.randint(1, max(2, min(N, 20))))
        pseed = int(rng.randint(1 << 30))
        np.random.seed(pseed)
        p = gen(M=M, N=N, K=K)
        opts = dict(verbose=False, evaluate_objective=bool(rng.randint(2)), max_iters=int(rng.randint(1, 120)),
                    adaptive=bool(rng.randint(2)), accelerate=bool(rng.randint(2)), restart=bool(rng.randint(2)),
                    backtrack=bool(rng.randint(4)), window=int(rng.randint(1, 15)), tolerance=float(10.0 ** -rng.randint(2, 8)))
        if rng.randint(3) == 0:
            opts["stepsize_shrink"] = float(rng.uniform(0.1, 0.9))
        rule = ("residual", "norm_residual", "ratio_residual", "hybrid_residual")[rng.randint(4)]
        f, gradf, g, proxg = problems.numpy_callables(p)
        apply, adjoint, vshape, wshape = problems.numpy_operator(p)
        opts["verbose"] = (k % 3 == 0)              # every third problem also compares the printed text byte for byte
        state = np.random.get_state()
        texts = []
        with np.errstate(all="ignore"):
            for which in ("ref", "oracle"):
                buf = io.StringIO()
                with contextlib.redirect_stdout(buf):
                    if which == "ref":
                        want = ref.fasta(ref.linalg.LinearMap(apply, adjoint, vshape, wshape), f, gradf, g, proxg, p.x0,
                                         stop_rule=getattr(ref.stopping, rule), **opts)
                        np.random.set_state(state)
                    else:
                        got = fasta_oracle.solve(apply, adjoint, f, gradf, g, proxg, p.x0,
                                                 stop_rule=getattr(fasta_oracle, "stop_" + rule), **opts)
                texts.append(buf.getvalue())
        ok = (texts[0] == texts[1] and got.iteration_count == want.iteration_count and got.backtracks == want.backtracks
              and np.array_equal(got.solution, want.solution, equal_nan=True)
              and np.array_equal(got.residuals, want.residuals, equal_nan=True)
              and np.array_equal(got.norm_residuals, want.norm_residuals, equal_nan=True)
              and np.array_equal(got.stepsizes, want.stepsizes, equal_nan=True)
              and ((got.objectives is None and want.objectives is None)
                   or np.array_equal(got.objectives, want.objectives, equal_nan=True)))
        print(f"{k:3d} {gen.__name__:22s} {M:4d}x{N:<4d} K={K:<3d} {rule:16s} n={want.iteration_count:4d} bt={want.backtracks:3d} "
              f"{'ok' if ok else 'DIFFERENT'}  {opts}", flush=True)
        if not ok:
            return 1
    print(f"live_check ok {count}")
    return 0


if __name__ == "__main__":
    sys.exit(main(*(int(a) for a in sys.argv[1:3])))

"""CPU, world_size 2 (gloo): the row-sharded driver's collective choreography.

Two processes each hold half the rows of A and b; the product's ``ShardedDriver`` combines the
loss partial sums and the A^T r partial vectors with all-reduce.  Both ranks must reproduce the
single-process golden trajectory (same counts, same values) and agree bitwise with each other.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def _worker(rank, world, port, case, mode, q, speculate=False):
    for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cpu_backend import CpuDenseDriver, backend_for
        from fasta import _loop
        from fasta._backends import ShardedDriver
        from fasta.distributed import row_slice
        from helpers import load_golden
        from oracle import problems

        gold = load_golden(case, mode)
        p = problems.build(case, 0)
        if rank != 0:
            np.random.seed(12345 + rank)     # ranks deliberately disagree: the driver must broadcast the probes
        rows = row_slice(p.A.shape[0], rank, world)
        local = problems.Problem(p.kind, p.loss, p.penalty, p.mu, p.x0, A=p.A[rows], b=p.b[rows])
        drv = ShardedDriver(CpuDenseDriver(local.A))
        be = backend_for(local, gold["opts"]["accelerate"], driver=drv, speculate=speculate)
        be.load()
        res = _loop.run(be, p.x0.shape, **gold["opts"])
        n = res.iteration_count
        q.put((rank, n, res.backtracks, res.solution, res.objectives[:n + 1], drv.collectives))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("speculate", [False, True])
@pytest.mark.parametrize("case,mode", [("lasso_200x1000_k50", "adaptive"), ("lasso_200x1000_k10", "accelerated"),
                                       ("logistic_1000x2000", "adaptive")])
def test_row_sharded_two_ranks_match_golden(case, mode, speculate):
    """speculate=True drives the speculative run-ahead protocol of the loop (trials queued ahead, dropped on a
    rejection / restart): both ranks must take the same decisions, or the collectives of the two would not pair up."""
    from helpers import load_golden
    gold = load_golden(case, mode)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + (7 if speculate else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, case, mode, q, speculate)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, n, bt, sol, obj, ncoll in out:
        assert n == gold["iteration_count"] and bt == gold["backtracks"]
        assert np.linalg.norm(sol - gold["solution"]) <= 1e-9 * np.linalg.norm(gold["solution"])
        assert np.max(np.abs(obj - gold["objectives"]) / np.abs(gold["objectives"])) <= 1e-10
        assert ncoll >= 2 * n
    assert np.array_equal(out[0][3], out[1][3]) and np.array_equal(out[0][4], out[1][4])   # replicas agree bitwise


def test_row_slice_partition():
    from fasta.distributed import row_slice
    for M in (1, 7, 200, 40000):
        for world in (1, 2, 3, 8):
            rows = [row_slice(M, r, world) for r in range(world)]
            assert rows[0].start == 0 and rows[-1].stop == M
            assert all(a.stop == b.start for a, b in zip(rows, rows[1:]))
            sizes = [r.stop - r.start for r in rows]
            assert max(sizes) - min(sizes) <= 1

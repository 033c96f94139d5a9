"""GPU, two ranks: the row-sharded solve with the fused peer-memory all-reduce + BB kernel (NVLink) must follow
the same golden trajectory as the single-GPU path.  Skipped on boxes with fewer than two GPUs."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, mode, peer, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "fasta-python_b200"), os.path.join(root, "tests")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), FASTA_B200_PEER="1" if peer else "0")
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import fasta
    from helpers import load_golden
    from oracle import problems
    gold = load_golden(case, mode)
    p = problems.build(case, int(gold["seed"]))
    rows = fasta.distributed.row_slice(p.A.shape[0], rank, world)
    A = fasta.distributed.RowShardedMatrix(np.ascontiguousarray(p.A[rows]))
    loss, pen = fasta.losses.LeastSquares(p.b[rows]), fasta.proximal.L1Norm(p.mu)
    res = fasta.fasta(A, loss.f, loss.gradf, pen.g, pen.prox, p.x0, **gold["opts"])
    if rank == 0:
        np.savez(out, iteration_count=res.iteration_count, backtracks=res.backtracks, solution=res.solution,
                 objectives=res.objectives, residuals=res.residuals, stepsizes=res.stepsizes,
                 norm_residuals=res.norm_residuals, peer=res.peer_reductions, single_pass=res.single_pass)
    dist.destroy_process_group()


@pytest.mark.parametrize("peer", [True, False])
@pytest.mark.parametrize("case,mode", [("lasso_200x1000_k50", "adaptive"), ("lasso_4000x10000_k500", "adaptive"),
                                       ("lasso_200x1000_k10", "plain"), ("lasso_200x1000_k50", "accelerated"),
                                       ("lasso_4000x10000_k500", "accelerated")])
def test_two_rank_sharded_solve_matches_golden(case, mode, peer, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from helpers import assert_trajectory, load_golden
    out = str(tmp_path / "rank0.npz")
    mp.spawn(_worker, args=(2, _free_port(), case, mode, peer, out), nprocs=2, join=True)
    with np.load(out) as z:
        class R:
            pass
        res = R()
        for k in z.files:
            setattr(res, k, z[k])
        res.iteration_count, res.backtracks = int(res.iteration_count), int(res.backtracks)
    assert bool(res.single_pass)
    assert (int(res.peer) > 0) == peer
    assert_trajectory(res, load_golden(case, mode), label=f"2-rank/{case}/{mode}/peer={peer}")

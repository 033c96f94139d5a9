"""GPU, two / four ranks: the row-sharded solve must follow the same golden trajectory as the single-GPU path with
each of its three exchanges -- "fused": the sweep followed by ONE kernel of ours that sums, signals, awaits and reduces
over NVLink peer memory chunk by chunk and takes the loop's decisions (fb200_dense_sweep_exchange); "peer": barrier +
fb200_peer_allreduce_bb; "nccl": ncclAllReduce + bb kernel.  Lasso (incl. backtracks, FISTA) and logistic losses.
Skipped on boxes with too few GPUs."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


EXCHANGES = {"fused": dict(FASTA_B200_PEER="1", FASTA_B200_FUSED_EXCHANGE="1"),
             "peer": dict(FASTA_B200_PEER="1", FASTA_B200_FUSED_EXCHANGE="0"),
             "nccl": dict(FASTA_B200_PEER="0", FASTA_B200_FUSED_EXCHANGE="0")}


def _worker(rank, world, port, case, mode, exchange, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "fasta-python_b200"), os.path.join(root, "tests")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), **EXCHANGES[exchange])
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
    Loss = fasta.losses.Logistic if p.loss == "logistic" else fasta.losses.LeastSquares
    loss, pen = Loss(p.b[rows]), fasta.proximal.L1Norm(p.mu)
    res = fasta.fasta(A, loss.f, loss.gradf, pen.g, pen.prox, p.x0, **gold["opts"])
    if rank == 0:
        np.savez(out, iteration_count=res.iteration_count, backtracks=res.backtracks, solution=res.solution,
                 objectives=res.objectives, residuals=res.residuals, stepsizes=res.stepsizes,
                 norm_residuals=res.norm_residuals, peer=res.peer_reductions, single_pass=res.single_pass,
                 launches=res.kernel_launches)
    dist.destroy_process_group()


CASES = [("lasso_200x1000_k50", "adaptive"), ("lasso_4000x10000_k500", "adaptive"), ("lasso_200x1000_k10", "plain"),
         ("lasso_200x1000_k50", "accelerated"), ("lasso_4000x10000_k500", "accelerated"),
         ("logistic_1000x2000", "adaptive"), ("logistic_1000x2000", "accelerated")]


def _run(world, case, mode, exchange, tmp_path):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    from helpers import assert_trajectory, load_golden
    out = str(tmp_path / "rank0.npz")
    mp.spawn(_worker, args=(world, _free_port(), case, mode, exchange, out), nprocs=world, join=True)
    with np.load(out) as z:
        class R:
            pass
        res = R()
        for k in z.files:
            setattr(res, k, z[k])
        res.iteration_count, res.backtracks = int(res.iteration_count), int(res.backtracks)
    assert bool(res.single_pass)
    assert (int(res.peer) > 0) == (exchange != "nccl")
    assert_trajectory(res, load_golden(case, mode), label=f"{world}-rank/{case}/{mode}/{exchange}")
    return res


@pytest.mark.parametrize("exchange", list(EXCHANGES))
@pytest.mark.parametrize("case,mode", CASES)
def test_two_rank_sharded_solve_matches_golden(case, mode, exchange, tmp_path):
    _run(2, case, mode, exchange, tmp_path)


@pytest.mark.parametrize("case,mode", [CASES[0], CASES[1], CASES[4], CASES[5]])
def test_four_rank_sharded_solve_matches_golden(case, mode, tmp_path):
    _run(4, case, mode, "fused", tmp_path)


def test_two_devices_in_one_process():
    """Per-device kernel attributes / plan caches (cudaFuncSetAttribute is per device): a solve on cuda:1 after one on
    cuda:0 in the SAME process, every kernel family that raises its shared-memory limit (dense sweep, streaming pair,
    resident loop, TV, device randn)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import fasta
    from helpers import assert_trajectory, load_golden
    from oracle import problems
    for case, mode, env in (("lasso_333x1414_k40", "adaptive", {}), ("lasso_200x1000_k50", "adaptive", {"FASTA_B200_RESIDENT": "0"}),
                            ("lasso_200x1000_k50", "plain", {"FASTA_B200_RESIDENT": "0", "FASTA_B200_SWEEP": "0"}),
                            ("tv_64", "accelerated", {}), ("tv_128", "plain", {"FASTA_B200_DEVICE_RNG": "force"})):
        gold = load_golden(case, mode)
        for dev in (0, 1, 0):
            old = {k: os.environ.get(k) for k in env}
            os.environ.update(env)
            try:
                with torch.cuda.device(dev):
                    p = problems.build(case, int(gold["seed"]))
                    if p.kind == "dense":
                        A = fasta.linalg.LinearMap.from_matrix(torch.from_numpy(p.A).to(f"cuda:{dev}"))
                        loss, pen = fasta.losses.LeastSquares(torch.from_numpy(p.b).to(f"cuda:{dev}")), fasta.proximal.L1Norm(p.mu)
                    else:
                        A = fasta.tv.divergence_map(p.x0.shape[:2])
                        loss, pen = fasta.losses.LeastSquares(torch.from_numpy(p.b).to(f"cuda:{dev}")), fasta.proximal.TVBall()
                    x0 = torch.from_numpy(p.x0).to(f"cuda:{dev}")
                    res = fasta.fasta(A, loss.f, loss.gradf, pen.g, pen.prox, x0, **gold["opts"])
                    assert res.solution.device.index == dev
                    res.solution = res.solution.cpu().numpy()
                    assert_trajectory(res, gold, label=f"cuda:{dev}/{case}/{mode}")
            finally:
                for k, v in old.items():
                    os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)


def _path_worker(rank, world, port, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "fasta-python_b200"), os.path.join(root, "tests")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import fasta
    from oracle import problems
    p = problems.build("lasso_200x1000_k10", 0)
    mus = p.mu * np.logspace(-0.3, 0.7, 6)
    np.random.seed(3)
    cols, full = fasta.batched.lasso_path_sharded(p.A, p.b, mus, gather=True, tolerance=1e-5, evaluate_objective=True)
    np.random.seed(3)
    mine_cols, mine = fasta.batched.lasso_path_sharded(p.A, p.b, mus, gather=False, tolerance=1e-5, evaluate_objective=True)
    assert list(mine_cols) == list(range(rank, len(mus), world)) and len(mine) == len(mine_cols)
    if rank == 0:
        np.savez(out, iters=[r.iteration_count for r in full], bts=[r.backtracks for r in full],
                 sols=np.stack([np.asarray(r.solution) for r in full]), objs=np.stack([r.objectives for r in full]))
    dist.destroy_process_group()


def test_column_sharded_regularisation_path():
    """fasta.batched.lasso_path_sharded: the columns of a path dealt round-robin over two ranks (independent units, no
    exchange during the solve) reproduce, column by column, the single-process path."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import tempfile
    import torch.multiprocessing as mp
    import fasta
    from oracle import problems
    out = os.path.join(tempfile.mkdtemp(), "path.npz")
    mp.spawn(_path_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    # (a well-conditioned path: the k50 problem at small penalties amplifies the last-bit differences between batch
    # widths into different iteration counts, like TV + adaptive)
    p = problems.build("lasso_200x1000_k10", 0)
    mus = p.mu * np.logspace(-0.3, 0.7, 6)
    np.random.seed(3)
    ref = fasta.batched.lasso_path(p.A, p.b, mus, tolerance=1e-5, evaluate_objective=True)
    with np.load(out) as z:
        assert list(z["iters"]) == [r.iteration_count for r in ref] and list(z["bts"]) == [r.backtracks for r in ref]
        for j, r in enumerate(ref):
            n = r.iteration_count
            assert np.linalg.norm(z["sols"][j] - np.asarray(r.solution)) <= 1e-9 * max(np.linalg.norm(np.asarray(r.solution)), 1e-300)
            assert np.max(np.abs(z["objs"][j][:n + 1] - r.objectives[:n + 1]) / np.abs(r.objectives[:n + 1])) <= 1e-10

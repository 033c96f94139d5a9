"""Quick device timing of the dense streaming kernels (developer tool, not the contract bench)."""
import os
import sys
import json

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fasta-python_b200")]

import torch
import fasta
from fasta import _backends, _cabi, _device


def time_op(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(iters)]
    return min(ts), sum(ts) / len(ts)


def main():
    shapes = [(20000, 50000), (40000, 100000), (5000, 100000), (100000, 20000)]
    if len(sys.argv) > 2:
        shapes = [(int(sys.argv[1]), int(sys.argv[2]))]
    out = []
    for M, N in shapes:
        A = torch.randn(M, N, dtype=torch.float64, device="cuda")
        x = torch.randn(N, dtype=torch.float64, device="cuda")
        b = torch.randn(M, dtype=torch.float64, device="cuda")
        z, r = torch.empty_like(b), torch.empty_like(b)
        g = torch.empty_like(x)
        drv = _backends.DenseDriver(A)
        ws = _device.Workspace(M, N)
        nbytes = M * N * 8
        t_min, t_avg = time_op(lambda: drv.forward(x, _cabi.LOSS_LEAST_SQUARES, b, z, r, ws))
        rec = dict(M=M, N=N, gemv_ms=t_min, gemv_avg_ms=t_avg, gemv_GBs=nbytes / t_min / 1e6)
        t_min, t_avg = time_op(lambda: drv.adjoint(r, g, 1, None, None, None, 0.0, ws))
        rec.update(gemvT_ms=t_min, gemvT_avg_ms=t_avg, gemvT_GBs=nbytes / t_min / 1e6)
        if drv.sweep_ok:
            t_min, t_avg = time_op(lambda: drv.sweep(x, _cabi.LOSS_LEAST_SQUARES, b, z, r, g, 1, None, None, None, 0.0, ws))
            rec.update(sweep_ms=t_min, sweep_avg_ms=t_avg, sweep_dram_GBs=nbytes / t_min / 1e6,
                       sweep_algorithmic_GBs=2 * nbytes / t_min / 1e6, sweep_cluster=drv.sweep_cluster)
            import ctypes
            plan = (ctypes.c_int * 5)()
            _cabi.load().fb200_sweep_plan(M, N, plan)
            rec.update(sweep_plan=dict(cs=plan[0], clusters=plan[1], stages=plan[2], nc=plan[3], cpt=plan[4]))
            g2 = torch.mv(A.T, torch.mv(A, x) - b)
            rec.update(sweep_g_relerr=float((g - g2).norm() / g2.norm()))
        t_min, _ = time_op(lambda: torch.mv(A, x))
        rec.update(torch_mv_ms=t_min, torch_mv_GBs=nbytes / t_min / 1e6)
        t_min, _ = time_op(lambda: torch.mv(A.T, b))
        rec.update(torch_mvT_ms=t_min, torch_mvT_GBs=nbytes / t_min / 1e6)
        err = float((z - torch.mv(A, x)).norm() / z.norm())
        rec.update(gemv_relerr=err)
        print(json.dumps(rec), flush=True)
        out.append(rec)
        del A
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

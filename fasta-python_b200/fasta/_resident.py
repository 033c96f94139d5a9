"""Device-resident solve for small dense problems: the prologue as in ``_loop.run``, then the WHOLE
forward-backward-splitting loop (reference ``fasta/__init__.py:172-313``) in one cooperative kernel
(``fb200_resident_fbs``, csrc/resident_loop.cu) -- no host round trip per iteration.

Used by ``fasta.fasta`` when the problem is eligible: dense tagged operator that fits the L2, tagged loss and
elementwise penalty, a built-in stop rule, no per-iterate hooks (all three modes: plain, adaptive, FISTA).  Everything else takes the
host-driven loop.  ``FASTA_B200_RESIDENT=0`` disables it.
"""

import os
from time import time

import numpy as np

from . import _cabi, _device, stopping
from ._loop import Convergence, lipschitz_probes

_RULES = {stopping.residual: 0, stopping.norm_residual: 1, stopping.ratio_residual: 2, stopping.hybrid_residual: 3}
_PROX_OK = (_cabi.PROX_SHRINK, _cabi.PROX_NONNEG, _cabi.PROX_BOX, _cabi.PROX_IDENTITY)


def eligible(be, opts):
    from . import _backends
    if os.environ.get("FASTA_B200_RESIDENT", "1") == "0":
        return False
    if not isinstance(be, _backends.FusedBackend) or not isinstance(be.drv, _backends.DenseDriver):
        return False
    if opts.get("record_iterates", False) or opts.get("func", None) is not None:
        return False
    if opts.get("stop_rule", stopping.hybrid_residual) not in _RULES:
        return False
    if be.loss.tag not in (_cabi.LOSS_LEAST_SQUARES, _cabi.LOSS_LOGISTIC) or be.pen.tag not in _PROX_OK:
        return False
    if be.pen.tag == _cabi.PROX_SHRINK and np.ndim(be.pen.mu) != 0:
        return False
    if opts.get("max_iters", 1000) < 1:
        return False
    return int(_cabi.load().fb200_resident_blocks(be.drv.M, be.drv.N)) > 0


def run(be, x0_shape, *, adaptive=True, accelerate=False, verbose=True, max_iters=1000, tolerance=1e-5,
        stop_rule=stopping.hybrid_residual, L=None, tau0=None, backtrack=True, stepsize_shrink=None,
        window=10, max_backtracks=20, restart=True, evaluate_objective=False, record_iterates=False,
        func=None) -> Convergence:
    t = _device.torch()
    lib = _cabi.load()
    S = _cabi
    if stepsize_shrink is None and backtrack:               # ref :92-97
        stepsize_shrink = 0.2 if adaptive else 0.5
    if not L or not tau0:                                   # ref :100-113, pipelined as in _loop.run
        dgrad, dpoint = lipschitz_probes(be, x0_shape)
        L = dgrad / dpoint
        tau0 = (2 / L) / 10
    if not tau0:
        tau0 = 1 / L
    if verbose:                                             # ref :118-120
        print("Initializing FASTA...\n")
        print("Iteration #\tResidual\tStepsize\tAccel. param\tBacktracks\tObjective")

    s = be.start()                                          # ref :135-143
    dev = be.X[0].device
    # every history the kernel fills lives in ONE zero-initialised device block (one memset, one D2H copy at the end):
    # doubles resid | nresid | tau | alpha [max_iters each] | f | obj [max_iters + 1 each] | out [4] | clk (int64)
    # [max_iters + 1] | bt (int32) [max_iters, padded to a multiple of 8 bytes]
    mi = int(max_iters)
    nd = 4 * mi + 2 * (mi + 1) + 4 + (mi + 1) + (mi + 1) // 2 + 1
    block = t.zeros(nd, dtype=t.float64, device=dev)
    o = 0
    resid_d, nresid_d, tau_d, alpha_all = (block[o + k * mi: o + (k + 1) * mi] for k in range(4))
    o += 4 * mi
    f_d, obj_d = block[o: o + mi + 1], block[o + mi + 1: o + 2 * (mi + 1)]
    o += 2 * (mi + 1)
    out_d = block[o: o + 4]
    o += 4
    clk_d = block[o: o + mi + 1].view(t.int64)
    o += mi + 1
    bt_d = block[o:].view(t.int32)[:mi]
    part = t.empty(int(lib.fb200_resident_scratch_doubles(be.drv.M, be.drv.N)), dtype=t.float64, device=dev)
    head = be.ws.stage(0, 2)
    head[0] = float(s.f)
    head[1] = float(s.f + s.pen) if evaluate_objective else 0.0
    f_d[0:1].copy_(head[0:1], non_blocking=True)
    obj_d[0:1].copy_(head[1:2], non_blocking=True)

    pen = be.pen
    mu = float(pen.mu) if pen.tag == S.PROX_SHRINK else 0.0
    lo, hi = (float(pen.lo), float(pen.hi)) if pen.tag == S.PROX_BOX else (0.0, 0.0)
    ia, ib_, ibest = be.ic, (be.ic + 1) % 4, (be.ic + 2) % 4
    xa, xb, best = be.X[ia], be.X[ib_], be.X[ibest]
    best.copy_(xa)                                          # stays the answer if no iterate ever improves (nan quality)
    ga, gb = be.G[be.gc], be.G[(be.gc + 1) % 3]
    alpha_d = alpha_all if accelerate else None
    fista = [be.XA[be.ac], be.XA[(be.ac + 1) % 3], be.ZA[be.ac], be.ZA[(be.ac + 1) % 3], alpha_d] if accelerate else [None] * 5
    t_launch = time()
    _cabi.check(lib.fb200_resident_fbs(
        be.drv.A.data_ptr(), be.drv.lda, be.drv.M, be.drv.N, be.loss.b.data_ptr(), be.loss.tag, pen.tag, mu, lo, hi,
        xa.data_ptr(), xb.data_ptr(), ga.data_ptr(), gb.data_ptr(), be.XH.data_ptr(), be.DX.data_ptr(), best.data_ptr(),
        be.Z.data_ptr(), be.R.data_ptr(), part.data_ptr(), resid_d.data_ptr(), nresid_d.data_ptr(), tau_d.data_ptr(),
        f_d.data_ptr(), obj_d.data_ptr(), bt_d.data_ptr(), clk_d.data_ptr(), out_d.data_ptr(), float(tau0), float(s.g_sq),
        float(tolerance), float(stepsize_shrink if stepsize_shrink is not None else 1.0), int(bool(adaptive)),
        int(bool(backtrack)), int(window), int(max_backtracks), int(max_iters), _RULES[stop_rule],
        int(bool(evaluate_objective)), int(bool(accelerate)), int(bool(restart)), *[_device.ptr(v) for v in fista],
        _device.stream_ptr()), "fb200_resident_fbs")
    be.launches += 1
    host = t.empty(nd, dtype=t.float64).pin_memory() if getattr(be.ws, "_resident_host", None) is None or be.ws._resident_host.numel() != nd \
        else be.ws._resident_host
    be.ws._resident_host = host
    host.copy_(block, non_blocking=True)
    _device.current_stream().synchronize()                  # the one sync of the solve
    hb = host.numpy().copy()                                # the pinned mirror is reused by the next solve
    o = 0
    residual_hist, norm_residual_hist, tau_hist, alphas_all = (hb[o + k * mi: o + (k + 1) * mi] for k in range(4))
    o += 4 * mi
    f_hist_h, obj_h = hb[o: o + mi + 1], hb[o + mi + 1: o + 2 * (mi + 1)]
    o += 2 * (mi + 1)
    out = hb[o: o + 4]
    o += 4
    clk = hb[o: o + mi + 1].view(np.int64).astype(np.float64)
    o += mi + 1
    bts_all = hb[o:].view(np.int32)[:mi]
    n = int(out[0])
    be.ic, be.ip = (ib_, ia) if int(out[2]) != 0 else (ia, ib_)      # which buffer holds the last iterate
    be.ib = ibest                                           # the kernel copies an improving iterate there (N is small)
    objective_hist = obj_h if evaluate_objective else None
    times = np.zeros(max_iters + 1)
    times[:n + 1] = t_launch + (clk[:n + 1] - clk[0]) * 1e-9     # %globaltimer stamps mapped onto the host clock
    if verbose:                                             # ref :235,302-306, printed after the fact
        bts = bts_all
        alphas = alphas_all if accelerate else None
        for i in range(n):
            if bts[i] & (1 << 30):
                print("Restarted acceleration.")
            print("[{:<6}]\t{:e}\t{:e}\t{:e}\t{:6}\t{:e}".format(
                i, residual_hist[i], tau_hist[i], alphas[i] if accelerate else 0.0,
                int(bts[i] & 0xFFFFF) if backtrack else 0, objective_hist[i] if evaluate_objective else 0))
    res = Convergence(residual_hist, norm_residual_hist, tau_hist, int(out[1]), times, n, be.solution(),
                      objective_hist, None, None)
    res.resident = True
    res.resident_cluster = bool(out[3])                     # single-cluster variant: A stayed in shared memory
    return res

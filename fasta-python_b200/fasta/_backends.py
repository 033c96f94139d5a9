"""Device back-ends of the FBS loop (see ``_loop.py`` for the protocol).

``FusedBackend``   tagged operator + tagged loss + tagged penalty.  Dense maps: an iteration is
                   fbs_step -> single-pass sweep (z = A x, loss, g = A^H r in ONE read of A) -> BB sums ->
                   device-side decisions -> snapshot, queued ahead of the host's bookkeeping (two-pass
                   fallback: fbs_step -> A x + loss -> sync -> A^H r + BB -> sync); TV: one kernel per
                   trial.  All hand-written sm_100a kernels behind the C ABI (include/fasta_b200.h).
``GenericBackend`` arbitrary user callables (f, gradf, g, proxg, A) that accept torch CUDA
                   tensors; the library's own pieces (forward step, reductions, extrapolation,
                   dense contractions) still run as the same kernels.
Operator drivers:  ``DenseDriver`` (single-pass sweep, TMA streaming GEMV / GEMV-T), ``TVDriver`` (stencils),
                   ``ShardedDriver`` (row-partitioned A over torch.distributed ranks: the local sweep + ONE
                   exchange kernel over NVLink peer memory; barrier + all-reduce kernel or NCCL as fallbacks).
There is no CPU implementation anywhere in this module.
"""

import os

import numpy as np

from . import _cabi, _device
from ._loop import Scalars

S = _cabi


# =================================================================================================
# operator drivers: z = A x (+ loss epilogue) and g = A^H r (+ BB epilogue) on raw device tensors
# =================================================================================================
class DenseDriver:
    """Dense row-major M x N matrix in HBM (reference linalg.py:37-41)."""

    def __init__(self, matrix, lda=None):
        self.A = matrix
        self.M, self.N = int(matrix.shape[0]), int(matrix.shape[1])
        self.lda = int(lda if lda is not None else matrix.stride(0))
        self.xshape, self.zshape = (self.N,), (self.M,)
        self.lib = _cabi.load()
        self.launches = 0
        # single-pass sweep (csrc/dense_sweep.cu): cluster size it would use, 0 = not eligible
        self.sweep_cluster = int(self.lib.fb200_sweep_supported(self.A.data_ptr(), self.lda, self.M, self.N))
        self.sweep_ok = self.sweep_cluster > 0 and os.environ.get("FASTA_B200_SWEEP", "1") != "0"

    def workspace_dims(self):
        return self.M, self.N

    def sweep(self, x, loss_tag, b, z, r, g, bb, x0, xhat, dx, tau, ws):
        """z = A x, r = gradf(z), S_F and g = A^T r (+BB epilogue) in ONE pass over A."""
        _cabi.check(self.lib.fb200_dense_sweep(self.A.data_ptr(), self.lda, self.M, self.N, x.data_ptr(), loss_tag,
                                               _device.ptr(b), z.data_ptr(), _device.ptr(r), _device.ptr(g), bb,
                                               _device.ptr(x0), _device.ptr(xhat), _device.ptr(dx), float(tau),
                                               ws.scal.data_ptr(), ws.buf.data_ptr(), ws.nbytes,
                                               _device.stream_ptr()), "fb200_dense_sweep")
        self.launches += 3

    def sweep_accel(self, xa1, loss_tag, b, za0, c, za1, z, r, g, bb, x0, xhat, dx, tau, ws):
        """FISTA mode of the single pass: za1 = A xa1, z = za1 + c (za1 - za0), r = gradf(z), g = A^T r;
        S_F = f(za1) (line search), S_AUX3 = f(z)."""
        _cabi.check(self.lib.fb200_dense_sweep_accel(self.A.data_ptr(), self.lda, self.M, self.N, xa1.data_ptr(), loss_tag,
                                                     _device.ptr(b), za0.data_ptr(), float(c), za1.data_ptr(), z.data_ptr(),
                                                     _device.ptr(r), _device.ptr(g), bb, _device.ptr(x0), _device.ptr(xhat),
                                                     _device.ptr(dx), float(tau), ws.scal.data_ptr(), ws.buf.data_ptr(),
                                                     ws.nbytes, _device.stream_ptr()), "fb200_dense_sweep_accel")
        self.launches += 4

    def forward(self, x, loss_tag, b, z, r, ws):
        _cabi.check(self.lib.fb200_gemv_loss(self.A.data_ptr(), self.lda, self.M, self.N, x.data_ptr(), loss_tag,
                                             _device.ptr(b), z.data_ptr(), _device.ptr(r), ws.scal.data_ptr(),
                                             ws.buf.data_ptr(), ws.nbytes, _device.stream_ptr()), "fb200_gemv_loss")
        self.launches += 2

    def adjoint(self, r, g, bb, x0, xhat, dx, tau, ws):
        _cabi.check(self.lib.fb200_gemvT_bb(self.A.data_ptr(), self.lda, self.M, self.N, r.data_ptr(), g.data_ptr(),
                                            bb, _device.ptr(x0), _device.ptr(xhat), _device.ptr(dx), float(tau),
                                            ws.scal.data_ptr(), ws.buf.data_ptr(), ws.nbytes, _device.stream_ptr()),
                    "fb200_gemvT_bb")
        self.launches += 2

    def bb_reduce(self, g, x0, xhat, dx, tau, adaptive, ws):
        _cabi.check(self.lib.fb200_bb_reduce(g.data_ptr(), _device.ptr(x0), _device.ptr(xhat), _device.ptr(dx),
                                             float(tau), g.numel(), 1 if adaptive else 0, ws.scal.data_ptr(),
                                             ws.buf.data_ptr(), _device.stream_ptr()), "fb200_bb_reduce")
        self.launches += 1

    def sync_point(self, v1, v2):
        pass


class TVDriver:
    """A = div, A^H = grad on an n0 x n1 image (reference tv_denoising.py:26-63,99)."""

    def __init__(self, n0, n1):
        self.n0, self.n1 = int(n0), int(n1)
        self.xshape, self.zshape = (self.n0, self.n1, 2), (self.n0, self.n1)
        self.lib = _cabi.load()
        self.launches = 0

    def workspace_dims(self):
        return 1, 1

    def forward(self, x, loss_tag, b, z, r, ws):
        _cabi.check(self.lib.fb200_tv_div_loss(x.data_ptr(), self.n0, self.n1, loss_tag, _device.ptr(b), z.data_ptr(),
                                               _device.ptr(r), ws.scal.data_ptr(), ws.buf.data_ptr(),
                                               _device.stream_ptr()), "fb200_tv_div_loss")
        self.launches += 1

    def adjoint(self, r, g, bb, x0, xhat, dx, tau, ws):
        _cabi.check(self.lib.fb200_tv_grad_bb(r.data_ptr(), self.n0, self.n1, g.data_ptr(), bb, _device.ptr(x0),
                                              _device.ptr(xhat), _device.ptr(dx), float(tau), ws.scal.data_ptr(),
                                              ws.buf.data_ptr(), _device.stream_ptr()), "fb200_tv_grad_bb")
        self.launches += 1

    # fused iteration (non-accelerated modes): step + ball projection + div + loss in one pass, and
    # grad + BB with xhat / dx recomputed -- xhat, dx and z are never materialised
    fused_step_ok = os.environ.get("FASTA_B200_TV_FUSED", "1") != "0"

    def step_forward(self, x0, g0, tau, loss_tag, b, x1, r, ws):
        _cabi.check(self.lib.fb200_tv_step_div_loss(x0.data_ptr(), g0.data_ptr(), float(tau), self.n0, self.n1, loss_tag,
                                                    b.data_ptr(), x1.data_ptr(), r.data_ptr(), ws.scal.data_ptr(),
                                                    ws.buf.data_ptr(), _device.stream_ptr()), "fb200_tv_step_div_loss")
        self.launches += 1

    def adjoint_fused(self, r, g, bb, x0, g0, x1, tau, ws):
        _cabi.check(self.lib.fb200_tv_grad_bb_fused(r.data_ptr(), self.n0, self.n1, g.data_ptr(), bb, x0.data_ptr(),
                                                    g0.data_ptr(), x1.data_ptr(), float(tau), ws.scal.data_ptr(),
                                                    ws.buf.data_ptr(), _device.stream_ptr()), "fb200_tv_grad_bb_fused")
        self.launches += 1

    # the whole iteration in one kernel: step + projection + div + loss + speculative grad + BB sums (9U bytes)
    iter_fused_ok = os.environ.get("FASTA_B200_TV_ITER", "1") != "0"

    def iterate_fused(self, x0, g0, tau, loss_tag, b, x1, g1, ws):
        _cabi.check(self.lib.fb200_tv_iter_fused(x0.data_ptr(), g0.data_ptr(), float(tau), self.n0, self.n1, loss_tag,
                                                 b.data_ptr(), x1.data_ptr(), g1.data_ptr(), ws.scal.data_ptr(),
                                                 ws.buf.data_ptr(), _device.stream_ptr()), "fb200_tv_iter_fused")
        self.launches += 1

    # the whole accelerated (FISTA) trial in one kernel (15U bytes), extrapolation weight c speculated by the caller
    fista_fused_ok = os.environ.get("FASTA_B200_TV_FISTA", "1") != "0"

    def fista_fused(self, x0, g0, tau, c, loss_tag, b, xa0, za0, xa1, za1, x1, g1, ws):
        _cabi.check(self.lib.fb200_tv_fista_fused(x0.data_ptr(), g0.data_ptr(), float(tau), float(c), self.n0, self.n1,
                                                  loss_tag, b.data_ptr(), xa0.data_ptr(), za0.data_ptr(), xa1.data_ptr(),
                                                  za1.data_ptr(), x1.data_ptr(), g1.data_ptr(), ws.scal.data_ptr(),
                                                  ws.buf.data_ptr(), _device.stream_ptr()), "fb200_tv_fista_fused")
        self.launches += 1

    def sync_point(self, v1, v2):
        pass


class ShardedDriver:
    """Row-partitioned map: this rank holds a row block of A (and of b) in ``local``.

    A x is local; the loss partial sum and the A^H r partial vector are combined with
    ``torch.distributed.all_reduce`` (NCCL over NVLink on GPUs; gloo in the CPU tests of the
    host logic).  x, the gradient, the step size and all histories are replicated, so every rank
    runs the identical scalar control flow.
    """

    def __init__(self, local, group=None):
        import torch.distributed as dist
        self.local = local
        self.dist = dist
        self.group = group
        self.xshape, self.zshape = local.xshape, local.zshape
        self.collectives = 0
        self.peer_reductions = 0

    @property
    def launches(self):
        return self.local.launches

    @property
    def sweep_ok(self):
        return getattr(self.local, "sweep_ok", False)

    def workspace_dims(self):
        return self.local.workspace_dims()

    # -- peer-memory path: the partial gradients live in NVLink-mapped ("symmetric") buffers, and ONE kernel of ours
    # (fb200_peer_allreduce_bb) reads every rank's partial, adds them in rank order and forms the BB sums.  torch
    # supplies the mapping and the barrier; two buffers alternate so that one barrier per call suffices (a rank
    # rewrites a buffer only after a later barrier, which every rank reaches after it finished reading).
    _peer_cache = {}
    peer_ok = os.environ.get("FASTA_B200_PEER", "1") != "0"

    PEER_CHUNKS = 128          # csrc/vector_kernels.cu: blocks (= flag words per rank) of the exchange kernel
    fused_ok = os.environ.get("FASTA_B200_FUSED_EXCHANGE", "1") != "0"

    def _peer(self, n, device):
        key = (id(self.group), int(n), device.index)
        if key in ShardedDriver._peer_cache:
            return ShardedDriver._peer_cache[key]
        st = None
        try:
            if self.peer_ok and device.type == "cuda" and self.dist.get_backend(self.group) == "nccl" \
                    and self.dist.get_world_size(self.group) <= 16:
                import ctypes
                import torch.distributed._symmetric_memory as symm
                t = _device.torch()
                pitch = (int(n) + 8 + 1) // 2 * 2
                P0 = self.dist.get_world_size(self.group)
                flag_doubles = P0 * self.PEER_CHUNKS // 2            # P x 128 32-bit words behind the two data slots
                buf = symm.empty(2 * pitch + flag_doubles, dtype=t.float64, device=device)
                hdl = symm.rendezvous(buf, self.group if self.group is not None else self.dist.group.WORLD)
                P = hdl.world_size
                ptrs = [(ctypes.c_uint64 * P)(*[int(hdl.buffer_ptrs[k]) + 8 * pitch * slot for k in range(P)])
                        for slot in (0, 1)]
                flags = (ctypes.c_uint64 * P)(*[int(hdl.buffer_ptrs[k]) + 8 * 2 * pitch for k in range(P)])
                buf.zero_()
                hdl.barrier(channel=0)
                st = dict(buf=buf, hdl=hdl, P=P, pitch=pitch, ptrs=ptrs, flags=flags, calls=0, rank=int(hdl.rank),
                          di=(ctypes.c_int * 8)(), dd=(ctypes.c_double * 3)())
        except Exception as exc:                       # no peer mapping on this system: NCCL path below
            import warnings
            warnings.warn(f"fasta-b200: peer-memory all-reduce unavailable ({exc}); using ncclAllReduce")
            st = None
        ShardedDriver._peer_cache[key] = st
        return st

    # -- fused path (csrc: fb200_dense_sweep_exchange): the local sweep and ONE kernel that sums the band partials into
    # the mapped buffer, signals / awaits the peers chunk by chunk, adds the P partials in rank order, forms the BB sums
    # and takes the loop's decisions -- no scalar copy, no barrier kernel, no separate all-reduce / decide launches
    def _fused(self, g):
        if not (self.fused_ok and g.is_cuda and self.sweep_ok and hasattr(self.local, "A")):
            return None
        peer = self._peer(g.numel(), g.device)
        if peer is None or not self.local.lib.fb200_sweep_exchange_supported():
            return None
        return peer

    @property
    def fuses_decide(self):
        """Whether sweep(..., decide=...) takes the loop's decisions itself (then no fb200_trial_decide follows)."""
        t = _device.torch()
        probe = getattr(self, "_fused_probe", None)
        if probe is None:
            A = getattr(self.local, "A", None)
            probe = self._fused_probe = bool(A is not None and A.is_cuda and self.fused_ok and self.sweep_ok
                                             and self._peer(self.local.N, A.device) is not None
                                             and self.local.lib.fb200_sweep_exchange_supported())
        return probe

    def _sweep_exchange(self, peer, x, loss_tag, b, z, r, za0, c, za1, g, bb, x0, xhat, dx, tau, ws, decide, host_out=0):
        peer["calls"] += 1
        epoch = peer["calls"]
        di = dd = None
        if decide is not None:
            di, dd = peer["di"], peer["dd"]
            di[:] = [int(v) for v in decide[:8]]
            dd[:] = [float(v) for v in decide[8:11]]
        L = self.local
        _cabi.check(L.lib.fb200_dense_sweep_exchange(L.A.data_ptr(), L.lda, L.M, L.N, x.data_ptr(), loss_tag, _device.ptr(b),
                                                     z.data_ptr(), _device.ptr(r), _device.ptr(za0), float(c), _device.ptr(za1),
                                                     peer["ptrs"][epoch & 1], peer["flags"], peer["rank"], peer["P"],
                                                     epoch & 0xFFFFFFFF, g.data_ptr(), int(bb), _device.ptr(x0),
                                                     _device.ptr(xhat), _device.ptr(dx), float(tau), di, dd, host_out,
                                                     ws.scal.data_ptr(), ws.buf.data_ptr(), ws.nbytes, _device.stream_ptr()),
                    "fb200_dense_sweep_exchange")
        L.launches += 2
        self.collectives += 1
        self.peer_reductions += 1

    def sweep(self, x, loss_tag, b, z, r, g, bb, x0, xhat, dx, tau, ws, decide=None, host_out=0):
        n = g.numel()
        fused = self._fused(g)
        if fused is not None:
            self._sweep_exchange(fused, x, loss_tag, b, z, r, None, 0.0, None, g, bb, x0, xhat, dx, tau, ws, decide, host_out)
            return
        assert decide is None
        peer = self._peer(n, g.device) if g.is_cuda else None
        if peer is not None:
            slot = peer["calls"] & 1
            peer["calls"] += 1
            part = peer["buf"][slot * peer["pitch"]: slot * peer["pitch"] + n + 1]
            # (a speculative trial, tau = NaN, must reach the sweep as such: it returns at once when it should not run)
            self.local.sweep(x, loss_tag, b, z, r, part[:n], 0, None, None, None, tau if tau != tau else 0.0, ws)
            with_loss = 1 if loss_tag != S.LOSS_NONE else 0
            if with_loss:
                part[n:n + 1].copy_(ws.scal[S.S_F:S.S_F + 1])
            peer["hdl"].barrier(channel=0)
            _cabi.check(self.local.lib.fb200_peer_allreduce_bb(peer["ptrs"][slot], peer["P"], n, g.data_ptr(), int(bb),
                                                               _device.ptr(x0), _device.ptr(xhat), _device.ptr(dx),
                                                               float(tau), with_loss, ws.scal.data_ptr(), ws.buf.data_ptr(),
                                                               _device.stream_ptr()), "fb200_peer_allreduce_bb")
            self.local.launches += 1
            self.collectives += 1
            self.peer_reductions += 1
            return
        self.local.sweep(x, loss_tag, b, z, r, g, 0, None, None, None, tau if tau != tau else 0.0, ws)
        base = getattr(g, "_base", None)
        packed = None
        if loss_tag != S.LOSS_NONE and base is not None and base.numel() > n and base.data_ptr() == g.data_ptr():
            packed = base[:n + 1]                       # [g ; raw loss partial] in one message
            packed[n:n + 1].copy_(ws.scal[S.S_F:S.S_F + 1])
            self.dist.all_reduce(packed, group=self.group)
            ws.scal[S.S_F:S.S_F + 1].copy_(packed[n:n + 1])
            self.collectives += 1
        else:
            self.dist.all_reduce(g, group=self.group)
            self.collectives += 1
            if loss_tag != S.LOSS_NONE:
                self.reduce_loss(ws)
        if bb:
            self.local.bb_reduce(g, x0, xhat, dx, tau, bb >= 2, ws)

    def sweep_accel(self, xa1, loss_tag, b, za0, c, za1, z, r, g, bb, x0, xhat, dx, tau, ws):
        """FISTA mode of the single pass on the row shard (see DenseDriver.sweep_accel): the partial gradient and BOTH
        loss partials (prox point: line search; extrapolated point) are summed over the ranks in one exchange."""
        n = g.numel()
        fused = self._fused(g)
        if fused is not None:
            self._sweep_exchange(fused, xa1, loss_tag, b, z, r, za0, c, za1, g, bb, x0, xhat, dx, tau, ws, None)
            return
        peer = self._peer(n, g.device) if g.is_cuda else None
        if peer is not None:
            slot = peer["calls"] & 1
            peer["calls"] += 1
            part = peer["buf"][slot * peer["pitch"]: slot * peer["pitch"] + n + 2]
            self.local.sweep_accel(xa1, loss_tag, b, za0, c, za1, z, r, part[:n], 0, None, None, None, 0.0, ws)
            part[n:n + 1].copy_(ws.scal[S.S_F:S.S_F + 1])
            part[n + 1:n + 2].copy_(ws.scal[S.S_AUX3:S.S_AUX3 + 1])
            peer["hdl"].barrier(channel=0)
            _cabi.check(self.local.lib.fb200_peer_allreduce_bb(peer["ptrs"][slot], peer["P"], n, g.data_ptr(), int(bb),
                                                               _device.ptr(x0), _device.ptr(xhat), _device.ptr(dx),
                                                               float(tau), 2, ws.scal.data_ptr(), ws.buf.data_ptr(),
                                                               _device.stream_ptr()), "fb200_peer_allreduce_bb")
            self.local.launches += 1
            self.collectives += 1
            self.peer_reductions += 1
            return
        self.local.sweep_accel(xa1, loss_tag, b, za0, c, za1, z, r, g, 0, None, None, None, 0.0, ws)
        base = getattr(g, "_base", None)
        if base is not None and base.numel() >= n + 2 and base.data_ptr() == g.data_ptr():
            packed = base[:n + 2]                       # [g ; f(prox point) ; f(extrapolated point)] in one message
            packed[n:n + 1].copy_(ws.scal[S.S_F:S.S_F + 1])
            packed[n + 1:n + 2].copy_(ws.scal[S.S_AUX3:S.S_AUX3 + 1])
            self.dist.all_reduce(packed, group=self.group)
            ws.scal[S.S_F:S.S_F + 1].copy_(packed[n:n + 1])
            ws.scal[S.S_AUX3:S.S_AUX3 + 1].copy_(packed[n + 1:n + 2])
            self.collectives += 1
        else:
            self.dist.all_reduce(g, group=self.group)
            self.dist.all_reduce(ws.scal[S.S_F:S.S_F + 1], group=self.group)
            self.dist.all_reduce(ws.scal[S.S_AUX3:S.S_AUX3 + 1], group=self.group)
            self.collectives += 3
        if bb:
            self.local.bb_reduce(g, x0, xhat, dx, tau, bb >= 2, ws)

    def _root(self):
        return self.dist.get_global_rank(self.group, 0) if self.group is not None else 0

    def forward(self, x, loss_tag, b, z, r, ws):
        self.local.forward(x, loss_tag, b, z, r, ws)
        if loss_tag != S.LOSS_NONE:
            self.reduce_loss(ws)

    def adjoint(self, r, g, bb, x0, xhat, dx, tau, ws):
        self.local.adjoint(r, g, 0, None, None, None, 0.0, ws)
        self.dist.all_reduce(g, group=self.group)
        self.collectives += 1
        if bb:
            self.local.bb_reduce(g, x0, xhat, dx, tau, bb >= 2, ws)

    def reduce_loss(self, ws):
        self.dist.all_reduce(ws.scal[S.S_F:S.S_F + 1], group=self.group)
        self.collectives += 1

    def sync_point(self, v1, v2):
        # every rank must use the same two random probes for the Lipschitz estimate
        self.dist.broadcast(v1, src=self._root(), group=self.group)
        self.dist.broadcast(v2, src=self._root(), group=self.group)

    def sync_probe(self, v):
        self.dist.broadcast(v, src=self._root(), group=self.group)

    def sync_state(self, s):
        # device-side probes (fasta/_rng.py): every rank continues rank 0's numpy stream -- one 2.5 KB message
        self.dist.broadcast(s, src=self._root(), group=self.group)
        self.collectives += 1


# =================================================================================================
# fused back-end
# =================================================================================================
class FusedBackend:
    def __init__(self, driver, loss, penalty, x0, accelerate):
        t = _device.torch()
        self.t = t
        self.lib = _cabi.load()
        self.drv = driver
        self.loss = loss
        self.pen = penalty
        self.accelerate = bool(accelerate)
        self.x0_in = x0
        self.shape = tuple(x0.shape)
        assert self.shape == tuple(driver.xshape), f"x0 shape {self.shape} != operator domain {driver.xshape}"
        assert tuple(loss.b.shape) == tuple(driver.zshape), "loss data does not match the operator range"
        dev = loss.b.device
        self.n = int(np.prod(driver.xshape))
        self.m = int(np.prod(driver.zshape))
        self.ws = _device.acquire_workspace(*driver.workspace_dims(), device=dev)
        new = lambda k: t.empty(k, dtype=t.float64, device=dev)
        # four iterate buffers: the best iterate so far (reference :298-300) is kept by INDEX, never copied, and a trial
        # queued ahead of the line-search decision must not overwrite the x0 of the trial it speculates on: a new trial
        # is written to the buffer that is none of {its x0, the previous x0, the best}
        self.X = [new(self.n) for _ in range(4)]
        # gradient buffers carry 8 spare doubles: a sharded driver packs the loss partial behind the
        # gradient so that ONE all-reduce per iteration moves both
        # (three of them: a trial queued ahead of the line-search decision must not overwrite the gradient at x0)
        self._Gbuf = [new(self.n + 8), new(self.n + 8), new(self.n + 8)]
        self.G = [g[:self.n] for g in self._Gbuf]
        self.XH, self.DX = new(self.n), new(self.n)
        self.Z, self.R = new(self.m), new(self.m)
        if self.accelerate:
            # three each: a FISTA trial queued ahead must not overwrite the x_accel0 / z_accel0 of the trial before it,
            # which is repeated with weight 0 when its restart test fires
            self.XA = [new(self.n) for _ in range(3)]
            self.ZA = [new(self.m) for _ in range(3)]
        self.ic = self.ip = 0      # X[ic] current iterate, X[ip] previous
        self.ib = 0                # X[ib] best iterate so far
        self.gc = self.gp = 0      # G[gc] current gradient, G[gp] previous
        self.ac = self.ap = 0      # XA/ZA[ac] current prox point, [ap] previous
        self.launches = 0          # kernels launched by this backend (vector kernels; + driver.launches)
        # Single-pass mode: every trial also produces g = A^T gradf(A x1) speculatively in the same
        # pass over A (the line search accepts the first trial in the vast majority of iterations;
        # a rejected trial costs exactly what the two-pass path would have paid for it).
        self.use_sweep = (not self.accelerate) and bool(getattr(driver, "sweep_ok", False))
        # FISTA: the extrapolated z is a per-row function of A xa1 and the previous prox image, so the same single
        # pass serves the accelerated mode once the restart decision (one scalar) is known
        self.use_sweep_accel = (self.accelerate and bool(getattr(driver, "sweep_ok", False))
                                and hasattr(driver, "sweep_accel") and loss.tag != S.LOSS_NONE
                                and os.environ.get("FASTA_B200_SWEEP_ACCEL", "1") != "0")
        # Lipschitz prologue: one probe pass instead of two when the gradient of the loss is affine (see lipschitz_push)
        self.affine_probe = (loss.tag == S.LOSS_LEAST_SQUARES
                             and os.environ.get("FASTA_B200_AFFINE_PROBE", "1") != "0")
        self._spec = None
        self._ahead = False
        self._pending = None
        self._spec_mode = False
        self._decide = None
        # TV: one fused kernel per half-iteration (see TVDriver.step_forward)
        self.use_tv_fused = ((not self.accelerate) and isinstance(driver, TVDriver) and driver.fused_step_ok
                             and penalty.tag == S.PROX_TV_BALL and loss.tag != S.LOSS_NONE)
        # Speculative run-ahead: the step size of the next trial stays on the device (fb200_trial_decide), so the next
        # iteration's trial is queued BEFORE the host has seen this trial's sums; see _loop.run
        tv_iter = self.use_tv_fused and driver.iter_fused_ok
        elementwise = penalty.tag in (S.PROX_NONNEG, S.PROX_BOX, S.PROX_IDENTITY) or \
            (penalty.tag == S.PROX_SHRINK and np.ndim(getattr(penalty, "mu", 0.0)) == 0)
        self.speculate_ok = ((tv_iter or (self.use_sweep and elementwise))
                             and os.environ.get("FASTA_B200_SPECULATE", "1") != "0")
        # TV + FISTA: the whole accelerated trial in one kernel (see _queue_accel)
        self.use_tv_accel = (self.accelerate and isinstance(driver, TVDriver) and driver.fused_step_ok
                             and driver.fista_fused_ok and penalty.tag == S.PROX_TV_BALL
                             and loss.tag in (S.LOSS_LEAST_SQUARES, S.LOSS_LOGISTIC))
        # FISTA trials can be queued ahead of the collect (see _queue_accel and _loop.run)
        self.accel_speculate_ok = ((self.use_tv_accel or self.use_sweep_accel)
                                   and os.environ.get("FASTA_B200_SPECULATE", "1") != "0")

    # -- helpers --------------------------------------------------------------------------------
    def _st(self):
        return _device.stream_ptr()

    def total_launches(self):
        return self.launches + self.drv.launches

    def close(self):
        _device.release_workspace(self.ws)
        self.ws = None

    # -- protocol -------------------------------------------------------------------------------
    def load(self):
        x0d = _device.to_device(self.x0_in, self.X[0].device).reshape(-1)
        self.X[self.ic].copy_(x0d)
        self.ib = self.ic
        if self.accelerate:
            self.XA[self.ac].copy_(x0d)

    # -- prologue, pipelined: the F-2 work (z = A x0, f, gradf1; reference :135-139) does not depend on the two
    # random probes of the Lipschitz estimate (:102-110), so it is queued first and the host draws the probes
    # while the device works; each probe is uploaded from pinned staging and its sweep queued before the next
    # draw.  The order of the draws from numpy's global RNG is the reference's.
    def start_async(self):
        """Queue start(); its sums are snapshotted on the stream, start() later only fetches them."""
        self._queue_start()
        self.ws.scal_saved.copy_(self.ws.scal, non_blocking=True)
        self.ws._saved_ready = False
        self._start_queued = True

    def _probe_outputs(self):
        # X[other] and G[other] are free until the first advance(); Z / R are scratch here
        return self.X[(self.ic + 1) % 4], self.G[(self.gc + 1) % 3]

    def lipschitz_push(self, k, v):
        """Upload probe k (0/1) and queue d_k = A^H gradf(A v_k).

        Least-squares losses have an AFFINE gradient, gradf(z) = z - b, so the difference the estimate needs,
        A^H gradf(A v1) - A^H gradf(A v2) (reference :106-110), is A^H A (v1 - v2): once both probes are on the device one
        contraction pair with a zero right-hand side replaces two (``affine_probe``; FASTA_B200_AFFINE_PROBE=0 evaluates
        the two terms separately as the reference does -- the results agree to rounding)."""
        t = self.t
        dst = self.XH if k == 0 else self.DX
        stage = self.ws.stage(k, self.n)
        stage.copy_(t.from_numpy(np.ascontiguousarray(v.reshape(-1))))
        dst.copy_(stage, non_blocking=True)
        if hasattr(self.drv, "sync_probe"):
            self.drv.sync_probe(dst)
        self._probe_queue(k)

    def _probe_queue(self, k):
        """Probe k is in XH (k = 0) / DX (k = 1): queue its contraction pair (see lipschitz_push)."""
        t = self.t
        dst = self.XH if k == 0 else self.DX
        if self.affine_probe:
            if k == 0:
                return
            w, d = self._probe_outputs()
            _cabi.check(self.lib.fb200_forward_step(self.XH.data_ptr(), self.DX.data_ptr(), 1.0, self.n, w.data_ptr(),
                                                    self._st()), "fb200_forward_step")          # w = v1 - v2
            self.launches += 1
            zero = t.zeros_like(self.loss.b)
            if self.use_sweep or self.use_sweep_accel:
                self.drv.sweep(w, self.loss.tag, zero, self.Z, self.R, d, 1, None, None, None, 0.0, self.ws)
            else:
                self.drv.forward(w, self.loss.tag, zero, self.Z, self.R, self.ws)
                self.drv.adjoint(self.R, d, 1, None, None, None, 0.0, self.ws)
            return
        d = self._probe_outputs()[k]
        if self.use_sweep or self.use_sweep_accel:
            self.drv.sweep(dst, self.loss.tag, self.loss.b, self.Z, self.R, d, 0, None, None, None, 0.0, self.ws)
        else:
            self.drv.forward(dst, self.loss.tag, self.loss.b, self.Z, self.R, self.ws)
            self.drv.adjoint(self.R, d, 0, None, None, None, 0.0, self.ws)

    # -- the same prologue with the probes drawn ON THE DEVICE from numpy's global stream (fasta/_rng.py): no host
    # draw, no upload, and on a row-sharded map one 2.5 KB state broadcast instead of two N-vector broadcasts
    def lipschitz_device_ok(self):
        from . import _rng
        return self.XH.is_cuda and _rng.usable(self.XH.device, self.n)

    def lipschitz_device(self):
        """Returns (|dgrad|, |dpoint|) like lipschitz_finish(), or None when the device draw has to be repeated on the
        host (numpy's generator is then untouched)."""
        from . import _rng
        gen = getattr(self, "_randn", None)
        if gen is None:
            gen = self._randn = _rng.DeviceRandn(self.XH.device)
        gen.begin(sync_fn=getattr(self.drv, "sync_state", None))
        for k, dst in enumerate((self.XH, self.DX)):
            self.launches += gen.draw(dst)
            if k == 1:
                gen.finish_async()               # the end state travels back while the probe sweeps run
            self._probe_queue(k)
        out = self.lipschitz_finish()
        if gen.finish() is None:
            return None
        return out

    def lipschitz_finish(self):
        a, b = self.XH, self.DX
        sc = self.ws.scal
        if self.affine_probe:                       # |A^H A (v1 - v2)|^2 is the probe sweep's S_G1_SQ
            _cabi.check(self.lib.fb200_diff_nrm2sq(a.data_ptr(), b.data_ptr(), self.n, sc[S.S_AUX1:].data_ptr(),
                                                   self.ws.buf.data_ptr(), self._st()), "fb200_diff_nrm2sq")
            self.launches += 1
            s = self.ws.fetch(with_saved=getattr(self, "_start_queued", False))
            return np.sqrt(s[S.S_G1_SQ]), np.sqrt(s[S.S_AUX1])
        d1, d2 = self._probe_outputs()
        _cabi.check(self.lib.fb200_diff_nrm2sq(d1.data_ptr(), d2.data_ptr(), self.n, sc[S.S_AUX0:].data_ptr(),
                                               self.ws.buf.data_ptr(), self._st()), "fb200_diff_nrm2sq")
        _cabi.check(self.lib.fb200_diff_nrm2sq(a.data_ptr(), b.data_ptr(), self.n, sc[S.S_AUX1:].data_ptr(),
                                               self.ws.buf.data_ptr(), self._st()), "fb200_diff_nrm2sq")
        self.launches += 2
        s = self.ws.fetch(with_saved=getattr(self, "_start_queued", False))
        return np.sqrt(s[S.S_AUX0]), np.sqrt(s[S.S_AUX1])

    def lipschitz(self, v1, v2):
        self.lipschitz_push(0, v1)
        self.lipschitz_push(1, v2)
        return self.lipschitz_finish()

    def _penalty_of(self, x):
        """raw penalty reduction of a vector (only the l1 norm needs one)."""
        if self.pen.tag == S.PROX_SHRINK:
            _cabi.check(self.lib.fb200_asum(x.data_ptr(), self.n, self.ws.scal[S.S_PEN:].data_ptr(),
                                            self.ws.buf.data_ptr(), self._st()), "fb200_asum")
            self.launches += 1

    def _queue_start(self):
        z = self.ZA[self.ac] if self.accelerate else self.Z
        if self.use_sweep or self.use_sweep_accel:
            self.drv.sweep(self.X[self.ic], self.loss.tag, self.loss.b, z, self.R, self.G[self.gc], 1, None, None, None,
                           0.0, self.ws)
        else:
            self.drv.forward(self.X[self.ic], self.loss.tag, self.loss.b, z, self.R, self.ws)
            self.drv.adjoint(self.R, self.G[self.gc], 1, None, None, None, 0.0, self.ws)
        self._penalty_of(self.X[self.ic])

    def start(self):
        if getattr(self, "_start_queued", False):
            self._start_queued = False
            s = self.ws.fetch(saved=True)
        else:
            self._queue_start()
            s = self.ws.fetch()
        return Scalars(f=self.loss.finalize(s[S.S_F]), pen=self.pen.value(s[S.S_PEN]), g_sq=s[S.S_G1_SQ])

    def advance(self):
        prev = self.ip
        self.ip = self.ic
        self.ic = next(k for k in (0, 1, 2, 3) if k != self.ip and k != self.ib and k != prev)
        self.gp, self.gc = self.gc, (self.gc + 1) % 3
        if self.accelerate:
            self.ap, self.ac = self.ac, (self.ac + 1) % 3

    # A trial is queued (all kernels asynchronous) and collected (one fetch).  The host loop uses the split to queue
    # the NEXT iteration's trial as soon as the new step size is known and to do its bookkeeping (histories, best
    # iterate, stop rule) while the device already works; trial() is queue + collect.
    def rotation(self):
        return self.ic, self.ip, self.gc, self.gp, self.ac, self.ap, self._ahead

    def restore(self, state):
        self.ic, self.ip, self.gc, self.gp, self.ac, self.ap, self._ahead = state

    def speculate_begin(self, f0, g0_sq, adaptive, backtrack, max_backtracks, window, stop_rule_id, tolerance):
        """From now on every trial is followed by fb200_trial_decide (the loop's decisions repeated on the device: the
        next step size stays there, a rejected trial or a fired stop rule makes speculative successors return at once)
        and an in-stream snapshot of its sums, so trials can be queued with tau=None before the previous one was read."""
        self._spec_mode = True
        self._decide = (self.loss.tag, 1 if adaptive else 0, 1 if backtrack else 0, int(max_backtracks), int(window),
                        int(stop_rule_id), float(tolerance))
        _cabi.check(self.lib.fb200_decide_init(self.ws.scal.data_ptr(), float(f0), float(g0_sq), self._st()),
                    "fb200_decide_init")
        self.launches += 1

    def _queue_trial(self, tau, bt=0, host=(0, -np.inf, 0.0)):
        """Queue one trial (reference :181-188, plus the speculative gradient of the single-pass kernels).  tau=None:
        the kernels read the step size fb200_trial_decide left in scal[S_TAU].  Returns a handle for _collect_trial."""
        x0, g0 = self.X[self.ip], self.G[self.gp]
        x1 = self.XA[self.ac] if self.accelerate else self.X[self.ic]
        z1 = self.ZA[self.ac] if self.accelerate else self.Z
        xa_prev = self.XA[self.ap] if self.accelerate else None
        if tau is None:
            tau = float("nan")
            p0, p1 = (0.0, float(self.pen.mu)) if self.pen.tag == S.PROX_SHRINK else self.pen.params(1.0)
        else:
            p0, p1 = self.pen.params(tau)
        st = self._st()
        kind = None
        fused_decide = None
        slot, host_out = (None, 0)
        if self._spec_mode and self.ws.zero_copy:
            slot, host_out = self.ws.snapshot_begin()     # the deciding kernel writes the sums straight into this pinned slot
        if self.use_tv_fused and self.drv.iter_fused_ok:
            # one kernel per trial; the gradient of an accepted trial is already in G[gc] (speculative, like the sweep)
            self.drv.iterate_fused(x0, g0, tau, self.loss.tag, self.loss.b, x1, self.G[self.gc], self.ws)
            kind = "tv_iter"
        elif self.use_tv_fused:
            self.drv.step_forward(x0, g0, tau, self.loss.tag, self.loss.b, x1, self.R, self.ws)
            kind = "tv_step"
        else:
            if self.pen.tag == S.PROX_L1BALL:
                _cabi.check(self.lib.fb200_forward_step(x0.data_ptr(), g0.data_ptr(), float(tau), self.n,
                                                        self.XH.data_ptr(), st), "fb200_forward_step")
                _cabi.check(self.lib.fb200_l1ball_threshold(self.XH.data_ptr(), self.n, float(self.pen.radius),
                                                            self.ws.scal.data_ptr(), self.ws.buf.data_ptr(), st),
                            "fb200_l1ball_threshold")
                self.launches += 2
            _cabi.check(self.lib.fb200_fbs_step(x0.data_ptr(), g0.data_ptr(), float(tau), self.pen.tag, float(p0),
                                                float(p1), _device.ptr(xa_prev), self.n, self.XH.data_ptr(), x1.data_ptr(),
                                                self.DX.data_ptr(), self.ws.scal.data_ptr(), self.ws.buf.data_ptr(), st),
                        "fb200_fbs_step")
            self.launches += 1
            if self.use_sweep:
                if self._spec_mode and getattr(self.drv, "fuses_decide", False):
                    # row-sharded driver: the exchange kernel behind the sweep also takes the loop's decisions
                    fused_decide = self._decide[:3] + (int(bt),) + self._decide[3:6] + (int(host[0]), self._decide[6],
                                                                                         float(host[1]), float(host[2]))
                    self.drv.sweep(x1, self.loss.tag, self.loss.b, z1, self.R, self.G[self.gc], 2, x0, self.XH, self.DX, tau,
                                   self.ws, decide=fused_decide, host_out=host_out)
                else:
                    self.drv.sweep(x1, self.loss.tag, self.loss.b, z1, self.R, self.G[self.gc], 2, x0, self.XH, self.DX, tau,
                                   self.ws)
                kind = "sweep"
            else:
                self.drv.forward(x1, self.loss.tag, self.loss.b, z1, self.R, self.ws)
                kind = "forward"
        if not self._spec_mode:
            return kind, None
        if fused_decide is not None:
            return kind, (self.ws.snapshot_end(slot) if slot is not None else self.ws.snapshot())
        loss_tag, adaptive, backtrack, max_bt, window, rule, tol = self._decide
        _cabi.check(self.lib.fb200_trial_decide(self.ws.scal.data_ptr(), float(tau), loss_tag, adaptive, backtrack, int(bt),
                                                max_bt, window, rule, tol, int(host[0]), float(host[1]), float(host[2]),
                                                host_out, st), "fb200_trial_decide")
        self.launches += 1
        return kind, (self.ws.snapshot_end(slot) if slot is not None else self.ws.snapshot())

    def _collect_trial(self, handle):
        kind, ticket = handle
        s = self.ws.fetch() if ticket is None else self.ws.collect(ticket)
        if kind in ("tv_iter", "sweep"):
            self._spec = Scalars(dx_dg=s[S.S_DX_DG], dg_sq=s[S.S_DG_SQ], g_sq=s[S.S_G1_SQ])
        tau_used = s[S.S_TAU_USED] if ticket is not None else None
        skipped = ticket is not None and s[S.S_SKIPPED] != 0
        if kind in ("tv_iter", "tv_step"):
            return Scalars(f=self.loss.finalize(s[S.S_F]), dx_g0=s[S.S_DX_G0], dx_sq=s[S.S_DX_SQ],
                           xmxh_sq=s[S.S_XMXH_SQ], pen=self.pen.value(0.0), restart=np.float64(0.0), tau_used=tau_used,
                           skipped=skipped)
        return Scalars(f=self.loss.finalize(s[S.S_F]), dx_g0=s[S.S_DX_G0], dx_sq=s[S.S_DX_SQ],
                       xmxh_sq=s[S.S_XMXH_SQ], pen=self.pen.value(s[S.S_PEN]), restart=s[S.S_RESTART], tau_used=tau_used,
                       skipped=skipped)

    def trial(self, tau, bt=0, host=(0, -np.inf, 0.0)):
        return self._collect_trial(self._queue_trial(tau, bt, host))

    # ---- FISTA trials with the contractions in a single pass (reference :181-188 and :220-249) ------------------
    # The extrapolation weight c = (alpha0 - 1) / alpha1 depends on the restart test of reference :231, a sum of the
    # very forward step being computed, but it has only two possible values: the regular one (alpha0 = the previous
    # alpha1), known before the launch, and 0.  A trial is therefore queued with the weight the caller expects and
    # reports the restart dot; the caller repeats it with c = 0 in the rare iterations that restart.  Queue / collect
    # are split (in-stream snapshot of the sums) so that the host loop can queue the NEXT trial before collecting this one.
    @staticmethod
    def accel_weight(alpha_prev):
        alpha1 = (1 + np.sqrt(1 + 4 * alpha_prev ** 2)) / 2          # reference :238-240 with alpha0 = alpha_prev
        return (alpha_prev - 1) / alpha1

    def _queue_accel(self, tau, c):
        x0, g0 = self.X[self.ip], self.G[self.gp]
        xa1, xa0 = self.XA[self.ac], self.XA[self.ap]
        if self.use_tv_accel:
            # TV: the whole trial in ONE kernel (15U bytes)
            self.drv.fista_fused(x0, g0, tau, c, self.loss.tag, self.loss.b, xa0, self.ZA[self.ap], xa1, self.ZA[self.ac],
                                 self.X[self.ic], self.G[self.gc], self.ws)
            return "tv", self.ws.snapshot(), c
        p0, p1 = self.pen.params(tau)
        st = self._st()
        if self.pen.tag == S.PROX_L1BALL:
            _cabi.check(self.lib.fb200_forward_step(x0.data_ptr(), g0.data_ptr(), float(tau), self.n,
                                                    self.XH.data_ptr(), st), "fb200_forward_step")
            _cabi.check(self.lib.fb200_l1ball_threshold(self.XH.data_ptr(), self.n, float(self.pen.radius),
                                                        self.ws.scal.data_ptr(), self.ws.buf.data_ptr(), st),
                        "fb200_l1ball_threshold")
            self.launches += 2
        _cabi.check(self.lib.fb200_fbs_step(x0.data_ptr(), g0.data_ptr(), float(tau), self.pen.tag, float(p0), float(p1),
                                            xa0.data_ptr(), self.n, self.XH.data_ptr(), xa1.data_ptr(),
                                            self.DX.data_ptr(), self.ws.scal.data_ptr(), self.ws.buf.data_ptr(), st),
                    "fb200_fbs_step")
        # x1 = x_accel1 + c (x_accel1 - x_accel0), |x1 - x1hat|^2, sum |x1|   (m = 0: the z part is the sweep's)
        _cabi.check(self.lib.fb200_accel_step(float(c), xa1.data_ptr(), xa0.data_ptr(), self.XH.data_ptr(), self.n,
                                              self.X[self.ic].data_ptr(), 0, 0, 0, 0, S.LOSS_NONE, self.pen.tag, 0, 0,
                                              self.ws.scal.data_ptr(), self.ws.buf.data_ptr(), st), "fb200_accel_step")
        self.launches += 2
        self.drv.sweep_accel(xa1, self.loss.tag, self.loss.b, self.ZA[self.ap], c, self.ZA[self.ac], self.Z, self.R,
                             self.G[self.gc], 2, x0, self.XH, self.DX, tau, self.ws)
        return "dense", self.ws.snapshot(), c

    def _collect_accel(self, handle):
        kind, ticket, c = handle
        s = self.ws.collect(ticket)
        self._spec = Scalars(dx_dg=s[S.S_DX_DG], dg_sq=s[S.S_DG_SQ], g_sq=s[S.S_G1_SQ])
        pen = self.pen.value(0.0) if kind == "tv" else self.pen.value(s[S.S_PEN])
        extrap = Scalars(f=self.loss.finalize(s[S.S_AUX3]), xmxh_sq=s[S.S_XMXH_SQ], pen=pen)
        return Scalars(f=self.loss.finalize(s[S.S_F]), dx_g0=s[S.S_DX_G0], dx_sq=s[S.S_DX_SQ], xmxh_sq=s[S.S_XMXH_SQ],
                       pen=pen, restart=s[S.S_RESTART], extrap=extrap, c=c)

    def trial_accel(self, tau, alpha_prev, restart):
        """One FISTA trial, collected at once: the regular weight first, repeated with c = 0 if the restart test fires."""
        t = self._collect_accel(self._queue_accel(tau, self.accel_weight(alpha_prev)))
        if restart and t.restart > 1E-30 and t.c != 0.0:
            t = self._collect_accel(self._queue_accel(tau, 0.0))
        return t

    def trial_launch(self, tau):
        """Queue the next iteration's first trial; until trial_finish() the 'current' iterate is X[ip]."""
        self._pending = self._queue_trial(tau)
        self._ahead = True

    def trial_finish(self):
        self._ahead = False
        return self._collect_trial(self._pending)

    def extrapolate(self, c):
        _cabi.check(self.lib.fb200_accel_step(float(c), self.XA[self.ac].data_ptr(), self.XA[self.ap].data_ptr(),
                                              self.XH.data_ptr(), self.n, self.X[self.ic].data_ptr(),
                                              self.ZA[self.ac].data_ptr(), self.ZA[self.ap].data_ptr(),
                                              self.loss.b.data_ptr(), self.m, self.loss.tag, self.pen.tag,
                                              self.Z.data_ptr(), self.R.data_ptr(), self.ws.scal.data_ptr(),
                                              self.ws.buf.data_ptr(), self._st()), "fb200_accel_step")
        self.launches += 1
        if hasattr(self.drv, "reduce_loss"):
            self.drv.reduce_loss(self.ws)
        s = self.ws.fetch()
        return Scalars(f=self.loss.finalize(s[S.S_F]), xmxh_sq=s[S.S_XMXH_SQ], pen=self.pen.value(s[S.S_PEN]))

    def gradient(self, tau, adaptive):
        if self._spec is not None:
            spec, self._spec = self._spec, None      # produced by the accepted trial's single pass: no device work
            return spec
        if self.use_tv_fused:
            self.drv.adjoint_fused(self.R, self.G[self.gc], 2 if adaptive else 1, self.X[self.ip], self.G[self.gp],
                                   self.X[self.ic], tau, self.ws)
            s = self.ws.fetch()
            return Scalars(dx_dg=s[S.S_DX_DG], dg_sq=s[S.S_DG_SQ], g_sq=s[S.S_G1_SQ])
        self.drv.adjoint(self.R, self.G[self.gc], 2 if adaptive else 1, self.X[self.ip], self.XH, self.DX, tau, self.ws)
        s = self.ws.fetch()
        return Scalars(dx_dg=s[S.S_DX_DG], dg_sq=s[S.S_DG_SQ], g_sq=s[S.S_G1_SQ])

    def _current(self):
        # after trial_launch() the buffers are already rotated for the next iteration
        return self.X[self.ip] if self._ahead else self.X[self.ic]

    def keep_best(self):
        self.ib = self.ip if self._ahead else self.ic      # by index: no copy (advance() never hands out X[ib])

    def iterate(self):
        cur = self._current().view(self.shape)
        if not isinstance(self.x0_in, np.ndarray) and _device.is_array(self.x0_in):
            cur = cur.clone()          # the rotating buffer is overwritten by later trials: a hook may keep its argument
        return _device.like_input(cur, self.x0_in)

    def solution(self):
        # X[ib] belongs to this backend alone and the backend ends with the solve: hand the buffer over
        # instead of cloning it (no allocation at the end of a solve)
        return _device.like_input(self.X[self.ib].view(self.shape), self.x0_in)


# =================================================================================================
# generic back-end: user callables on torch CUDA tensors
# =================================================================================================
def _scalar(v):
    if hasattr(v, "item"):
        v = v.item()
    return np.float64(v)


class GenericBackend:
    def __init__(self, A, f, gradf, g, proxg, x0, accelerate, need_objective):
        t = _device.torch()
        self.t = t
        self.lib = _cabi.load()
        self.A, self.f, self.gradf, self.g, self.proxg = A, f, gradf, g, proxg
        self.accelerate = bool(accelerate)
        self.need_objective = bool(need_objective)
        self.x0_in = x0
        self.shape = tuple(x0.shape)
        self.n = int(np.prod(self.shape))
        self.ws = _device.acquire_workspace(1, 1)
        self.launches = 0
        self.x1 = self.x0 = self.g1 = self.g0 = self.z1 = None
        self.xhat = self.dx = self.best = None
        self.xa1 = self.xa0 = self.za1 = self.za0 = None
        self._scratch, self._flip = None, 0

    def total_launches(self):
        return self.launches

    def close(self):
        _device.release_workspace(self.ws)
        self.ws = None

    def _dev(self, a):
        return _device.to_device(a)

    def _flat(self, a):
        return a.contiguous().view(-1)

    def load(self):
        self.x1 = self._dev(self.x0_in).clone()
        self.best = self.x1
        if self.accelerate:
            self.xa1 = self.x1

    def lipschitz(self, v1, v2):
        a, b = self._dev(v1), self._dev(v2)
        d1 = self._dev(self.A.H(self.gradf(self.A(a))))
        d2 = self._dev(self.A.H(self.gradf(self.A(b))))
        sc = self.ws.scal
        st = _device.stream_ptr()
        _cabi.check(self.lib.fb200_diff_nrm2sq(self._flat(d1).data_ptr(), self._flat(d2).data_ptr(), self.n,
                                               sc[S.S_AUX0:].data_ptr(), self.ws.buf.data_ptr(), st), "fb200_diff_nrm2sq")
        _cabi.check(self.lib.fb200_diff_nrm2sq(self._flat(a).data_ptr(), self._flat(b).data_ptr(), self.n,
                                               sc[S.S_AUX1:].data_ptr(), self.ws.buf.data_ptr(), st), "fb200_diff_nrm2sq")
        self.launches += 2
        s = self.ws.fetch()
        return np.sqrt(s[S.S_AUX0]), np.sqrt(s[S.S_AUX1])

    def _grad(self):
        self.g1 = self._dev(self.A.H(self.gradf(self.z1)))

    def start(self):
        self.z1 = self._dev(self.A(self.x1))
        if self.accelerate:
            self.za1 = self.z1
        f1 = _scalar(self.f(self.z1))
        self._grad()
        g = self._flat(self.g1)
        _cabi.check(self.lib.fb200_dot(g.data_ptr(), g.data_ptr(), self.n, self.ws.scal[S.S_G1_SQ:].data_ptr(),
                                       self.ws.buf.data_ptr(), _device.stream_ptr()), "fb200_dot")
        self.launches += 1
        s = self.ws.fetch()
        pen = _scalar(self.g(self.x1)) if self.need_objective else 0
        return Scalars(f=f1, pen=pen, g_sq=s[S.S_G1_SQ])

    def advance(self):
        self.x0, self.g0 = self.x1, self.g1
        if self.accelerate:
            self.xa0, self.za0 = self.xa1, self.za1

    def trial(self, tau):
        t = self.t
        st = _device.stream_ptr()
        x0f, g0f = self._flat(self.x0), self._flat(self.g0)
        # two scratch pairs alternate (the previous trial's xhat / dx may still be referenced by its x1 when the user's prox
        # returns its argument): no allocation per trial
        if self._scratch is None:
            self._scratch = [(t.empty_like(self.x0), t.empty_like(self.x0)) for _ in range(2)]
        self._flip ^= 1
        self.xhat, self.dx = self._scratch[self._flip]
        _cabi.check(self.lib.fb200_forward_step(x0f.data_ptr(), g0f.data_ptr(), float(tau), self.n,
                                                self.xhat.data_ptr(), st), "fb200_forward_step")
        x1 = self._dev(self.proxg(self.xhat, tau))
        if x1.data_ptr() == self.xhat.data_ptr():       # identity prox returns its argument
            x1 = x1.clone()
        xa_prev = self._flat(self.xa0) if self.accelerate else None
        _cabi.check(self.lib.fb200_step_reduce(x0f.data_ptr(), self._flat(x1).data_ptr(), self.xhat.data_ptr(),
                                               g0f.data_ptr(), _device.ptr(xa_prev), self.n, self.dx.data_ptr(),
                                               self.ws.scal.data_ptr(), self.ws.buf.data_ptr(), st), "fb200_step_reduce")
        self.launches += 2
        self.x1 = x1
        self.z1 = self._dev(self.A(x1))
        if self.accelerate:
            self.xa1, self.za1 = self.x1, self.z1
        f1 = _scalar(self.f(self.z1))
        s = self.ws.fetch()
        pen = _scalar(self.g(self.x1)) if (self.need_objective and not self.accelerate) else 0
        return Scalars(f=f1, dx_g0=s[S.S_DX_G0], dx_sq=s[S.S_DX_SQ], xmxh_sq=s[S.S_XMXH_SQ], pen=pen,
                       restart=s[S.S_RESTART])

    def extrapolate(self, c):
        t = self.t
        x1 = t.empty_like(self.xa1)
        z1 = t.empty_like(self.za1)
        m = z1.numel()
        _cabi.check(self.lib.fb200_accel_step(float(c), self._flat(self.xa1).data_ptr(), self._flat(self.xa0).data_ptr(),
                                              self.xhat.data_ptr(), self.n, x1.data_ptr(),
                                              self._flat(self.za1).data_ptr(), self._flat(self.za0).data_ptr(), 0, m,
                                              S.LOSS_NONE, S.PROX_IDENTITY, z1.data_ptr(), 0, self.ws.scal.data_ptr(),
                                              self.ws.buf.data_ptr(), _device.stream_ptr()), "fb200_accel_step")
        self.launches += 1
        self.x1, self.z1 = x1, z1
        f1 = _scalar(self.f(self.z1))
        s = self.ws.fetch()
        pen = _scalar(self.g(self.x1)) if self.need_objective else 0
        return Scalars(f=f1, xmxh_sq=s[S.S_XMXH_SQ], pen=pen)

    def gradient(self, tau, adaptive):
        self._grad()
        _cabi.check(self.lib.fb200_bb_reduce(self._flat(self.g1).data_ptr(), self._flat(self.x0).data_ptr(),
                                             self.xhat.data_ptr(), self.dx.data_ptr(), float(tau), self.n,
                                             1 if adaptive else 0, self.ws.scal.data_ptr(), self.ws.buf.data_ptr(),
                                             _device.stream_ptr()), "fb200_bb_reduce")
        self.launches += 1
        s = self.ws.fetch()
        return Scalars(dx_dg=s[S.S_DX_DG], dg_sq=s[S.S_DG_SQ], g_sq=s[S.S_G1_SQ])

    def keep_best(self):
        self.best = self.x1

    def iterate(self):
        return _device.like_input(self.x1, self.x0_in)

    def solution(self):
        return _device.like_input(self.best.clone(), self.x0_in)

"""Batched FASTA: B independent forward-backward-splitting solves that share the operator A
(a regularisation path over B penalty weights, or B right-hand sides) run in lock-step, so that
the two contractions per iteration become fp64 GEMMs (BASELINE config 5).

    results = fasta.batched.lasso_path(A, b, mus, tolerance=1e-5)           # list of Convergence, one per mu
    results = fasta.batched.fasta_batched(A, loss, penalty, X0, **options)  # general form

Every column is exactly one reference run (reference ``fasta/__init__.py:38-320``): it has its own
step size, its own non-monotone line search (columns that fail the test are re-tried alone, the
others keep their accepted point), its own Barzilai-Borwein update, its own best iterate and its own
stopping time; the scalar algebra is the vectorised ``np.float64`` transcription of ``_loop.run``.
A column's trajectory therefore matches ``fasta.fasta`` on that column (same counts; values to
reduction-order rounding).  The randomised Lipschitz estimate (reference ``:100-113``) is drawn once:
all columns share A, and for the least-squares / logistic losses the estimate does not depend on the
column (the data term cancels in ``gradf1 - gradf2`` up to rounding), so each column sees what a single
run seeded identically would see.

Device work per iteration: one batched step/prox kernel, the forward GEMM, one loss kernel, the
adjoint GEMM, one BB kernel (csrc/batched_vector.cu), plus masked column copies.  The GEMMs run on the
tcgen05 tensor cores as int8 digit-plane products with int32 TMEM accumulators recombined in fp64
(csrc/ozaki_gemm.cu; A is split into digit planes once per batch, the iterates once per product);
small problems, and ``FASTA_B200_GEMM=dmma``, use the fp64 DMMA kernel (csrc/batched_gemm.cu) instead.
Accelerated (FISTA) mode is not available in the batched loop yet.
"""

import os
from time import time

import numpy as np

from . import _cabi, _device, linalg, losses, proximal, stopping
from ._loop import Convergence, EPSILON

__all__ = ["fasta_batched", "lasso_path", "column_shard", "lasso_path_sharded"]

# measurement hook (tools/bench_batched.py): when set to a list, every GEMM appends
# (start_event, end_event, adjoint, active_columns) recorded on the launching stream
GEMM_EVENTS = None


# reference stopping.py:15,27,39,51 on whole columns of np.float64 (nan comparisons are False, as for scalars)
_VECTOR_RULES = {
    stopping.residual: lambda r, nr, mr, tol: r < tol,
    stopping.norm_residual: lambda r, nr, mr, tol: nr < tol,
    stopping.ratio_residual: lambda r, nr, mr, tol: r / mr < tol,
    stopping.hybrid_residual: lambda r, nr, mr, tol: (r / mr < tol) | (nr < tol),
}


class _Batch:
    """Device state of a batch (all matrices row-major, batch index fastest)."""

    def __init__(self, A, loss, pen, X0, Bu, spare=0):
        t = _device.torch()
        self.t, self.lib = t, _cabi.load()
        self.A = A.matrix
        self.M, self.N, self.lda = A.M, A.N, A.lda
        # columns 0..Bu-1 are the caller's problems; columns Bu..B-1 are spare slots in which the line search
        # evaluates further shrunken step sizes of a column in the same pass (candidate fan-out)
        self.Bu, self.spare = Bu, int(spare)
        B = self.B = Bu + self.spare
        dev = self.A.device
        self.loss, self.pen = loss, pen
        b = loss.b
        assert b.shape[0] == self.M and (b.ndim == 1 or b.shape[1] == Bu)
        if b.ndim == 1:
            self.b, self.b_ld = b.contiguous(), 0
        else:
            self.b, self.b_ld = t.zeros((self.M, B), dtype=t.float64, device=dev), B
            self.b[:, :Bu].copy_(b)
        new = lambda r: t.empty((r, B), dtype=t.float64, device=dev)
        self.X0, self.X1, self.XH, self.DX, self.G0, self.G1, self.BEST = (new(self.N) for _ in range(7))
        self.Z, self.R = new(self.M), new(self.M)
        self.ws = t.empty(int(self.lib.fb200_batched_workspace_bytes(self.M, self.N, B)), dtype=t.uint8, device=dev)
        self.out = t.zeros(5 * B, dtype=t.float64, device=dev)
        self.launches = 0
        self._stage_w = max(B, 64)
        self._stage = t.empty((64, 8 * self._stage_w), dtype=t.uint8).pin_memory()
        self._stage_next = 0
        mode = os.environ.get("FASTA_B200_GEMM", "auto")
        if mode not in ("auto", "dmma", "ozaki"):
            raise ValueError("FASTA_B200_GEMM must be auto, dmma or ozaki")
        self.ozaki = mode == "ozaki" or (mode == "auto" and self.M * self.N >= (1 << 24) and B >= 16)
        if self.ozaki:
            self._ozaki_setup()
        else:
            self.sf = int(self.lib.fb200_gemm_splits(self.M, Bu, self.N))      # forward: (M x B) over K = N
            self.sa = int(self.lib.fb200_gemm_splits(self.N, Bu, self.M))      # adjoint: (N x B) over K = M
            self.ZP = t.empty((self.sf, self.M, B), dtype=t.float64, device=dev)
            self.GP = t.empty((self.sa, self.N, B), dtype=t.float64, device=dev)
        x0 = _device.to_device(X0, dev)
        if x0.ndim == 1:
            x0 = x0[:, None].expand(self.N, Bu)
        assert tuple(x0.shape) == (self.N, Bu)
        for buf in (self.X0, self.X1, self.XH, self.DX, self.G0, self.G1, self.BEST, self.Z, self.R):
            buf[:, Bu:].zero_()
        self.X1[:, :Bu].copy_(x0)
        self.BEST[:, :Bu].copy_(x0)

    def _ozaki_setup(self):
        """Digit planes of A for both products (once per batch) and the buffers of the per-product planes."""
        t, lib, dev = self.t, self.lib, self.A.device
        pad = lambda n, tile: int(lib.fb200_ozaki_pad(n, tile))
        M, N, B = self.M, self.N, self.B
        mp, np_, bp = pad(M, 128), pad(N, 128), pad(B, 64)
        i8 = lambda *shape: t.empty(shape, dtype=t.int8, device=dev)
        f8 = lambda n: t.empty(n, dtype=t.float64, device=dev)
        self.oz_scratch = t.empty(max(B, N), dtype=t.int64, device=dev)
        self.AF, self.af = i8(8, mp, np_), f8(mp)            # row-scaled planes of A     (forward: contraction over N)
        self.AT, self.at = i8(8, np_, mp), f8(np_)           # column-scaled planes of A^T (adjoint: contraction over M)
        _cabi.check(lib.fb200_ozaki_slice_rows(self.A.data_ptr(), self.lda, M, N, self.AF.data_ptr(), self.af.data_ptr(), self._st()),
                    "fb200_ozaki_slice_rows")
        _cabi.check(lib.fb200_ozaki_slice_cols(self.A.data_ptr(), self.lda, M, 0, N, 128, self.AT.data_ptr(), self.at.data_ptr(),
                                               self.oz_scratch.data_ptr(), self._st()), "fb200_ozaki_slice_cols")
        self.XS, self.xs = i8(8, bp, np_), f8(bp)            # planes of the iterate X (N x B), transposed
        self.RS, self.rs = i8(8, bp, mp), f8(bp)             # planes of the residual R (M x B), transposed
        # ONE K-split count per product, whatever the number of active columns: a column's result must not depend
        # on which other columns happen to be multiplied with it (the partials are added in split order)
        widths = range(64, pad(self.Bu, 64) + 1, 64)
        self.sf = max(int(lib.fb200_ozaki_splits(M, w, N)) for w in widths)
        self.sa = max(int(lib.fb200_ozaki_splits(N, w, M)) for w in widths)
        self.ZP = t.empty((self.sf, M, B), dtype=t.float64, device=dev)
        self.GP = t.empty((self.sa, N, B), dtype=t.float64, device=dev)
        self.launches += 3

    def _gemm_ozaki(self, adjoint, src, part, rows, K, act):
        lib, t = self.lib, self.t
        cols = np.nonzero(act)[0]
        n_act = len(cols)
        cm = None if n_act == self.B else self._vec(cols, np.int32)
        cmp_ = 0 if cm is None else cm.data_ptr()
        LS, ls, RS, rs = (self.AT, self.at, self.RS, self.rs) if adjoint else (self.AF, self.af, self.XS, self.xs)
        _cabi.check(lib.fb200_ozaki_slice_cols(src.data_ptr(), self.B, K, cmp_, n_act, 64, RS.data_ptr(), rs.data_ptr(),
                                               self.oz_scratch.data_ptr(), self._st()), "fb200_ozaki_slice_cols")
        s = self.sa if adjoint else self.sf
        _cabi.check(lib.fb200_ozaki_gemm(LS.data_ptr(), ls.data_ptr(), rows, RS.data_ptr(), rs.data_ptr(), n_act, K, part.data_ptr(),
                                         self.B, cmp_, s, rows * self.B, self._st()), "fb200_ozaki_gemm")
        self.launches += 4
        return s

    def _st(self):
        return _device.stream_ptr()

    def _vec(self, a, dtype):
        """Upload a short per-column host vector without blocking the host: pinned staging ring + async copy (a
        pageable cudaMemcpy would wait for everything already queued on the stream).  Every round of the loop ends
        in a device->host fetch, so far fewer than the ring's 64 uploads are ever in flight."""
        t = self.t
        a = np.asarray(a, dtype=dtype)
        if a.size > self._stage_w:
            return t.as_tensor(np.array(a, copy=True), device=self.A.device)
        slot = self._stage_next
        self._stage_next = (slot + 1) % 64
        view = self._stage[slot, :a.size * a.itemsize].numpy().view(dtype)
        view[:] = a.ravel()
        out = t.empty(a.size, dtype=t.float64 if dtype == np.float64 else t.int32, device=self.A.device)
        out.view(t.uint8).copy_(self._stage[slot, :a.size * a.itemsize], non_blocking=True)
        return out

    def _gemm(self, adjoint, src, part, rows, K, splits, act):
        """Columns `act` of (A or A^T) . src into `part`; returns the number of split partials the
        epilogue has to add (1 when the product was computed on a compacted column subset)."""
        lib, t = self.lib, self.t
        n_act = int(np.count_nonzero(act))
        if n_act == 0:
            return 1
        if GEMM_EVENTS is not None:
            e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
            e0.record()
            ns = self._gemm_inner(adjoint, src, part, rows, K, splits, act, n_act)
            e1.record()
            GEMM_EVENTS.append((e0, e1, adjoint, n_act))
            return ns
        return self._gemm_inner(adjoint, src, part, rows, K, splits, act, n_act)

    def _gemm_inner(self, adjoint, src, part, rows, K, splits, act, n_act):
        lib, t = self.lib, self.t
        if self.ozaki:
            return self._gemm_ozaki(adjoint, src, part, rows, K, act)
        if n_act > 0.75 * self.B or self.B <= 8:
            _cabi.check(lib.fb200_gemm_f64(adjoint, self.A.data_ptr(), self.lda, src.data_ptr(), self.B, part.data_ptr(),
                                           self.B, rows, self.B, K, splits, rows * self.B, self._st()), "fb200_gemm_f64")
            self.launches += 1
            return splits
        # few active columns (late iterations, or a handful of columns re-trying a shorter step):
        # gather them, multiply the narrow matrix, scatter the result columns back
        cols = np.nonzero(act)[0]
        if len(cols) % 2:
            cols = np.append(cols, cols[-1])
        idx = t.as_tensor(cols, device=src.device)
        nb = len(cols)
        sub = src.index_select(1, idx).contiguous()
        s = splits                              # the same K split as the full-width product: results independent of the active set
        out = t.empty((s, rows, nb), dtype=t.float64, device=src.device)
        _cabi.check(lib.fb200_gemm_f64(adjoint, self.A.data_ptr(), self.lda, sub.data_ptr(), nb, out.data_ptr(), nb, rows, nb, K,
                                       s, rows * nb, self._st()), "fb200_gemm_f64")
        self.launches += 1
        res = out[0]
        for k in range(1, s):
            res = res + out[k]                 # fixed order
        part[0].index_copy_(1, idx, res)
        return 1

    def forward(self, act):
        """Z = A X1 on the active columns, then z, r = gradf(z), f for them -> f (B,)"""
        lib = self.lib
        ns = self._gemm(0, self.X1, self.ZP, self.M, self.N, self.sf, act)
        a = self._vec(act, np.int32)
        _cabi.check(lib.fb200_batched_loss(self.loss.tag, self.ZP.data_ptr(), ns, self.M * self.B, self.b.data_ptr(), self.b_ld,
                                           a.data_ptr(), self.M, self.B, self.Z.data_ptr(), self.R.data_ptr(), self.out.data_ptr(),
                                           self.ws.data_ptr(), self._st()), "fb200_batched_loss")
        self.launches += 2
        return self.out[:self.B].cpu().numpy()

    def adjoint(self, act, tau, bb):
        """G1 = A^T R on the active columns, per-column BB sums -> (dx_dg, dg_sq, g_sq) each (B,)"""
        lib = self.lib
        ns = self._gemm(1, self.R, self.GP, self.N, self.M, self.sa, act)
        a = self._vec(act, np.int32)
        tv = self._vec(tau, np.float64)
        _cabi.check(lib.fb200_batched_bb(self.GP.data_ptr(), ns, self.N * self.B, self.X0.data_ptr(), self.XH.data_ptr(),
                                         self.DX.data_ptr(), tv.data_ptr(), a.data_ptr(), bb, self.N, self.B, self.G1.data_ptr(),
                                         self.out.data_ptr(), self.ws.data_ptr(), self._st()), "fb200_batched_bb")
        self.launches += 2
        o = self.out[:3 * self.B].cpu().numpy().reshape(3, self.B)
        return o[0], o[1], o[2]

    def step(self, act, tau, owner=None):
        """xhat, x1 = prox, dx for the active columns -> (dx_g0, dx_sq, xmxh_sq, pen_raw) each (B,);
        owner[j] = the caller's column whose penalty a (spare) column j uses"""
        tau = np.asarray(tau, dtype=np.float64)
        if isinstance(self.pen, proximal.L1Norm):
            mu = np.asarray(self.pen.mu, dtype=np.float64)
            mu = np.concatenate([mu, np.zeros(self.B - len(mu))]) if mu.ndim else np.full(self.B, mu)
            if owner is not None:
                mu = mu[owner]
            p0, p1 = tau * mu, 0.0                            # as L1Norm.params: shrink threshold t * mu
        else:
            p0, p1 = self.pen.params(tau)
        p0 = np.broadcast_to(np.asarray(p0, dtype=np.float64), (self.B,))
        p1 = np.broadcast_to(np.asarray(p1, dtype=np.float64), (self.B,))
        a, tv = self._vec(act, np.int32), self._vec(tau, np.float64)
        p0d, p1d = self._vec(p0, np.float64), self._vec(p1, np.float64)
        _cabi.check(self.lib.fb200_batched_fbs_step(self.X0.data_ptr(), self.G0.data_ptr(), tv.data_ptr(), self.pen.tag,
                                                    p0d.data_ptr(), p1d.data_ptr(), a.data_ptr(), self.N, self.B,
                                                    self.XH.data_ptr(), self.X1.data_ptr(), self.DX.data_ptr(),
                                                    self.out.data_ptr(), self.ws.data_ptr(), self._st()), "fb200_batched_fbs_step")
        self.launches += 2
        o = self.out[:4 * self.B].cpu().numpy().reshape(4, self.B)
        return o[0], o[1], o[2], o[3]

    def copy_cols(self, arrays, scol, dcol):
        """arr[:, dcol[k]] = arr[:, scol[k]] for every array in `arrays`."""
        if len(scol) == 0:
            return
        sc, dc = self._vec(scol, np.int32), self._vec(dcol, np.int32)
        for a in arrays:
            _cabi.check(self.lib.fb200_batched_copy_cols(a.data_ptr(), a.data_ptr(), a.shape[0], a.shape[1], a.shape[1],
                                                         sc.data_ptr(), dc.data_ptr(), len(scol), self._st()),
                        "fb200_batched_copy_cols")
            self.launches += 1

    def select(self, dst, src, mask):
        if not np.any(mask):
            return
        m = self._vec(mask, np.int32)
        _cabi.check(self.lib.fb200_batched_select(dst.data_ptr(), src.data_ptr(), m.data_ptr(), dst.shape[0], self.B, self._st()),
                    "fb200_batched_select")
        self.launches += 1


def _lipschitz(A, loss, x_shape_n):
    """The randomised estimate L = |A^T gradf(A v1) - A^T gradf(A v2)| / |v1 - v2| (reference :100-113), two probes drawn
    once for the whole batch.  For least squares the data term cancels in the difference (gradf is affine), so one
    estimate serves every column exactly as a single run seeded identically would see it; for other losses with one
    right-hand side per column (logistic: gradf is not affine in b) the estimate is formed per column."""
    v1 = np.random.randn(x_shape_n)
    v2 = np.random.randn(x_shape_n)
    t = _device.torch()
    dv = np.float64(np.linalg.norm(v1 - v2))
    d1v, d2v = A(_device.to_device(v1)), A(_device.to_device(v2))
    cols = [None] if (loss.b.ndim == 1 or isinstance(loss, losses.LeastSquares)) else range(loss.b.shape[1])
    Ls = []
    for j in cols:
        bj = loss.b if loss.b.ndim == 1 else loss.b[:, 0 if j is None else j].contiguous()
        single = type(loss)(bj)
        num = float(t.linalg.norm(A.H(single.gradf(d1v)) - A.H(single.gradf(d2v))))
        Ls.append(np.float64(num) / dv)
    L = Ls[0] if len(Ls) == 1 else np.array(Ls)
    return L, (2 / L) / 10


def fasta_batched(A, loss, penalty, X0, *, adaptive=True, accelerate=False, verbose=False, max_iters=1000,
                  tolerance=1e-5, stop_rule=stopping.hybrid_residual, L=None, tau0=None, backtrack=True,
                  stepsize_shrink=None, window=10, max_backtracks=20, evaluate_objective=False):
    """Solve B problems min_x f_j(A x) + g_j(x) in lock-step; returns a list of B ``Convergence``."""
    _cabi.require_cuda()
    if accelerate:
        raise NotImplementedError("fasta_batched: accelerated mode is not available in the batched loop")
    if not isinstance(A, linalg.DenseMap):
        A = linalg.LinearMap.from_matrix(A)
    if penalty is None:
        penalty = proximal._Penalty()
    if isinstance(penalty, proximal.L1Norm):
        mus = np.atleast_1d(np.asarray(penalty.mu, dtype=np.float64))
    else:
        mus = np.zeros(1)
    x0a = X0 if _device.is_array(X0) else np.asarray(X0)
    B = x0a.shape[1] if x0a.ndim == 2 else (loss.b.shape[1] if loss.b.ndim == 2 else len(mus))
    mus = np.broadcast_to(mus, (B,)).copy()
    if isinstance(penalty, proximal.L1Norm):
        penalty = proximal.L1Norm(mus)
    if B % 2:
        raise ValueError("fasta_batched: the batch width must be even (pad with a duplicate column)")
    # candidate fan-out of the line search: up to `fan` consecutive shrunken step sizes of a column are evaluated
    # in ONE pass (the extra candidates live in `spare` scratch columns), because a retry round costs a whole
    # stream over A however few columns take part in it
    fan = max(1, int(os.environ.get("FASTA_B200_FANOUT", "3"))) if backtrack else 1
    spare = int(os.environ.get("FASTA_B200_FANOUT_SLOTS", "0" if fan == 1 else ("64" if B >= 32 else "8")))
    spare += (B + spare) % 2
    st = _Batch(A, loss, penalty, x0a, B, spare)
    BT = st.B

    if stepsize_shrink is None and backtrack:
        stepsize_shrink = 0.2 if adaptive else 0.5
    if L is None or tau0 is None or not np.all(L) or not np.all(tau0):
        L, tau0 = _lipschitz(A, loss, st.N)
    if verbose:
        print(f"Initializing batched FASTA: {B} columns\n")

    def ext(v, fill=0):
        """host vector over the caller's columns -> over all device columns (spare slots filled)"""
        out = np.full(BT, fill, dtype=np.asarray(v).dtype)
        out[:B] = v
        return out

    ones = np.ones(B, dtype=bool)
    resid_h = np.zeros((B, max_iters))
    nresid_h = np.zeros((B, max_iters))
    tau_h = np.zeros((B, max_iters))
    f_h = np.zeros((B, max_iters + 1))
    obj_h = np.zeros((B, max_iters + 1)) if evaluate_objective else None
    times = np.zeros(max_iters + 1)

    tau1 = np.full(B, tau0, dtype=np.float64)
    f1 = loss.finalize(st.forward(ext(ones, False)))[:B]
    _, _, g1_sq = st.adjoint(ext(ones, False), ext(tau1, 1.0), 1)
    g1_sq = g1_sq[:B].copy()
    f_h[:, 0] = f1
    pen_raw0 = np.zeros(B)
    if evaluate_objective:
        pen_raw0 = st.X1[:, :B].abs().sum(dim=0).cpu().numpy() if isinstance(penalty, proximal.L1Norm) else np.zeros(B)
        obj_h[:, 0] = f1 + penalty.value(pen_raw0)

    done = np.zeros(B, dtype=bool)
    iters = np.zeros(B, dtype=np.int64)
    total_bt = np.zeros(B, dtype=np.int64)
    max_resid = np.full(B, -np.inf)
    best_q = np.full(B, np.inf)
    dx_g0, dx_sq, xmxh_sq, pen_raw = (np.zeros(B) for _ in range(4))
    f1 = f1.copy()
    bt_prev = np.zeros(B, dtype=np.int64)
    fan_rounds = fan_hits = 0
    state_arrays = (st.XH, st.X1, st.DX, st.Z, st.R)

    i = 0
    while i < max_iters and not done.all():
        times[i] = time()
        act = ~done
        st.select(st.X0, st.X1, ext(act, False))          # x0 <- x1, gradf0 <- gradf1 (reference :176-177)
        st.select(st.G0, st.G1, ext(act, False))
        g0_sq = g1_sq.copy()
        tau_cur = tau1.copy()
        lo = max(i - window + 1, 0)
        f_max = np.max(f_h[:, lo:i + 1], axis=1)

        def fails(f, dxg, dxsq, tau, fm):
            with np.errstate(all="ignore"):
                return f - (fm + dxg + np.sqrt(dxsq) ** 2 / (2 * tau)) > EPSILON      # reference :200

        def run_round(cols, first_shrunk, depth):
            """For every column j in `cols`, evaluate depth[j] consecutive candidates tau, tau*s, tau*s*s, ...
            (starting at tau*s when first_shrunk) in one pass: candidate 0 in the column itself, the others in
            spare slots; adopt the first one that passes the line-search test (else the last).  Returns the
            number of shrinks applied per column of `cols`."""
            nonlocal fan_rounds, fan_hits
            mask = np.zeros(BT, dtype=bool)
            tau_e = np.ones(BT)
            owner = np.arange(BT)
            owner[B:] = 0
            slots = {}
            nxt = B
            for j in cols:
                tj = tau_cur[j] * stepsize_shrink if first_shrunk else tau_cur[j]
                mask[j], tau_e[j] = True, tj
                slots[j] = [(j, tj)]
                for _ in range(1, int(depth[j])):
                    tj = tj * stepsize_shrink                 # the same sequence of products as repeated tau0 *= shrink
                    mask[nxt], tau_e[nxt], owner[nxt] = True, tj, j
                    slots[j].append((nxt, tj))
                    nxt += 1
            mirror = [(j, q) for j in cols for q, _ in slots[j][1:]]
            if mirror:
                src, dst = [m[0] for m in mirror], [m[1] for m in mirror]
                st.copy_cols((st.X0, st.G0) + ((st.b,) if st.b_ld else ()), src, dst)
                fan_rounds += 1
            a, b_, c, d = st.step(mask, tau_e, owner)
            fm = loss.finalize(st.forward(mask))
            shrinks = {}
            win_src, win_dst = [], []
            for j in cols:
                pick = len(slots[j]) - 1
                for k, (q, tq) in enumerate(slots[j]):
                    if not backtrack or not fails(fm[q], a[q], b_[q], tq, f_max[j]):
                        pick = k
                        break
                q, tq = slots[j][pick]
                dx_g0[j], dx_sq[j], xmxh_sq[j], pen_raw[j], f1[j] = a[q], b_[q], c[q], d[q], fm[q]
                tau_cur[j] = tq
                shrinks[j] = pick + (1 if first_shrunk else 0)
                if q != j:
                    win_src.append(q)
                    win_dst.append(j)
                    fan_hits += 1
            st.copy_cols(state_arrays, win_src, win_dst)
            return shrinks

        cols = np.nonzero(act)[0]
        depth = np.ones(B, dtype=np.int64)
        if fan > 1 and spare:
            # predictive fan-out: a column that had to backtrack in the previous iteration brings its next
            # candidates along in the first trial, as long as spare slots last
            budget = spare
            for j in cols[np.argsort(-bt_prev[cols], kind="stable")]:
                if bt_prev[j] == 0 or budget <= 0:
                    break
                depth[j] = 1 + min(fan - 1, budget, max_backtracks)
                budget -= depth[j] - 1
        sh = run_round(cols, False, depth)
        bt = np.zeros(B, dtype=np.int64)
        for j, k in sh.items():
            bt[j] = k
        if backtrack:
            while True:
                need = act & fails(f1, dx_g0, dx_sq, tau_cur, f_max) & (bt < max_backtracks)
                if not need.any():
                    break
                cols = np.nonzero(need)[0]
                depth = np.ones(B, dtype=np.int64)
                if fan > 1 and spare:
                    per = 1 + min(fan - 1, spare // len(cols))
                    depth[cols] = np.minimum(per, max_backtracks - bt[cols])
                sh = run_round(cols, True, depth)
                for j, k in sh.items():
                    bt[j] += k
            total_bt += bt
        bt_prev = bt

        act_e, tau_e = ext(act, False), ext(tau_cur, 1.0)
        dx_dg, dg_sq, gsq = (v[:B] for v in st.adjoint(act_e, tau_e, 2 if adaptive else 1))
        g1_sq[act] = gsq[act]
        tau1[act] = tau_cur[act]
        dx_norm = np.sqrt(dx_sq)
        if adaptive:
            with np.errstate(all="ignore"):
                tau_s = dx_norm ** 2 / dx_dg
                q = dx_dg / np.sqrt(dg_sq) ** 2
                tau_m = np.where(np.isnan(q), q, np.maximum(q, 0))      # python max(q, 0) keeps a nan first argument
                new = np.where(2 * tau_m > tau_s, tau_m, tau_s - .5 * tau_m)
                bad = (new <= 0) | np.isinf(new) | np.isnan(new)
                new = np.where(bad, tau_cur * 1.5, new)
            tau1[act] = new[act]

        with np.errstate(all="ignore"):
            resid = dx_norm / tau_cur
            na, nb = np.sqrt(g0_sq), np.sqrt(xmxh_sq) / tau_cur
            normalizer = np.where(nb > na, nb, na) + EPSILON               # python max(a, b): a unless b > a (keeps a nan a)
            nresid = resid / normalizer
        resid_h[act, i] = resid[act]
        tau_h[act, i] = tau_cur[act]
        nresid_h[act, i] = nresid[act]
        f_h[act, i + 1] = f1[act]
        max_resid[act] = np.where(resid[act] > max_resid[act], resid[act], max_resid[act])     # python max(max_residual, resid)
        if evaluate_objective:
            obj = f1 + penalty.value(pen_raw)
            obj_h[act, i + 1] = obj[act]
            quality = obj
        else:
            quality = resid
        better = act & (quality < best_q)
        st.select(st.BEST, st.X1, ext(better, False))
        best_q[better] = quality[better]

        iters[act] = i + 1
        with np.errstate(all="ignore"):
            if stop_rule in _VECTOR_RULES:       # the built-in rules, evaluated for all columns at once (same truth tables)
                stop = _VECTOR_RULES[stop_rule](resid_h[:, i], nresid_h[:, i], max_resid, tolerance)
                done |= act & stop
            else:
                for j in np.nonzero(act)[0]:
                    if stop_rule(i, resid_h[j, i], nresid_h[j, i], max_resid[j], tolerance):
                        done[j] = True
        if verbose:
            print(f"[{i:<6}]\tactive {int(act.sum()):4d}\tmax residual {np.max(resid[act]):e}")
        i += 1
    times[i] = time()

    best = st.BEST[:, :B].t().contiguous()   # (B, N)
    results = []
    for j in range(B):
        n = int(iters[j])
        tj = np.zeros(max_iters + 1)
        tj[:n] = times[:n]
        tj[n] = times[min(n, i)] if n < i else times[i]
        sol = _device.like_input(best[j].clone(), x0a)
        c = Convergence(resid_h[j], nresid_h[j], tau_h[j], int(total_bt[j]), tj, n, sol,
                        obj_h[j] if evaluate_objective else None, None, None)
        results.append(c)
    results_meta = dict(kernel_launches=st.launches, gemm_splits=(st.sf, st.sa), iterations_lockstep=i, loop_s=float(times[i] - times[0]),
                        fanout=dict(depth=fan, spare_columns=spare, rounds_with_fanout=fan_rounds, candidates_adopted_from_slots=fan_hits),
                        gemm="tcgen05-i8-digit-planes" if st.ozaki else "dmma-f64")
    for c in results:
        c.batch = results_meta
    return results


def lasso_path(A, b, mus, x0=None, **options):
    """The regularisation path min mu_j |x|_1 + .5 |A x - b|^2 for every mu_j in ``mus`` (config 5)."""
    mus = np.asarray(mus, dtype=np.float64)
    if not isinstance(A, linalg.DenseMap):
        A = linalg.LinearMap.from_matrix(A)
    if x0 is None:
        x0 = np.zeros((A.N, len(mus)))
    return fasta_batched(A, losses.LeastSquares(b), proximal.L1Norm(mus), x0, **options)


def column_shard(n_columns, rank, world):
    """Columns of a batch that rank ``rank`` of ``world`` solves: dealt round-robin, because neighbouring penalty
    weights need about the same number of iterations (columns are independent units, no exchange)."""
    return np.arange(int(rank), int(n_columns), int(world))


def lasso_path_sharded(A, b, mus, group=None, gather=False, **options):
    """The regularisation path split by columns over the ranks of ``group`` (one process per GPU, A replicated on every
    GPU): each rank solves ``column_shard(len(mus), rank, world)`` with ``lasso_path``; nothing is exchanged during the
    solve.  Returns ``(columns, results)`` of this rank, or with ``gather=True`` the full list of ``Convergence`` in
    column order on every rank (solutions travel as host arrays)."""
    import torch.distributed as dist
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if dist.is_initialized() else (0, 1)
    mus = np.asarray(mus, dtype=np.float64)
    cols = column_shard(len(mus), rank, world)
    pad = len(cols) % 2 == 1                          # the batched kernels need an even width: duplicate the last column
    local = mus[np.append(cols, cols[-1])] if pad else mus[cols]
    res = lasso_path(A, b, local, **options)[:len(cols)]
    if not gather or world == 1:
        return cols, res
    for r in res:
        if hasattr(r.solution, "detach"):
            r.solution = r.solution.detach().cpu().numpy()
    box = [None] * world
    dist.all_gather_object(box, (cols.tolist(), res), group=group)
    full = [None] * len(mus)
    for cs, rs in box:
        for c, r in zip(cs, rs):
            full[c] = r
    return np.arange(len(mus)), full

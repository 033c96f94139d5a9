"""Proximal operators: drop-in for the reference ``fasta/proximal.py`` plus tagged penalties.

Functions (same names, arguments and results as the reference, computed on the GPU):
  shrink(x, t)               soft threshold, any shape                      proximal.py:58-67
  project_L1_ball(x, t)      Euclidean projection onto {|x|_1 <= t}, 1-D    proximal.py:34-41
  project_Linf_ball(x, t)    (misnamed in the reference) prox of t*|.|_inf  proximal.py:12-31
  project_Lnuc_ball(X, t)    (misnamed) singular-value soft threshold       proximal.py:44-55

Penalty objects pair ``g`` with ``proxg`` so that ``fasta()`` can fuse the backward step into
the forward-step kernel (K1-K3): pass ``pen.g`` and ``pen.prox``.
  L1Norm(mu)        g=mu*|x|_1,  prox=shrink(x, t*mu)          sparse_least_squares.py:43-44
  L1Ball(radius)    g=0,         prox=project_L1_ball(x, r)    lasso.py:44-45
  NonNegative()     g=0,         prox=max(x, 0)                nn_least_squares.py:41-42
  Box(lo, hi)       g=0,         prox=clip(x, lo, hi)          svm.py:71
  TVBall()          g=0,         prox=Y/max(|Y|_2, 1) on the last axis of size 2   tv_denoising.py:87-96
  RowGroupL2(mu)    g=mu*sum_i|X_i|_2, prox=shrink_rows(X, t*mu)                   mmv.py:51-63
  RowNormBall(mu)   g=0,         prox=project_rows_L2_ball(X, mu)                  max_norm.py:51-59
  LinfNorm(mu)      g=mu*|x|_inf, prox=project_Linf_ball(x, t*mu)                   democratic_representation.py:43-45

Row-wise operators on 2-D iterates (the prox bodies the matrix examples define inline):
  shrink_rows(X, t)            X_i * shrink(|X_i|_2, t) / (|X_i|_2 + (|X_i|_2 == 0))     mmv.py:53-61
  project_rows_L2_ball(X, mu)  mu * X_i / (max(|X_i|_2, mu) + (|X_i|_2 == 0))           max_norm.py:53-59
"""

import numpy as np

from . import _cabi, _device

__all__ = ["project_Linf_ball", "project_L1_ball", "project_Lnuc_ball", "shrink", "shrink_rows",
           "project_rows_L2_ball", "row_norms", "L1Norm", "L1Ball", "NonNegative", "Box", "TVBall", "RowGroupL2",
           "RowNormBall", "LinfNorm"]


def _apply(x, tag, p0=0.0, p1=0.0, radius=None):
    t = _device.torch()
    lib = _cabi.load()
    xd = _device.to_device(x)
    out = t.empty_like(xd)
    ws = _device.shared_workspace(xd.numel(), 1)
    st = _device.stream_ptr()
    if tag == _cabi.PROX_L1BALL:
        _cabi.check(lib.fb200_l1ball_threshold(xd.data_ptr(), xd.numel(), float(radius), ws.scal.data_ptr(),
                                               ws.buf.data_ptr(), st), "fb200_l1ball_threshold")
    _cabi.check(lib.fb200_prox_apply(xd.data_ptr(), tag, float(p0), float(p1), xd.numel(), out.data_ptr(),
                                     ws.scal.data_ptr(), st), "fb200_prox_apply")
    return _device.like_input(out, x)


def shrink(x, t):
    """sign(x) * max(|x| - t, 0): the prox of t*|.|_1 (reference proximal.py:58-67)."""
    return _apply(x, _cabi.PROX_SHRINK, t)


def project_L1_ball(x, t):
    """Euclidean projection of a vector onto the l1 ball of radius t (reference proximal.py:34-41)."""
    assert x.ndim == 1
    return _apply(x, _cabi.PROX_L1BALL, radius=t)


def project_Linf_ball(x, t):
    """The reference's function of this name: prox of t*|.|_inf, by Moreau's identity
    x - project_L1_ball(x, t) (reference proximal.py:12-31 clips |x| at the same threshold)."""
    assert x.ndim == 1
    return x - project_L1_ball(x, t)


def _nuclear(X, t, want_out):
    tt = _device.torch()
    lib = _cabi.load()
    assert X.ndim == 2
    Xd = _device.to_device(X).contiguous()
    M, N = Xd.shape
    out = tt.empty_like(Xd) if want_out else None
    sv = tt.empty(min(M, N), dtype=tt.float64, device=Xd.device)
    need = int(lib.fb200_prox_nuclear_scratch_doubles(M, N))
    scratch = tt.empty(need, dtype=tt.float64, device=Xd.device) if need else None
    _cabi.check(lib.fb200_prox_nuclear(Xd.data_ptr(), M, N, N, float(t), _device.ptr(out), N, sv.data_ptr(),
                                       _device.ptr(scratch), None, _device.stream_ptr()), "fb200_prox_nuclear")
    return out, sv


def project_Lnuc_ball(X, t):
    """The reference's function of this name: singular-value soft threshold U diag(shrink(s,t)) V
    (reference proximal.py:44-55), by the one-sided Jacobi kernel of csrc/jacobi_svd.cu (no library SVD)."""
    out, _ = _nuclear(X, t, True)
    return _device.like_input(out, X)


def singular_values(X):
    """Singular values of a matrix iterate, descending (what `la.svd(X)[1]` gives the nuclear-norm penalty of
    reference logistic_matrix_completion.py:43), from the same Jacobi kernel."""
    _, sv = _nuclear(X, 0.0, False)
    return _device.like_input(_device.torch().sort(sv, descending=True).values, X)


def _rows(X, mode, p, want_out=True, want_norms=False):
    t = _device.torch()
    lib = _cabi.load()
    assert X.ndim == 2
    Xd = _device.to_device(X).contiguous()
    out = t.empty_like(Xd) if want_out else None
    norms = t.empty(Xd.shape[0], dtype=t.float64, device=Xd.device) if want_norms else None
    _cabi.check(lib.fb200_prox_rows(Xd.data_ptr(), Xd.shape[0], Xd.shape[1], mode, float(p), _device.ptr(out),
                                    _device.ptr(norms), _device.stream_ptr()), "fb200_prox_rows")
    return out, norms


def shrink_rows(X, t):
    """Row-group soft threshold, the prox of t*sum_i |X_i|_2 (reference mmv.py:53-61)."""
    out, _ = _rows(X, 0, t)
    return _device.like_input(out, X)


def project_rows_L2_ball(X, mu):
    """Every row scaled onto the l2 ball of radius mu (reference max_norm.py:53-59)."""
    out, _ = _rows(X, 1, mu)
    return _device.like_input(out, X)


def row_norms(X):
    """|X_i|_2 per row (la.norm(X, axis=1) of the reference's matrix examples)."""
    _, norms = _rows(X, 0, 0.0, want_out=False, want_norms=True)
    return _device.like_input(norms, X)


class _Penalty:
    tag = _cabi.PROX_IDENTITY

    def params(self, t):
        """(p0, p1) for the device prox at step size t, computed in np.float64 like the reference."""
        return 0.0, 0.0

    def value(self, raw):
        """g(x1) from the raw device reduction (sum |x1|)."""
        return 0

    def g(self, x):
        return 0

    def prox(self, x, t):
        p0, p1 = self.params(t)
        return _apply(x, self.tag, p0, p1)


class L1Norm(_Penalty):
    """g(x) = mu*|x|_1 (reference sparse_least_squares.py:43-44, sparse_logistic.py:49-50)."""
    tag = _cabi.PROX_SHRINK

    def __init__(self, mu):
        self.mu = mu

    def params(self, t):
        return t * self.mu, 0.0

    def value(self, raw):
        return self.mu * raw

    def g(self, x):
        lib = _cabi.load()
        xd = _device.to_device(x)
        ws = _device.shared_workspace(xd.numel(), 1)
        _cabi.check(lib.fb200_asum(xd.data_ptr(), xd.numel(), ws.scal[_cabi.S_AUX0:].data_ptr(), ws.buf.data_ptr(),
                                   _device.stream_ptr()), "fb200_asum")
        return self.mu * ws.fetch()[_cabi.S_AUX0]


class L1Ball(_Penalty):
    """Indicator of {|x|_1 <= radius}; g evaluates to 0 like the reference (lasso.py:44-45)."""
    tag = _cabi.PROX_L1BALL

    def __init__(self, radius):
        self.radius = radius

    def prox(self, x, t):
        return project_L1_ball(x, self.radius)


class NonNegative(_Penalty):
    """Indicator of the non-negative orthant (reference nn_least_squares.py:41-42)."""
    tag = _cabi.PROX_NONNEG


class Box(_Penalty):
    """Indicator of the box [lo, hi] (reference svm.py:71)."""
    tag = _cabi.PROX_BOX

    def __init__(self, lo, hi):
        self.lo, self.hi = lo, hi

    def params(self, t):
        return self.lo, self.hi


class TVBall(_Penalty):
    """Indicator of {|Y_ij|_2 <= 1} on a trailing axis of size 2 (reference tv_denoising.py:87-96)."""
    tag = _cabi.PROX_TV_BALL

    def prox(self, x, t):
        if x.shape[-1] == 2:
            return _apply(x, self.tag)
        tt = _device.torch()                     # other ranks: Y / max(|Y|_2 along the last axis, 1), tv_denoising.py:89-96
        lib = _cabi.load()
        xd = _device.to_device(x).contiguous()
        out = tt.empty_like(xd)
        k = int(xd.shape[-1])
        _cabi.check(lib.fb200_tv_ball_nd(xd.data_ptr(), xd.numel() // k, k, out.data_ptr(), _device.stream_ptr()), "fb200_tv_ball_nd")
        return _device.like_input(out, x)


class RowGroupL2(_Penalty):
    """g(X) = mu * sum_i |X_i|_2 on a 2-D iterate (reference mmv.py:51,63); generic back-end only."""
    tag = None

    def __init__(self, mu):
        self.mu = mu

    def g(self, X):
        _, n = _rows(X, 0, 0.0, want_out=False, want_norms=True)        # device row norms (non-negative: asum is their sum)
        lib = _cabi.load()
        ws = _device.shared_workspace(n.numel(), 1)
        _cabi.check(lib.fb200_asum(n.data_ptr(), n.numel(), ws.scal[_cabi.S_AUX0:].data_ptr(), ws.buf.data_ptr(),
                                   _device.stream_ptr()), "fb200_asum")
        return self.mu * ws.fetch()[_cabi.S_AUX0]

    def prox(self, X, t):
        return shrink_rows(X, t * self.mu)


class RowNormBall(_Penalty):
    """Indicator of {|X_i|_2 <= mu for every row} (reference max_norm.py:51-59); generic back-end only."""
    tag = None

    def __init__(self, mu):
        self.mu = mu

    def prox(self, X, t):
        return project_rows_L2_ball(X, self.mu)


class LinfNorm(_Penalty):
    """g(x) = mu * |x|_inf (reference democratic_representation.py:43-45); generic back-end only."""
    tag = None

    def __init__(self, mu):
        self.mu = mu

    def g(self, x):
        lib = _cabi.load()
        xd = _device.to_device(x).contiguous()
        ws = _device.shared_workspace(1, 1)
        _cabi.check(lib.fb200_amax(xd.data_ptr(), xd.numel(), ws.scal[_cabi.S_AUX0:].data_ptr(), _device.stream_ptr()), "fb200_amax")
        return self.mu * ws.fetch()[_cabi.S_AUX0]

    def prox(self, x, t):
        return project_Linf_ball(x, t * self.mu)

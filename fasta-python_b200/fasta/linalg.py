"""Linear-map objects for fasta(): drop-in for the reference ``fasta/linalg.py``.

``LinearMap`` keeps the reference's surface (constructor, ``from_matrix``, ``identity``,
``__call__``, ``.H``, ``@ * - + **``, ``is_operator``, ``eigs``; reference linalg.py:13-160) and its
bare-``assert`` shape checks.  What differs is where the arithmetic runs: ``from_matrix`` returns a
``DenseMap`` whose two contractions are the hand-written sm_100a streaming kernels behind
``fb200_gemv_loss`` / ``fb200_gemvT_bb`` (csrc/dense_stream.cu), and which the solver recognises
so that it can fuse the loss / step / prox epilogues around them.

Array types: maps accept numpy arrays or torch tensors and answer in the same type; all compute
is on the current CUDA device (numpy inputs are uploaded, results downloaded).
"""

from functools import reduce
from operator import mul
from typing import Callable, Tuple

import numpy as np

from . import _cabi, _device

Matrix = np.ndarray
Vector = np.ndarray

__all__ = ["LinearMap", "LinearOperator", "DenseMap", "Matrix", "Vector"]


class LinearMap:
    """A linear map V -> W on n-dimensional arrays with an adjoint (reference linalg.py:13-35)."""

    def __init__(self, map_func: Callable, adj_func: Callable, Vshape: Tuple[int, ...], Wshape: Tuple[int, ...] = None):
        # 3-argument legacy form LinearOperator(map, adj, shape) (democratic_representation.py:80-82)
        if Wshape is None:
            Wshape = Vshape
        self.map_func = map_func
        self.adj_func = adj_func
        self.Vshape = tuple(Vshape)
        self.Wshape = tuple(Wshape)

    # -- constructors ---------------------------------------------------------------------------
    @staticmethod
    def from_matrix(A) -> "DenseMap":
        """Linear map of a 2-D array, x -> A @ x with adjoint y -> A.T @ y (reference linalg.py:37-41)."""
        assert A.ndim == 2
        return DenseMap(A)

    @staticmethod
    def identity(shape: Tuple[int, ...]) -> "LinearMap":
        """The identity on arrays of ``shape``; returns its argument object itself (reference linalg.py:43-50)."""
        return LinearMap(lambda x: x, lambda x: x, shape, shape)

    # -- application ----------------------------------------------------------------------------
    def __call__(self, v):
        assert tuple(v.shape) == self.Vshape
        w = self.map_func(v)
        assert tuple(w.shape) == self.Wshape
        return w

    @property
    def H(self) -> "LinearMap":
        """The adjoint map; a fresh object per access (reference linalg.py:63-69)."""
        return LinearMap(self.adj_func, self.map_func, self.Wshape, self.Vshape)

    # -- algebra (reference linalg.py:71-139, including its labelling of compositions) ------------
    def __matmul__(self, B: "LinearMap") -> "LinearMap":
        assert isinstance(B, LinearMap) and self.Wshape == B.Vshape
        return LinearMap(lambda x: self(B(x)), lambda x: B.H(self.H(x)), self.Vshape, B.Wshape)

    def __rmul__(self, k) -> "LinearMap":
        assert np.isscalar(k)
        return LinearMap(lambda x: k * self(x), lambda x: k * self.H(x), self.Vshape, self.Wshape)

    def __mul__(self, k) -> "LinearMap":
        return k * self

    def __neg__(self) -> "LinearMap":
        return (-1) * self

    def __add__(self, B: "LinearMap") -> "LinearMap":
        assert isinstance(B, LinearMap) and self.Vshape == B.Vshape and self.Wshape == B.Wshape
        return LinearMap(lambda x: self(x) + B(x), lambda x: self.H(x) + B.H(x), self.Vshape, self.Wshape)

    def __sub__(self, B: "LinearMap") -> "LinearMap":
        return self + (-B)

    @property
    def is_operator(self) -> bool:
        return self.Vshape == self.Wshape

    def __pow__(self, n: int, modulo=None) -> "LinearMap":
        assert self.is_operator
        out = LinearMap.identity(self.Vshape)
        for _ in range(n):
            out @= self
        return out

    # -- scipy bridge (reference linalg.py:141-160); off the hot path -----------------------------
    @property
    def _scipy(self):
        from scipy.sparse import linalg as sla
        m = reduce(mul, self.Vshape, 1)
        n = reduce(mul, self.Wshape, 1)
        return sla.LinearOperator((m, n), matvec=lambda x: np.ravel(self(x)), rmatvec=lambda x: np.ravel(self.H(x)))

    def eigs(self, k: int = 1):
        assert self.is_operator
        from scipy.sparse import linalg as sla
        values, vectors = sla.eigs(self._scipy, k)
        return values, np.reshape(vectors, (k,) + self.Wshape)


class DenseMap(LinearMap):
    """x -> A @ x for a dense row-major fp64 matrix resident in HBM (what ``from_matrix`` returns).

    The matrix is uploaded once (numpy input) or borrowed (CUDA tensor with unit column stride);
    ``.H`` shares the same buffer -- like the reference's ``A.T`` view, both contractions stream
    the one row-major copy.  ``_fb200_dense`` is the tag the solver looks for.
    """

    def __init__(self, A, _transposed=False, _dev=None, columns=None):
        if _dev is None:
            t = _device.torch()
            if isinstance(A, t.Tensor) and A.is_cuda and A.dtype == t.float64 and A.stride(1) == 1 and A.stride(0) >= A.shape[1]:
                _dev = A                       # borrow: may be a row-slice view of a larger matrix
            else:
                _dev = _device.to_device(A)
        self.matrix = _dev
        self.transposed = bool(_transposed)
        M, N = self.matrix.shape
        self.M, self.N, self.lda = int(M), int(N), int(self.matrix.stride(0))
        # columns = L: the map acts on matrix iterates, X (N x L) -> A @ X (M x L), as the reference's
        # `A @ x` does when the examples pass 2-D unknowns (mmv.py:65); the contractions are then GEMMs
        self.cols = None if columns is None else int(columns)
        tail = () if self.cols is None else (self.cols,)
        V, W = ((N,) + tail, (M,) + tail) if not self.transposed else ((M,) + tail, (N,) + tail)
        self._padded = None
        super().__init__(self._apply, self._apply_adjoint, V, W)

    _fb200_dense = True

    @property
    def H(self) -> "DenseMap":
        return DenseMap(None, _transposed=not self.transposed, _dev=self.matrix, columns=self.cols)

    def with_columns(self, L) -> "DenseMap":
        """The same matrix acting on N x L matrix iterates."""
        return DenseMap(None, _transposed=self.transposed, _dev=self.matrix, columns=L)

    @property
    def uses_tma(self) -> bool:
        lib = _cabi.load()
        return bool(lib.fb200_dense_uses_tma(self.matrix.data_ptr(), self.lda, self.M, self.N))

    # raw device contractions (used by the solver back-ends too)
    def gemv_into(self, x_dev, z_dev, ws):
        lib = _cabi.load()
        _cabi.check(lib.fb200_gemv_loss(self.matrix.data_ptr(), self.lda, self.M, self.N, x_dev.data_ptr(),
                                        _cabi.LOSS_NONE, 0, z_dev.data_ptr(), 0, ws.scal.data_ptr(),
                                        ws.buf.data_ptr(), ws.nbytes, _device.stream_ptr()), "fb200_gemv_loss")

    def gemvT_into(self, r_dev, g_dev, ws):
        lib = _cabi.load()
        _cabi.check(lib.fb200_gemvT_bb(self.matrix.data_ptr(), self.lda, self.M, self.N, r_dev.data_ptr(),
                                       g_dev.data_ptr(), 0, 0, 0, 0, 0.0, ws.scal.data_ptr(), ws.buf.data_ptr(),
                                       ws.nbytes, _device.stream_ptr()), "fb200_gemvT_bb")

    def gram_norm(self, iters: int = 500, tol: float = 1e-12, check_every: int = 8):
        """Largest eigenvalue of A^T A, i.e. |A|_2^2 -- the Lipschitz constant of the least-squares gradient that the
        reference only ESTIMATES from two random probes (``fasta/__init__.py:100-113``) -- by power iteration on the
        device (exact constant, no dependence on the global RNG).  An iteration is ONE pass over A:
        the single-pass sweep with a zero right-hand side gives ``g = A^T (A x)`` and ``|A x|^2`` (the Rayleigh
        quotient of the normalised x) together; matrices the sweep cannot take use the two streaming contractions.
        The start vector is fixed, the reductions are fixed-order: the result is reproducible bit for bit.

            L = A.gram_norm();  fasta.fasta(A, f, gradf, g, proxg, x0, L=L, tau0=(2 / L) / 10)   # reference :113
        """
        t = _device.torch()
        lib = _cabi.load()
        S = _cabi
        dev = self.matrix.device
        M, N, A, lda = self.M, self.N, self.matrix, self.lda
        ws = _device.shared_workspace(M, N)
        new = lambda k: t.empty(k, dtype=t.float64, device=dev)
        x, g, z, r = new(N), new(N), new(M), new(M)
        zero_n, zero_m = t.zeros(N, dtype=t.float64, device=dev), t.zeros(M, dtype=t.float64, device=dev)
        k = t.arange(N, dtype=t.float64, device=dev)
        g.copy_(t.cos(0.7 * k + 0.3) + 1.5)                   # fixed start: positive mean, no symmetry
        st = _device.stream_ptr
        sweep = bool(lib.fb200_sweep_supported(A.data_ptr(), lda, M, N))
        lam_prev, lam = 0.0, 0.0
        _cabi.check(lib.fb200_dot(g.data_ptr(), g.data_ptr(), N, ws.scal[S.S_AUX0:].data_ptr(), ws.buf.data_ptr(), st()), "fb200_dot")
        g_sq = float(ws.fetch()[S.S_AUX0])
        for it in range(1, iters + 1):
            # x = g / |g|  (xhat = x0 - tau * g0 with x0 = 0, tau = -1 / |g|)
            _cabi.check(lib.fb200_forward_step(zero_n.data_ptr(), g.data_ptr(), -1.0 / np.sqrt(g_sq), N, x.data_ptr(), st()),
                        "fb200_forward_step")
            if sweep:
                _cabi.check(lib.fb200_dense_sweep(A.data_ptr(), lda, M, N, x.data_ptr(), S.LOSS_LEAST_SQUARES,
                                                  zero_m.data_ptr(), z.data_ptr(), r.data_ptr(), g.data_ptr(), 1, 0, 0, 0,
                                                  0.0, ws.scal.data_ptr(), ws.buf.data_ptr(), ws.nbytes, st()),
                            "fb200_dense_sweep")
            else:
                _cabi.check(lib.fb200_gemv_loss(A.data_ptr(), lda, M, N, x.data_ptr(), S.LOSS_LEAST_SQUARES,
                                                zero_m.data_ptr(), z.data_ptr(), r.data_ptr(), ws.scal.data_ptr(),
                                                ws.buf.data_ptr(), ws.nbytes, st()), "fb200_gemv_loss")
                _cabi.check(lib.fb200_gemvT_bb(A.data_ptr(), lda, M, N, r.data_ptr(), g.data_ptr(), 1, 0, 0, 0, 0.0,
                                               ws.scal.data_ptr(), ws.buf.data_ptr(), ws.nbytes, st()), "fb200_gemvT_bb")
            s = ws.fetch()                                    # |A x|^2 = x^T A^T A x (Rayleigh quotient), |A^T A x|^2
            lam, g_sq = float(s[S.S_F]), float(s[S.S_G1_SQ])
            if g_sq == 0.0:
                return 0.0
            if it % check_every == 0:
                if abs(lam - lam_prev) <= tol * lam:
                    break
                lam_prev = lam
        return lam

    def _gemm(self, v, transposed):
        """A @ V or A.T @ V for a matrix V with the fp64 DMMA kernel (csrc/batched_gemm.cu); shapes the kernel
        cannot take directly (odd N, L or leading dimension) go through zero-padded copies."""
        t = _device.torch()
        lib = _cabi.load()
        dev = self.matrix.device
        V = _device.to_device(v, dev)
        L = self.cols
        Np, Lp = self.N + (self.N & 1), L + (L & 1)
        A, lda = self.matrix, self.lda
        if Np != self.N or (lda & 1) or (A.data_ptr() & 15):
            if self._padded is None:
                self._padded = t.zeros((self.M, Np), dtype=t.float64, device=dev)
                self._padded[:, :self.N].copy_(A)
            A, lda = self._padded, Np
        rows_in = self.M if transposed else Np
        if Lp != L or (not transposed and Np != self.N) or not V.is_contiguous() or (V.data_ptr() & 15):
            Vp = t.zeros((rows_in, Lp), dtype=t.float64, device=dev)
            Vp[:V.shape[0], :L].copy_(V)
        else:
            Vp = V
        rows_out = Np if transposed else self.M
        out = t.empty((rows_out, Lp), dtype=t.float64, device=dev)
        K = self.M if transposed else Np
        _cabi.check(lib.fb200_gemm_f64(1 if transposed else 0, A.data_ptr(), lda, Vp.data_ptr(), Lp, out.data_ptr(), Lp,
                                       rows_out, Lp, K, 1, rows_out * Lp, _device.stream_ptr()), "fb200_gemm_f64")
        res = out[:(self.N if transposed else self.M), :L]
        return _device.like_input(res if res.is_contiguous() else res.contiguous(), v)

    def _run(self, v, transposed):
        if self.cols is not None:
            return self._gemm(v, transposed)
        t = _device.torch()
        ws = _device.shared_workspace(self.M, self.N)
        x = _device.to_device(v, self.matrix.device).reshape(-1)
        out = t.empty(self.N if transposed else self.M, dtype=t.float64, device=self.matrix.device)
        if transposed:
            self.gemvT_into(x, out, ws)
        else:
            self.gemv_into(x, out, ws)
        return _device.like_input(out, v)

    def _apply(self, v):
        return self._run(v, self.transposed)

    def _apply_adjoint(self, v):
        return self._run(v, not self.transposed)


# legacy name used by every reference example (``from fasta.linalg import LinearOperator``)
LinearOperator = LinearMap

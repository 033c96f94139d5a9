"""Matrix-free total-variation operators on the GPU (K11 in SURVEY.md).

``grad`` / ``div`` restate the reference's periodic finite-difference pair
(reference examples/tv_denoising.py:26-40 and :43-63): 2-D images run the fused stencil kernels
(csrc/tv_stencil.cu), every other rank (the reference's functions are rank-generic) the N-d
kernels ``fb200_tv_grad_nd`` / ``fb200_tv_div_nd``.  ``divergence_map(shape)`` wraps them as the
``LinearMap`` the TV-denoising dual problem uses (A = div, A^H = grad; tv_denoising.py:99); for
2-D images it is tagged so that fasta() fuses the loss and Barzilai-Borwein epilogues into the
stencils, for other ranks the generic back-end drives the N-d kernels.
"""

import ctypes

from . import _cabi, _device
from .linalg import LinearMap

__all__ = ["grad", "div", "divergence_map", "TVDivergenceMap"]


def grad(X):
    """(n0, n1) -> (n0, n1, 2): G[...,d] = roll(X, +1, axis=d) - X (reference tv_denoising.py:26-40)."""
    t = _device.torch()
    lib = _cabi.load()
    Xd = _device.to_device(X)
    if Xd.ndim != 2:
        shape = (ctypes.c_int64 * Xd.ndim)(*Xd.shape)
        out = t.empty(tuple(Xd.shape) + (Xd.ndim,), dtype=t.float64, device=Xd.device)
        _cabi.check(lib.fb200_tv_grad_nd(Xd.data_ptr(), shape, Xd.ndim, out.data_ptr(), _device.stream_ptr()), "fb200_tv_grad_nd")
        return _device.like_input(out, X)
    n0, n1 = Xd.shape
    out = t.empty((n0, n1, 2), dtype=t.float64, device=Xd.device)
    ws = _device.shared_workspace(1, 1)
    _cabi.check(lib.fb200_tv_grad_bb(Xd.data_ptr(), n0, n1, out.data_ptr(), 0, 0, 0, 0, 0.0, ws.scal.data_ptr(),
                                     ws.buf.data_ptr(), _device.stream_ptr()), "fb200_tv_grad_bb")
    return _device.like_input(out, X)


def div(Y):
    """(n0, n1, 2) -> (n0, n1): sum_d roll(Y[...,d], -1, axis=d) - Y[...,d] (reference tv_denoising.py:43-63)."""
    t = _device.torch()
    lib = _cabi.load()
    Yd = _device.to_device(Y)
    assert Yd.shape[-1] == Yd.ndim - 1
    if Yd.ndim != 3:
        rank = Yd.ndim - 1
        shape = (ctypes.c_int64 * rank)(*Yd.shape[:-1])
        out = t.empty(tuple(Yd.shape[:-1]), dtype=t.float64, device=Yd.device)
        _cabi.check(lib.fb200_tv_div_nd(Yd.data_ptr(), shape, rank, out.data_ptr(), _device.stream_ptr()), "fb200_tv_div_nd")
        return _device.like_input(out, Y)
    n0, n1, _ = Yd.shape
    out = t.empty((n0, n1), dtype=t.float64, device=Yd.device)
    ws = _device.shared_workspace(1, 1)
    _cabi.check(lib.fb200_tv_div_loss(Yd.data_ptr(), n0, n1, _cabi.LOSS_NONE, 0, out.data_ptr(), 0,
                                      ws.scal.data_ptr(), ws.buf.data_ptr(), _device.stream_ptr()),
                "fb200_tv_div_loss")
    return _device.like_input(out, Y)


class TVDivergenceMap(LinearMap):
    """A = div : (n0, n1, 2) -> (n0, n1) with adjoint grad; recognised (fused) by fasta()."""
    _fb200_tv = True

    def __init__(self, image_shape):
        n0, n1 = image_shape
        self.n0, self.n1 = int(n0), int(n1)
        super().__init__(div, grad, (self.n0, self.n1, 2), (self.n0, self.n1))


def divergence_map(image_shape) -> LinearMap:
    """A = div with adjoint grad on arrays of shape `image_shape` (any rank); 2-D images get the fused, tagged map."""
    image_shape = tuple(int(n) for n in image_shape)
    if len(image_shape) == 2:
        return TVDivergenceMap(image_shape)
    return LinearMap(div, grad, image_shape + (len(image_shape),), image_shape)

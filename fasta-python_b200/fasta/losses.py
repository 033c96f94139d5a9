"""Tagged smooth losses f(z): objects whose ``.f`` / ``.gradf`` bound methods fasta() recognises
and fuses into the epilogue of the ``A x`` kernel (K5 in SURVEY.md).  They are also ordinary
callables on numpy arrays / torch tensors, computed on the GPU by ``fb200_loss_eval``.

The formulas are those of the reference's example problems:
  LeastSquares  f=.5*norm(z-b)**2, gradf=z-b                   sparse_least_squares.py:41-42
  Logistic      f=sum(log(1+exp(z)) - (b==1)*z), gradf=-b/(1+exp(b*z))   sparse_logistic.py:47-48
"""

import numpy as np

from . import _cabi, _device

__all__ = ["LeastSquares", "Logistic"]


class _Loss:
    tag = _cabi.LOSS_NONE

    def __init__(self, b):
        self._b_in = b
        self.b = _device.to_device(b)

    # host-side finish of the raw device reduction (np.float64 in, np.float64 out)
    def finalize(self, raw):
        return raw

    def _eval(self, z):
        t = _device.torch()
        lib = _cabi.load()
        zd = _device.to_device(z, self.b.device)
        assert zd.shape == self.b.shape
        r = t.empty_like(zd)
        ws = _device.shared_workspace(zd.numel(), 1)
        _cabi.check(lib.fb200_loss_eval(self.tag, zd.data_ptr(), self.b.data_ptr(), zd.numel(), r.data_ptr(),
                                        ws.scal.data_ptr(), ws.buf.data_ptr(), _device.stream_ptr()),
                    "fb200_loss_eval")
        return r, ws

    def f(self, z):
        _, ws = self._eval(z)
        return self.finalize(ws.fetch()[_cabi.S_F])

    def gradf(self, z):
        r, _ = self._eval(z)
        return _device.like_input(r, z)


class LeastSquares(_Loss):
    """f(z) = .5*|z - b|^2 (reference sparse_least_squares.py:41-42, lasso.py:42-43, tv_denoising.py:85-86)."""
    tag = _cabi.LOSS_LEAST_SQUARES

    def finalize(self, raw):
        # the reference squares a norm: .5 * la.norm(z - b)**2 = .5 * sqrt(sum r^2)**2
        return .5 * np.sqrt(raw) ** 2


class Logistic(_Loss):
    """f(z) = sum log(1+e^z) - (b==1) z, labels b in {-1,+1} (reference sparse_logistic.py:47-48)."""
    tag = _cabi.LOSS_LOGISTIC

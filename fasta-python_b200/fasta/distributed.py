"""Row-sharded dense maps over the GPUs of one box.

One process per GPU (``torch.distributed``, NCCL over NVLink 5 / NVSwitch).  Rank p holds rows
``[lo_p, hi_p)`` of A and the matching slice of the observation vector; ``A x`` is local, the
loss partial sums and the ``A^H r`` partial vectors are combined with all-reduce.  ``x``, the
gradient, the step size and every history are replicated, so each rank returns the same
``Convergence``.

    rows = fasta.distributed.row_slice(M, rank, world)
    A    = fasta.distributed.RowShardedMatrix(A_full[rows])        # or a matrix built per rank
    loss = fasta.losses.LeastSquares(b_full[rows])
    res  = fasta.fasta(A, loss.f, loss.gradf, pen.g, pen.prox, x0)
"""

from . import _device
from .linalg import LinearMap

__all__ = ["row_slice", "RowShardedMatrix"]


def row_slice(M, rank, world):
    """Contiguous, balanced row range of rank ``rank`` out of ``world``."""
    lo = (M * rank) // world
    hi = (M * (rank + 1)) // world
    return slice(lo, hi)


class RowShardedMatrix(LinearMap):
    """The local row block of a row-partitioned matrix; the global map is the stack over ranks."""

    def __init__(self, local_rows, group=None):
        t = _device.torch()
        if isinstance(local_rows, t.Tensor) and local_rows.is_cuda and local_rows.dtype == t.float64 \
                and local_rows.stride(1) == 1:
            self.matrix = local_rows
        else:
            self.matrix = _device.to_device(local_rows)
        self.group = group
        m, n = self.matrix.shape
        super().__init__(self._apply_local, self._adjoint_sum, (int(n),), (int(m),))

    def _fb200_driver(self):
        from ._backends import DenseDriver, ShardedDriver
        return ShardedDriver(DenseDriver(self.matrix), self.group)

    def _apply_local(self, x):
        from .linalg import DenseMap
        return DenseMap(None, _dev=self.matrix)(x)

    def _adjoint_sum(self, r):
        import torch.distributed as dist
        from .linalg import DenseMap
        g = _device.to_device(DenseMap(None, _dev=self.matrix).H(r))
        dist.all_reduce(g, group=self.group)
        return _device.like_input(g, r)

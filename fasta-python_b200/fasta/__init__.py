"""fasta -- B200-native drop-in for phasepack/fasta-python's forward-backward-splitting solver.

    min_x  f(A x) + g(x)          f smooth & convex, g convex with a cheap proximal operator

Entry point (same name, options and ``Convergence`` result as reference ``fasta/__init__.py:38-53``):

    fasta(A, f, gradf, g, proxg, x0, **options)            # A: fasta.linalg.LinearMap
    fasta(A, At, f, gradf, g, proxg, x0, **options)        # legacy 7-argument form: A/At arrays,
                                                           # callables, or None (identity)

All arithmetic runs on the current CUDA device as hand-written sm_100a kernels reached through
the C ABI in ``include/fasta_b200.h``; the ``while`` loop and its scalar algebra stay in Python
(``_loop.py``).  When the operator, the loss and the penalty are *tagged* objects --

    A    = fasta.linalg.LinearMap.from_matrix(M)      or  fasta.tv.divergence_map((n, n))
    loss = fasta.losses.LeastSquares(b)               or  fasta.losses.Logistic(b)
    pen  = fasta.proximal.L1Norm(mu) | L1Ball(r) | NonNegative() | Box(lo, hi) | TVBall()
    res  = fasta.fasta(A, loss.f, loss.gradf, pen.g, pen.prox, x0)

-- every iteration is one pass over A (z = A x, the loss and g = A^T gradf(z) in a single sweep, the
step / prox / reductions in two small kernels around it).  Any other callables run through the
generic path and receive torch CUDA tensors.
There is no CPU fallback: without a CUDA device or the built library, fasta() raises.
"""

import numpy as np

from . import linalg, losses, proximal, stopping, tv
from . import _cabi, _device, _resident
from ._loop import Convergence, EPSILON, run as _run

__all__ = ["fasta", "Convergence", "EPSILON", "linalg", "losses", "proximal", "stopping", "tv"]

_OPTION_ORDER = ("adaptive", "accelerate", "verbose", "max_iters", "tolerance", "stop_rule", "L", "tau0",
                 "backtrack", "stepsize_shrink", "window", "max_backtracks", "restart", "evaluate_objective",
                 "record_iterates", "func")


def __getattr__(name):
    if name == "plots":            # matplotlib is optional and never needed by the solver
        import importlib
        return importlib.import_module(".plots", __name__)
    if name in ("distributed", "batched"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)


class _LooseMap(linalg.LinearMap):
    """Legacy (map, adjoint) callables whose range shape is not declared (tv_denoising.py:99)."""

    def __init__(self, map_func, adj_func, Vshape):
        self.map_func, self.adj_func = map_func, adj_func
        self.Vshape = tuple(Vshape) if Vshape is not None else None
        self.Wshape = None

    def __call__(self, v):
        return self.map_func(v)

    @property
    def H(self):
        return _LooseMap(self.adj_func, self.map_func, None)


def _owner(fn, cls):
    obj = getattr(fn, "__self__", None)
    return obj if isinstance(obj, cls) else None


def _split_arguments(args, kwargs):
    """Accept both call conventions: 6 leading positionals (new) or 7 (legacy)."""
    args = list(args)
    if "At" in kwargs:
        legacy = True
        args.insert(1, kwargs.pop("At"))
    elif len(args) >= 7 and not _device.is_array(args[5]):
        legacy = True
    elif len(args) >= 6:
        legacy = False
    else:
        raise TypeError("fasta(A, f, gradf, g, proxg, x0, ...) or fasta(A, At, f, gradf, g, proxg, x0, ...)")
    lead = 7 if legacy else 6
    head, rest = args[:lead], args[lead:]
    if len(rest) > len(_OPTION_ORDER):
        raise TypeError("too many positional arguments")
    for name, value in zip(_OPTION_ORDER, rest):
        if name in kwargs:
            raise TypeError(f"fasta() got multiple values for argument '{name}'")
        kwargs[name] = value
    unknown = set(kwargs) - set(_OPTION_ORDER)
    if unknown:
        raise TypeError(f"fasta() got unexpected keyword arguments {sorted(unknown)}")
    if legacy:
        A, At, f, gradf, g, proxg, x0 = head
        A = _legacy_operator(A, At, x0)
    else:
        A, f, gradf, g, proxg, x0 = head
    return A, f, gradf, g, proxg, x0, kwargs


def _legacy_operator(A, At, x0):
    if A is None:                                             # svm.py:74 and friends
        return linalg.LinearMap.identity(tuple(x0.shape))
    if isinstance(A, linalg.LinearMap):
        return A
    if _device.is_array(A):                                   # sparse_least_squares.py:46,76
        if len(x0.shape) == 2:                                # matrix unknowns: A @ X (mmv.py:65)
            return linalg.LinearMap.from_matrix(A).with_columns(int(x0.shape[1]))
        return linalg.LinearMap.from_matrix(A)
    if callable(A):
        if A is tv.div and At is tv.grad and len(x0.shape) == 3:
            return tv.divergence_map(tuple(x0.shape[:-1]))    # tv_denoising.py:99
        return _LooseMap(A, At, tuple(x0.shape))
    raise TypeError(f"cannot use {type(A)} as a linear operator")


def _driver_for(A):
    from . import _backends
    if hasattr(A, "_fb200_driver"):
        return A._fb200_driver()
    if getattr(A, "_fb200_dense", False) and not A.transposed and A.cols is None:
        return _backends.DenseDriver(A.matrix)
    if getattr(A, "_fb200_tv", False):
        return _backends.TVDriver(A.n0, A.n1)
    return None


def _make_backend(A, f, gradf, g, proxg, x0, accelerate, evaluate_objective):
    from . import _backends
    loss = _owner(f, losses._Loss)
    if loss is not None and _owner(gradf, losses._Loss) is not loss:
        loss = None
    if g is None:
        pen = proximal._Penalty()
    else:
        pen = _owner(proxg, proximal._Penalty)
        if pen is not None and (_owner(g, proximal._Penalty) is not pen or pen.tag is None):
            pen = None                                         # untagged / row-wise penalties: generic back-end
    driver = _driver_for(A) if (loss is not None and pen is not None) else None
    if driver is not None:
        return _backends.FusedBackend(driver, loss, pen, x0, accelerate)
    if g is None:                                             # reference __init__.py:88-90
        g = lambda x: 0
        proxg = lambda x, t: x
    return _backends.GenericBackend(A, f, gradf, g, proxg, x0, accelerate, evaluate_objective)


def fasta(*args, **kwargs) -> Convergence:
    """Run FASTA.  See the module docstring for the two call forms; options (defaults as in the
    reference, ``fasta/__init__.py:42-53``):

    adaptive=True, accelerate=False, verbose=True, max_iters=1000, tolerance=1e-5,
    stop_rule=stopping.hybrid_residual, L=None, tau0=None, backtrack=True, stepsize_shrink=None,
    window=10, max_backtracks=20, restart=True, evaluate_objective=False, record_iterates=False,
    func=None.  Returns a ``Convergence`` whose ``solution`` is the best iterate, in the array type
    of ``x0`` (numpy in -> numpy out, torch in -> CUDA tensor out).
    """
    A, f, gradf, g, proxg, x0, opts = _split_arguments(args, kwargs)
    _cabi.require_cuda()
    _cabi.load()
    be = _make_backend(A, f, gradf, g, proxg, x0, opts.get("accelerate", False),
                       opts.get("evaluate_objective", False))
    try:
        be.load()
        if _resident.eligible(be, opts):
            result = _resident.run(be, tuple(x0.shape), **opts)      # small problem: the whole loop in one kernel
        else:
            result = _run(be, tuple(x0.shape), **opts)
    finally:
        if hasattr(be, "close"):
            be.close()
    result.backend = type(be).__name__
    result.single_pass = bool(getattr(be, "use_sweep", False) or getattr(be, "use_sweep_accel", False))
    result.tv_fused = bool(getattr(be, "use_tv_fused", False) or getattr(be, "use_tv_accel", False))
    result.resident = bool(getattr(result, "resident", False))
    result.speculative = bool(getattr(be, "_spec_mode", False))
    result.kernel_launches = be.total_launches()
    result.peer_reductions = int(getattr(getattr(be, "drv", None), "peer_reductions", 0))   # fused NVLink all-reduce + BB calls
    return result

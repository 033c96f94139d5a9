"""TEST INFRASTRUCTURE ONLY -- synthetic problem generators for the FASTA hot-path oracle.

Each generator restates the ``construct()`` body of one reference example (cited per function),
drawing from numpy's *global legacy* RNG in exactly the reference's order so that, after
``np.random.seed(s)``, the same arrays come out as from the reference example.  The example
modules themselves cannot be imported, hence the restatement.

A problem is a plain ``Problem`` record: the data (``A``/``b``/``image``), tags naming the loss
and the penalty, the regularisation weight and the starting point.  ``numpy_callables`` turns a
record into the four numpy lambdas the reference's ``solve()`` builds.
"""

from dataclasses import dataclass, field
from typing import Optional

import numpy as np
from numpy import linalg as la


@dataclass
class Problem:
    kind: str                      # "dense" (explicit matrix) | "tv" (matrix-free div/grad)
    loss: str                      # "least_squares" | "logistic"
    penalty: str                   # "l1" | "l1ball" | "nonneg" | "none" | "tv_ball"
    mu: float
    x0: np.ndarray
    A: Optional[np.ndarray] = None
    b: Optional[np.ndarray] = None          # dense: observation vector; tv: image / mu
    x_true: Optional[np.ndarray] = None
    image: Optional[np.ndarray] = None      # tv only: the noisy image
    meta: dict = field(default_factory=dict)


# ----------------------------------------------------------------------------------------------
# Dense problems
# ----------------------------------------------------------------------------------------------

def _sparse_signal(N: int, K: int) -> np.ndarray:
    x = np.zeros(N)
    x[np.random.permutation(N)[:K]] = 1
    return x


def _scale_matrix(A: np.ndarray, scaling: str) -> None:
    """In-place column/row scaling of the measurement matrix.

    "svd":       A /= ||A||_2 as in the reference (sparse_least_squares.py:68) -- full SVD.
    "asymptote": A /= sqrt(M) + sqrt(N), the Gaussian spectral-norm asymptote SURVEY.md 8d
                 substitutes where the SVD is infeasible (configs 2 and 5).
    """
    if scaling == "svd":
        A /= la.norm(A, 2)
    elif scaling == "asymptote":
        A /= (np.sqrt(A.shape[0]) + np.sqrt(A.shape[1]))
    elif scaling != "none":
        raise ValueError(scaling)


def sparse_least_squares(M=200, N=1000, K=10, sigma=0.01, mu=0.02, scaling="svd") -> Problem:
    """Penalised lasso, min mu|x|_1 + .5|Ax-b|^2  (sparse_least_squares.py:51-76; configs 1, 2)."""
    x = _sparse_signal(N, K)
    A = np.random.randn(M, N)
    _scale_matrix(A, scaling)
    b = A @ x + sigma * np.random.randn(M)
    return Problem("dense", "least_squares", "l1", mu, np.zeros(N), A=A, b=b, x_true=x,
                   meta=dict(M=M, N=N, K=K, sigma=sigma, scaling=scaling))


def l1ball_lasso(M=200, N=1000, K=10, sigma=0.01, mu=0.8, scaling="svd") -> Problem:
    """Constrained lasso, min .5|Ax-b|^2 s.t. |x|_1 <= mu*|x_true|_1  (lasso.py:52-79)."""
    x = _sparse_signal(N, K)
    radius = mu * la.norm(x, 1)
    A = np.random.randn(M, N)
    _scale_matrix(A, scaling)
    b = A @ x + sigma * np.random.randn(M)
    return Problem("dense", "least_squares", "l1ball", radius, np.zeros(N), A=A, b=b, x_true=x,
                   meta=dict(M=M, N=N, K=K, sigma=sigma, scaling=scaling))


def nonneg_least_squares(M=200, N=1000, K=10, sigma=0.005, scaling="svd") -> Problem:
    """Non-negative least squares  (nn_least_squares.py:51-72)."""
    x = _sparse_signal(N, K)
    A = np.random.randn(M, N)
    _scale_matrix(A, scaling)
    b = A @ x + sigma * np.random.randn(M)
    return Problem("dense", "least_squares", "nonneg", 0.0, np.zeros(N), A=A, b=b, x_true=x,
                   meta=dict(M=M, N=N, K=K, sigma=sigma, scaling=scaling))


def sparse_logistic(M=1000, N=2000, K=5, mu=40.0) -> Problem:
    """l1-penalised logistic regression  (sparse_logistic.py:57-80; config 3)."""
    x = _sparse_signal(N, K)
    A = np.random.randn(M, N)
    p = 1 / (1 + np.exp(-A @ x))
    b = 2.0 * (np.random.rand(M) < p) - 1
    return Problem("dense", "logistic", "l1", mu, np.zeros(N), A=A, b=b, x_true=x,
                   meta=dict(M=M, N=N, K=K))


# ----------------------------------------------------------------------------------------------
# Total-variation denoising (matrix-free)
# ----------------------------------------------------------------------------------------------

def tv_grad(X: np.ndarray) -> np.ndarray:
    """Periodic forward-difference gradient, (n..)->(n..,ndim)  (tv_denoising.py:26-40)."""
    out = np.zeros(X.shape + (X.ndim,))
    for d in range(X.ndim):
        out[..., d] = np.roll(X, 1, axis=d) - X
    return out


def tv_div(Y: np.ndarray) -> np.ndarray:
    """Adjoint of ``tv_grad``: summed periodic backward differences  (tv_denoising.py:43-63)."""
    nd = Y.shape[-1]
    assert nd == Y.ndim - 1
    out = np.zeros(Y.shape[:-1])
    for d in range(nd):
        comp = Y[..., d]
        out += np.roll(comp, -1, axis=d) - comp
    return out


def checkerboard(n: int, cell: int = 64) -> np.ndarray:
    """Deterministic stand-in for scipy.misc.ascent (unavailable offline): cell-px checkerboard in {0,1}."""
    idx = np.arange(n) // cell
    return ((idx[:, None] + idx[None, :]) % 2).astype(float)


def tv_denoising(n=128, sigma=0.1, mu=0.1, cell=None) -> Problem:
    """Dual TV denoising, min_Y .5|div(Y) - M/mu|^2 s.t. |Y_ij|_2 <= 1  (tv_denoising.py:85-125).

    The image is a synthetic checkerboard normalised to max 1 like the
    reference (``M /= max(M)``, :117) plus ``sigma * randn(n, n)`` (:120).
    """
    cell = cell or max(n // 8, 1)
    img = checkerboard(n, cell)
    img /= np.max(img)
    img += sigma * np.random.randn(*img.shape)
    return Problem("tv", "least_squares", "tv_ball", mu, np.zeros(img.shape + (2,)),
                   b=img / mu, image=img, meta=dict(n=n, sigma=sigma, cell=cell))


# ----------------------------------------------------------------------------------------------
# numpy lambdas, exactly as the reference's solve() methods build them
# ----------------------------------------------------------------------------------------------

def _shrink(x, t):
    return np.sign(x) * np.maximum(np.abs(x) - t, 0)        # proximal.py:67


def _prox_tinf(x, t):
    """prox of t*|.|_inf  (proximal.py:12-31, the misnamed project_Linf_ball)."""
    n = len(x)
    mag = np.abs(x)
    desc = mag.copy()
    desc[::-1].sort()
    alpha = np.max((np.cumsum(desc) - t) / np.arange(1, n + 1))
    if alpha > 0:
        return np.minimum(mag, alpha) * np.sign(x)
    return np.zeros(n)


def _project_l1_ball(x, t):
    return x - _prox_tinf(x, t)                                # proximal.py:34-41


def _tv_ball(Y, t):
    nrm = np.maximum(la.norm(Y, axis=Y.ndim - 1), 1)            # tv_denoising.py:89-96
    return Y / nrm[..., np.newaxis]


def numpy_callables(p: Problem):
    """(f, gradf, g, proxg) as numpy lambdas, one-to-one with the reference's solve() bodies."""
    b = p.b
    mu = p.mu
    if p.loss == "least_squares":
        f = lambda z: .5 * la.norm((z - b).ravel()) ** 2          # sparse_least_squares.py:41
        gradf = lambda z: z - b                                    # :42
    elif p.loss == "logistic":
        f = lambda z: np.sum(np.log(1 + np.exp(z)) - (b == 1) * z)  # sparse_logistic.py:47
        gradf = lambda z: -b / (1 + np.exp(b * z))                  # :48
    else:
        raise ValueError(p.loss)

    if p.penalty == "l1":
        g = lambda x: mu * la.norm(x.ravel(), 1)                   # sparse_least_squares.py:43
        proxg = lambda x, t: _shrink(x, t * mu)                    # :44
    elif p.penalty == "l1ball":
        g = lambda x: 0                                            # lasso.py:44
        proxg = lambda x, t: _project_l1_ball(x, mu)               # lasso.py:45
    elif p.penalty == "nonneg":
        g = lambda x: 0                                            # nn_least_squares.py:41
        proxg = lambda x, t: np.maximum(x, 0)                      # :42
    elif p.penalty == "tv_ball":
        g = lambda Y: 0                                            # tv_denoising.py:87
        proxg = _tv_ball
    elif p.penalty == "none":
        g, proxg = None, None
    else:
        raise ValueError(p.penalty)
    return f, gradf, g, proxg


def numpy_operator(p: Problem):
    """(apply, adjoint, Vshape, Wshape) numpy functions for the problem's linear map."""
    if p.kind == "dense":
        A = p.A
        return (lambda x: A @ x), (lambda y: A.T @ y), (A.shape[1],), (A.shape[0],)   # linalg.py:41
    if p.kind == "tv":
        shape = p.x0.shape
        return tv_div, tv_grad, shape, shape[:-1]                                      # tv_denoising.py:99
    raise ValueError(p.kind)


# name -> (generator, kwargs); seeds are applied by the caller with np.random.seed(seed)
CASES = {
    # config 1a: reference defaults
    "lasso_200x1000_k10": (sparse_least_squares, dict()),
    # config 1b: 5% sparsity -> exercises the backtracking redo path
    "lasso_200x1000_k50": (sparse_least_squares, dict(K=50)),
    # ragged sizes (not multiples of any tile), asymptote scaling as configs 2/5
    "lasso_333x1414_k40": (sparse_least_squares, dict(M=333, N=1414, K=40, scaling="asymptote")),
    # mid-size twin of config 2 (1/10 scale per SURVEY 8d)
    "lasso_4000x10000_k500": (sparse_least_squares, dict(M=4000, N=10000, K=500, scaling="asymptote")),
    "l1ball_200x1000": (l1ball_lasso, dict()),
    "nnls_200x1000": (nonneg_least_squares, dict()),
    # config 3 twin: reference defaults
    "logistic_1000x2000": (sparse_logistic, dict()),
    # config 4 twins
    "tv_64": (tv_denoising, dict(n=64)),
    "tv_128": (tv_denoising, dict(n=128)),
}

# the three mode dicts of the reference harness (examples/__init__.py:74,80,86)
MODES = {
    "adaptive": dict(adaptive=True, accelerate=False),
    "accelerated": dict(adaptive=False, accelerate=True),
    "plain": dict(adaptive=False, accelerate=False),
}
HARNESS_OPTS = dict(tolerance=1e-5, evaluate_objective=True, verbose=False)


def build(case: str, seed: int = 0) -> Problem:
    gen, kw = CASES[case]
    np.random.seed(seed)
    return gen(**kw)

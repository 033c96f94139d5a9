"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference FASTA hot path.

This module is the CPU oracle the CUDA path is checked against, and the ``cpu_baseline`` arm of
``bench.py``.  It is NOT part of the product: nothing under ``fasta-python_b200/`` imports it.

What it restates (all line numbers are in /root/reference/fasta/):
  * ``solve``            the forward-backward-splitting loop           __init__.py:38-320
  * ``Trace``            the ``Convergence`` record                    __init__.py:323-351
  * ``shrink`` ...       the proximal operators                        proximal.py:12-67
  * ``stop_*``           the four stopping rules                       stopping.py:6-51

Arithmetic follows the reference expression by expression (same operand order, ``sqrt`` then
square for squared norms, un-conjugated dots, ``np.float64`` scalars) so that on the same numpy
build it reproduces the live reference BIT FOR BIT; ``tests/test_oracle_golden.py`` pins that
against ``tests/golden/*.npz`` (written by ``oracle/make_golden.py`` from the live reference).
Parity status: pinned against live-reference output; the reference has no own golden vectors.
"""

from time import time

import numpy as np
from numpy import linalg as la

EPS = 1e-12          # __init__.py:32


# ---------------------------------------------------------------------------------------------
# stopping.py
# ---------------------------------------------------------------------------------------------

def stop_residual(i, resid, norm_resid, max_resid, tol):           # stopping.py:15
    return resid < tol


def stop_norm_residual(i, resid, norm_resid, max_resid, tol):      # stopping.py:27
    return norm_resid < tol


def stop_ratio_residual(i, resid, norm_resid, max_resid, tol):     # stopping.py:39
    return resid / max_resid < tol


def stop_hybrid_residual(i, resid, norm_resid, max_resid, tol):    # stopping.py:51
    return resid / max_resid < tol or norm_resid < tol


# ---------------------------------------------------------------------------------------------
# proximal.py
# ---------------------------------------------------------------------------------------------

def shrink(x, t):                                                  # proximal.py:67
    return np.sign(x) * np.maximum(np.abs(x) - t, 0)


def prox_tinf(x, t):
    """The reference's ``project_Linf_ball`` (really prox of t*|.|_inf)   proximal.py:12-31."""
    n = len(x)
    mag = np.abs(x)
    desc = mag.copy()
    desc[::-1].sort()                       # ascending sort of the reversed view = descending
    alpha = np.max((np.cumsum(desc) - t) / np.arange(1, n + 1))
    if alpha > 0:
        return np.minimum(mag, alpha) * np.sign(x)
    return np.zeros(n)


def project_l1_ball(x, t):                                         # proximal.py:34-41 (Moreau)
    return x - prox_tinf(x, t)


def prox_nuclear(X, t):
    """The reference's ``project_Lnuc_ball`` (singular-value soft threshold)  proximal.py:44-55."""
    U, s, Vh = la.svd(X)
    S = np.zeros(X.shape)
    S[:len(s), :len(s)] = np.diag(shrink(s, t))
    return U @ S @ Vh


# ---------------------------------------------------------------------------------------------
# __init__.py
# ---------------------------------------------------------------------------------------------

class Trace:
    """Same fields as the reference ``Convergence``  (__init__.py:323-351)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def _nrm(v):
    return la.norm(v.ravel())


def _dot(a, b):
    return np.real(a.ravel().T @ b.ravel())


def estimate_lipschitz(op, adj, gradf, shape):
    """Randomised Lipschitz estimate, draws v1 then v2 from the global RNG  (__init__.py:100-113)."""
    v1 = np.random.randn(*shape)
    v2 = np.random.randn(*shape)
    d1 = adj(gradf(op(v1)))
    d2 = adj(gradf(op(v2)))
    L = _nrm(d1 - d2) / _nrm(v1 - v2)
    return L, (2 / L) / 10


def solve(op, adj, f, gradf, g, proxg, x0, *, adaptive=True, accelerate=False, max_iters=1000,
          tolerance=1e-5, stop_rule=stop_hybrid_residual, L=None, tau0=None, backtrack=True,
          stepsize_shrink=None, window=10, max_backtracks=20, restart=True,
          evaluate_objective=False, record_iterates=False, func=None, verbose=False):
    """Forward-backward splitting exactly as the reference runs it.

    ``op`` / ``adj`` are plain numpy callables for A and its adjoint (the reference reaches them
    through ``LinearMap.__call__`` / ``.H``, linalg.py:52-69, which only add shape asserts).
    """
    if g is None:                                                   # :88-90
        g = lambda x: 0
        proxg = lambda x, t: x
    if stepsize_shrink is None and backtrack:                       # :92-97
        stepsize_shrink = 0.2 if adaptive else 0.5
    if not L or not tau0:                                           # :100 (both needed to skip)
        L, tau0 = estimate_lipschitz(op, adj, gradf, x0.shape)

    if verbose:                                                     # :118-120
        print("Initializing FASTA...\n")
        print("Iteration #\tResidual\tStepsize\tAccel. param\tBacktracks\tObjective")

    resid_h = np.zeros(max_iters)                                   # :123-127
    nresid_h = np.zeros(max_iters)
    tau_h = np.zeros(max_iters)
    f_h = np.zeros(max_iters + 1)
    times = np.zeros(max_iters + 1)
    obj_h = np.zeros(max_iters + 1) if evaluate_objective else None
    it_h = np.zeros((max_iters + 1,) + x0.shape) if record_iterates else None
    fn_h = np.zeros(max_iters + 1) if func else None

    x_cur = x0                                                      # :132-139
    tau_next = tau0
    z_cur = op(x_cur)
    f_cur = f(z_cur)
    grad_cur = adj(gradf(z_cur))
    f_h[0] = f_cur
    if evaluate_objective:
        obj_h[0] = f_cur + g(x_cur)
    if record_iterates:
        it_h[0] = x_cur
    if func:
        fn_h[0] = func(x_cur)
    if accelerate:                                                  # :154-157
        xa_cur, za_cur, alpha_cur = x_cur, z_cur, 1.0

    n_backtracks = 0
    max_resid = -np.inf                                             # :165-167
    best_q = np.inf
    best_x = x0

    i = 0
    while i < max_iters:
        times[i] = time()
        x_prev, grad_prev, tau = x_cur, grad_cur, tau_next          # :176-178

        x_hat = x_prev - tau * grad_cur                             # :181
        x_cur = proxg(x_hat, tau)                                   # :184
        dx = x_cur - x_prev                                         # :186
        z_cur = op(x_cur)
        f_cur = f(z_cur)

        bt = 0
        if backtrack:                                               # :195-217
            f_max = np.max(f_h[max(i - window + 1, 0):(i + 1)])
            while f_cur - (f_max + _dot(dx, grad_prev) + _nrm(dx) ** 2 / (2 * tau)) > EPS \
                    and bt < max_backtracks:
                tau *= stepsize_shrink
                x_hat = x_prev - tau * grad_prev
                x_cur = proxg(x_hat, tau)
                dx = x_cur - x_prev
                z_cur = op(x_cur)
                f_cur = f(z_cur)
                bt += 1
            n_backtracks += bt

        if accelerate:                                              # :220-245
            xa_prev, za_prev = xa_cur, za_cur
            xa_cur, za_cur = x_cur, z_cur
            alpha_prev = alpha_cur
            if restart and (x_prev - x_cur).ravel().T @ (x_cur - xa_prev).ravel() > 1e-30:
                alpha_prev = 1.0
                if verbose:                                         # :235
                    print("Restarted acceleration.")
            alpha_cur = (1 + np.sqrt(1 + 4 * alpha_prev ** 2)) / 2
            x_cur = x_cur + (alpha_prev - 1) / alpha_cur * (xa_cur - xa_prev)
            z_cur = z_cur + (alpha_prev - 1) / alpha_cur * (za_cur - za_prev)
            f_cur = f(z_cur)

        grad_cur = adj(gradf(z_cur))                                # :248-249
        tau_next = tau

        if adaptive:                                                # :253-270
            dg = grad_cur + (x_hat - x_prev) / tau
            dd = _dot(dx, dg)
            tau_s = _nrm(dx) ** 2 / dd
            tau_m = max(dd / _nrm(dg) ** 2, 0)
            tau_next = tau_m if 2 * tau_m > tau_s else tau_s - .5 * tau_m
            if tau_next <= 0 or np.isinf(tau_next) or np.isnan(tau_next):
                tau_next = tau * 1.5

        resid_h[i] = _nrm(dx) / tau                                 # :272-281
        normalizer = max(_nrm(grad_prev), _nrm(x_cur - x_hat) / tau) + EPS
        tau_h[i] = tau
        nresid_h[i] = resid_h[i] / normalizer
        f_h[i + 1] = f_cur
        max_resid = max(max_resid, resid_h[i])

        if evaluate_objective:                                      # :284-300
            obj_h[i + 1] = f_cur + g(x_cur)
            quality = obj_h[i + 1]
        else:
            quality = resid_h[i]
        if record_iterates:
            it_h[i + 1, ...] = x_cur
        if func:
            fn_h[i + 1] = func(x_cur)
        if quality < best_q:
            best_x, best_q = x_cur, quality

        if verbose:                                                 # :302-306 (alpha0; objective of the PREVIOUS iterate)
            print("[{:<6}]\t{:e}\t{:e}\t{:e}\t{:6}\t{:e}".format(
                i, resid_h[i], tau_h[i], alpha_prev if accelerate else 0.0, bt if backtrack else 0,
                obj_h[i] if evaluate_objective else 0))

        if stop_rule(i, resid_h[i], nresid_h[i], max_resid, tolerance):   # :308-312
            i += 1
            break
        i += 1

    times[i] = time()                                               # :315
    return Trace(residuals=resid_h, norm_residuals=nresid_h, stepsizes=tau_h,
                 backtracks=n_backtracks, times=times, iteration_count=i, solution=best_x,
                 objectives=obj_h, iterates=it_h, function_hist=fn_h, f_hist=f_h)


def solve_problem(problem, **opts):
    """Run the oracle on an ``oracle.problems.Problem`` (draws tau0 from the global RNG)."""
    from . import problems as _p
    op, adj, _, _ = _p.numpy_operator(problem)
    f, gradf, g, proxg = _p.numpy_callables(problem)
    return solve(op, adj, f, gradf, g, proxg, problem.x0, **opts)

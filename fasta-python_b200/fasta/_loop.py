"""Host side of the forward-backward-splitting loop.

The control flow, the scalar step-size algebra, the histories and the stopping test stay on the
host in ``np.float64`` exactly as in the reference (``fasta/__init__.py:38-320``); every array
operation is delegated to a *backend* whose methods enqueue sm_100a kernels and hand back the
handful of reduction scalars a decision needs.  One backend call = one decision point = one
device->host sync:

    load(x0)                 copy the start point into device buffers
    lipschitz(v1, v2)        -> (|A^H gradf(A v1) - A^H gradf(A v2)|, |v1 - v2|)      ref :106-110
    start()                  z = A x, f(z), gradf1 = A^H gradf(z)  -> Scalars(f, pen, g_sq)   ref :135-143
    advance()                x0 <- x1, gradf0 <- gradf1 (buffer rotation)            ref :176-177
    trial(tau)               x1hat, x1 = prox, Dx, z1 = A x1, f1  -> Scalars(f, dx_g0, dx_sq,
                             xmxh_sq, pen, restart)                                   ref :181-188
    extrapolate(c)           FISTA step on x1 and z1, f1 again -> Scalars(f, xmxh_sq, pen)   ref :242-245
    gradient(tau, adaptive)  gradf1 = A^H gradf(z1) -> Scalars(dx_dg, dg_sq, g_sq)    ref :248-260
    keep_best()              best <- x1                                               ref :298-300
    iterate()                current x1 in the caller's array type                    ref :292,296
    solution()               best iterate in the caller's array type                  ref :317

Optional protocols of the fused back-ends, used when present (see ``run``):

    lipschitz_push / lipschitz_finish, start_async     pipelined prologue
    trial_accel(tau, alpha_prev, restart)              FISTA trial incl. extrapolation and gradient, one collect
    trial_launch / trial_finish                        next trial queued right after the decision (run-ahead)
    speculate_begin, _queue_trial(tau | None, bt, host_state), _collect_trial, rotation / restore
                                                       next trial queued BEFORE this trial's sums are read: the device
                                                       repeats the decisions (fb200_trial_decide), tau=None = "use the
                                                       step size left on the device, return at once if told to skip"
    _queue_accel(tau, c), _collect_accel, accel_weight the same for FISTA trials, step size and weight by value
"""

import os
from time import time

import numpy as np

from . import stopping

EPSILON = 1E-12      # reference __init__.py:32


class Scalars:
    """Attribute bag of np.float64 reduction results."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


class Convergence:
    """Convergence record: same attributes as the reference's (``fasta/__init__.py:323-351``).

    residuals, norm_residuals, stepsizes: arrays of length max_iters (zero past iteration_count)
    backtracks: total number of line-search backtracks (int)
    times: wall-clock at the start of each iteration, and at exit in times[iteration_count]
    iteration_count: iterations performed
    solution: the BEST iterate (lowest objective if evaluate_objective else smallest residual)
    objectives / iterates / function_hist: optional histories (None when not requested)
    """

    def __init__(self, residuals, norm_residuals, stepsizes, backtracks, times, iteration_count, solution,
                 objectives=None, iterates=None, function_hist=None):
        self.residuals = residuals
        self.norm_residuals = norm_residuals
        self.stepsizes = stepsizes
        self.backtracks = backtracks
        self.times = times
        self.iteration_count = iteration_count
        self.solution = solution
        self.objectives = objectives
        self.iterates = iterates
        self.function_hist = function_hist


_BUILTIN_RULES = {stopping.residual: 0, stopping.norm_residual: 1, stopping.ratio_residual: 2,
                  stopping.hybrid_residual: 3}


def _sq(norm_squared):
    """la.norm(v)**2 the way the reference forms it: sqrt of the dot, then squared."""
    return np.sqrt(norm_squared) ** 2


def lipschitz_probes(be, x0_shape):
    """|A^H gradf(A x1) - A^H gradf(A x2)| and |x1 - x2| for two standard-normal probes (ref :102-110).

    The probes are the next 2 N values of numpy's global legacy stream, x1 first, exactly as the reference's two
    ``np.random.randn(*x0.shape)`` calls -- either drawn on the device (fused back-ends, ``fasta/_rng.py``: the kernels
    continue the very stream and hand numpy its end state back) or on the host and uploaded."""
    if hasattr(be, "lipschitz_push"):
        # pipelined prologue: the device starts on z = A x0 / gradf1 (ref :135-139, independent of the probes)
        be.start_async()
        if be.lipschitz_device_ok():
            out = be.lipschitz_device()
            if out is not None:
                return out
        # ... and on each probe's sweep while the host draws the next probe
        be.lipschitz_push(0, np.random.randn(*x0_shape))
        be.lipschitz_push(1, np.random.randn(*x0_shape))
        return be.lipschitz_finish()
    v1 = np.random.randn(*x0_shape)                         # ref :102-103
    v2 = np.random.randn(*x0_shape)
    return be.lipschitz(v1, v2)


def run(be, x0_shape, *, adaptive=True, accelerate=False, verbose=True, max_iters=1000, tolerance=1e-5,
        stop_rule=stopping.hybrid_residual, L=None, tau0=None, backtrack=True, stepsize_shrink=None,
        window=10, max_backtracks=20, restart=True, evaluate_objective=False, record_iterates=False,
        func=None) -> Convergence:
    """Drive ``be`` through FASTA.  Options and defaults are the reference's (``__init__.py:42-53``)."""
    if stepsize_shrink is None and backtrack:               # ref :92-97
        stepsize_shrink = 0.2 if adaptive else 0.5

    if not L or not tau0:                                   # ref :100 -- both are needed to skip
        dgrad, dpoint = lipschitz_probes(be, x0_shape)
        L = dgrad / dpoint                                  # ref :110
        tau0 = (2 / L) / 10                                 # ref :113
    if not tau0:                                            # ref :115-116 (unreachable, kept)
        tau0 = 1 / L

    if verbose:                                             # ref :118-120
        print("Initializing FASTA...\n")
        print("Iteration #\tResidual\tStepsize\tAccel. param\tBacktracks\tObjective")

    residual_hist = np.zeros(max_iters)                     # ref :123-127
    norm_residual_hist = np.zeros(max_iters)
    tau_hist = np.zeros(max_iters)
    f_hist = np.zeros(max_iters + 1)
    times = np.zeros(max_iters + 1)
    objective_hist = np.zeros(max_iters + 1) if evaluate_objective else None
    iterate_hist = np.zeros((max_iters + 1,) + tuple(x0_shape)) if record_iterates else None
    function_hist = np.zeros(max_iters + 1) if func else None

    tau1 = tau0
    s = be.start()                                          # ref :135-139
    f1 = s.f
    g1_sq = s.g_sq
    f_hist[0] = f1
    if evaluate_objective:
        objective_hist[0] = f1 + s.pen                      # ref :143
    if record_iterates:
        iterate_hist[0] = _host(be.iterate())
    if func:
        function_hist[0] = func(be.iterate())
    alpha1 = 1.0                                            # ref :157
    alpha0 = 0.0
    total_backtracks = 0
    max_residual = -np.inf                                  # ref :165-166
    best_quality = np.inf

    # Back-ends that can queue a trial without waiting for it get the NEXT iteration's trial queued as soon as the
    # new step size is known (end of the loop body); histories, best-iterate copy and stop rule then run on the host
    # while the device works.  A trial queued before a stop is simply never collected.
    # FISTA trial incl. extrapolation in one pass (dense single-pass sweep / TV whole-iteration kernel)
    fused_accel = accelerate and (getattr(be, "use_sweep_accel", False) or getattr(be, "use_tv_accel", False))
    run_ahead = hasattr(be, "trial_launch") and not fused_accel and os.environ.get("FASTA_B200_RUN_AHEAD", "1") != "0"
    queued = False
    # Speculative run-ahead (back-ends with speculate_ok): the Barzilai-Borwein step size of a trial is formed on the
    # device right behind it (fb200_trial_decide, the algebra of :253-270), so the NEXT iteration's trial -- which in
    # the common case differs from this one only in the step size and the buffer rotation -- is queued before the host
    # has read this trial's sums.  The device never waits for the host.  If the line search rejects the trial, the
    # speculated one is dropped (it only wrote scratch buffers) and the backtracking runs as usual.
    speculate = run_ahead and not accelerate and getattr(be, "speculate_ok", False) and 1 <= window < 40
    if speculate:
        rule_id = _BUILTIN_RULES.get(stop_rule, -1)          # a user's rule is evaluated by the host only
        be.speculate_begin(f1, g1_sq, adaptive, backtrack, max_backtracks, window, rule_id, tolerance)
    # FISTA trials (fused back-ends, fixed step size): the next trial's step size and -- unless this trial restarts the
    # acceleration -- its extrapolation weight are known before this trial's sums are: it is queued by value ahead of
    # the collect, and dropped (it only wrote scratch buffers) if this trial restarts or is rejected by the line search.
    spec_accel = (fused_accel and not adaptive and getattr(be, "accel_speculate_ok", False)
                  and os.environ.get("FASTA_B200_RUN_AHEAD", "1") != "0")
    spec_stats = dict(speculated=0, dropped=0, mismatched=0)
    by_value = True           # whether `pending` was queued with the host's step size (else: the device's)
    pending = None            # handle of the trial queued for iteration i
    ahead = None              # handle of the trial speculatively queued for iteration i + 1

    i = 0
    while i < max_iters:
        times[i] = time()                                   # ref :173
        g0_sq = g1_sq
        tau0 = tau1
        if speculate:
            host_state = (i, max_residual, g0_sq)           # re-arms the device's copy of the loop state (by-value trials)
            if pending is None:
                be.advance()                                # ref :176-178
                pending = be._queue_trial(tau0, 0, host_state)     # ref :181-188
                by_value = True
            be._ahead = False
            rot = be.rotation()
            ahead = None
            if i + 1 < max_iters:
                be.advance()
                be._ahead = True
                ahead = be._queue_trial(None)               # step size and go / no-go are on the device
                spec_stats["speculated"] += 1
            t = be._collect_trial(pending)
            pending = None
            if t.skipped:                                   # the device's decision differed from the host's (it never
                spec_stats["mismatched"] += 1
                if ahead is not None:                       # should): run the trial now, by value
                    be.restore(rot)
                    ahead = None
                t = be.trial(tau0, 0, host_state)
                by_value = True
            if not by_value:
                tau0 = t.tau_used                           # the device's value of the :253-270 algebra, the one it used
        elif spec_accel:
            if pending is None:
                be.advance()                                # ref :176-178
                pending = be._queue_accel(tau0, be.accel_weight(alpha1))
            be._ahead = False
            rot = be.rotation()
            ahead = None
            if i + 1 < max_iters:
                be.advance()
                be._ahead = True                            # weight of trial i+1 if trial i does not restart (ref :238-240)
                ahead = be._queue_accel(tau0, be.accel_weight((1 + np.sqrt(1 + 4 * alpha1 ** 2)) / 2))
                spec_stats["speculated"] += 1
            t = be._collect_accel(pending)
            pending = None
            if restart and t.restart > 1E-30 and t.c != 0.0:            # ref :231 fires: the weight must be 0, repeat
                if ahead is not None:
                    be.restore(rot)
                    ahead = None
                    spec_stats["dropped"] += 1
                t = be._collect_accel(be._queue_accel(tau0, 0.0))
        elif queued:
            t = be.trial_finish()
            queued = False
        else:
            be.advance()                                    # ref :176-178
            t = be.trial_accel(tau0, alpha1, restart) if fused_accel else be.trial(tau0)   # ref :181-188
        f1 = t.f

        backtrack_count = 0
        if backtrack:                                       # ref :195-217
            f_window_max = np.max(f_hist[max(i - window + 1, 0):(i + 1)])
            while f1 - (f_window_max + t.dx_g0 + _sq(t.dx_sq) / (2 * tau0)) > EPSILON \
                    and backtrack_count < max_backtracks:
                tau0 *= stepsize_shrink
                if ahead is not None:                       # the speculated trial assumed acceptance: drop it
                    be.restore(rot)                         # (on the device it returned at once)
                    ahead = None
                    spec_stats["dropped"] += 1
                if speculate:
                    t = be.trial(tau0, backtrack_count + 1, host_state)
                else:
                    t = be.trial_accel(tau0, alpha1, restart) if fused_accel else be.trial(tau0)
                f1 = t.f
                backtrack_count += 1
            total_backtracks += backtrack_count

        dx_norm = np.sqrt(t.dx_sq)
        xmxh_sq, pen = t.xmxh_sq, t.pen

        if accelerate:                                      # ref :220-245
            alpha0 = alpha1
            if restart and t.restart > 1E-30:
                alpha0 = 1.0
                if verbose:
                    print("Restarted acceleration.")
            alpha1 = (1 + np.sqrt(1 + 4 * alpha0 ** 2)) / 2
            # fused FISTA trial: the accepted trial already extrapolated with this very weight
            e = t.extrap if fused_accel else be.extrapolate((alpha0 - 1) / alpha1)
            f1, xmxh_sq, pen = e.f, e.xmxh_sq, e.pen

        gr = be.gradient(tau0, adaptive)                    # ref :248-249
        g1_sq = gr.g_sq
        tau1 = tau0

        if adaptive:                                        # ref :253-270
            dotprod = gr.dx_dg
            tau_s = dx_norm ** 2 / dotprod
            tau_m = max(dotprod / _sq(gr.dg_sq), 0)
            if 2 * tau_m > tau_s:
                tau1 = tau_m
            else:
                tau1 = tau_s - .5 * tau_m
            if tau1 <= 0 or np.isinf(tau1) or np.isnan(tau1):
                tau1 = tau0 * 1.5

        residual_hist[i] = dx_norm / tau0                   # ref :272-281
        normalizer = max(np.sqrt(g0_sq), np.sqrt(xmxh_sq) / tau0) + EPSILON
        tau_hist[i] = tau0
        norm_residual_hist[i] = residual_hist[i] / normalizer
        f_hist[i + 1] = f1
        max_residual = max(max_residual, residual_hist[i])

        # ref :308 -- evaluated here (it depends only on the residuals above) so that no trial is queued past the end
        stop = stop_rule(i, residual_hist[i], norm_residual_hist[i], max_residual, tolerance)
        if spec_accel:
            if ahead is None and not stop and i + 1 < max_iters:
                be.advance()                                # after a restart / backtracked iteration: queue the next trial now
                be._ahead = True
                ahead = be._queue_accel(tau1, be.accel_weight(alpha1))
            pending, ahead = ahead, None
        elif speculate:
            if ahead is None and not stop and i + 1 < max_iters:
                be.advance()                                # after a backtracked iteration: queue the next trial now
                be._ahead = True
                ahead = be._queue_trial(tau1, 0, (i + 1, max_residual, g1_sq))
                by_value = True
            else:
                by_value = False                            # the successor (if any) was queued speculatively
            pending, ahead = ahead, None
        elif run_ahead and not stop and i + 1 < max_iters:
            be.advance()                                    # next iteration's ref :176-188, queued now
            be.trial_launch(tau1)
            queued = True

        if evaluate_objective:                              # ref :284-300
            objective_hist[i + 1] = f1 + pen
            quality = objective_hist[i + 1]
        else:
            quality = residual_hist[i]
        if record_iterates:
            iterate_hist[i + 1, ...] = _host(be.iterate())
        if func:
            function_hist[i + 1] = func(be.iterate())
        if quality < best_quality:
            be.keep_best()
            best_quality = quality

        if verbose:                                         # ref :302-306
            print("[{:<6}]\t{:e}\t{:e}\t{:e}\t{:6}\t{:e}".format(
                i, residual_hist[i], tau_hist[i], alpha0 if accelerate else 0.0,
                backtrack_count if backtrack else 0, objective_hist[i] if evaluate_objective else 0))

        if stop:                                            # ref :308-312
            i += 1
            break
        i += 1

    times[i] = time()                                       # ref :315
    res = Convergence(residual_hist, norm_residual_hist, tau_hist, total_backtracks, times, i, be.solution(),
                      objective_hist, iterate_hist, function_hist)
    res.speculation = spec_stats if (speculate or spec_accel) else None
    return res


def _host(a):
    if isinstance(a, np.ndarray):
        return a
    return a.detach().cpu().numpy()

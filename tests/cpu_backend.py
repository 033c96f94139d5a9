"""TEST DOUBLE (test infrastructure, never shipped): a CPU implementation of the back-end protocol
of ``fasta._loop`` so that the HOST logic -- the FBS control flow in ``fasta/_loop.py`` and the
collective choreography of ``fasta._backends.ShardedDriver`` -- can be exercised without a GPU
(``-m "not gpu"`` tests; gloo world_size-2 tests).  The product never imports this module and
``fasta.fasta()`` cannot select it.

Buffers are torch CPU tensors (so ``torch.distributed`` works on them); the arithmetic restates
the same reference expressions the CUDA kernels implement.
"""

import numpy as np
import torch

from fasta import _cabi as S
from fasta._loop import Scalars


class CpuWorkspace:
    def __init__(self):
        self.scal = torch.zeros(S.NSCAL, dtype=torch.float64)

    def fetch(self):
        return self.scal.numpy().copy()


def _loss(tag, z, b):
    if tag == S.LOSS_LEAST_SQUARES:
        r = z - b
        return r, torch.dot(r, r)
    if tag == S.LOSS_LOGISTIC:
        f = torch.sum(torch.log(1 + torch.exp(z)) - (b == 1).double() * z)
        return -b / (1 + torch.exp(b * z)), f
    return None, torch.tensor(0.0, dtype=torch.float64)


class CpuDenseDriver:
    """Same interface as fasta._backends.DenseDriver, torch CPU arithmetic."""

    def __init__(self, matrix):
        self.A = torch.as_tensor(matrix, dtype=torch.float64)
        self.M, self.N = self.A.shape
        self.xshape, self.zshape = (self.N,), (self.M,)
        self.launches = 0

    def workspace_dims(self):
        return self.M, self.N

    def forward(self, x, loss_tag, b, z, r, ws):
        z.copy_(self.A @ x)
        rr, f = _loss(loss_tag, z, b)
        if rr is not None:
            r.copy_(rr)
            ws.scal[S.S_F] = f
        self.launches += 2

    def adjoint(self, r, g, bb, x0, xhat, dx, tau, ws):
        g.copy_(self.A.T @ r)
        if bb:
            self.bb_reduce(g, x0, xhat, dx, tau, bb >= 2, ws)
        self.launches += 2

    def bb_reduce(self, g, x0, xhat, dx, tau, adaptive, ws):
        ws.scal[S.S_G1_SQ] = torch.dot(g, g)
        if adaptive:
            dg = g + (xhat - x0) / tau
            ws.scal[S.S_DX_DG] = torch.dot(dx, dg)
            ws.scal[S.S_DG_SQ] = torch.dot(dg, dg)

    def sync_point(self, v1, v2):
        pass


class CpuTVDriver:
    def __init__(self, n0, n1):
        self.n0, self.n1 = n0, n1
        self.xshape, self.zshape = (n0, n1, 2), (n0, n1)
        self.launches = 0

    def workspace_dims(self):
        return 1, 1

    def forward(self, x, loss_tag, b, z, r, ws):
        Y = x.view(self.n0, self.n1, 2)
        out = torch.zeros(self.n0, self.n1, dtype=torch.float64)
        for d in range(2):
            out += torch.roll(Y[..., d], -1, dims=d) - Y[..., d]
        z.copy_(out.reshape(-1))
        rr, f = _loss(loss_tag, z, b.reshape(-1) if b is not None else None)
        if rr is not None:
            r.copy_(rr)
            ws.scal[S.S_F] = f

    def adjoint(self, r, g, bb, x0, xhat, dx, tau, ws):
        R = r.view(self.n0, self.n1)
        G = torch.stack([torch.roll(R, 1, dims=0) - R, torch.roll(R, 1, dims=1) - R], dim=-1)
        g.copy_(G.reshape(-1))
        if bb:
            CpuDenseDriver.bb_reduce(self, g, x0, xhat, dx, tau, bb >= 2, ws)

    def sync_point(self, v1, v2):
        pass


def _prox(tag, h, p0, p1, radius=None):
    if tag == S.PROX_SHRINK:
        return torch.sign(h) * torch.clamp(torch.abs(h) - p0, min=0)
    if tag == S.PROX_NONNEG:
        return torch.clamp(h, min=0)
    if tag == S.PROX_BOX:
        return torch.clamp(h, min=p0, max=p1)
    if tag == S.PROX_TV_BALL:
        Y = h.view(-1, 2)
        nrm = torch.clamp(torch.linalg.norm(Y, dim=1), min=1)
        return (Y / nrm[:, None]).reshape(-1)
    if tag == S.PROX_L1BALL:
        x = h.numpy()
        mag = np.abs(x)
        if mag.sum() <= radius:
            return h.clone()
        desc = np.sort(mag)[::-1]
        theta = np.max((np.cumsum(desc) - radius) / np.arange(1, len(x) + 1))
        return torch.from_numpy(np.sign(x) * np.maximum(mag - theta, 0))
    return h.clone()


class CpuFusedBackend:
    """Mirror of fasta._backends.FusedBackend on torch CPU tensors."""

    def __init__(self, driver, loss_tag, b, penalty, x0, accelerate):
        self.drv = driver
        self.loss_tag = loss_tag
        self.b = torch.as_tensor(np.asarray(b, dtype=np.float64)).reshape(-1)
        self.pen = penalty
        self.accelerate = accelerate
        self.shape = tuple(x0.shape)
        self.x0_in = x0
        self.n = int(np.prod(driver.xshape))
        self.m = int(np.prod(driver.zshape))
        self.ws = CpuWorkspace()
        new = lambda k: torch.zeros(k, dtype=torch.float64)
        self.x1, self.g1 = new(self.n), new(self.n)
        self.z, self.r = new(self.m), new(self.m)

    def _finalize_f(self, raw):
        return .5 * np.sqrt(raw) ** 2 if self.loss_tag == S.LOSS_LEAST_SQUARES else raw

    def total_launches(self):
        return 0

    def load(self):
        self.x1 = torch.from_numpy(np.array(self.x0_in, dtype=np.float64).reshape(-1))
        self.best = self.x1.clone()
        self.xa1 = self.x1.clone()

    def lipschitz(self, v1, v2):
        a = torch.from_numpy(np.ascontiguousarray(v1.reshape(-1)))
        b = torch.from_numpy(np.ascontiguousarray(v2.reshape(-1)))
        self.drv.sync_point(a, b)
        ds = []
        for v in (a, b):
            d = torch.zeros(self.n, dtype=torch.float64)
            self.drv.forward(v, self.loss_tag, self.b, self.z, self.r, self.ws)
            self.drv.adjoint(self.r, d, 0, None, None, None, 0.0, self.ws)
            ds.append(d)
        return np.float64(torch.linalg.norm(ds[0] - ds[1])), np.float64(torch.linalg.norm(a - b))

    def start(self):
        z = torch.zeros(self.m, dtype=torch.float64)
        self.drv.forward(self.x1, self.loss_tag, self.b, z, self.r, self.ws)
        self.za1 = z.clone()
        self.g1 = torch.zeros(self.n, dtype=torch.float64)
        self.drv.adjoint(self.r, self.g1, 1, None, None, None, 0.0, self.ws)
        s = self.ws.fetch()
        return Scalars(f=self._finalize_f(s[S.S_F]), pen=self.pen.value(np.float64(self.x1.abs().sum())),
                       g_sq=s[S.S_G1_SQ])

    def advance(self):
        self.x0, self.g0 = self.x1, self.g1
        self.xa0, self.za0 = self.xa1, self.za1

    def trial(self, tau):
        tau = float(tau)
        self.xhat = self.x0 - tau * self.g0
        p0, p1 = self.pen.params(np.float64(tau))
        x1 = _prox(self.pen.tag, self.xhat, float(p0), float(p1), getattr(self.pen, "radius", None))
        self.dx = x1 - self.x0
        z = torch.zeros(self.m, dtype=torch.float64)
        self.drv.forward(x1, self.loss_tag, self.b, z, self.r, self.ws)
        s = self.ws.fetch()
        restart = np.float64(torch.dot(self.x0 - x1, x1 - self.xa0))
        self.x1, self.xa1, self.za1 = x1, x1, z
        return Scalars(f=self._finalize_f(s[S.S_F]), dx_g0=np.float64(torch.dot(self.dx, self.g0)),
                       dx_sq=np.float64(torch.dot(self.dx, self.dx)),
                       xmxh_sq=np.float64(torch.dot(x1 - self.xhat, x1 - self.xhat)),
                       pen=self.pen.value(np.float64(x1.abs().sum())), restart=restart)

    def extrapolate(self, c):
        c = float(c)
        self.x1 = self.xa1 + c * (self.xa1 - self.xa0)
        z = self.za1 + c * (self.za1 - self.za0)
        rr, f = _loss(self.loss_tag, z, self.b)
        self.r = rr
        self.ws.scal[S.S_F] = f
        if hasattr(self.drv, "reduce_loss"):
            self.drv.reduce_loss(self.ws)
        s = self.ws.fetch()
        e = self.x1 - self.xhat
        return Scalars(f=self._finalize_f(s[S.S_F]), xmxh_sq=np.float64(torch.dot(e, e)),
                       pen=self.pen.value(np.float64(self.x1.abs().sum())))

    def gradient(self, tau, adaptive):
        self.g1 = torch.zeros(self.n, dtype=torch.float64)
        self.drv.adjoint(self.r, self.g1, 2 if adaptive else 1, self.x0, self.xhat, self.dx, float(tau), self.ws)
        s = self.ws.fetch()
        return Scalars(dx_dg=s[S.S_DX_DG], dg_sq=s[S.S_DG_SQ], g_sq=s[S.S_G1_SQ])

    def keep_best(self):
        self.best = self.x1.clone()

    def iterate(self):
        return self.x1.numpy().reshape(self.shape)

    def solution(self):
        return self.best.numpy().reshape(self.shape).copy()


class CpuSpeculatingBackend(CpuFusedBackend):
    """The speculative run-ahead protocol of fasta._backends.FusedBackend (speculate_ok, _queue_trial(tau or None),
    _collect_trial, rotation / restore) with eager CPU arithmetic: a queued trial is computed at once, its sums are
    kept in its handle, and 'the step size on the device' is whatever the most recently QUEUED trial left there --
    including trials the loop later drops -- exactly as the stream order makes it on the GPU."""

    speculate_ok = True

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self._ahead = False
        self._spec = None
        self.dev_tau = None
        self.dropped = 0
        self.queued = 0
        self.skipped = 0
        self.mismatch = 0

    def trial_launch(self, tau):               # the plain run-ahead protocol (used when speculation is not possible)
        assert not hasattr(self, "dev"), "the speculative protocol does not use trial_launch"
        self._pending_t = CpuFusedBackend.trial(self, tau)
        self._ahead = True

    def trial_finish(self):
        self._ahead = False
        return self._pending_t

    def speculate_begin(self, f0, g0_sq, adaptive, backtrack, max_backtracks, window, stop_rule_id, tolerance):
        self.adaptive = bool(adaptive)
        self.cfg = (bool(backtrack), max_backtracks, window, stop_rule_id, tolerance)
        self.dev = dict(skip=False, it=0, maxres=-np.inf, g0sq=np.float64(g0_sq), f=[np.float64(f0)])

    _ROT = ("x0", "g0", "x1", "g1", "xa0", "za0", "xa1", "za1", "_ahead")

    def rotation(self):
        return tuple(getattr(self, k, None) for k in self._ROT)

    def restore(self, state):
        for k, v in zip(self._ROT, state):
            setattr(self, k, v)
        self.dropped += 1

    # ---- FISTA: queue / collect with the extrapolation weight passed by value (FusedBackend._queue_accel) ----------
    @property
    def use_sweep_accel(self):
        return self.accelerate

    accel_speculate_ok = True

    @staticmethod
    def accel_weight(alpha_prev):
        alpha1 = (1 + np.sqrt(1 + 4 * alpha_prev ** 2)) / 2
        return (alpha_prev - 1) / alpha1

    def _queue_accel(self, tau, c):
        self.queued += 1
        t = CpuFusedBackend.trial(self, tau)
        e = CpuFusedBackend.extrapolate(self, c)
        g = CpuFusedBackend.gradient(self, tau, True)       # the fused trial always forms the BB sums (bb = 2)
        return "dense", (t, e, g), c

    def _collect_accel(self, handle):
        t, e, g = handle[1]
        t.extrap, t.c = e, handle[2]
        self._spec = g
        return t

    def trial_accel(self, tau, alpha_prev, restart):
        t = self._collect_accel(self._queue_accel(tau, self.accel_weight(alpha_prev)))
        if restart and t.restart > 1E-30 and t.c != 0.0:
            t = self._collect_accel(self._queue_accel(tau, 0.0))
        return t

    def _queue_trial(self, tau, bt=0, host=(0, -np.inf, 0.0)):
        """Eager: the trial's kernels and fb200_trial_decide, with the device-side state in self.dev."""
        self.queued += 1
        d = self.dev
        if tau is not None:                    # by-value trials re-arm the device's loop state from the host's
            d["it"], d["maxres"], d["g0sq"] = int(host[0]), np.float64(host[1]), np.float64(host[2])
            del d["f"][d["it"] + 1:]
        if tau is None and d["skip"]:
            self.skipped += 1
            return "sweep", (Scalars(skipped=True), None)
        tau0 = np.float64(self.dev_tau if tau is None else tau)
        t = CpuFusedBackend.trial(self, tau0)
        g = CpuFusedBackend.gradient(self, tau0, self.adaptive)
        t.skipped = False
        backtrack, max_bt, window, rule, tol = self.cfg
        with np.errstate(all="ignore"):
            if backtrack and bt < max_bt:
                fwin = np.max(d["f"][max(d["it"] - window + 1, 0):d["it"] + 1])
                if t.f - (fwin + t.dx_g0 + np.sqrt(t.dx_sq) ** 2 / (2 * tau0)) > 1e-12:
                    d["skip"] = True
                    t.tau_used = tau0
                    return "sweep", (t, g)
            tau1 = tau0
            dx_norm = np.sqrt(t.dx_sq)
            if self.adaptive:
                tau_s = dx_norm ** 2 / g.dx_dg
                tau_m = max(g.dx_dg / np.sqrt(g.dg_sq) ** 2, 0)
                tau1 = tau_m if 2 * tau_m > tau_s else tau_s - .5 * tau_m
                if tau1 <= 0 or np.isinf(tau1) or np.isnan(tau1):
                    tau1 = tau0 * 1.5
            resid = dx_norm / tau0
            nres = resid / (max(np.sqrt(d["g0sq"]), np.sqrt(t.xmxh_sq) / tau0) + 1e-12)
            d["maxres"] = max(d["maxres"], resid)
            stop = {0: resid < tol, 1: nres < tol, 2: resid / d["maxres"] < tol,
                    3: resid / d["maxres"] < tol or nres < tol}.get(rule, False)
        d["f"].append(t.f)
        d["it"] += 1
        d["g0sq"] = g.g_sq
        d["skip"] = bool(stop)
        self.dev_tau = tau1
        t.tau_used = tau0
        return "sweep", (t, g)

    def _collect_trial(self, handle):
        t, g = handle[1]
        if t.skipped:
            self.mismatch += 1                 # the host wants a trial the device skipped
        self._spec = g
        return t

    def trial(self, tau, bt=0, host=(0, -np.inf, 0.0)):
        if not hasattr(self, "dev"):
            return CpuFusedBackend.trial(self, tau)
        return self._collect_trial(self._queue_trial(tau, bt, host))

    def gradient(self, tau, adaptive):
        if self._spec is None:
            return CpuFusedBackend.gradient(self, tau, adaptive)
        g, self._spec = self._spec, None
        return g

    def keep_best(self):
        self.best = (self.x0 if self._ahead else self.x1).clone()

    def iterate(self):
        return (self.x0 if self._ahead else self.x1).numpy().reshape(self.shape)


class HostPenalty:
    """params()/value() of fasta.proximal penalties without touching the GPU."""

    def __init__(self, kind, mu):
        self.kind, self.mu = kind, mu
        self.tag = {"l1": S.PROX_SHRINK, "l1ball": S.PROX_L1BALL, "nonneg": S.PROX_NONNEG,
                    "tv_ball": S.PROX_TV_BALL, "none": S.PROX_IDENTITY}[kind]
        self.radius = mu

    def params(self, t):
        return (t * self.mu, 0.0) if self.kind == "l1" else (0.0, 0.0)

    def value(self, raw):
        return self.mu * raw if self.kind == "l1" else 0


def backend_for(problem, accelerate, driver=None, speculate=False):
    """CPU test-double back-end for an oracle.problems.Problem."""
    tag = S.LOSS_LEAST_SQUARES if problem.loss == "least_squares" else S.LOSS_LOGISTIC
    if driver is None:
        driver = CpuDenseDriver(problem.A) if problem.kind == "dense" else CpuTVDriver(*problem.x0.shape[:2])
    cls = CpuSpeculatingBackend if speculate else CpuFusedBackend
    return cls(driver, tag, problem.b, HostPenalty(problem.penalty, problem.mu), problem.x0, accelerate)

"""GPU: every C-ABI kernel against its numpy expression (through the public toolkits / ctypes)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _t():
    import torch
    return torch


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


# M, N pairs: tile multiples, ragged edges, tiny, single row/col, tall, wide
SHAPES = [(200, 1000), (16, 256), (333, 1414), (1, 2), (5, 4098), (4099, 6), (1000, 2000), (2048, 4096),
          (17, 258), (640, 5120)]


@pytest.mark.parametrize("M,N", SHAPES)
def test_dense_map_tma_path(M, N):
    import fasta
    rng = np.random.RandomState(M * 7 + N)
    A = rng.randn(M, N)
    x, r = rng.randn(N), rng.randn(M)
    op = fasta.linalg.LinearMap.from_matrix(A)
    assert op.uses_tma
    assert op.Vshape == (N,) and op.Wshape == (M,)
    z = op(x)
    g = op.H(r)
    assert isinstance(z, np.ndarray) and z.shape == (M,) and g.shape == (N,)
    assert _rel(z, A @ x) < 1e-14
    assert _rel(g, A.T @ r) < 1e-14
    # bit-reproducible run to run (fixed-order, atomics-free reductions)
    assert np.array_equal(z, op(x)) and np.array_equal(g, op.H(r))
    with pytest.raises(AssertionError):
        op(np.zeros(N + 1))


@pytest.mark.parametrize("M,N", [(200, 1001), (33, 77), (1, 1), (513, 2049)])
def test_dense_map_plain_path_odd_leading_dimension(M, N):
    import fasta
    rng = np.random.RandomState(M + N)
    A = rng.randn(M, N)
    x, r = rng.randn(N), rng.randn(M)
    op = fasta.linalg.LinearMap.from_matrix(A)
    assert not op.uses_tma            # odd lda: TMA needs 16-byte row pitch
    assert _rel(op(x), A @ x) < 1e-14
    assert _rel(op.H(r), A.T @ r) < 1e-14


def test_dense_map_borrows_cuda_tensor_and_row_slices():
    import fasta
    torch = _t()
    rng = np.random.RandomState(3)
    A = rng.randn(300, 512)
    Ad = torch.from_numpy(A).cuda()
    op = fasta.linalg.LinearMap.from_matrix(Ad)
    assert op.matrix.data_ptr() == Ad.data_ptr()
    x = torch.from_numpy(rng.randn(512)).cuda()
    z = op(x)
    assert isinstance(z, torch.Tensor) and z.is_cuda
    assert _rel(z.cpu().numpy(), A @ x.cpu().numpy()) < 1e-14
    sl = fasta.linalg.LinearMap.from_matrix(Ad[100:260])          # view: rows 100..259
    assert sl.matrix.data_ptr() == Ad[100:260].data_ptr()
    assert _rel(sl(x).cpu().numpy(), A[100:260] @ x.cpu().numpy()) < 1e-14
    r = rng.randn(160)
    assert _rel(sl.H(r), A[100:260].T @ r) < 1e-14


def test_adjoint_identity_large():
    """<A x, r> == <x, A^T r> at a size where both kernels run many tiles per CTA."""
    import fasta
    torch = _t()
    g = torch.Generator(device="cuda").manual_seed(1)
    M, N = 6000, 20000
    A = torch.randn(M, N, dtype=torch.float64, device="cuda", generator=g)
    x = torch.randn(N, dtype=torch.float64, device="cuda", generator=g)
    r = torch.randn(M, dtype=torch.float64, device="cuda", generator=g)
    op = fasta.linalg.LinearMap.from_matrix(A)
    z, gr = op(x), op.H(r)
    assert _rel(z.cpu().numpy(), (A @ x).cpu().numpy()) < 1e-13
    assert _rel(gr.cpu().numpy(), (A.T @ r).cpu().numpy()) < 1e-13
    lhs, rhs = float(torch.dot(z, r)), float(torch.dot(x, gr))
    assert abs(lhs - rhs) <= 1e-12 * abs(lhs)


def test_prox_known_answers(golden_dir):
    import fasta
    with np.load(f"{golden_dir}/kat_prox.npz") as z:
        for i in range(int(z["count"])):
            x, t = z[f"x{i}"], float(z[f"t{i}"])
            got = fasta.proximal.shrink(x, t)
            assert np.array_equal(got, z[f"shrink{i}"])                       # bit exact, incl. -0.0
            assert np.array_equal(np.signbit(got), np.signbit(z[f"shrink{i}"]))
            want = z[f"l1ball{i}"]
            got = fasta.proximal.project_L1_ball(x, t)
            assert np.max(np.abs(got - want)) <= 4e-16 * max(1.0, np.max(np.abs(x)))
            assert np.max(np.abs(fasta.proximal.project_Linf_ball(x, t) - z[f"tinf{i}"])) <= 4e-16 * max(1.0, np.max(np.abs(x)))
        assert _rel(fasta.proximal.project_Lnuc_ball(z["X"], float(z["Xt"])), z["nuc"]) < 1e-12
    assert np.array_equal(fasta.proximal.shrink(np.ones((3, 5)) * -2, 0.5), -1.5 * np.ones((3, 5)))  # any shape
    assert np.array_equal(fasta.proximal.shrink(np.array([1.0, -1.0]), -1.0), np.array([2.0, -2.0]))  # negative t expands
    x = np.array([0.1, -0.2, 0.05])
    assert np.array_equal(fasta.proximal.project_L1_ball(x, 1.0), x)           # inside the ball: unchanged


def test_l1_ball_projection_large_and_penalties():
    import fasta
    from oracle import fasta_oracle
    rng = np.random.RandomState(11)
    x = rng.randn(100003) * 3
    for radius in (1.0, 50.0, 5000.0, 1e9):
        got = fasta.proximal.project_L1_ball(x, radius)
        want = fasta_oracle.project_l1_ball(x, radius)
        assert np.max(np.abs(got - want)) < 1e-13
        assert np.abs(got).sum() <= radius * (1 + 1e-12) or radius >= np.abs(x).sum()
    pen = fasta.proximal.L1Norm(0.3)
    assert abs(pen.g(x) - 0.3 * np.abs(x).sum()) < 1e-9
    assert np.array_equal(pen.prox(x, 0.5), fasta_oracle.shrink(x, 0.5 * 0.3))
    assert np.array_equal(fasta.proximal.NonNegative().prox(x, 1.0), np.maximum(x, 0))
    assert np.array_equal(fasta.proximal.Box(-1.0, 0.5).prox(x, 1.0), np.clip(x, -1.0, 0.5))
    Y = rng.randn(37, 41, 2) * 2
    nrm = np.maximum(np.linalg.norm(Y, axis=2), 1)
    assert np.array_equal(fasta.proximal.TVBall().prox(Y, 0.1), Y / nrm[..., None])


@pytest.mark.parametrize("M,N", [(40, 60), (60, 40), (7, 7), (1, 5), (5, 1), (33, 90), (120, 300), (300, 120)])
def test_nuclear_prox_jacobi_svd(M, N):
    """Singular-value soft threshold by the one-sided Jacobi kernel vs numpy's SVD (reference proximal.py:44-55);
    the two larger shapes take the global-scratch path, (M > N) the transposed one; one case is rank deficient."""
    import fasta
    import torch
    rng = np.random.default_rng(M * 1000 + N)
    X = rng.standard_normal((M, N))
    if M > 5:
        X[3] = 2 * X[2]
    for t in (0.0, 1.5, 1e3):
        U, s, V = np.linalg.svd(X, full_matrices=False)
        want = U @ np.diag(np.maximum(s - t, 0)) @ V
        got = fasta.proximal.project_Lnuc_ball(X, t)
        assert isinstance(got, np.ndarray) and got.shape == X.shape
        assert np.abs(got - want).max() <= 1e-12 * max(s[0], 1.0)
    sv = fasta.proximal.singular_values(X)
    assert np.abs(sv - s).max() <= 1e-12 * s[0]
    Xd = torch.from_numpy(X).cuda()
    a, b = fasta.proximal.project_Lnuc_ball(Xd, 1.5), fasta.proximal.project_Lnuc_ball(Xd, 1.5)
    assert a.is_cuda and torch.equal(a, b)                 # bit-reproducible run to run
    # non-contiguous input (a transposed view) is handled
    assert np.abs(fasta.proximal.project_Lnuc_ball(Xd.t(), 1.5).cpu().numpy() - a.cpu().numpy().T).max() <= 1e-12 * s[0]


@pytest.mark.parametrize("M,N", [(300, 800), (301, 803), (64, 32)])
def test_gram_norm_power_iteration(M, N):
    """|A|_2^2 by power iteration on the device (one single-pass sweep per iteration when the matrix is eligible, the two
    streaming contractions otherwise) against numpy's SVD; reproducible bit for bit; usable as the exact Lipschitz
    constant in place of the reference's randomized estimate (fasta/__init__.py:100-113) without touching the RNG."""
    import fasta
    rng = np.random.RandomState(M + N)
    u, v = rng.randn(M), rng.randn(N)
    A = rng.randn(M, N) + 3.0 * np.outer(u, v) / np.sqrt(M * N) * np.sqrt(M + N)      # a clear spectral gap
    op = fasta.linalg.LinearMap.from_matrix(A)
    want = np.linalg.norm(A, 2) ** 2
    got = op.gram_norm()
    assert abs(got - want) <= 1e-9 * want
    assert op.gram_norm() == got
    # a plain Gaussian matrix (tiny gap): the Rayleigh quotient is a lower bound that converges from below
    B = rng.randn(M, N)
    lam = fasta.linalg.LinearMap.from_matrix(B).gram_norm(iters=3000)
    assert 0.999 * np.linalg.norm(B, 2) ** 2 <= lam <= (1 + 1e-12) * np.linalg.norm(B, 2) ** 2
    if (M, N) == (300, 800):
        b = A @ (rng.rand(N) < 0.02) + 0.01 * rng.randn(M)
        loss, pen = fasta.losses.LeastSquares(b), fasta.proximal.L1Norm(0.5)
        state = np.random.get_state()[1].copy()
        res = fasta.fasta(op, loss.f, loss.gradf, pen.g, pen.prox, np.zeros(N), L=got, tau0=(2 / got) / 10, verbose=False,
                          evaluate_objective=True)
        assert np.array_equal(np.random.get_state()[1], state)                 # no draws from the global RNG
        ref = fasta.fasta(op, loss.f, loss.gradf, pen.g, pen.prox, np.zeros(N), verbose=False, evaluate_objective=True)
        fa, fb = res.objectives[res.iteration_count], ref.objectives[ref.iteration_count]
        assert abs(fa - fb) <= 1e-6 * abs(fb)


def test_losses():
    import fasta
    rng = np.random.RandomState(5)
    z, b = rng.randn(7777) * 3, rng.randn(7777)
    ls = fasta.losses.LeastSquares(b)
    assert abs(ls.f(z) - .5 * np.linalg.norm(z - b) ** 2) <= 1e-12 * ls.f(z)
    assert np.array_equal(ls.gradf(z), z - b)
    lab = np.sign(rng.randn(7777))
    lg = fasta.losses.Logistic(lab)
    want = np.sum(np.log(1 + np.exp(z)) - (lab == 1) * z)
    assert abs(lg.f(z) - want) <= 1e-12 * abs(want)
    assert np.max(np.abs(lg.gradf(z) - (-lab / (1 + np.exp(lab * z))))) <= 1e-15


@pytest.mark.parametrize("n0,n1", [(64, 64), (128, 96), (33, 130), (1, 7), (5, 1), (300, 257)])
def test_tv_stencils(n0, n1):
    import fasta
    from oracle import problems
    rng = np.random.RandomState(n0 + n1)
    X, Y = rng.randn(n0, n1), rng.randn(n0, n1, 2)
    assert np.array_equal(fasta.tv.grad(X), problems.tv_grad(X))        # elementwise: bit exact
    assert np.array_equal(fasta.tv.div(Y), problems.tv_div(Y))
    lhs = np.sum(fasta.tv.div(Y) * X)
    rhs = np.sum(Y * fasta.tv.grad(X))
    assert abs(lhs - rhs) <= 1e-12 * max(abs(lhs), 1.0)                  # adjointness


@pytest.mark.parametrize("tma", ["0", "force"])
@pytest.mark.parametrize("n0,n1", [(64, 64), (128, 96), (33, 130), (1, 7), (5, 1), (2, 2), (300, 257), (70, 64),
                                   (130, 360), (5, 362), (200, 724), (97, 1084), (3, 4)])
def test_tv_whole_iteration_kernel(n0, n1, tma, monkeypatch):
    """fb200_tv_iter_fused: x1 and g1 bit-identical to the numpy expressions of the reference lines
    (tv_denoising.py:26-63,85-96; __init__.py:181-188,248-260), the seven sums to reduction rounding -- with the
    register-marching kernel and with the bulk-copy-fed one (one, two, three column tiles, image edges inside a tile,
    strips shorter than the ring; odd widths fall back to the marching kernel by design)."""
    from fasta import _cabi, _device
    monkeypatch.setenv("FASTA_B200_TV_TMA", tma)
    from oracle import problems
    torch = _t()
    lib = _cabi.load()
    rng = np.random.RandomState(3 * n0 + n1)
    x0, g0, b = rng.randn(n0, n1, 2), rng.randn(n0, n1, 2), rng.randn(n0, n1)
    tau = 0.3
    d = {k: torch.from_numpy(v).cuda() for k, v in dict(x0=x0, g0=g0, b=b).items()}
    x1, g1 = (torch.full((n0, n1, 2), np.nan, dtype=torch.float64, device="cuda") for _ in range(2))
    ws = _device.Workspace(1, 1)
    _cabi.check(lib.fb200_tv_iter_fused(d["x0"].data_ptr(), d["g0"].data_ptr(), tau, n0, n1, _cabi.LOSS_LEAST_SQUARES,
                                        d["b"].data_ptr(), x1.data_ptr(), g1.data_ptr(), ws.scal.data_ptr(),
                                        ws.buf.data_ptr(), _device.stream_ptr()))
    s = ws.fetch().copy()
    h = x0 - tau * g0
    nrm = np.maximum(np.sqrt(h[..., 0] * h[..., 0] + h[..., 1] * h[..., 1]), 1.0)
    y = h / nrm[..., None]
    r = problems.tv_div(y) - b
    g = problems.tv_grad(r)
    assert np.array_equal(x1.cpu().numpy(), y)
    assert np.array_equal(g1.cpu().numpy(), g)
    dx = y - x0
    dg = g + (h - x0) / tau
    dot = lambda u, v: float(np.sum(u * v))
    for slot, want in ((_cabi.S_DX_G0, dot(dx, g0)), (_cabi.S_DX_SQ, dot(dx, dx)), (_cabi.S_XMXH_SQ, dot(y - h, y - h)),
                       (_cabi.S_F, dot(r, r)), (_cabi.S_DX_DG, dot(dx, dg)), (_cabi.S_DG_SQ, dot(dg, dg)),
                       (_cabi.S_G1_SQ, dot(g, g))):
        assert abs(s[slot] - want) <= 1e-12 * max(abs(want), 1e-3), (n0, n1, slot, s[slot], want)


@pytest.mark.parametrize("c", [0.0, 0.4375, 0.83])
@pytest.mark.parametrize("n0,n1", [(64, 64), (33, 130), (1, 7), (5, 1), (2, 2), (300, 257), (70, 64)])
def test_tv_fista_iteration_kernel(n0, n1, c):
    """fb200_tv_fista_fused: prox point, its image, the extrapolated x1 and the gradient at the extrapolated z are
    bit-identical to the numpy expressions of the reference lines (__init__.py:181-188,242-248), the nine sums to
    reduction rounding."""
    from fasta import _cabi, _device
    from oracle import problems
    torch = _t()
    lib = _cabi.load()
    rng = np.random.RandomState(5 * n0 + n1)
    x0, g0, xa0 = (rng.randn(n0, n1, 2) for _ in range(3))
    b, za0 = rng.randn(n0, n1), rng.randn(n0, n1)
    tau = 0.3
    d = {k: torch.from_numpy(v).cuda() for k, v in dict(x0=x0, g0=g0, b=b, xa0=xa0, za0=za0).items()}
    xa1, x1, g1 = (torch.full((n0, n1, 2), np.nan, dtype=torch.float64, device="cuda") for _ in range(3))
    za1 = torch.full((n0, n1), np.nan, dtype=torch.float64, device="cuda")
    ws = _device.Workspace(1, 1)
    _cabi.check(lib.fb200_tv_fista_fused(d["x0"].data_ptr(), d["g0"].data_ptr(), tau, c, n0, n1, _cabi.LOSS_LEAST_SQUARES,
                                         d["b"].data_ptr(), d["xa0"].data_ptr(), d["za0"].data_ptr(), xa1.data_ptr(),
                                         za1.data_ptr(), x1.data_ptr(), g1.data_ptr(), ws.scal.data_ptr(),
                                         ws.buf.data_ptr(), _device.stream_ptr()))
    s = ws.fetch().copy()
    h = x0 - tau * g0
    nrm = np.maximum(np.sqrt(h[..., 0] * h[..., 0] + h[..., 1] * h[..., 1]), 1.0)
    y = h / nrm[..., None]
    zp = problems.tv_div(y)
    xe = y + c * (y - xa0)
    ze = zp + c * (zp - za0)
    r = ze - b
    g = problems.tv_grad(r)
    assert np.array_equal(xa1.cpu().numpy(), y)
    assert np.array_equal(za1.cpu().numpy(), zp)
    assert np.array_equal(x1.cpu().numpy(), xe)
    assert np.array_equal(g1.cpu().numpy(), g)
    dx = y - x0
    dg = g + (h - x0) / tau
    dot = lambda u, v: float(np.sum(u * v))
    for slot, want in ((_cabi.S_DX_G0, dot(dx, g0)), (_cabi.S_DX_SQ, dot(dx, dx)), (_cabi.S_XMXH_SQ, dot(xe - h, xe - h)),
                       (_cabi.S_F, dot(zp - b, zp - b)), (_cabi.S_AUX3, dot(r, r)), (_cabi.S_DX_DG, dot(dx, dg)),
                       (_cabi.S_DG_SQ, dot(dg, dg)), (_cabi.S_G1_SQ, dot(g, g)), (_cabi.S_RESTART, dot(x0 - y, y - xa0))):
        assert abs(s[slot] - want) <= 1e-12 * max(abs(want), 1e-3), (n0, n1, slot, s[slot], want)


def test_step_kernels_through_ctypes():
    """fbs_step / bb_reduce / accel_step scalars against numpy on ragged sizes."""
    import fasta
    from fasta import _cabi, _device
    torch = _t()
    lib = _cabi.load()
    rng = np.random.RandomState(9)
    for n in (1, 31, 1000, 100003):
        x0, g0, xa = rng.randn(n), rng.randn(n), rng.randn(n)
        tau, thr = 0.37, 0.21
        d = {k: torch.from_numpy(v).cuda() for k, v in dict(x0=x0, g0=g0, xa=xa).items()}
        xhat, x1, dx = (torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(3))
        ws = _device.Workspace(1, 1)
        _cabi.check(lib.fb200_fbs_step(d["x0"].data_ptr(), d["g0"].data_ptr(), tau, _cabi.PROX_SHRINK, thr, 0.0,
                                       d["xa"].data_ptr(), n, xhat.data_ptr(), x1.data_ptr(), dx.data_ptr(),
                                       ws.scal.data_ptr(), ws.buf.data_ptr(), _device.stream_ptr()))
        s = ws.fetch().copy()
        h = x0 - tau * g0
        y = np.sign(h) * np.maximum(np.abs(h) - thr, 0)
        assert np.array_equal(xhat.cpu().numpy(), h) and np.array_equal(x1.cpu().numpy(), y)
        assert np.array_equal(dx.cpu().numpy(), y - x0)
        for slot, want in ((_cabi.S_DX_G0, (y - x0) @ g0), (_cabi.S_DX_SQ, (y - x0) @ (y - x0)),
                           (_cabi.S_XMXH_SQ, (y - h) @ (y - h)), (_cabi.S_PEN, np.abs(y).sum()),
                           (_cabi.S_RESTART, (x0 - y) @ (y - xa))):
            assert abs(s[slot] - want) <= 1e-12 * max(abs(want), 1e-3), (n, slot)
        g1 = torch.from_numpy(rng.randn(n)).cuda()
        _cabi.check(lib.fb200_bb_reduce(g1.data_ptr(), d["x0"].data_ptr(), xhat.data_ptr(), dx.data_ptr(), tau, n, 1,
                                        ws.scal.data_ptr(), ws.buf.data_ptr(), _device.stream_ptr()))
        s = ws.fetch().copy()
        g1n = g1.cpu().numpy()
        dg = g1n + (h - x0) / tau
        for slot, want in ((_cabi.S_DX_DG, (y - x0) @ dg), (_cabi.S_DG_SQ, dg @ dg), (_cabi.S_G1_SQ, g1n @ g1n)):
            assert abs(s[slot] - want) <= 1e-12 * max(abs(want), 1e-3), (n, slot)


SWEEP_SHAPES = [(200, 1000), (37, 6656), (333, 1414), (1000, 20000), (64, 26624), (500, 53248), (257, 100000),
                (3, 2), (1, 106496), (40, 13314), (4001, 100000), (150, 212992)]


@pytest.mark.parametrize("loss", ["least_squares", "logistic", "none"])
@pytest.mark.parametrize("M,N", SWEEP_SHAPES)
def test_single_pass_sweep(M, N, loss):
    """fb200_dense_sweep: z, r, f and g = A^T r from ONE pass over A, for every cluster size."""
    import fasta
    from fasta import _backends, _cabi, _device
    torch = _t()
    rng = np.random.RandomState(M + N)
    A = rng.randn(M, N) / np.sqrt(N)
    x = rng.randn(N)
    b = np.sign(rng.randn(M)) if loss == "logistic" else rng.randn(M)
    tag = {"least_squares": _cabi.LOSS_LEAST_SQUARES, "logistic": _cabi.LOSS_LOGISTIC, "none": _cabi.LOSS_NONE}[loss]
    Ad, xd, bd = (torch.from_numpy(v).cuda() for v in (A, x, b))
    drv = _backends.DenseDriver(Ad)
    assert 1 <= drv.sweep_cluster <= 32          # column slabs per row band (cluster size with the cluster kernel)
    ws = _device.Workspace(M, N)
    z, r = torch.zeros(M, dtype=torch.float64, device="cuda"), torch.zeros(M, dtype=torch.float64, device="cuda")
    g = torch.zeros(N, dtype=torch.float64, device="cuda")
    drv.sweep(xd, tag, bd, z, r, g, 1, None, None, None, 0.0, ws)
    s = ws.fetch().copy()
    zr = A @ x
    if loss == "least_squares":
        rr, fr = zr - b, np.sum((zr - b) ** 2)
    elif loss == "logistic":
        rr, fr = -b / (1 + np.exp(b * zr)), np.sum(np.log(1 + np.exp(zr)) - (b == 1) * zr)
    else:
        rr, fr = zr, None
    gr = A.T @ rr
    assert _rel(z.cpu().numpy(), zr) < 1e-14
    if loss != "none":
        assert _rel(r.cpu().numpy(), rr) < 1e-13
        assert abs(s[_cabi.S_F] - fr) <= 1e-13 * abs(fr)
    assert _rel(g.cpu().numpy(), gr) < 1e-13
    assert abs(s[_cabi.S_G1_SQ] - gr @ gr) <= 1e-12 * (gr @ gr)
    # bit-reproducible
    z2, g2 = torch.zeros_like(z), torch.zeros_like(g)
    drv.sweep(xd, tag, bd, z2, r, g2, 1, None, None, None, 0.0, ws)
    torch.cuda.synchronize()
    assert torch.equal(z, z2) and torch.equal(g, g2)


def test_sweep_not_eligible_for_odd_shapes():
    import fasta
    from fasta import _backends
    torch = _t()
    assert _backends.DenseDriver(torch.zeros(10, 1001, dtype=torch.float64, device="cuda")).sweep_cluster == 0
    assert _backends.DenseDriver(torch.zeros(4, 212994, dtype=torch.float64, device="cuda")).sweep_cluster == 0


@pytest.mark.parametrize("shape", [(7,), (1,), (5, 6, 4), (3, 4, 5, 6), (2, 1, 3), (17, 9, 11)])
def test_tv_operators_of_any_rank(shape):
    """fasta.tv.grad / div / TVBall.prox on N-d arrays (the reference's functions are rank-generic, tv_denoising.py:26-63,
    89-96): bit-identical to the numpy loops, grad adjoint to div."""
    import fasta
    from oracle import problems
    rng = np.random.RandomState(sum(shape))
    X = rng.randn(*shape)
    Y = rng.randn(*(shape + (len(shape),))) * 1.5
    G = fasta.tv.grad(X)
    D = fasta.tv.div(Y)
    assert G.shape == shape + (len(shape),) and D.shape == shape
    assert np.array_equal(G, problems.tv_grad(X)) and np.array_equal(D, problems.tv_div(Y))
    lhs, rhs = float(np.sum(D * X)), float(np.sum(Y * G))
    assert abs(lhs - rhs) <= 1e-12 * max(abs(lhs), 1.0)
    P = fasta.proximal.TVBall().prox(Y, 0.3)
    assert np.array_equal(P, problems._tv_ball(Y, 0.3))
    A = fasta.tv.divergence_map(shape)
    assert A.Vshape == shape + (len(shape),) and A.Wshape == shape and np.array_equal(A(Y), D) and np.array_equal(A.H(X), G)


def test_tv_denoising_of_a_volume_matches_the_oracle():
    """The reference's TV-denoising call (legacy form with the div / grad callables, tv_denoising.py:99) on a 3-D volume:
    the generic back-end drives the N-d kernels; trajectory against the oracle on the same seed."""
    import fasta
    from oracle import fasta_oracle, problems
    rng = np.random.RandomState(4)
    n = (12, 10, 8)
    vol = np.zeros(n)
    vol[3:9, 2:7, 1:6] = 1.0
    vol += 0.1 * rng.randn(*n)
    mu = 0.1
    b = vol / mu
    Y0 = np.zeros(n + (3,))
    opts = dict(adaptive=False, accelerate=True, verbose=False, tolerance=1e-5, max_iters=60, evaluate_objective=True)
    np.random.seed(2)
    ref = fasta_oracle.solve(problems.tv_div, problems.tv_grad, lambda Z: .5 * np.linalg.norm((Z - b).ravel()) ** 2, lambda Z: Z - b,
                             lambda Y: 0, problems._tv_ball, Y0, **opts)
    loss, pen = fasta.losses.LeastSquares(b), fasta.proximal.TVBall()
    np.random.seed(2)
    res = fasta.fasta(fasta.tv.div, fasta.tv.grad, loss.f, loss.gradf, pen.g, pen.prox, Y0, **opts)
    assert res.backend == "GenericBackend"
    assert (res.iteration_count, res.backtracks) == (ref.iteration_count, ref.backtracks)
    n_it = ref.iteration_count
    assert np.linalg.norm(res.solution - ref.solution) <= 1e-9 * np.linalg.norm(ref.solution)
    assert np.max(np.abs(res.objectives[:n_it + 1] - ref.objectives[:n_it + 1]) / np.abs(ref.objectives[:n_it + 1])) <= 1e-10

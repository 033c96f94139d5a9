"""GPU: np.random.randn continued on the device (csrc/legacy_rng.cu through fasta/_rng.py) -- values and end
state bit for bit against numpy itself, and the solver prologue with either source of the probes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def _device_draws(entry, sizes):
    import torch
    from fasta import _rng
    dev = torch.device("cuda", torch.cuda.current_device())
    g = _rng.DeviceRandn(dev)
    outs = [torch.empty(n, dtype=torch.float64, device=dev) for n in sizes]
    g.begin(entry)
    for o in outs:
        g.draw(o)
    g.finish_async()
    end = g.finish(set_global=False)
    return [o.cpu().numpy() for o in outs], end


@pytest.mark.parametrize("seed,warm,sizes", [
    (0, 0, [1]), (0, 0, [2]), (0, 1, [1]), (0, 1, [2, 3]), (1, 0, [311, 312]), (2, 3, [1000, 1000]),
    (0, 0, [100000, 100000]),                  # config 2: the two probes of a 100000-vector
    (9, 5, [155, 1]), (4, 0, [156, 156, 157]), # end positions at / next to an MT19937 block boundary (624 = 4 * 156 words)
    (3, 2, [1 << 20, 12345]),
])
def test_device_randn_is_numpy_randn(seed, warm, sizes):
    rs = np.random.RandomState(seed)
    if warm:
        rs.randn(warm)
    entry = rs.get_state()
    refs = [rs.randn(n) for n in sizes]
    end = rs.get_state()
    got, got_end = _device_draws(entry, sizes)
    assert got_end is not None
    for a, b in zip(got, refs):
        assert np.array_equal(_bits(a), _bits(b))
    assert np.array_equal(got_end[1], end[1]) and tuple(got_end[2:]) == tuple(end[2:])


def test_device_randn_tv_sized_draw():
    """2 x 4096 x 4096 x 2 values (config 4): spot-checked against numpy on a prefix / suffix and by the end state."""
    n = 4096 * 4096 * 2
    rs = np.random.RandomState(0)
    entry = rs.get_state()
    ref = rs.randn(n)
    end = rs.get_state()
    got, got_end = _device_draws(entry, [n])
    assert np.array_equal(_bits(got[0]), _bits(ref))
    assert np.array_equal(got_end[1], end[1]) and tuple(got_end[2:]) == tuple(end[2:])


def test_global_stream_stays_in_step_with_the_reference():
    """After a solve numpy's global generator is where the reference's two host draws would have left it."""
    import fasta
    from oracle import problems
    p = problems.build("lasso_333x1414_k20", 0) if "lasso_333x1414_k20" in problems.CASES else problems.build("lasso_200x1000_k10", 0)
    A = fasta.linalg.LinearMap.from_matrix(p.A)
    loss, pen = fasta.losses.LeastSquares(p.b), fasta.proximal.L1Norm(p.mu)
    opts = dict(verbose=False, max_iters=5)
    ends, sols = [], []
    import os
    for mode in ("force", "0"):
        os.environ["FASTA_B200_DEVICE_RNG"] = mode
        try:
            np.random.seed(42)
            res = fasta.fasta(A, loss.f, loss.gradf, pen.g, pen.prox, p.x0, **opts)
        finally:
            os.environ.pop("FASTA_B200_DEVICE_RNG", None)
        ends.append(np.random.get_state())
        sols.append((res.stepsizes.copy(), np.asarray(res.solution).copy()))
    np.random.seed(42)
    np.random.randn(*p.x0.shape)
    np.random.randn(*p.x0.shape)
    want = np.random.get_state()
    for e in ends:
        assert np.array_equal(e[1], want[1]) and tuple(e[2:]) == tuple(want[2:])
    # identical probes -> identical trajectories whichever side drew them
    assert np.array_equal(sols[0][0], sols[1][0]) and np.array_equal(sols[0][1], sols[1][1])

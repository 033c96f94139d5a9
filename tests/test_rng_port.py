"""CPU: the restatement of numpy's legacy Gaussian stream (oracle/np_legacy_rng.c, which compiles the very
glibc_log.h the CUDA kernels use) against np.random.randn and against libm's log, bit for bit."""
import ctypes

import numpy as np
import pytest

from oracle import build_oracle


@pytest.fixture(scope="module")
def lib():
    return build_oracle.load()


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def test_restated_log_is_libm_log_bit_for_bit(lib):
    rng = np.random.default_rng(11)
    a, b = 2 * rng.random(4_000_000) - 1, 2 * rng.random(4_000_000) - 1
    r2 = a * a + b * b
    cases = {
        "polar r2": r2[(r2 < 1) & (r2 > 0)],
        "uniform": rng.random(2_000_000),
        "near one": 1 - rng.random(2_000_000) * 0.07,        # the |x - 1| < 2^-4 branch
        "tiny": rng.random(500_000) * 2.0 ** -90,
        "wide": np.exp(rng.uniform(-72, 0, 1_000_000)),
        "edges": np.array([1.0, 1 - 2.0 ** -53, 0.9375, np.nextafter(0.9375, 0), 2.0 ** -104, 0.5, 0.6875,
                           np.nextafter(0.6875, 0)]),
    }
    for name, x in cases.items():
        x = np.ascontiguousarray(x)
        worst = ctypes.c_double(0.0)
        bad = lib.fb200_ref_log_mismatches(x.ctypes.data, x.size, ctypes.byref(worst))
        assert bad == 0, f"{name}: {bad} of {x.size} differ from libm log, first at {worst.value!r}"


@pytest.mark.parametrize("seed,warm,n", [(0, 0, 1), (0, 0, 10), (0, 0, 100001), (123, 7, 250000), (7, 1, 33333), (5, 623, 4)])
def test_restated_generator_is_np_random_randn(seed, warm, n):
    np.random.seed(seed)
    if warm:
        np.random.randn(warm)               # odd counts leave a cached deviate, any count a mid-block position
    entry = np.random.get_state()
    ref = np.random.randn(n)
    ref_next = np.random.randn(5)
    end = np.random.get_state()
    got, st = build_oracle.randn(entry, n)
    got_next, st2 = build_oracle.randn(st, 5)
    assert np.array_equal(_bits(got), _bits(ref)) and np.array_equal(_bits(got_next), _bits(ref_next))
    assert np.array_equal(st2[1], end[1]) and tuple(st2[2:]) == tuple(end[2:])


def test_state_packing_round_trip():
    import fasta._rng as _rng
    np.random.seed(3)
    np.random.randn(3)
    st = np.random.get_state()
    words = np.zeros(_rng.OUT_WORDS, dtype=np.uint32)
    _rng.pack_state(st, words)
    back = _rng.unpack_state(words)
    assert np.array_equal(back[1], st[1]) and back[2:] == tuple(st[2:])


def test_jump_polynomial_table_matches_recomputation_and_numpy():
    """fasta/mt19937_jump.npz (tools/make_mt_jump.py): the characteristic polynomial recovered here again by
    Berlekamp-Massey, two table entries recomputed from it, and the jump y[n+J] = XOR_{g_j} y[n+j] carried out in numpy
    against numpy's own MT19937 advanced J words -- the table the device's parallel stream relies on is right
    independently of any GPU run."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("make_mt_jump", os.path.join(root, "tools", "make_mt_jump.py"))
    mj = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mj)
    with np.load(os.path.join(root, "fasta-python_b200", "fasta", "mt19937_jump.npz")) as z:
        polys = z["polys"]
        assert int(z["unit_blocks"]) == 64 and polys.shape == (4, 16, 624)
    y = mj.raw_words(4357, 2 * mj.DEG + 700)
    C, L = mj.berlekamp_massey((y[:2 * mj.DEG + 64] & 1).tolist())
    assert L == mj.DEG
    phi = sum(((C >> i) & 1) << (mj.DEG - i) for i in range(mj.DEG + 1))
    assert bin(phi).count("1") == 135                                   # the known weight of MT19937's characteristic polynomial
    for level, digit in ((0, 1), (1, 3)):
        J = 624 * 64 * 16 ** level * digit
        assert np.array_equal(mj.poly_words(mj.x_pow_mod(J, phi)), polys[level, digit])
    # the jump itself, on the stream of another seed: 7 x 64 blocks ahead by one convolution
    pre = mj.raw_words(99, mj.DEG + 624)
    J = 624 * 64 * 7
    ref = mj.raw_words(99, J + 624)
    assert np.array_equal(mj.jump_numpy(pre, polys[0, 7]), ref[J:J + 624])

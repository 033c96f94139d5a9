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

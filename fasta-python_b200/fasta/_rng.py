"""numpy's global legacy Gaussian stream, continued on the device (csrc/legacy_rng.cu).

The reference draws the two probes of its Lipschitz estimate with ``np.random.randn(*x0.shape)`` twice
(``fasta/__init__.py:102-103``) -- 2 N draws from the GLOBAL ``RandomState``.  ``DeviceRandn`` takes that state
(``np.random.get_state()``), lets the GPU produce the very same values (``fb200_randn_legacy``: MT19937 words, 53-bit
doubles, polar method, glibc's ``log``, all bit for bit), and puts the state numpy would have ended in back with
``np.random.set_state()``: a caller cannot tell the difference from the host draws, except that no 2 N-vector is
drawn on one host core and pushed through PCIe per solve.

The first use per device checks the kernels against ``numpy.random.RandomState`` on a private stream (the host's libm
is the one thing this depends on that the image could change); on any mismatch the device path switches itself off and
the probes are drawn on the host as before -- that is the reference's own behaviour, not a compute fallback.
``FASTA_B200_DEVICE_RNG=0`` forces the host draws, ``=force`` uses the device for every size.
"""

import os
import warnings

import numpy as np

from . import _cabi, _device

STATE_WORDS = 628        # key[624], pos, has_gauss, cached gauss (double at word 626)
OUT_WORDS = 632          # + status word [628], tries used [630..632)
MIN_DEVICE_DRAWS = 4096  # below this the host's two draws cost less than ten kernel launches

_checked = {}            # device index -> bool
_polys = {}              # device index -> device tensor of the MT19937 jump polynomials (or None)


def jump_polys(device):
    """fasta/mt19937_jump.npz ([4][16][624] words: t^J mod phi for J = 64 * 16^level * digit blocks) on `device`."""
    key = device.index if device.index is not None else _device.torch().cuda.current_device()
    if key not in _polys:
        t = _device.torch()
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mt19937_jump.npz")
        try:
            with np.load(path) as z:
                table = np.ascontiguousarray(z["polys"], dtype=np.uint32)
            assert table.shape == (4, 16, 624)
            _polys[key] = t.from_numpy(table.view(np.int32)).to(device)
        except Exception as exc:               # without the table large draws come from one thread block: slower, same values
            warnings.warn(f"fasta-b200: {path} not usable ({exc}); large device draws stay sequential")
            _polys[key] = None
    return _polys[key]


def mode():
    return os.environ.get("FASTA_B200_DEVICE_RNG", "1")


def pack_state(state, out_words):
    """('MT19937', key, pos, has_gauss, cached) -> 628 uint32 words in ``out_words`` (a numpy uint32 view)."""
    name, key, pos, has_gauss, cached = state
    if name != "MT19937":
        raise ValueError(f"numpy legacy state of kind {name!r}")
    out_words[:624] = key
    out_words[624] = int(pos)
    out_words[625] = int(has_gauss)
    out_words[626:628].view(np.float64)[0] = float(cached)


def unpack_state(words):
    key = np.array(words[:624], dtype=np.uint32)
    return ("MT19937", key, int(words[624]), int(words[625]), float(words[626:628].view(np.float64)[0]))


class DeviceRandn:
    """One chained sequence of draws: begin() -> draw(out) [-> draw(out) ...] -> finish()."""

    def __init__(self, device):
        t = _device.torch()
        self.t = t
        self.lib = _cabi.load()
        self.device = device
        self._host_in = t.zeros(OUT_WORDS, dtype=t.int32).pin_memory()
        self._host_in_np = self._host_in.numpy().view(np.uint32)
        self._states = []        # device state buffers of this sequence: [entry, after draw 1, after draw 2, ...]
        self._host_out = None
        self._event = t.cuda.Event()

    def begin(self, state=None, sync_fn=None):
        """Upload the entry state (default: numpy's global one).  ``sync_fn(tensor)`` lets a sharded driver make
        every rank start from rank 0's state."""
        t = self.t
        pack_state(np.random.get_state() if state is None else state, self._host_in_np)
        s0 = t.empty(OUT_WORDS, dtype=t.int32, device=self.device)
        s0.copy_(self._host_in, non_blocking=True)
        if sync_fn is not None:
            sync_fn(s0)
        self._states = [s0]
        self._scratch = []

    def draw(self, out):
        """Queue ``out.numel()`` standard normals into the contiguous fp64 device tensor ``out``."""
        t = self.t
        n = int(out.numel())
        nbytes = int(self.lib.fb200_randn_scratch_bytes(n))
        scratch = t.empty(nbytes, dtype=t.uint8, device=self.device)
        nxt = t.empty(OUT_WORDS, dtype=t.int32, device=self.device)
        polys = jump_polys(self.device) if os.environ.get("FASTA_B200_MT_JUMP", "1") != "0" else None
        _cabi.check(self.lib.fb200_randn_legacy(self._states[-1].data_ptr(), n, out.data_ptr(), scratch.data_ptr(), nbytes,
                                                nxt.data_ptr(), _device.ptr(polys), _device.stream_ptr()), "fb200_randn_legacy")
        self._states.append(nxt)
        self._scratch.append(scratch)      # keep alive until the stream has passed (finish)
        return 5                           # kernels launched

    def finish_async(self):
        """Queue the D2H copy of every draw's end state + status."""
        t = self.t
        k = len(self._states) - 1
        self._host_out = t.empty((k, OUT_WORDS), dtype=t.int32).pin_memory() if (
            self._host_out is None or self._host_out.shape[0] != k) else self._host_out
        for i in range(k):
            self._host_out[i].copy_(self._states[i + 1], non_blocking=True)
        self._event.record()

    def finish(self, set_global=True):
        """Wait for the end state.  Returns the numpy state tuple after the draws, or None if a draw reported too few
        accepted candidate points (then the outputs must not be used and numpy's state is left untouched)."""
        self._event.synchronize()
        words = self._host_out.numpy().view(np.uint32)
        self._scratch = []
        self._states = []
        if any(int(words[i, STATE_WORDS]) != 0 for i in range(words.shape[0])):
            return None
        state = unpack_state(words[-1])
        if set_global:
            np.random.set_state(state)
        return state


def _selfcheck(device):
    """Device draws against numpy.random.RandomState on a private stream: values and end state, bit for bit."""
    t = _device.torch()
    rs = np.random.RandomState(20260229 % (2 ** 32))
    rs.randn(3)                                    # leaves a cached deviate and a mid-block position
    entry = rs.get_state()
    ref_a, ref_b = rs.randn(1001), rs.randn(778)
    end = rs.get_state()
    g = DeviceRandn(device)
    a = t.empty(1001, dtype=t.float64, device=device)
    b = t.empty(778, dtype=t.float64, device=device)
    g.begin(entry)
    g.draw(a)
    g.draw(b)
    g.finish_async()
    got = g.finish(set_global=False)
    if got is None:
        return False
    same = (np.array_equal(a.cpu().numpy().view(np.uint64), ref_a.view(np.uint64))
            and np.array_equal(b.cpu().numpy().view(np.uint64), ref_b.view(np.uint64))
            and np.array_equal(got[1], end[1]) and got[2:] == end[2:])
    return bool(same)


def usable(device, draws):
    """Whether the device stream may stand in for np.random.randn on this device for `draws` values per call."""
    m = mode()
    if m == "0":
        return False
    if m != "force" and draws < MIN_DEVICE_DRAWS:
        return False
    key = device.index if device.index is not None else _device.torch().cuda.current_device()
    ok = _checked.get(key)
    if ok is None:
        try:
            ok = _selfcheck(device)
        except Exception as exc:                   # pragma: no cover
            warnings.warn(f"fasta-b200: device randn self-check raised {exc!r}; drawing the probes on the host")
            ok = False
        if not ok:
            warnings.warn("fasta-b200: device randn does not reproduce numpy's stream on this system "
                          "(different libm?); drawing the Lipschitz probes on the host")
        _checked[key] = ok
    return ok

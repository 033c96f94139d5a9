"""Device-buffer plumbing: torch tensors are the buffer type, ctypes pointers cross the C ABI."""

import os

import numpy as np

from . import _cabi


def torch():
    return _cabi.require_cuda()


def is_array(x) -> bool:
    if isinstance(x, np.ndarray):
        return True
    try:
        import torch as _t
        return isinstance(x, _t.Tensor)
    except Exception:       # pragma: no cover
        return False


def is_numpy(x) -> bool:
    return isinstance(x, np.ndarray)


def to_device(x, device=None):
    """fp64, contiguous, CUDA copy/borrow of a numpy array or torch tensor."""
    t = torch()
    device = device or t.device("cuda", t.cuda.current_device())
    if isinstance(x, np.ndarray):
        return t.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(device)
    if isinstance(x, t.Tensor):
        return x.to(device=device, dtype=t.float64).contiguous()
    return t.as_tensor(np.asarray(x, dtype=np.float64), device=device)


def like_input(dev_tensor, template):
    """Hand a result back in the caller's array type (numpy in -> numpy out)."""
    if isinstance(template, np.ndarray) or not is_array(template):
        return dev_tensor.detach().cpu().numpy()
    return dev_tensor


def ptr(tensor):
    return 0 if tensor is None else tensor.data_ptr()


def stream_ptr():
    """cudaStream_t of torch's current stream on the current device (the raw accessor: torch.cuda.current_stream()
    costs ~14 us per call, and a solve asks nine times)."""
    t = torch()
    try:
        return t._C._cuda_getCurrentRawStream(t._C._cuda_getDevice())
    except AttributeError:                      # pragma: no cover  (other torch builds)
        return t.cuda.current_stream().cuda_stream


_stream_objs = {}


def current_stream():
    """torch's current stream object, cached by (device, raw handle)."""
    t = torch()
    try:
        key = (t._C._cuda_getDevice(), t._C._cuda_getCurrentRawStream(t._C._cuda_getDevice()))
    except AttributeError:                      # pragma: no cover
        return t.cuda.current_stream()
    s = _stream_objs.get(key)
    if s is None:
        s = _stream_objs[key] = t.cuda.current_stream()
    return s


class Workspace:
    """Zero-initialised device workspace + scalar block + pinned host mirror for one solve."""

    def __init__(self, M, N, device=None):
        t = torch()
        self.lib = _cabi.load()
        self.device = device or t.device("cuda", t.cuda.current_device())
        self.nbytes = int(self.lib.fb200_workspace_bytes(int(M), int(N)))
        self.buf = t.zeros(self.nbytes, dtype=t.uint8, device=self.device)
        self.scal = t.zeros(_cabi.NSCAL, dtype=t.float64, device=self.device)
        self.host = t.zeros(_cabi.NSCAL, dtype=t.float64).pin_memory()
        self._np = self.host.numpy()
        self.host_saved = t.zeros(_cabi.NSCAL, dtype=t.float64).pin_memory()
        self._np_saved = self.host_saved.numpy()
        self._saved_ready = False
        self.scal_saved = t.zeros(_cabi.NSCAL, dtype=t.float64, device=self.device)   # snapshot of the start() sums
        self._stage = None                                                            # pinned upload staging (2 vectors)
        # in-stream snapshots of the scalar block: a trial queued ahead overwrites `scal` before the host has read the
        # previous trial's sums, so each trial's sums are copied to their own pinned slot, fenced by an event
        self._slots = t.zeros((4, _cabi.NSCAL), dtype=t.float64).pin_memory()
        self._slots_np = self._slots.numpy()
        self._events = [t.cuda.Event() for _ in range(4)]
        self._ticket = 0

    def grow(self, M, N):
        need = int(self.lib.fb200_workspace_bytes(int(M), int(N)))
        if need > self.nbytes:
            t = torch()
            self.buf = t.zeros(need, dtype=t.uint8, device=self.device)
            self.nbytes = need

    def fetch(self, saved=False, with_saved=False):
        """One D2H copy of the scalar block + one stream sync: the per-decision-point sync.  with_saved: the snapshot of
        the start() sums rides along in the same sync (fetch_saved() then returns it without another one)."""
        if saved and self._saved_ready:
            self._saved_ready = False
            return self._np_saved
        self.host.copy_(self.scal_saved if saved else self.scal, non_blocking=True)
        if with_saved:
            self.host_saved.copy_(self.scal_saved, non_blocking=True)
        current_stream().synchronize()
        self._saved_ready = bool(with_saved)
        return self._np      # np.float64 elements

    def snapshot(self):
        """Queue a D2H copy of the scalar block into the next pinned slot; returns the ticket for collect()."""
        k = self._ticket & 3
        self._ticket += 1
        self._slots[k].copy_(self.scal, non_blocking=True)
        self._events[k].record()
        return k

    # The same snapshot WITHOUT a memcpy node on the compute stream.  FASTA_B200_SNAPSHOT selects how:
    #   stream   (the round-1 way) D2H copy of `scal` queued on the compute stream behind the deciding kernel
    #   side     the deciding kernel copies the block into a device ring slot; a side stream waits for it and copies the
    #            slot to pinned memory (fb200_snapshot_copy): the next trial's kernel starts right behind the deciding one
    #   zerocopy the deciding kernel writes the block straight into the pinned slot (device-visible under unified
    #            addressing); measured slower: it retires only after its PCIe writes (profiles/r02_zerocopy_ab.log)
    mode = os.environ.get("FASTA_B200_SNAPSHOT", "side")
    if os.environ.get("FASTA_B200_ZEROCOPY_SNAPSHOT", "0") != "0":
        mode = "zerocopy"
    zero_copy = mode in ("side", "zerocopy")          # the deciding kernel takes a destination pointer

    def _side_setup(self):
        t = torch()
        self._ring = t.zeros((4, _cabi.NSCAL), dtype=t.float64, device=self.device)
        self._side = t.cuda.Stream(device=self.device)
        self._ev_main = [t.cuda.Event() for _ in range(4)]
        for e in self._ev_main + self._events:       # a torch event gets its CUDA handle on first record
            e.record()
        t.cuda.current_stream().synchronize()

    def snapshot_begin(self):
        """Next pinned slot: (ticket, destination pointer for the deciding kernel)."""
        k = self._ticket & 3
        self._ticket += 1
        if self.mode == "side":
            if getattr(self, "_ring", None) is None:
                self._side_setup()
            return k, self._ring[k].data_ptr()
        return k, self._slots[k].data_ptr()

    def snapshot_end(self, k):
        if self.mode == "side":
            _cabi.check(self.lib.fb200_snapshot_copy(self._slots[k].data_ptr(), self._ring[k].data_ptr(), 8 * _cabi.NSCAL,
                                                     stream_ptr(), self._side.cuda_stream, self._ev_main[k].cuda_event,
                                                     self._events[k].cuda_event), "fb200_snapshot_copy")
        else:
            self._events[k].record()
        return k

    def collect(self, ticket):
        self._events[ticket].synchronize()
        return self._slots_np[ticket]

    def stage(self, slot, n):
        """Pinned host staging vector `slot` (0/1) of n doubles, so that an upload neither blocks the host
        behind the work already queued on the stream nor pays a cudaHostAlloc per solve."""
        t = torch()
        if self._stage is None or self._stage.shape[1] < n:
            self._stage = t.empty((2, n), dtype=t.float64).pin_memory()
        return self._stage[slot, :n]


# ---- per-device pool: a solve borrows a workspace (device scratch + pinned scalar mirror) and gives
# it back, so back-to-back solves pay no cudaHostAlloc / memset (kernels re-arm their own counters)
_pool = {}


def acquire_workspace(M, N, device=None):
    t = torch()
    device = device or t.device("cuda", t.cuda.current_device())
    key = (device.index if device.index is not None else t.cuda.current_device())
    free = _pool.setdefault(key, [])
    need = int(_cabi.load().fb200_workspace_bytes(int(M), int(N)))
    for i, ws in enumerate(free):
        if ws.nbytes >= need:
            return free.pop(i)
    return Workspace(M, N, device=device)


def release_workspace(ws):
    if ws is None:
        return
    key = ws.device.index if ws.device.index is not None else torch().cuda.current_device()
    free = _pool.setdefault(key, [])
    if len(free) < 4:
        free.append(ws)


_shared_ws = {}


def shared_workspace(M=1, N=1):
    """Per-device workspace for stand-alone operator / prox calls outside a solve."""
    t = torch()
    dev = t.cuda.current_device()
    ws = _shared_ws.get(dev)
    if ws is None:
        ws = _shared_ws[dev] = Workspace(M, N)
    else:
        ws.grow(M, N)
    return ws

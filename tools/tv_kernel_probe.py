"""Developer tool: CUDA-event timing of the two whole-iteration TV kernels at n x n (default 4096), one process per
variant knob (FASTA_B200_TVM_VARIANT / FASTA_B200_TVF_VARIANT are read once per process)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fasta-python_b200")]
import torch
from fasta import _cabi, _device
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
lib = _cabi.load()
g = torch.Generator(device="cuda").manual_seed(0)
rnd = lambda *s: torch.randn(*s, dtype=torch.float64, device="cuda", generator=g)
x0, g0, xa0 = rnd(n, n, 2), rnd(n, n, 2), rnd(n, n, 2)
b, za0 = rnd(n, n), rnd(n, n)
xa1, x1, g1, za1 = torch.empty_like(x0), torch.empty_like(x0), torch.empty_like(x0), torch.empty_like(b)
ws = _device.Workspace(1, 1)
st = _device.stream_ptr()
def it():
    _cabi.check(lib.fb200_tv_iter_fused(x0.data_ptr(), g0.data_ptr(), 0.3, n, n, _cabi.LOSS_LEAST_SQUARES, b.data_ptr(),
                                        x1.data_ptr(), g1.data_ptr(), ws.scal.data_ptr(), ws.buf.data_ptr(), st))
def fi():
    _cabi.check(lib.fb200_tv_fista_fused(x0.data_ptr(), g0.data_ptr(), 0.3, 0.6, n, n, _cabi.LOSS_LEAST_SQUARES, b.data_ptr(),
                                         xa0.data_ptr(), za0.data_ptr(), xa1.data_ptr(), za1.data_ptr(), x1.data_ptr(),
                                         g1.data_ptr(), ws.scal.data_ptr(), ws.buf.data_ptr(), st))
U = n * n * 8
for name, fn, units in (("tv_iter", it, 9), ("tv_fista", fi, 15)):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print(f"{name} TVM={os.environ.get('FASTA_B200_TVM_VARIANT', '0')} TVF={os.environ.get('FASTA_B200_TVF_VARIANT', '0')} "
          f"{us:.1f} us/launch  {units * U / us / 1e3:.0f} GB/s algorithmic ({units}U)")

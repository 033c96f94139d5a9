"""Developer tool: timing of the device randn pipeline (csrc/legacy_rng.cu) next to np.random.randn."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fasta-python_b200")]

import numpy as np
import torch

from fasta import _rng

dev = torch.device("cuda", 0)
print("selfcheck:", _rng._selfcheck(dev))
for n in [1000, 100000, 1 << 20, 4096 * 4096 * 2]:
    g = _rng.DeviceRandn(dev)
    a, b = torch.empty(n, dtype=torch.float64, device=dev), torch.empty(n, dtype=torch.float64, device=dev)
    for rep in range(3):
        np.random.seed(0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.begin()
        g.draw(a)
        g.draw(b)
        g.finish_async()
        e1.record()
        t1 = time.perf_counter()
        ok = g.finish() is not None
        t2 = time.perf_counter()
        torch.cuda.synchronize()
        dev_ms = e0.elapsed_time(e1)
    np.random.seed(0)
    t3 = time.perf_counter()
    x, y = np.random.randn(n), np.random.randn(n)
    t4 = time.perf_counter()
    same = np.array_equal(a.cpu().numpy().view(np.uint64), x.view(np.uint64)) and np.array_equal(b.cpu().numpy().view(np.uint64), y.view(np.uint64))
    print(f"n={n}: device {dev_ms:.3f} ms (host queue {1e3 * (t1 - t0):.3f} ms, to end state {1e3 * (t2 - t0):.3f} ms), "
          f"numpy 2 draws {1e3 * (t4 - t3):.3f} ms, identical={same}, ok={ok}", flush=True)

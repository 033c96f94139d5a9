"""CPU, build container only: a randomized cross-check of the numpy oracle against the UNMODIFIED live reference
(oracle/live_check.py: random problem family, shape, sparsity, mode, stop rule and solver options; bit-for-bit incl.
nan patterns).  Skipped where /root/reference does not exist (e.g. the GPU box) -- the committed fixtures of
tests/golden/ are the portable pin, this is the wider net cast where the reference can be imported."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT
from oracle import ref_loader


@pytest.mark.skipif(not ref_loader.available(), reason="live reference not present")
def test_oracle_equals_live_reference_on_random_problems():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "live_check.py"), "60", "31337"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0 and "live_check ok 60" in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]

"""N-sharded greedy loop on >= 2 GPUs (skipped on a single-GPU box): launches tests/mgpu_check.py under
torchrun, which compares every rank's result with the single-process oracle."""
import os
import subprocess
import sys
import pytest
from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_two_gpu_sharded_matches_oracle():
  import bayesiancoresets_b200._native as nat
  import ctypes
  n = ctypes.c_int(0)
  nat.check(nat.lib().bcg_device_count(ctypes.byref(n)))
  if n.value < 2:
    pytest.skip('needs >= 2 GPUs')
  world = 2
  cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
         '--master-addr', '127.0.0.1', '--master-port', '29517', os.path.join(ROOT, 'tests', 'mgpu_check.py')]
  out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
  assert out.returncode == 0 and 'MGPU OK' in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]

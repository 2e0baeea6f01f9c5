"""N-sharded greedy loop on >= 2 GPUs (skipped on a single-GPU box): spawns one tests/mgpu_check.py process per GPU
(plain subprocesses with the launcher environment -- no torchrun, no torch in the workers), which compare every rank's
result with the single-process oracle.  Runs at every world size in {2, 4, 8} the box offers."""
import os
import subprocess
import sys
import pytest
from conftest import ROOT

pytestmark = pytest.mark.gpu


def device_count():
  import bayesiancoresets_b200._native as nat
  import ctypes
  n = ctypes.c_int(0)
  nat.check(nat.lib().bcg_device_count(ctypes.byref(n)))
  return n.value


@pytest.mark.parametrize('world', [2, 4, 8])
def test_sharded_matches_oracle(world):
  if device_count() < world:
    pytest.skip('needs >= %d GPUs' % world)
  procs = []
  for r in range(world):
    env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR='127.0.0.1',
               MASTER_PORT=str(29517 + world))
    procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, 'tests', 'mgpu_check.py')], env=env,
                                  stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
  outs = []
  for p in procs:
    try:
      outs.append(p.communicate(timeout=900))
    except subprocess.TimeoutExpired:
      for q in procs:
        q.kill()
      raise
  assert all(p.returncode == 0 for p in procs) and 'MGPU OK' in outs[0][0], \
      '\n'.join(o[0][-1500:] + o[1][-1500:] for o in outs)

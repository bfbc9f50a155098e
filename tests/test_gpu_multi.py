"""Multi-GPU invariants (needs >= 2 visible GPUs; skipped otherwise): launches tests/multigpu_check.py
under torchrun, one process per GPU over NCCL."""
import os
import subprocess
import sys

import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu


def test_two_rank_sharding_invariants():
  if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
    pytest.skip('needs >= 2 GPUs')
  n = 2
  here = os.path.dirname(os.path.abspath(__file__))
  cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(n),
         '--master-addr', '127.0.0.1', '--master-port', '29517', os.path.join(here, 'multigpu_check.py')]
  out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
  assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
  assert 'MULTIGPU OK' in out.stdout

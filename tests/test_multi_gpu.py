"""-m gpu, needs >= 2 GPUs (skipped otherwise): the NCCL merges of the order-dependent
collectors, one process per GPU, against the oracle's single pass (tests/mgpu_worker.py)."""
import os
import subprocess
import sys

import pytest

from tests.helpers import ROOT

pytestmark = pytest.mark.gpu


def _n_gpus():
    from sequali_b200 import _lib
    try:
        return _lib.load().sq_device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_merges_match_the_oracle(world):
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tests", "mgpu_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0 and "MGPU PARITY OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]

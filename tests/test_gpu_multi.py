"""Channel-sharded run on REAL GPUs (needs >= 2 visible devices; skipped otherwise): two ranks under torchrun, collectives inside
libnmb200 (nm_comm_*: NCCL all-reduce of the common-average sums, result blocks through the shared page-locked matrix and through
the grouped send / receive gather) -- the merged matrix must equal the un-sharded run of the same recording (tools/sharded_check.py)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.gpu
@pytest.mark.parametrize("native", [True, False])
def test_two_rank_sharded_run_equals_unsharded(native):
    from py_neuromodulation_b200 import _lib

    _lib._LIB = None
    _lib.load()
    if _lib.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29611" if native else "29631", str(ROOT / "tools" / "sharded_check.py")] + (["--native"] if native else [])
    env = dict(os.environ, NCCL_DEBUG_FILE="/dev/stderr")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-3000:]
    assert "SHARDED CHECK PASSED" in r.stdout, r.stdout[-3000:]

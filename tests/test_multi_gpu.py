"""Multi-GPU parity on real devices (needs >= 2 GPUs; skipped otherwise): tools/multi_gpu_check.py under
torchrun — view sharding + sr_comm_allgather_views + cross-check, and row sharding + sr_comm_allgather_rows,
each compared bit for bit with the same job on one GPU (sharding must be invisible).  The host-side
bookkeeping is covered on CPU by tests/test_sharding.py (gloo, world_size 2)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _num_gpus():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:  # noqa: BLE001
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_job_equals_single_gpu(world):
    if _num_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29530 + world), os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MULTI-GPU CHECK OK" in r.stdout

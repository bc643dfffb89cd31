"""Multi-GPU test (needs >= 2 GPUs on the box, skipped otherwise): one rank per GPU over NCCL runs the byte-range sharded
scan + strip + header parse (record all_gather, stitch, parameter-set hand-over) and checks it against the unsharded result."""
import os
import subprocess
import sys

import pytest

from oracle import ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
def test_nccl_sharded_scan_and_parse():
    n = _gpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    env = dict(os.environ, PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1", "--master-port",
           "29541", os.path.join(ROOT, "tests", "shard_nccl_worker.py")]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "SHARD_NCCL_OK" in out.stdout

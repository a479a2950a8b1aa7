"""The N > 1 plumbing (one process per GPU, no data-path collective) on CPU with gloo, world size 2,
plus the in-process utterance sharding of the Python API."""
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT


def test_two_ranks_gloo():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.join(ROOT, "tests", "_dist_worker.py")]
    res = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-3000:]
    assert "rank 0 ok" in res.stdout and "rank 1 ok" in res.stdout


def test_reference_arm_only_rank0_prints(tmp_path):
    """--impl reference under torchrun: rank 0 alone works and prints, the others exit 0 silently."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--workload", "cfg1", "--gpus", "2"], env=env, stdout=subprocess.PIPE,
                         stderr=subprocess.PIPE, text=True, timeout=300)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_api_sharding_is_a_partition():
    from lws_b200 import api
    for n in (1, 2, 7, 64, 65, 256):
        for k in (1, 2, 4, 8):
            parts = api._shard(n, k)
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1 and min(sizes) >= 1 and len(parts) == min(n, k)

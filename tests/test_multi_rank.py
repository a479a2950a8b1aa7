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
    """equal lengths: contiguous, balanced split (BASELINE configs[3]: 256 over 8 -> 32 per GPU; configs[4]: 32 -> 4)"""
    from lws_b200 import api
    for n in (1, 2, 7, 32, 64, 65, 256):
        for k in (1, 2, 4, 8):
            parts = api._shard([628] * n, k)
            assert sorted(i for p in parts for i in p) == list(range(n))
            assert all(p == list(range(p[0], p[-1] + 1)) for p in parts)
            sizes = [len(p) for p in parts]
            assert max(sizes) - min(sizes) <= 1 and min(sizes) >= 1 and len(parts) == min(n, k)
    assert [len(p) for p in api._shard([628] * 256, 8)] == [32] * 8
    assert [len(p) for p in api._shard([5632] * 32, 8)] == [4] * 8


def test_api_sharding_balances_ragged_batches_by_frames():
    """ragged lengths: longest-processing-time-first over the frame counts -- the busiest device
    carries at most 4/3 - 1/(3k) of the optimum (Graham's bound), far better than the split by count"""
    from lws_b200 import api
    rng = np.random.default_rng(8)
    for n, k in ((9, 2), (40, 4), (100, 8), (17, 8), (3, 8)):
        frames = [int(f) for f in rng.integers(20, 3000, n)]
        frames[0] = 9000  # one long utterance
        parts = api._shard(frames, k)
        assert sorted(i for p in parts for i in p) == list(range(n)) and all(p == sorted(p) for p in parts)
        loads = [sum(frames[i] for i in p) for p in parts]
        lower = max(max(frames), -(-sum(frames) // min(k, n)))
        assert max(loads) <= lower * (4.0 / 3.0) + 1, (loads, lower)
        by_count = [(n * i) // min(k, n) for i in range(min(k, n) + 1)]
        naive = max(sum(frames[a:b]) for a, b in zip(by_count, by_count[1:]))
        assert max(loads) <= naive


def test_api_rejects_duplicate_devices():
    from lws_b200 import api
    import pytest
    with pytest.raises(ValueError):
        api._devices([0, 0])
    assert api._devices([1, 0]) == [1, 0] and api._devices(None) == [0] and api._devices(3) == [3]

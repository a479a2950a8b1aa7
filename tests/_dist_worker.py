"""Worker of tests/test_multi_rank.py: run under torchrun with the gloo backend (CPU)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench

r = bench.Ranks("gloo")
assert r.world == int(os.environ["WORLD_SIZE"]) and r.rank == int(os.environ["RANK"])
r.barrier()
# the slowest rank sets the time, the job's units are the sum over ranks
t = r.max(0.010 * (r.rank + 1))
assert abs(t - 0.010 * r.world) < 1e-12
units = r.sum(1000.0 + r.rank)
assert units == sum(1000.0 + k for k in range(r.world))
v = bench.throughput(units / r.world, 4, r.world, t)
assert abs(v - units * 4 / t) < 1e-6 * v
# every rank synthesises different utterances (weak scaling), reproducibly
x = bench.signals("cfg1", r.rank)
assert x.shape == (1, 32000)
h = float(np.abs(x).sum())
hs = [None] * r.world
r.dist.all_gather_object(hs, h)
assert len(set(hs)) == r.world
assert np.array_equal(x, bench.signals("cfg1", r.rank))
r.barrier()
r.close()
print("rank %d ok" % r.rank)

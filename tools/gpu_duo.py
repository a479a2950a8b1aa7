"""GPU experiment: the strip kernel with two lanes per task (variant 3, DUO) against one thread per task (variant 2)
over plans, BASELINE configs[1] (64 x 10 s, 1024/256, 100 sweeps), default and zero thresholds."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import lws_b200
from lws_b200 import api
ctx = api._context(0)
p = lws_b200.lws(1024, 256)
B = int(os.environ.get("DUO_B", "64"))
x = np.stack([np.random.default_rng(2000 + b).standard_normal(160000) for b in range(B)])
A = np.abs(p.stft(x))
ref = {}
plans = [(2, 0, 0, 0), (3, 0, 0, 0), (3, 2, 7, 0), (3, 2, 6, 5), (3, 4, 12, 0), (3, 4, 16, 0), (3, 4, 20, 0), (3, 4, 16, 5), (3, 8, 24, 0), (3, 8, 32, 0),
         (2, 4, 16, 0)]
if len(sys.argv) > 1:
    plans = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for var, cl, sw, lag in plans:
    for name, thr in (("default", None), ("zero", np.zeros(100))):
        try:
            ctx.set_tuning(0, cl, sw)
            ctx.set_variant(lag, var)
            ms = []
            for _ in range(3):
                Y = p.batch_lws(A, thresholds=thr)
                ms.append(ctx.last_compute_ms())
            pl = ctx.last_batch_plan()
            cyc = ctx.last_batch_cycles()
            w = max(cyc["warps"], 1)
            key = name
            if key not in ref:
                ref[key] = Y
            same = np.array_equal(Y, ref[key])
            print("var %d force(C=%d,G=%d,lag=%d) %-7s -> C=%d NS=%d G=%d lag=%d gfast=%d thr=%d var=%d: %.2f ms  same=%s  (work %.1f waitS %.1f waitN %.1f Mclk/warp)" % (
                var, cl, sw, lag, name, pl["cluster"], pl["frame_slots"], pl["sweeps_per_pass"], pl["sweep_lag"], pl["sweep_fastest"], pl["threads"],
                pl["tensor_memory"], min(ms), same, cyc["warp_work"] / w / 1e6, cyc["warp_wait_strip"] / w / 1e6, cyc["warp_wait_neighbours"] / w / 1e6), flush=True)
        except Exception as ex:
            print("var %d C=%d G=%d lag=%d %s EXC %s" % (var, cl, sw, lag, name, str(ex)[-200:]), flush=True)
ctx.set_tuning(0, 0, 0); ctx.set_variant(0, 0)

"""GPU timing experiment: what a macro-step of the strip kernel is made of.  Runs the cfg2 batch (zero thresholds) with the
library given by LWSB_LIB_PATH (builds with -DLWSB_ABLATE=1: no neighbour-frame loads, 2: no sqrt / division, 3: both)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import lws_b200
from lws_b200 import api
ctx = api._context(0)
p = lws_b200.lws(1024, 256)
B = 64
x = np.stack([np.random.default_rng(2000 + b).standard_normal(160000) for b in range(B)])
A = np.abs(p.stft(x))
for var, cl, sw, lag in ((2, 2, 7, 4), (3, 2, 7, 4), (2, 2, 5, 5), (3, 2, 5, 5), (3, 4, 14, 5)):
    ctx.set_tuning(0, cl, sw); ctx.set_variant(lag, var)
    ms = []
    for _ in range(2):
        p.batch_lws(A, thresholds=np.zeros(100)); ms.append(ctx.last_compute_ms())
    pl = ctx.last_batch_plan(); cyc = ctx.last_batch_cycles(); w = max(cyc["warps"], 1)
    print("%s var %d C=%d G=%d lag=%d gfast=%d: %.2f ms (work %.1f waitS %.1f waitN %.1f Mclk/warp)" % (
        os.environ.get("LWSB_LIB_PATH", "default")[-12:], var, pl["cluster"], pl["sweeps_per_pass"], pl["sweep_lag"], pl["sweep_fastest"], min(ms),
        cyc["warp_work"] / w / 1e6, cyc["warp_wait_strip"] / w / 1e6, cyc["warp_wait_neighbours"] / w / 1e6), flush=True)

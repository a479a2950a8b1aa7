"""GPU experiment: variants of the cluster strip kernel at BASELINE configs[1] shape (64 x 10 s, 1024/256, 100 sweeps).

For every variant (one thread per task = 2, pair-split modes 10..15 = window mode + 3 * explicit pipelining) and a few
forced plans: sweep-kernel time (CUDA events inside the library) and bit-equality with the one-thread-per-task result,
which the parity tests pin to the oracle.  Modes other than the default one need a build with -DLWSB_PAIR_EXPERIMENTS.

    python tools/gpu_pair.py [B] [variants,comma] [plans: C:G;C:G...]
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import lws_b200
from lws_b200 import api

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
variants = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [2, 10, 11, 12, 13, 14, 15]
plans = [tuple(int(q) for q in p.split(":")) for p in sys.argv[3].split(";")] if len(sys.argv) > 3 else [(0, 0), (2, 0), (4, 0), (8, 0)]
plans = [p if len(p) == 3 else p + (0,) for p in plans]   # cluster : sweeps per pass : bins per block

p = lws_b200.lws(1024, 256)
x = np.stack([np.random.default_rng(2000 + b).standard_normal(160000) for b in range(B)])
A = np.abs(p.stft(x))
ctx = api._context(0)
its = 100
for thrname, thr in (("default", lws_b200.get_thresholds(its, 100.0, 0.1, 1)), ("zero", np.zeros(its))):
    ctx.set_variant(0, 2); ctx.set_tuning(0, 0, 0); ctx.set_block_bins(8)
    ref = p.batch_lws(A, thresholds=thr)
    for var in variants:
        for cl, sw, bk in plans:
            ctx.set_variant(0, var); ctx.set_tuning(0, cl, sw); ctx.set_block_bins(bk)
            try:
                Y = p.batch_lws(A, thresholds=thr)
                ms = []
                for _ in range(2):
                    Y = p.batch_lws(A, thresholds=thr)
                    ms.append(ctx.last_compute_ms())
                plan = ctx.last_batch_plan()
                cyc = ctx.last_batch_cycles() if hasattr(ctx, "last_batch_cycles") else None
                print("%-7s variant %2d force(C=%d,G=%d,bk=%d) -> bk=%d C=%d NS=%d G=%d lag=%d gfast=%d thr=%d var=%d: %.1f ms  equal=%s%s"
                      % (thrname, var, cl, sw, bk, plan["block_bins"], plan["cluster"], plan["frame_slots"], plan["sweeps_per_pass"], plan["sweep_lag"],
                         plan["sweep_fastest"], plan["threads"], plan["tensor_memory"], min(ms), np.array_equal(Y, ref),
                         "" if cyc is None else "  work/waitS/waitN per warp (Mclk): %.0f/%.0f/%.0f" % tuple(
                             cyc[k] / max(cyc["warps"], 1) / 1e6 for k in ("warp_work", "warp_wait_strip", "warp_wait_neighbours"))), flush=True)
            except Exception as ex:
                print("%-7s variant %2d force(C=%d,G=%d,bk=%d): EXC %s" % (thrname, var, cl, sw, bk, ex), flush=True)
ctx.set_variant(0, 0); ctx.set_tuning(0, 0, 0); ctx.set_block_bins(0)

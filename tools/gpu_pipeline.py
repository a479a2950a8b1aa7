"""GPU experiment: pass pipelining across clusters -- kernel time of batch_lws vs batch size and sweeps per pass."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import lws_b200
from lws_b200 import api
ctx = api._context(0)
def run(fs, hop, nsamp, B, its, plans, tag):
    p = lws_b200.lws(fs, hop)
    x = np.stack([np.random.default_rng(5000 + b).standard_normal(nsamp) for b in range(B)])
    A = np.abs(p.stft(x))
    for cl, sw in plans:
        ctx.set_tuning(0, cl, sw)
        try:
            ms = []
            for _ in range(2):
                p.batch_lws(A, iterations=its)
                ms.append(ctx.last_compute_ms())
            pl = ctx.last_batch_plan()
            cyc = ctx.last_batch_cycles()
            w = max(cyc["warps"], 1)
            print("%s B=%d force(C=%d,G=%d) -> C=%d NS=%d G=%d lag=%d thr=%d: %.2f ms (work %.1f waitN %.1f Mclk/warp, ctrl %.1f/%.1f/%.1f)" % (
                tag, B, cl, sw, pl["cluster"], pl["frame_slots"], pl["sweeps_per_pass"], pl["sweep_lag"], pl["threads"], min(ms),
                cyc["warp_work"] / w / 1e6, cyc["warp_wait_neighbours"] / w / 1e6, cyc["ctrl_publish"] / 1e6, cyc["ctrl_poll"] / 1e6, cyc["ctrl_tma"] / 1e6), flush=True)
        except Exception as ex:
            print(tag, B, cl, sw, "EXC", ex, flush=True)
    ctx.set_tuning(0, 0, 0)
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "cfg1"):
    run(512, 128, 32000, 1, 100, [(0, 0), (4, 5), (4, 14), (2, 5), (2, 10), (2, 20), (1, 10), (1, 34)], "cfg1")
if which in ("all", "cfg2"):
    for B in (1, 8, 64):
        run(1024, 256, 160000, B, 100, [(0, 0), (2, 7), (2, 4), (4, 16), (4, 8)], "cfg2")
if which in ("all", "cfg5"):
    run(2048, 256, 1440000, 4, 200, [(0, 0), (8, 9), (8, 5), (4, 3)], "cfg5")
if which in ("all", "trace"):
    p = lws_b200.lws(1024, 256)
    x = np.stack([np.random.default_rng(5000 + b).standard_normal(160000) for b in range(1)])
    A = np.abs(p.stft(x))
    ctx.batch_trace(True)
    for thr in (None, np.zeros(100)):
        p.batch_lws(A, thresholds=thr); p.batch_lws(A, thresholds=thr)
        print("plan", ctx.last_batch_plan(), ctx.last_compute_ms(), "ms")
        for row in ctx.batch_trace(True):
            print("utt %d pass %2d: taken %8.3f primed %8.3f computed %8.3f written %8.3f ms; Mclk rows-wait %.2f nbr-poll %.2f work %.2f wait-ctrl %.2f" % (
                row[0], row[1], row[2] / 1e6, row[3] / 1e6, row[4] / 1e6, row[5] / 1e6, row[6] / 1e6, row[7] / 1e6, row[8] / 1e6, row[9] / 1e6))
    ctx.batch_trace(False)
if which in ("trace64",):
    p = lws_b200.lws(1024, 256)
    x = np.stack([np.random.default_rng(2000 + b).standard_normal(160000) for b in range(64)])
    A = np.abs(p.stft(x))
    ctx.batch_trace(True)
    for cl, sw in ((0, 0),):
        ctx.set_tuning(0, cl, sw)
        p.batch_lws(A); p.batch_lws(A)
        rows = ctx.batch_trace(True)
        pl = ctx.last_batch_plan()
        print("plan C=%d G=%d" % (pl["cluster"], pl["sweeps_per_pass"]), ctx.last_compute_ms(), "ms;", len(rows), "items")
        end = max(r[5] for r in rows)
        # clusters: items taken in order by the same cluster are those whose 'taken' follows another's 'written'
        prime = sum(r[3] - r[2] for r in rows) / 1e6
        comp = sum(r[4] - r[3] for r in rows) / 1e6
        print("  sum over items: priming %.1f ms, computing %.1f ms; kernel end %.2f ms; mean item %.2f ms" % (prime, comp, end / 1e6, comp / len(rows)))
        bypass = {}
        for r in rows:
            bypass.setdefault(r[1], []).append(r)
        for ps in sorted(bypass):
            rs = bypass[ps]
            print("  pass %2d: items %3d taken %.2f..%.2f  priming mean %.2f max %.2f  computing mean %.2f ms  Mclk: rows-wait %.2f nbr-poll %.2f work %.2f wait-ctrl %.2f" % (
                ps, len(rs), min(r[2] for r in rs) / 1e6, max(r[2] for r in rs) / 1e6, np.mean([r[3] - r[2] for r in rs]) / 1e6,
                max(r[3] - r[2] for r in rs) / 1e6, np.mean([r[4] - r[3] for r in rs]) / 1e6, np.mean([r[6] for r in rs]) / 1e6,
                np.mean([r[7] for r in rs]) / 1e6, np.mean([r[8] for r in rs]) / 1e6, np.mean([r[9] for r in rs]) / 1e6))
    ctx.batch_trace(False); ctx.set_tuning(0, 0, 0)

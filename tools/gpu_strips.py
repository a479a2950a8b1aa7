"""GPU diagnostic for the cluster strip kernel: bit-exactness vs the oracle over cluster sizes / passes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import lws_oracle, lws_b200
from lws_b200 import _native, api
from conftest import relF, make_signal

def main():
    ctx = api._context(0)
    cases = [(512, 128, 6000, 6), (512, 128, 32000, 100), (1024, 256, 30000, 12), (2048, 256, 40000, 8), (128, 64, 9000, 10)]
    for fs, hop, n, its in cases:
        po, pg = lws_oracle.lws(fs, hop), lws_b200.lws(fs, hop)
        A = np.abs(po.stft(make_signal("tonal", 3, n)))
        for thrname, thr in (("zero", np.zeros(its)), ("default", lws_b200.get_thresholds(its, 100 if its > 50 else 2.0, 0.1, 1))):
            Yo = po.batch_lws(A, thresholds=thr)
            for cl, sw, sm in ((1, 0, 0), (2, 0, 0), (4, 0, 0), (8, 0, 0), (0, 0, 0), (2, 3, 0), (4, 2, 60000)):
                ctx.set_tuning(sm, cl, sw)
                t0 = time.time()
                try:
                    Y = pg.batch_lws(A, thresholds=thr)
                    plan = ctx.last_batch_plan()
                    print("%d/%d T=%d its=%d %-7s force(C=%d,G=%d,smem=%d) -> plan %s: equal=%s relF=%.2e  %.1f ms"
                          % (fs, hop, A.shape[0], its, thrname, cl, sw, sm,
                             None if plan is None else (plan["cluster"], plan["blocks_per_strip"], plan["sweeps_per_pass"], plan["ring_rows"]),
                             np.array_equal(Y, Yo), relF(Y, Yo), 1e3 * (time.time() - t0)), flush=True)
                except Exception as ex:
                    print("%d/%d its=%d %s force(C=%d,G=%d): EXC %s" % (fs, hop, its, thrname, cl, sw, ex), flush=True)
    ctx.set_tuning(0, 0, 0)
    # batch of ragged utterances, more utterances than clusters
    po, pg = lws_oracle.lws(512, 128), lws_b200.lws(512, 128)
    As = [np.abs(po.stft(make_signal("white" if i % 2 else "tonal", 50 + i, 3000 + 700 * (i % 9)))) for i in range(45)]
    thr = lws_b200.get_thresholds(9, 2.0, 0.2, 1)
    for cl in (8, 4, 0):
        ctx.set_tuning(0, cl, 0)
        Ys = pg.batch_lws(As, thresholds=thr)
        ok = [np.array_equal(Y, po.batch_lws(A, thresholds=thr)) for A, Y in zip(As, Ys)]
        print("ragged batch of 45, cluster %d: plan %s, all equal: %s (%d/45)" % (cl, ctx.last_batch_plan(), all(ok), sum(ok)), flush=True)
    ctx.set_tuning(0, 0, 0)

if __name__ == "__main__":
    main()

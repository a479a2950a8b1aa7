import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import lws_oracle, lws_b200
from lws_b200 import _native, api
ctx = api._context(0)
po, pg = lws_oracle.lws(512, 128), lws_b200.lws(512, 128)
rng = np.random.default_rng(0)
for T in (5, 10, 30, 50, 51, 70, 100):
    A = np.abs(rng.standard_normal((T, 257)))
    for its in (2, 3, 6, 7):
        for cl, sw in ((2, 1), (2, 2), (2, 3), (4, 2), (2, 0)):
            thr = np.zeros(its)
            ctx.set_tuning(0, cl, sw)
            try:
                Y = pg.batch_lws(A, thresholds=thr)
                ok = np.array_equal(Y, po.batch_lws(A, thresholds=thr))
                print("T=%d its=%d C=%d G=%d plan=%s ok=%s" % (T, its, cl, sw, ctx.last_batch_plan() and ctx.last_batch_plan()["sweeps_per_pass"], ok), flush=True)
            except Exception as ex:
                print("T=%d its=%d C=%d G=%d EXC %s" % (T, its, cl, sw, str(ex)[-60:]), flush=True)

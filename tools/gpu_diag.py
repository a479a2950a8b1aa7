"""GPU diagnostic (not a test): where does the CUDA path start to deviate from the oracle?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import lws_oracle, lws_b200
from lws_b200 import _native
from conftest import relF, make_signal

def main():
    for fs, hop, kind, n in [(512, 128, "white", 32000), (512, 128, "tonal", 32000), (64, 16, "tonal", 1500)]:
        po, pg = lws_oracle.lws(fs, hop), lws_b200.lws(fs, hop)
        A = np.abs(po.stft(make_signal(kind, 42, n)))
        ctx = _native.Context(0)
        ctx.set_weights(_native.W, pg.W)
        ctx.load([A], _native.F64)
        mean, mx = ctx.stats()
        print(fs, hop, kind, A.shape, "mean gpu-numpy", mean[0] - np.mean(A), "max", mx[0] - A.max(), flush=True)
        ctx.close()
        for its in (1, 2, 3, 5, 10, 40, 100):
            for name, thr in (("zero", np.zeros(its)), ("default", lws_b200.get_thresholds(its, 100, 0.1, 1)),
                              ("low", lws_b200.get_thresholds(its, 1.0, 0.1, 1))):
                Yo = po.batch_lws(A, thresholds=thr)
                Y1 = pg.batch_lws(A, thresholds=thr)
                Y2 = pg.batch_lws(A, thresholds=thr)
                Y3 = lws_b200.batch_lws(A, pg.W, thr, flags=_native.FORCE_ANYQ)
                d = np.abs(Y1 - Yo)
                worst = np.unravel_index(np.argmax(d), d.shape)
                print("  its %3d %-7s relF %.2e  rerun-equal %s  anyq-vs-folded %.2e  worst bin %s amp %.2e (mean %.2e) nbad %d"
                      % (its, name, relF(Y1, Yo), np.array_equal(Y1, Y2), relF(Y3, Y1), worst, A[worst], A.mean(),
                         int((d > 1e-6 * A.mean()).sum())), flush=True)

if __name__ == "__main__":
    main()

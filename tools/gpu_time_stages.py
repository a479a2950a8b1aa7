"""Wall-clock of the other stages at BASELINE configs[2] / [3] shape (64 x 10 s, 1024/256, mode='music')."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import lws_b200
from lws_b200 import api
p = lws_b200.lws(1024, 256, mode="music")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
x = np.stack([np.random.default_rng(3000 + b).standard_normal(160000) for b in range(B)])
A = np.abs(p.stft(x))
ctx = api._context(0)
for name, fn in (("nofuture_lws (1 it)", lambda: p.nofuture_lws(A)), ("online_lws (10 it, LA=3)", lambda: p.online_lws(A)),
                 ("batch_lws (100 it)", lambda: p.batch_lws(A)), ("run_lws (music)", lambda: p.run_lws(A))):
    fn()
    t0 = time.perf_counter(); fn(); dt = time.perf_counter() - t0
    print("%-28s %8.1f ms wall for %d utterances (%.3e bins/s), last kernel %.1f ms" % (name, 1e3 * dt, B, A.size / dt, ctx.last_compute_ms()), flush=True)
t0 = time.perf_counter(); X = p.stft(x); t1 = time.perf_counter(); y = p.istft(X); t2 = time.perf_counter()
print("stft %.1f ms, istft %.1f ms for %d x 10 s" % (1e3 * (t1 - t0), 1e3 * (t2 - t1), B))

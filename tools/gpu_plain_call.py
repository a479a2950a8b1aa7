"""GPU timing: the plain drop-in call p.batch_lws(A) (pageable numpy in, fresh numpy out) at BASELINE configs[1], call by call."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import lws_b200
from lws_b200 import api
p = lws_b200.lws(1024, 256)
x = np.stack([np.random.default_rng(2000 + b).standard_normal(160000) for b in range(64)])
A = np.array(np.abs(p.stft(x)))
ctx = api._context(0)
Y = None
for i in range(7):
    t0 = time.perf_counter()
    Y = p.batch_lws(A)
    dt = time.perf_counter() - t0
    print("call %d: %.1f ms wall, kernel %.1f ms (LWSB_PINNED_RESULTS=%s, LWSB_HOST_THREADS=%s)" % (
        i, 1e3 * dt, ctx.last_compute_ms(), os.environ.get("LWSB_PINNED_RESULTS", "1"), os.environ.get("LWSB_HOST_THREADS", "auto")), flush=True)
print("cpu count", os.cpu_count())

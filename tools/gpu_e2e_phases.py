"""Where the end-to-end time of batch_lws goes at BASELINE configs[1]: load (H2D + extend + stats), sweeps, store (crop + D2H)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import lws_b200
from lws_b200 import _native
p = lws_b200.lws(1024, 256)
x = np.stack([np.random.default_rng(2000 + b).standard_normal(160000) for b in range(64)])
A = np.abs(p.stft(x))
A_pin = torch.from_numpy(A).pin_memory().numpy()
Y_pin = torch.empty(A.shape, dtype=torch.complex128).pin_memory().numpy()
thr = lws_b200.get_thresholds(100, 100.0, 0.1, 1)
ctx = _native.Context(0)
ctx.set_weights(_native.W, p.W)
arrs = [A_pin[b] for b in range(64)]
outs = [Y_pin[b] for b in range(64)]
for rep in range(3):
    t0 = time.perf_counter(); ctx.load(arrs, _native.F64); t1 = time.perf_counter(); ctx.batch(thr); t2 = time.perf_counter(); ctx.store(outs); t3 = time.perf_counter()
    print("load %.2f ms (%.1f GB/s if all copy)  sweeps %.2f ms (kernel %.2f)  store %.2f ms (%.1f GB/s if all copy)  total %.2f" % (
        1e3 * (t1 - t0), A.nbytes / (t1 - t0) / 1e9, 1e3 * (t2 - t1), ctx.last_compute_ms(), 1e3 * (t3 - t2), Y_pin.nbytes / (t3 - t2) / 1e9, 1e3 * (t3 - t0)))
t0 = time.perf_counter(); p.batch_lws(A_pin, thresholds=thr, out=Y_pin); print("public API %.2f ms" % (1e3 * (time.perf_counter() - t0)))
t0 = time.perf_counter(); p.batch_lws(A_pin, thresholds=thr, out=Y_pin); print("public API %.2f ms" % (1e3 * (time.perf_counter() - t0)))

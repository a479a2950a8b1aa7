"""One batch_lws call at BASELINE configs[1] shape on B utterances for ncu (strip kernel variant from argv)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import lws_b200
from lws_b200 import api
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
var = int(sys.argv[2]) if len(sys.argv) > 2 else 0
cl = int(sys.argv[3]) if len(sys.argv) > 3 else 0
sw = int(sys.argv[4]) if len(sys.argv) > 4 else 0
thr = None if len(sys.argv) <= 5 or sys.argv[5] == "default" else np.zeros(100)
p = lws_b200.lws(1024, 256)
x = np.stack([np.random.default_rng(2000 + b).standard_normal(160000) for b in range(B)])
A = np.abs(p.stft(x))
ctx = api._context(0)
ctx.set_variant(0, var); ctx.set_tuning(0, cl, sw)
for _ in range(2):
    Y = p.batch_lws(A, thresholds=thr)
print(ctx.last_batch_plan(), ctx.last_compute_ms())

"""One batch_lws call at BASELINE configs[4] geometry (2048-pt, hop 256, Q = 8) on 4 utterances of 1/4 length for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import lws_b200
from lws_b200 import api
p = lws_b200.lws(2048, 256)
x = np.stack([np.random.default_rng(5000 + b).standard_normal(360000) for b in range(4)])
A = np.abs(p.stft(x))
ctx = api._context(0)
for _ in range(2):
    Y = p.batch_lws(A, iterations=200)
print(A.shape, ctx.last_batch_plan(), ctx.last_compute_ms())

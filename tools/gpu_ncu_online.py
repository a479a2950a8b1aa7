"""One online_lws call at BASELINE configs[2] shape on B utterances for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import lws_b200
from lws_b200 import api
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
p = lws_b200.lws(1024, 256, mode="music")
x = np.stack([np.random.default_rng(3000 + b).standard_normal(160000) for b in range(B)])
A = np.abs(p.stft(x))
ctx = api._context(0)
for _ in range(2):
    Y = p.online_lws(A)
print(ctx.last_online_kernel(), ctx.last_compute_ms())

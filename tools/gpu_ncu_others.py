"""One pass through the kernels around the sweeps (load / stats / no-future / transforms / consistency / crop) at BASELINE configs[1]
shape on 16 utterances, for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import lws_b200
p = lws_b200.lws(1024, 256, mode="music", online_iterations=0, batch_iterations=0)
x = np.stack([np.random.default_rng(4000 + b).standard_normal(160000) for b in range(16)])
for _ in range(2):
    X = p.stft(x)
    A = np.abs(X)
    Y = p.nofuture_lws(A)
    y = p.istft(Y)
    c = p.get_consistency(Y)
print(X.shape, y.shape, c[:2])

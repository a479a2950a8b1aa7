"""Small calls through the shipping kernel families, for compute-sanitizer (memcheck / racecheck / synccheck).
usage: gpu_sanitize.py [ops] [shapes]   ops: letters of n(ofuture) o(nline) b(atch) r(un_lws) t(ransforms); shapes: e.g. 64x16,128x64"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import lws_b200
from lws_b200 import api
ops = sys.argv[1] if len(sys.argv) > 1 else "nobrt"
shapes = [tuple(int(v) for v in s.split("x")) for s in (sys.argv[2] if len(sys.argv) > 2 else "64x16,128x64,256x32").split(",")]
rng = np.random.default_rng(1)
for fs, hop in shapes:
    p = lws_b200.lws(fs, hop, mode="music", batch_iterations=6, batch_alpha=1, online_iterations=3)
    x = rng.standard_normal((2, 1500))
    A = np.abs(p.stft(x))
    k = None
    if "n" in ops: p.nofuture_lws(A)
    if "o" in ops:
        p.online_lws(A); k = api._context(0).last_online_kernel()
    if "b" in ops: p.batch_lws(A)
    if "r" in ops: p.run_lws(A[0])
    if "t" in ops: p.istft(A.astype(np.complex128))
    print(fs, hop, A.shape, "online kernel", k)
print("done")

#!/usr/bin/env python
"""Generate tests/golden/*.npz from the COMPILED REFERENCE (oracle/_ref/lws_ref*.so).

Run in the build container only (needs /root/reference to have been compiled with
``make -C oracle ref``); the vectors travel with the repository, the reference does not.

    python tools/make_golden.py

Every case stores its constructor arguments, the synthetic input and what the reference
returned, so the tests can replay it against the oracle (CPU) and the CUDA path (GPU).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
import lws_ref  # noqa: E402  (the reference's Cython module, built as lws_ref)

OUT = os.path.join(ROOT, "tests", "golden")


def white(seed, n):
    return np.random.default_rng(seed).standard_normal(n)


def tonal(seed, n, sr=16000.0):
    """Harmonic, amplitude-modulated source: the hard case for parity."""
    t = np.arange(n) / sr
    f0 = 120 + 30 * np.sin(2 * np.pi * 3 * t)
    ph = 2 * np.pi * np.cumsum(f0) / sr
    x = sum(np.sin(k * ph) / k for k in range(1, 30))
    x = x * (0.5 + 0.5 * np.sin(2 * np.pi * 2 * t)) ** 2
    return x + 0.01 * np.random.default_rng(seed).standard_normal(n)


def main():
    os.makedirs(OUT, exist_ok=True)
    cases = []
    # (name, ctor args, ctor kwargs, signal kind, n samples)
    grid = [
        ("q2", (32, 16), {}, "white", 700),
        ("q4", (32, 8), {}, "white", 500),
        ("q4b", (64, 16), {}, "tonal", 1500),
        ("q8", (64, 8), {}, "white", 520),
        ("q8b", (32, 4), {}, "tonal", 300),
        ("q6", (48, 8), {}, "white", 500),
        ("q3", (36, 12), {}, "white", 500),
        ("q4_L3", (32, 8), {"L": 3}, "white", 400),
        ("q4_L7", (64, 16), {"L": 7}, "white", 900),
        ("q4_la0", (32, 8), {"look_ahead": 0}, "white", 400),
        ("q4_la1", (32, 8), {"look_ahead": 1}, "white", 400),
        ("q4_la5", (32, 8), {"look_ahead": 5}, "tonal", 400),
        ("q8_la2", (64, 8), {"look_ahead": 2}, "white", 400),
        ("q2_la4", (32, 16), {"look_ahead": 4}, "white", 500),
    ]
    for name, args, kw, kind, n in grid:
        seed = len(cases) + 1
        x = white(seed, n) if kind == "white" else tonal(seed, n)
        ref = lws_ref.lws(*args, mode="music", **kw)
        X = ref.stft(x)
        A = np.abs(X)
        rng = np.random.default_rng(100 + seed)
        Sc = A * np.exp(1j * rng.uniform(0, 2 * np.pi, A.shape))  # complex input with random phases
        thr_mid = lws_ref.get_thresholds(6, 2.0, 0.4, 1)
        d = dict(
            x=x, awin=ref.awin, swin=ref.swin, win_ai=ref.win_ai, win_af=ref.win_af,
            W=ref.W, W_ai=ref.W_ai, W_af=ref.W_af, X=X, xrec=ref.istft(X), Sc=Sc, thr_mid=thr_mid,
            consistency=np.float64(ref.get_consistency(Sc)),
            batch_zero=ref.batch_lws(A, thresholds=np.zeros(5)),
            batch_mid=ref.batch_lws(A, thresholds=thr_mid),
            batch_cplx=ref.batch_lws(Sc, thresholds=np.zeros(3)),
            nofuture_zero=ref.nofuture_lws(A, thresholds=np.zeros(2)),
            nofuture_def=ref.nofuture_lws(A),
            nofuture_cplx=ref.nofuture_lws(Sc, thresholds=np.array([0.5, 0.1])),
            online_def=ref.online_lws(A, iterations=3),
            online_zero=ref.online_lws(A, thresholds=np.zeros(2)),
            online_cplx=ref.online_lws(Sc, iterations=2),
            run=lws_ref.lws(*args, mode="music", batch_iterations=8, batch_alpha=1.0, **kw).run_lws(A),
        )
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
        cases.append(dict(name=name, args=list(args), kwargs=kw, signal=kind, n=n, seed=seed,
                          T=int(A.shape[0]), Nreal=int(A.shape[1])))

    # user-supplied (non-default) windows: changes the weight sparsity pattern
    fsize, hop = 64, 16
    awin = np.hamming(fsize) ** 0.5
    swin = np.hanning(fsize + 2)[1:-1]
    ref = lws_ref.lws(awin, hop, swin=swin, mode="music")
    x = white(77, 1200)
    X = ref.stft(x)
    A = np.abs(X)
    np.savez_compressed(
        os.path.join(OUT, "custom_win.npz"), x=x, awin_in=awin, swin_in=swin, awin=ref.awin, swin=ref.swin,
        W=ref.W, W_ai=ref.W_ai, W_af=ref.W_af, X=X, xrec=ref.istft(X),
        batch_zero=ref.batch_lws(A, thresholds=np.zeros(4)), online_def=ref.online_lws(A, iterations=2),
        nofuture_def=ref.nofuture_lws(A), run=lws_ref.lws(awin, hop, swin=swin, mode="music", batch_iterations=6,
                                                         batch_alpha=1.0).run_lws(A))
    cases.append(dict(name="custom_win", args=[fsize, hop], kwargs={}, signal="white", n=1200, seed=77,
                      T=int(A.shape[0]), Nreal=int(A.shape[1])))

    # the reference's headline CPU case (BASELINE.json configs[0]) at reduced length: 512/128, default ctor
    ref = lws_ref.lws(512, 128)
    x = tonal(5, 6000)
    X = ref.stft(x)
    A = np.abs(X)
    np.savez_compressed(os.path.join(OUT, "cfg1_short.npz"), x=x, X=X, W=ref.W,
                        batch_def=ref.batch_lws(A), run_music=lws_ref.lws(512, 128, mode="music").run_lws(A))
    cases.append(dict(name="cfg1_short", args=[512, 128], kwargs={}, signal="tonal", n=6000, seed=5,
                      T=int(A.shape[0]), Nreal=int(A.shape[1])))

    # the *fractionalQ paths (frame shift not dividing the frame size / use_simplifications=False): per-frequency weight
    # rows.  The compiled module reads one row past its table at the DC bin (undefined: its output varies from call to
    # call), so these vectors come from the reference's C functions driven on tables with a zero row appended
    # (tests/conftest.py::ref_fractional); windows, weights and transforms are the module's.
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import ref_fractional
    frac = []
    for name, args, kw, kind, n in [("frac_64_20", (64, 20), {}, "white", 1500), ("frac_nosimp_32_8", (32, 8), {"use_simplifications": False}, "tonal", 600),
                                    ("frac_48_10", (48, 10), {"look_ahead": 2}, "white", 900)]:
        seed = 200 + len(frac)
        x = white(seed, n) if kind == "white" else tonal(seed, n)
        ref = lws_ref.lws(*args, mode="music", **kw)
        X = ref.stft(x)
        A = np.abs(X)
        rng = np.random.default_rng(100 + seed)
        Sc = A * np.exp(1j * rng.uniform(0, 2 * np.pi, A.shape))
        thr_mid = lws_ref.get_thresholds(6, 2.0, 0.4, 1)
        on_thr = lws_ref.get_thresholds(3, 1, 0.1, 1)
        nf = ref_fractional(lws_ref, "nofuture", A, ref.W_ai, lws_ref.get_thresholds(1, 1, 0.1, 1))
        on = ref_fractional(lws_ref, "online", nf, ref.W, lws_ref.get_thresholds(10, 1, 0.1, 1), ref.W_ai, ref.W_af, ref.look_ahead)
        d = dict(x=x, awin=ref.awin, swin=ref.swin, W=ref.W, W_ai=ref.W_ai, W_af=ref.W_af, X=X, xrec=ref.istft(X), Sc=Sc, thr_mid=thr_mid,
                 batch_zero=ref_fractional(lws_ref, "batch", A, ref.W, np.zeros(5)),
                 batch_mid=ref_fractional(lws_ref, "batch", A, ref.W, thr_mid),
                 batch_cplx=ref_fractional(lws_ref, "batch", Sc, ref.W, np.zeros(3)),
                 nofuture_def=nf,
                 nofuture_zero=ref_fractional(lws_ref, "nofuture", A, ref.W_ai, np.zeros(2)),
                 online_def=ref_fractional(lws_ref, "online", A, ref.W, on_thr, ref.W_ai, ref.W_af, ref.look_ahead),
                 online_zero=ref_fractional(lws_ref, "online", Sc, ref.W, np.zeros(2), ref.W_ai, ref.W_af, ref.look_ahead),
                 run=ref_fractional(lws_ref, "batch", on, ref.W, lws_ref.get_thresholds(8, 1.0, 0.1, 1)))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
        frac.append(dict(name=name, args=list(args), kwargs=kw, signal=kind, n=n, seed=seed, T=int(A.shape[0]), Nreal=int(A.shape[1])))

    with open(os.path.join(OUT, "cases.json"), "w") as f:
        json.dump(dict(generator="tools/make_golden.py", reference="Jonathan-LeRoux/lws v%s" % lws_ref.__version__,
                       numpy=np.__version__, cases=cases, fractional_cases=frac), f, indent=1)
    tot = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print("wrote %d cases, %.2f MB" % (len(cases), tot / 1e6))


if __name__ == "__main__":
    main()

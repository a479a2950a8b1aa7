"""GPU parity tests: the CUDA path (through the C-ABI, via the reference-API mirror
``lws_b200``) against (1) the committed golden vectors generated from the compiled reference
and (2) the CPU oracle on fresh seeded inputs.

Parity bar: BIT-EXACT (np.array_equal).  north_star asks for <= 1e-5 relative, but the LWS
iteration amplifies a one-ulp difference into 1e-2 within ~50 sweeps on ordinary inputs
(unstable symmetric configurations kept alive by exact cancellations), so the
only parity that holds at every size is the reference's own operation sequence; the CUDA
path restates it (csrc/exact.cuh) and the tests demand identical bits.
"""
import numpy as np
import pytest

from conftest import CASES, SMALL_CASES, golden, relF, make_signal

pytestmark = pytest.mark.gpu

TOL = 1e-9
NAMES = [c["name"] for c in SMALL_CASES]


@pytest.fixture(scope="module")
def gpu():
    import lws_b200
    from lws_b200 import _native
    _native.lib()
    ctx = _native.Context(0)  # raises when no CUDA device: there is no fallback to test
    info = ctx.device_info()
    assert info["cc"][0] >= 10, "expected a Blackwell (sm_100) device, got cc %s" % (info["cc"],)
    ctx.close()
    return lws_b200


def _experiments():
    from lws_b200 import _native
    return _native.lib().lwsb_has_experiments() == 1


needs_experiments = pytest.mark.skipif("not __import__('lws_b200')._native.lib().lwsb_has_experiments()",
                                       reason="kernel variants built only with -DLWSB_EXPERIMENTS (LWSB_NVCC_EXTRA)")


def _ctor(mod, case, **extra):
    kw = dict(case["kwargs"])
    kw.update(extra)
    if case["name"] == "custom_win":
        g = golden("custom_win")
        return mod.lws(g["awin_in"], case["args"][1], swin=g["swin_in"], **kw)
    return mod.lws(*case["args"], **kw)


def _close(y, yref, what, tol=0.0):
    assert y.shape == yref.shape and y.dtype == np.complex128, what
    if tol == 0.0 and np.array_equal(y, yref):
        return
    e = relF(y, yref)
    assert e <= tol, "%s: relF %.3e, %d of %d bins differ (bit-exact expected)" % (
        what, e, int((y != yref).sum()), y.size) if tol == 0.0 else "%s: relF %.3e > %.1e" % (what, e, tol)


@pytest.mark.parametrize("case", SMALL_CASES, ids=NAMES)
def test_golden_sweeps(gpu, case, capsys):
    g = golden(case["name"])
    p = _ctor(gpu, case, mode="music")
    for k in ("W", "W_ai", "W_af"):
        assert np.array_equal(getattr(p, k), g[k]), k
    A = np.abs(g["X"])
    z = np.zeros
    full = "Sc" in g
    checks = {
        "batch_zero": lambda: p.batch_lws(A, thresholds=z(5 if full else 4)),
        "nofuture_def": lambda: p.nofuture_lws(A),
        "online_def": lambda: p.online_lws(A, iterations=3 if full else 2),
    }
    if full:
        Sc = g["Sc"]
        checks.update({
            "batch_mid": lambda: p.batch_lws(A, thresholds=g["thr_mid"]),
            "batch_cplx": lambda: p.batch_lws(Sc, thresholds=z(3)),
            "nofuture_zero": lambda: p.nofuture_lws(A, thresholds=z(2)),
            "nofuture_cplx": lambda: p.nofuture_lws(Sc, thresholds=np.array([0.5, 0.1])),
            "online_zero": lambda: p.online_lws(A, thresholds=z(2)),
            "online_cplx": lambda: p.online_lws(Sc, iterations=2),
        })
    for k, fn in checks.items():
        _close(fn(), g[k], case["name"] + ":" + k)
    nb = 8 if full else 6
    _close(_ctor(gpu, case, mode="music", batch_iterations=nb, batch_alpha=1.0).run_lws(A), g["run"],
           case["name"] + ":run")


def test_golden_cfg1_short(gpu):
    g = golden("cfg1_short")
    A = np.abs(g["X"])
    _close(gpu.lws(512, 128).batch_lws(A), g["batch_def"], "cfg1_short:batch_def")
    _close(gpu.lws(512, 128, mode="music").run_lws(A), g["run_music"], "cfg1_short:run_music")


@pytest.mark.parametrize("case", SMALL_CASES, ids=NAMES)
def test_golden_transforms(gpu, case, capsys):
    g = golden(case["name"])
    p = _ctor(gpu, case, mode="music")
    X = p.stft(g["x"])
    assert X.shape == g["X"].shape
    assert np.abs(X - g["X"]).max() <= 1e-12 * max(1.0, np.abs(g["X"]).max())
    xr = p.istft(g["X"])
    assert xr.shape == g["xrec"].shape
    assert np.abs(xr - g["xrec"]).max() <= 1e-12 * max(1.0, np.abs(g["xrec"]).max())
    if "consistency" in g:
        assert abs(p.get_consistency(g["Sc"]) - float(g["consistency"])) < 1e-6


from conftest import FRAC_CASES  # noqa: E402


@pytest.mark.parametrize("case", FRAC_CASES, ids=[c["name"] for c in FRAC_CASES])
def test_fractional_paths_match_golden(gpu, case):
    """hop not dividing the frame size / use_simplifications=False (LWSfractionalQ, NoFuture_LWSfractionalQ,
    Asym_UpdatePhasefractionalQ: per-frequency weight rows): golden vectors from the reference's C functions on tables with
    the zero row N that the reference reads out of bounds at the DC bin (tests/conftest.py::ref_fractional)."""
    g = golden(case["name"])
    p = gpu.lws(*case["args"], mode="music", **case["kwargs"])
    for k in ("W", "W_ai", "W_af"):
        assert np.array_equal(getattr(p, k), g[k]), k
    A = np.abs(g["X"])
    checks = {
        "batch_zero": lambda: p.batch_lws(A, thresholds=np.zeros(5)),
        "batch_mid": lambda: p.batch_lws(A, thresholds=g["thr_mid"]),
        "batch_cplx": lambda: p.batch_lws(g["Sc"], thresholds=np.zeros(3)),
        "nofuture_def": lambda: p.nofuture_lws(A),
        "nofuture_zero": lambda: p.nofuture_lws(A, thresholds=np.zeros(2)),
        "online_def": lambda: p.online_lws(A, iterations=3),
        "online_zero": lambda: p.online_lws(g["Sc"], thresholds=np.zeros(2)),
    }
    for k, fn in checks.items():
        _close(fn(), g[k], case["name"] + ":" + k)
    pr = gpu.lws(*case["args"], mode="music", batch_iterations=8, batch_alpha=1.0, **case["kwargs"])
    _close(pr.run_lws(A), g["run"], case["name"] + ":run")
    X = p.stft(g["x"])
    assert np.abs(X - g["X"]).max() <= 1e-12 * max(1.0, np.abs(g["X"]).max())


@pytest.mark.parametrize("fs,hop,kw,n", [(512, 100, {}, 16000), (256, 64, {"use_simplifications": False}, 9000), (200, 48, {"look_ahead": 1}, 7000)])
def test_fractional_paths_vs_oracle(gpu, oracle, fs, hop, kw, n):
    """The reference's README case lws.lws(512, 100) (Q = 5.12, W of shape (512, 6, 6)) and friends against the oracle,
    single and ragged batch."""
    po, pg = oracle.lws(fs, hop, mode="music", batch_iterations=20, batch_alpha=2, **kw), gpu.lws(fs, hop, mode="music", batch_iterations=20, batch_alpha=2, **kw)
    assert pg.W.shape[0] == fs and np.array_equal(pg.W, po.W)
    As = [np.abs(po.stft(make_signal(k, 41 + i, n - 900 * i))) for i, k in enumerate(("tonal", "white"))]
    for A, Y in zip(As, pg.batch_lws(As)):
        _close(Y, po.batch_lws(A), "fractional batch")
    for A, Y in zip(As, pg.run_lws(As)):
        _close(Y, po.run_lws(A), "fractional run_lws")
    _close(pg.online_lws(As[1]), po.online_lws(As[1]), "fractional online")


def test_forced_anyq_equals_folded(gpu):
    """Q2 / Q4 shortcuts vs the anyQ formulas: equal to ~1e-10."""
    from lws_b200 import _native
    for args in ((32, 16), (32, 8)):
        p = gpu.lws(*args)
        A = np.abs(p.stft(make_signal("white", 11, 700)))
        a = gpu.batch_lws(A, p.W, np.zeros(5))
        b = gpu.batch_lws(A, p.W, np.zeros(5), flags=_native.FORCE_ANYQ)
        assert relF(a, b) < 1e-8


@pytest.mark.parametrize("fs,hop,kind,n", [(512, 128, "white", 32000), (512, 128, "tonal", 32000),
                                           (256, 32, "tonal", 12000), (128, 64, "white", 20000)])
def test_vs_oracle_medium(gpu, oracle, fs, hop, kind, n):
    """BASELINE.json configs[0] (2 s at 16 kHz, 512/128, 100 default iterations) and friends."""
    x = make_signal(kind, 42, n)
    po, pg = oracle.lws(fs, hop, mode="music"), gpu.lws(fs, hop, mode="music")
    A = np.abs(po.stft(x))
    _close(pg.batch_lws(A), po.batch_lws(A), "batch")
    _close(pg.online_lws(A), po.online_lws(A), "online")
    _close(pg.nofuture_lws(A), po.nofuture_lws(A), "nofuture")
    _close(pg.run_lws(A), po.run_lws(A), "run")


def test_batched_ragged_equals_single(gpu, oracle):
    """The utterance batch is an extension: every member must equal its own single call."""
    po, pg = oracle.lws(64, 16, mode="music", batch_iterations=12), gpu.lws(64, 16, mode="music", batch_iterations=12)
    lens = [900, 2500, 64, 1300, 5000, 130, 3100]
    As = [np.abs(po.stft(make_signal("white" if i % 2 else "tonal", 100 + i, n))) for i, n in enumerate(lens)]
    outs = pg.run_lws(As)
    assert isinstance(outs, list) and len(outs) == len(As)
    for A, Y in zip(As, outs):
        _close(Y, po.run_lws(A), "ragged run_lws T=%d" % A.shape[0])
    outs = pg.batch_lws(As)
    for A, Y in zip(As, outs):
        _close(Y, po.batch_lws(A), "ragged batch_lws T=%d" % A.shape[0])
    B3 = np.stack([As[1][:50], As[4][:50], As[6][:50]])
    Y3 = pg.batch_lws(B3)
    assert Y3.shape == B3.shape
    for b in range(3):
        _close(Y3[b], po.batch_lws(B3[b]), "3-D batch member %d" % b)


def test_tiny_inputs(gpu, oracle):
    """Edge cases: fewer frames than Q, a single frame, thresholds that deactivate everything."""
    po, pg = oracle.lws(32, 8, mode="music"), gpu.lws(32, 8, mode="music")
    rng = np.random.default_rng(5)
    for T in (1, 2, 3, 4, 7):
        A = np.abs(rng.standard_normal((T, 17)))
        _close(pg.batch_lws(A, iterations=6), po.batch_lws(A, iterations=6), "batch T=%d" % T)
        _close(pg.online_lws(A), po.online_lws(A), "online T=%d" % T)
        _close(pg.nofuture_lws(A), po.nofuture_lws(A), "nofuture T=%d" % T)
        _close(pg.run_lws(A), po.run_lws(A), "run T=%d" % T)
    A = np.abs(rng.standard_normal((9, 17)))
    Y = pg.batch_lws(A, thresholds=np.full(3, 1e9))
    assert np.array_equal(Y, A.astype(np.complex128))
    Z = np.zeros((5, 17))
    assert np.array_equal(pg.batch_lws(Z, iterations=3), Z.astype(np.complex128))


def test_api_behaviours(gpu):
    """SURVEY.md section 9.9 on the CUDA path."""
    p = gpu.lws(32, 8)
    x = np.random.default_rng(3).standard_normal(300)
    A = np.abs(p.stft(x))
    A0 = A.copy()
    Y = p.batch_lws(A, thresholds=np.zeros(2))
    assert Y.dtype == np.complex128 and Y.flags.c_contiguous and np.array_equal(A, A0)
    assert np.allclose(np.abs(Y), A, rtol=1e-12, atol=1e-14)
    Sc = A.astype(np.complex128)
    assert p.batch_lws(Sc, iterations=0) is Sc
    with pytest.raises(ValueError):
        p.batch_lws(A[:, :-1], iterations=1)
    Y32 = p.batch_lws(A.astype(np.float32), thresholds=np.zeros(2))
    assert Y32.dtype == np.complex128
    xx = np.random.default_rng(4).standard_normal(1000)
    assert np.abs(p.istft(p.stft(xx))[:1000] - xx).max() < 1e-13
    # a synthesis window shorter than the frame is zero-padded up to fftsize (lws.pyx:111-112)
    import lws_b200 as mod
    Xs = p.stft(xx)
    fs = 2 * (Xs.shape[1] - 1)
    a = mod.istft(Xs, p.fshift, p.swin[: fs // 2])
    b = mod.istft(Xs, p.fshift, np.hstack([p.swin[: fs // 2], np.zeros(fs - fs // 2)]), fftsize=fs)
    assert np.array_equal(a, b) and np.abs(a).max() > 0
    assert p.get_consistency(p.stft(xx)) > 250.0
    assert gpu.lws(512, 100).W.shape == (512, 6, 6)  # per-frequency weights: the *fractionalQ path (test_fractional_*)


def test_pageable_buffers_are_staged_in_chunks(gpu):
    """Pageable host arrays go through the library's two pinned staging buffers in 24 MB chunks (host threads copying one
    chunk while the DMA engine moves the other): ~100 MB each way, chunk boundaries inside utterances; with thresholds no
    bin exceeds, batch_lws is the identity, so every byte must come back.  Results are fresh arrays (page-locked memory
    from the library's pool when large), never aliases of the input."""
    rng = np.random.default_rng(12)
    p = gpu.lws(1024, 256)
    S = rng.standard_normal((20, 628, 513)) + 1j * rng.standard_normal((20, 628, 513))
    thr = np.full(2, 1e9)
    Y = p.batch_lws(S, thresholds=thr)
    assert Y is not S and Y.dtype == np.complex128 and Y.flags.c_contiguous and np.array_equal(Y, S)
    Ys = p.batch_lws([S[b, : 100 + 25 * b] for b in range(20)], thresholds=thr)   # ragged list, pageable views
    assert all(np.array_equal(Ys[b], S[b, : 100 + 25 * b]) for b in range(20))
    Y2 = p.batch_lws(S, thresholds=thr)
    assert np.array_equal(Y2, S) and not np.shares_memory(Y2, Y)                   # a second result never reuses a live one
    out = np.empty_like(S)
    assert p.batch_lws(S, thresholds=thr, out=out) is out and np.array_equal(out, S)
    del Y, Y2
    Y3 = p.batch_lws(S[:3], thresholds=np.zeros(2))                               # pooled blocks are reused safely
    assert np.array_equal(Y3, p.batch_lws(np.ascontiguousarray(S[:3]), thresholds=np.zeros(2)))


def test_results_leave_while_the_kernel_runs(gpu, oracle):
    """One-shot batch_lws with page-locked result buffers: the utterances are worked on in groups and every utterance is copied
    out by the DMA engine as soon as its last pass has written its frames back, while later groups are still in flight.
    Same bits as with the copy after the kernel (LWSB_EARLY_STORE=0) and as the oracle; ragged batch of 40."""
    import os
    po, pg = oracle.lws(512, 128), gpu.lws(512, 128)
    rng = np.random.default_rng(23)
    As = [np.abs(po.stft(make_signal("white" if i % 3 else "tonal", 500 + i, int(rng.integers(30000, 60000))))) for i in range(40)]
    thr = gpu.get_thresholds(24, 3.0, 0.12, 1)   # some utterances drop leading sweeps, pass counts differ
    Ys = pg.batch_lws(As, thresholds=thr)
    os.environ["LWSB_EARLY_STORE"] = "0"
    try:
        Yl = pg.batch_lws(As, thresholds=thr)
    finally:
        del os.environ["LWSB_EARLY_STORE"]
    for i in range(40):
        assert np.array_equal(Ys[i], Yl[i]), i
    for i in (0, 7, 19, 39):
        _close(Ys[i], po.batch_lws(As[i], thresholds=thr), "early store, utterance %d" % i)
    big = np.stack([A[:230] for A in As])      # 3-D batch, one pinned block, thresholds nobody exceeds for half the sweeps
    Y3 = pg.batch_lws(big, thresholds=np.concatenate([np.full(3, 1e9), thr[:6]]))
    assert np.array_equal(Y3[5], pg.batch_lws(big[5], thresholds=thr[:6]))


def test_full_size_properties(gpu):
    """BASELINE.json configs[1] shape (628 x 513, Q = 4, 100 default iterations), 4 utterances:
    size-independent properties instead of an oracle run -- magnitudes preserved, result
    independent of batching, consistency improves by a wide margin."""
    p = gpu.lws(1024, 256)
    xs = np.stack([make_signal("white" if b % 2 else "tonal", 2000 + b, 160000) for b in range(4)])
    A = np.abs(p.stft(xs))
    assert A.shape == (4, 628, 513)
    Y = p.batch_lws(A)
    assert np.allclose(np.abs(Y), A, rtol=1e-11, atol=1e-12 * A.max())
    assert np.array_equal(p.batch_lws(A[2]), Y[2])
    c0 = p.get_consistency(A[0].astype(np.complex128))
    c1 = p.get_consistency(Y[0])
    assert c1 > c0 + 5.0, (c0, c1)


def test_strip_kernel_cooperative_launch(gpu, monkeypatch):
    """The clusters of a strip-kernel launch wait for one another.  LWSB_STRIP_COOP=1 makes the launch cooperative: the driver
    guarantees that all of them are resident (or refuses the launch) also when other work shares the GPU.  The default is the
    plain launch with the grid sized by the occupancy query (Nsight Compute cannot replay cooperative cluster launches)."""
    from lws_b200 import _native
    p = gpu.lws(512, 128, batch_iterations=6, batch_alpha=1)
    A = np.abs(p.stft(make_signal("white", 8, 9000)))
    Y = p.batch_lws(A)
    assert _native.lib().lwsb_strip_launch_mode() == 0
    monkeypatch.setenv("LWSB_STRIP_COOP", "1")
    assert np.array_equal(p.batch_lws([A, A[:40]])[0], Y)
    assert _native.lib().lwsb_strip_launch_mode() == 1


def test_cfg2_one_utterance_vs_oracle(gpu, oracle):
    """One utterance of configs[1] against the oracle at full size and 100 iterations (~1.5 s CPU)."""
    po, pg = oracle.lws(1024, 256), gpu.lws(1024, 256)
    A = np.abs(po.stft(make_signal("tonal", 2002, 160000)))
    _close(pg.batch_lws(A), po.batch_lws(A), "cfg2 batch")


def test_cfg3_cfg4_one_utterance_vs_oracle(gpu, oracle):
    """One utterance of configs[2] / configs[3] at full size (628 x 513, mode='music': NoFuture_LWSQ4, TF-RTISI-LA with
    look-ahead 3 and 10 iterations -- 25 688 row updates -- and the full run_lws chain) against the oracle (~3 s CPU)."""
    po, pg = oracle.lws(1024, 256, mode="music"), gpu.lws(1024, 256, mode="music")
    A = np.abs(po.stft(make_signal("white", 3003, 160000)))
    assert A.shape == (628, 513)
    _close(pg.nofuture_lws(A), po.nofuture_lws(A), "cfg3 nofuture")
    _close(pg.online_lws(A), po.online_lws(A), "cfg3 online")
    from lws_b200 import api
    assert api._context(0).last_online_kernel() == 4
    _close(pg.run_lws(A), po.run_lws(A), "cfg4 run_lws")
    # a batch of three: every member equals the single-utterance result
    Ys = pg.online_lws(np.stack([A, A[::-1], A]))
    assert np.array_equal(Ys[0], pg.online_lws(A)) and np.array_equal(Ys[0], Ys[2])


@pytest.mark.parametrize("fs,hop,la,its,n", [(1024, 256, 3, 10, 40000), (512, 128, 3, 4, 9000), (512, 128, 5, 3, 9000), (512, 128, 1, 5, 5000),
                                             (512, 128, 0, 6, 5000), (64, 16, 3, 10, 4000), (128, 64, 1, 7, 6000), (2048, 256, 3, 3, 30000)])
def test_online_ring_kernels(gpu, oracle, fs, hop, la, its, n):
    """The shared-memory ring kernel of online_lws (K warps per task taking turns) and the generic fallback against the
    oracle over look-aheads, iteration counts and ragged batches."""
    from lws_b200 import api
    ctx = api._context(0)
    kw = dict(look_ahead=la, online_iterations=its)
    po, pg = oracle.lws(fs, hop, **kw), gpu.lws(fs, hop, **kw)
    As = [np.abs(po.stft(make_signal(k, 31 + i, n + 700 * i))) for i, k in enumerate(("tonal", "white", "tonal"))]
    for thr in (None, np.zeros(its)):
        Ys = pg.online_lws(As, thresholds=thr)
        want = 0 if fs // hop > 4 else 4  # Q = 8 at 1025 bins: ring too large for shared memory, generic kernel
        assert ctx.last_online_kernel() == want, ctx.last_online_kernel()
        for A, Y in zip(As, Ys):
            _close(Y, po.online_lws(A, thresholds=thr), "online ring kernel, LA=%d" % la)
    for T in (1, 2, 3, 5):  # fewer frames than look-ahead + 1, a single frame
        _close(pg.online_lws(As[0][:T]), po.online_lws(As[0][:T]), "online T=%d" % T)


def test_nofuture_q4_shared_memory_kernel(gpu, oracle, monkeypatch):
    """NoFuture_LWSQ4 (the reference's doubled bin offset included) with the frames in a shared-memory window: a ragged batch,
    several sweeps, against the oracle and against the global-memory kernel (LWSB_NOFUTURE_RING=0)."""
    for fs, hop, n in ((512, 128, 9000), (64, 16, 3000), (1024, 256, 30000)):
        po, pg = oracle.lws(fs, hop, mode="music"), gpu.lws(fs, hop, mode="music")
        As = [np.abs(po.stft(make_signal(k, 5 + i, n + 517 * i))) for i, k in enumerate(("tonal", "white", "tonal"))]
        for thr in (None, np.zeros(3), np.array([0.7, 0.2])):
            Ys = pg.nofuture_lws(As, thresholds=thr)
            for A, Y in zip(As, Ys):
                _close(Y, po.nofuture_lws(A, thresholds=thr), "nofuture Q4 ring %d/%d" % (fs, hop))
            monkeypatch.setenv("LWSB_NOFUTURE_RING", "0")
            Yg = pg.nofuture_lws(As, thresholds=thr)
            monkeypatch.delenv("LWSB_NOFUTURE_RING")
            assert all(np.array_equal(a, b) for a, b in zip(Ys, Yg))
        for T in (1, 2, 4):
            _close(pg.nofuture_lws(As[0][:T]), po.nofuture_lws(As[0][:T]), "nofuture ring T=%d" % T)


_RAIL = lambda lag: pytest.param({"LWSB_ONLINE_RAIL": "1", "LWSB_ONLINE_RAIL_S": lag}, 5, 4, marks=needs_experiments)


@pytest.mark.parametrize("env,want,want_any", [({"LWSB_ONLINE_FLOW": "0"}, 3, 3), ({"LWSB_ONLINE_FLOW": "2"}, 4, 4), ({"LWSB_ONLINE_FLOW": "3"}, 4, 4),
                                               ({"LWSB_ONLINE_FLOW_S": "10"}, 4, 4), ({"LWSB_ONLINE_FLOW_S": "13"}, 4, 4),
                                               _RAIL("7"), _RAIL("9"), _RAIL("12")])
@pytest.mark.parametrize("fs,hop,la,its", [(512, 128, 3, 4), (128, 64, 2, 5)])
def test_online_kernel_choices(gpu, oracle, monkeypatch, env, want, want_any, fs, hop, la, its):
    """The other shapes of the online chain kernel -- two lanes per task, 2 or 3 warps per task taking turns, longer lags between
    row updates, and (experiments build) value warps + chain warps -- give the same bits."""
    from lws_b200 import api
    for k, val in env.items():
        monkeypatch.setenv(k, val)
    kw = dict(look_ahead=la, online_iterations=its)
    po, pg = oracle.lws(fs, hop, **kw), gpu.lws(fs, hop, **kw)
    As = [np.abs(po.stft(make_signal(k, 77 + i, 7000 + 900 * i))) for i, k in enumerate(("white", "tonal"))]
    for thr in (None, np.zeros(its)):
        Ys = pg.online_lws(As, thresholds=thr)
        assert api._context(0).last_online_kernel() == want
        for A, Y in zip(As, Ys):
            _close(Y, po.online_lws(A, thresholds=thr), "online kernel %s" % env)
    # the anyQ formulas (LWSB_FORCE_ANYQ) against the oracle made to take its anyQ branch
    from lws_b200 import _native
    thr = np.ones(its)
    Yg = gpu.online_lws(As[0], pg.W, pg.W_ai, pg.W_af, thr, la, hop, flags=_native.FORCE_ANYQ)
    assert api._context(0).last_online_kernel() == want_any
    monkeypatch.setattr(oracle, "_fold", lambda Q, Qprime, simp, Nreal: 0)
    _close(Yg, oracle.online_lws(As[0], po.W, po.W_ai, po.W_af, thr, la, hop), "online anyQ formulas, %s" % env)


@pytest.mark.parametrize("fs,hop,kw,chunks", [(512, 128, {}, (1, 2, 3, 5, 8, 13, 40)), (64, 16, {"look_ahead": 0}, (7,)), (64, 8, {"look_ahead": 2}, (1, 30)),
                                              (128, 64, {"look_ahead": 5}, (4, 1, 1, 9)), (64, 20, {}, (3, 11))])
def test_streaming_online_equals_offline(gpu, oracle, fs, hop, kw, chunks):
    """Frame-in / frame-out online_lws (lwsb_stream_*): frames pushed in chunks of any size, every frame returned as soon as it is
    final (look_ahead frames behind the newest); given the utterance's mean amplitude the concatenation is bit-identical
    to online_lws on the complete spectrogram -- oracle and offline CUDA path alike (incl. the per-frequency-weights path)."""
    po, pg = oracle.lws(fs, hop, mode="music", **kw), gpu.lws(fs, hop, mode="music", **kw)
    A = np.abs(po.stft(make_signal("tonal", 61, 9000)))
    T = A.shape[0]
    want = po.online_lws(A)
    st = pg.online_stream(A.shape[1], max_frames=T, mean_amp=float(np.mean(A)))
    got, m, i, lat = [], 0, 0, []
    while m < T:
        n = min(chunks[i % len(chunks)], T - m)
        out = st.push(A[m:m + n])
        m += n; i += 1
        got.append(out)
        lat.append(m - sum(len(g) for g in got))
    assert max(lat) <= pg.look_ahead and all(len(g) <= max(chunks) + pg.look_ahead for g in got)
    got.append(st.close())
    Y = np.concatenate(got)
    _close(Y, want, "streamed online_lws")
    assert np.array_equal(Y, pg.online_lws(A))
    # complex input, and an estimate of the mean instead of the true one: still a valid reconstruction of the magnitudes
    st = pg.online_stream(A.shape[1], max_frames=T, mean_amp=1.1 * float(np.mean(A)), complex_input=True)
    Z = np.concatenate([st.push(A[:T // 2].astype(np.complex128)), st.push(A[T // 2:].astype(np.complex128)), st.close()])
    assert Z.shape == want.shape and np.allclose(np.abs(Z), A, rtol=1e-10, atol=1e-12 * A.max())
    with pytest.raises(ValueError):
        pg.online_stream(A.shape[1], max_frames=4, mean_amp=1.0).push(A[:5])


def test_q8_large_frame_vs_oracle(gpu, oracle):
    """configs[4] geometry (2048-pt, hop 256, Q = 8) at reduced length and 10 iterations."""
    po, pg = oracle.lws(2048, 256), gpu.lws(2048, 256)
    A = np.abs(po.stft(make_signal("tonal", 7, 40000)))
    thr = gpu.get_thresholds(10, 2.0, 0.3, 1)
    _close(pg.batch_lws(A, thresholds=thr), po.batch_lws(A, thresholds=thr), "q8 batch")


def test_no_fallback_and_launch_counter(gpu):
    from lws_b200 import _native
    ctx = _native.Context(0)
    p = gpu.lws(32, 8)
    ctx.set_weights(_native.W, p.W)
    A = np.abs(np.random.default_rng(0).standard_normal((20, 17)))
    n0 = ctx.launch_count()
    ctx.batch_lws([A], _native.F64, np.zeros(3))
    assert ctx.launch_count() - n0 >= 3  # extend + stats, sweeps, crop
    assert ctx.last_compute_ms() > 0.0
    ctx.close()


# ------------------------------------------------------------------ cluster strip kernel
STRIP_CASES = [  # fsize, hop, samples, iterations, cluster, sweeps per pass, smem budget
    (512, 128, 6000, 6, 1, 0, 0), (512, 128, 6000, 6, 2, 0, 0), (512, 128, 6000, 6, 4, 0, 0),
    (512, 128, 6000, 7, 2, 3, 0), (512, 128, 6000, 7, 4, 2, 60000), (512, 128, 700, 5, 2, 2, 0),
    (1024, 256, 30000, 12, 2, 0, 0), (1024, 256, 30000, 12, 4, 0, 0), (1024, 256, 30000, 12, 8, 0, 0),
    (1024, 256, 30000, 9, 8, 2, 0), (2048, 256, 40000, 8, 4, 0, 0), (2048, 256, 40000, 8, 8, 0, 0),
    (128, 64, 9000, 10, 1, 0, 0), (128, 64, 9000, 10, 2, 3, 0), (512, 128, 32000, 100, 0, 0, 0),
    # one sweep per pass: as many passes as sweeps, all in flight on different clusters, each reading what the previous one
    # wrote a few frames earlier (the tightest use of the progress counters)
    (512, 128, 32000, 40, 2, 1, 0), (1024, 256, 60000, 24, 4, 1, 0), (512, 128, 32000, 30, 1, 1, 0),
]


@pytest.mark.parametrize("fs,hop,n,its,cluster,sweeps,smem", STRIP_CASES,
                         ids=["%d-%d-n%d-it%d-C%d-G%d-S%d" % c for c in STRIP_CASES])
def test_strip_kernel_bit_exact(gpu, oracle, fs, hop, n, its, cluster, sweeps, smem):
    """The cluster strip kernel (TMA ring, DSMEM halos) over cluster sizes, pass counts and ring
    sizes: identical bits to the oracle, and the plan asked for is the plan that ran."""
    from lws_b200 import api
    ctx = api._context(0)
    po, pg = oracle.lws(fs, hop), gpu.lws(fs, hop)
    A = np.abs(po.stft(make_signal("tonal", 3, n)))
    try:
        ctx.set_tuning(smem, cluster, sweeps)
        for thr in (np.zeros(its), gpu.get_thresholds(its, 100 if its > 50 else 2.0, 0.1, 1)):
            Y = pg.batch_lws(A, thresholds=thr)
            plan = ctx.last_batch_plan()
            assert plan is not None, "strip kernel did not run"
            if cluster:
                assert plan["cluster"] == cluster
            if sweeps:
                assert plan["sweeps_per_pass"] <= sweeps
            _close(Y, po.batch_lws(A, thresholds=thr), "strips %s" % (plan,))
    finally:
        ctx.set_tuning(0, 0, 0)


def test_strip_kernel_ragged_batch_persistent_clusters(gpu, oracle):
    """More utterances than resident clusters, different lengths: the persistent cluster loop."""
    from lws_b200 import api
    ctx = api._context(0)
    po, pg = oracle.lws(512, 128), gpu.lws(512, 128)
    As = [np.abs(po.stft(make_signal("white" if i % 2 else "tonal", 50 + i, 3000 + 700 * (i % 9)))) for i in range(45)]
    thr = gpu.get_thresholds(9, 2.0, 0.2, 1)
    want = [po.batch_lws(A, thresholds=thr) for A in As]
    try:
        for cl in (4, 2, 0):
            ctx.set_tuning(0, cl, 0)
            Ys = pg.batch_lws(As, thresholds=thr)
            assert ctx.last_batch_plan() is not None
            for i, (Y, W) in enumerate(zip(Ys, want)):
                _close(Y, W, "ragged member %d, cluster %d" % (i, cl))
    finally:
        ctx.set_tuning(0, 0, 0)


DUO_CASES = [  # fsize, hop, samples, iterations, cluster, sweeps per pass, sweep lag: many warp pairs, every register tier
    (1024, 256, 60000, 30, 2, 7, 0), (1024, 256, 60000, 30, 2, 6, 5), (1024, 256, 60000, 40, 4, 16, 0), (1024, 256, 60000, 40, 4, 20, 0),
    (1024, 256, 60000, 40, 4, 24, 0), (1024, 256, 60000, 40, 8, 32, 0), (512, 128, 32000, 100, 0, 0, 0), (128, 64, 9000, 30, 1, 0, 0),
    (512, 128, 700, 5, 1, 0, 0),
]


@needs_experiments
@pytest.mark.parametrize("fs,hop,n,its,cluster,sweeps,lag", DUO_CASES, ids=["%d-%d-n%d-it%d-C%d-G%d-lag%d" % c for c in DUO_CASES])
def test_strip_kernel_two_lanes_per_task(gpu, oracle, fs, hop, n, its, cluster, sweeps, lag):
    """The two-lanes-per-task kernel (lane l of warp w takes the even bins of a block, lane l of warp w + NWT the odd
    ones, a named barrier per hand-over) at plans with 2-7 warp pairs: identical bits to the oracle."""
    from lws_b200 import api
    ctx = api._context(0)
    po, pg = oracle.lws(fs, hop), gpu.lws(fs, hop)
    A = np.abs(po.stft(make_signal("tonal", 3, n)))
    try:
        ctx.set_tuning(0, cluster, sweeps)
        ctx.set_variant(lag, 3)
        for thr in (np.zeros(its), gpu.get_thresholds(its, 100 if its > 50 else 2.0, 0.1, 1)):
            Y = pg.batch_lws(A, thresholds=thr)
            plan = ctx.last_batch_plan()
            assert plan is not None and plan["tensor_memory"] == 3, plan
            assert cluster == 0 or plan["cluster"] == cluster
            _close(Y, po.batch_lws(A, thresholds=thr), "two lanes per task %s" % (plan,))
    finally:
        ctx.set_tuning(0, 0, 0)
        ctx.set_variant(0, 0)


def test_cfg5_plan_vs_oracle(gpu, oracle):
    """BASELINE configs[4] as planned for 4 utterances per GPU (Q = 8 LWSanyQ, cluster of 8 strips, >= 3 passes of the
    sweeps the ring holds, 200 default sweeps) on 1/16 of the frames (T = 352), 4 utterances, against the oracle."""
    from lws_b200 import api
    ctx = api._context(0)
    po, pg = oracle.lws(2048, 256), gpu.lws(2048, 256)
    x = make_signal("tonal", 5005, 90112 - 2048 + 256)
    A = np.abs(po.stft(x))
    assert A.shape == (352, 1025), A.shape
    thr = gpu.get_thresholds(200, 100, 0.1, 1)
    As = [A, A[::-1].copy(), np.abs(po.stft(make_signal("white", 5006, 90112 - 2048 + 256))), A]
    try:
        ctx.set_tuning(0, 8, 0)
        Ys = pg.batch_lws(As, thresholds=thr)
        plan = ctx.last_batch_plan()
        work = ctx.last_batch_work()
        assert plan is not None and plan["cluster"] == 8 and plan["sweeps_per_pass"] >= 7, plan
        assert work["passes"] >= 3, (plan, work)
    finally:
        ctx.set_tuning(0, 0, 0)
    want = po.batch_lws(A, thresholds=thr)
    _close(Ys[0], want, "cfg5 plan %s" % (plan,))
    assert np.array_equal(Ys[3], Ys[0])
    _close(Ys[2], po.batch_lws(As[2], thresholds=thr), "cfg5 plan, white utterance")


@pytest.mark.parametrize("fs,hop,n_sig,n_utt,its,cluster", [(2048, 512, 24000, 8, 30, 4), (1024, 512, 30000, 4, 24, 2), (4096, 1024, 50000, 1, 12, 8),
                                                          (1024, 256, 12000, 5, 16, 2)])
def test_rotating_lane_order_vs_oracle(gpu, oracle, monkeypatch, fs, hop, n_sig, n_utt, its, cluster):
    """Strips of 16 + 1 frame slots run the rotating lane order (lane -> frame slot changes with the macro-step): other spectra
    and cluster sizes than BASELINE configs[1], ragged batches, several passes -- against the oracle and against the
    frame-fastest order (LWSB_STRIP_NO_ROTATE)."""
    from lws_b200 import api
    ctx = api._context(0)
    po, pg = oracle.lws(fs, hop), gpu.lws(fs, hop)
    As = [np.abs(po.stft(make_signal("tonal" if i % 2 else "white", 700 + i, n_sig + 1500 * i))) for i in range(n_utt)]
    thr = gpu.get_thresholds(its, 1.5, 0.2, 1)
    try:
        ctx.set_tuning(0, cluster, 0)
        Ys = pg.batch_lws(As, thresholds=thr)
        plan = ctx.last_batch_plan()
        assert plan is not None and plan["frame_slots"] == 17 and plan["sweep_fastest"] == 2, plan
        monkeypatch.setenv("LWSB_STRIP_NO_ROTATE", "1")
        Yf = pg.batch_lws(As, thresholds=thr)
        assert ctx.last_batch_plan()["sweep_fastest"] != 2
    finally:
        ctx.set_tuning(0, 0, 0)
    assert all(np.array_equal(a, b) for a, b in zip(Ys, Yf))
    for A, Y in zip(As[:2], Ys[:2]):
        _close(Y, po.batch_lws(A, thresholds=thr), "rotating lane order %s" % (plan,))


def test_generic_and_strip_kernels_agree(gpu):
    from lws_b200 import _native
    p = gpu.lws(1024, 256)
    A = np.abs(p.stft(make_signal("white", 9, 40000)))
    thr = gpu.get_thresholds(20, 3.0, 0.15, 1)
    a = gpu.batch_lws(A, p.W, thr)
    b = gpu.batch_lws(A, p.W, thr, flags=_native.FORCE_GENERIC)
    assert np.array_equal(a, b)


@needs_experiments
@pytest.mark.parametrize("fs,hop,n,its,cluster,lag", [(512, 128, 9000, 9, 2, 0), (512, 128, 9000, 9, 4, 5), (1024, 256, 30000, 12, 8, 0),
                                                      (128, 64, 9000, 10, 1, 0), (128, 64, 9000, 10, 2, 3)])
def test_strip_kernel_tensor_memory_variant(gpu, oracle, fs, hop, n, its, cluster, lag):
    """The experimental producer / consumer variant (term values handed from warp w+4 to warp w through
    tensor memory, tcgen05.st / tcgen05.ld) and forced sweep lags: same bits."""
    from lws_b200 import api
    ctx = api._context(0)
    po, pg = oracle.lws(fs, hop), gpu.lws(fs, hop)
    A = np.abs(po.stft(make_signal("tonal", 4, n)))
    try:
        ctx.set_tuning(0, cluster, 0)
        ctx.set_variant(lag, 1)
        for thr in (np.zeros(its), gpu.get_thresholds(its, 2.0, 0.1, 1)):
            Y = pg.batch_lws(A, thresholds=thr)
            plan = ctx.last_batch_plan()
            assert plan is not None and plan["tensor_memory"] == 1 and plan["cluster"] == cluster
            assert lag == 0 or plan["sweep_lag"] == lag
            _close(Y, po.batch_lws(A, thresholds=thr), "tensor-memory variant %s" % (plan,))
    finally:
        ctx.set_tuning(0, 0, 0)
        ctx.set_variant(0, 0)


PAIR_CASES = [(512, 128, 9000, 9, 2, 0), (512, 128, 9000, 9, 4, 5), (1024, 256, 30000, 12, 2, 0), (1024, 256, 30000, 12, 8, 0),
              (128, 64, 9000, 10, 1, 0), (128, 64, 9000, 10, 2, 3), (512, 128, 700, 5, 1, 0)]


@pytest.mark.parametrize("variant", [2, 3, 10, 11, 12, 13, 14, 15])
@pytest.mark.parametrize("fs,hop,n,its,cluster,lag", PAIR_CASES)
def test_strip_kernel_variants(gpu, oracle, variant, fs, hop, n, its, cluster, lag):
    """One thread per task (2), two lanes per task alternating bins (3) and -- experimental builds -- the pair-split
    kernels (10 + window mode + 3 * explicit pipelining; modes that are not compiled into this build fall back to the
    default one): same bits, default and custom windows."""
    from lws_b200 import api
    if variant >= 3 and not _experiments():
        pytest.skip("the two-lane and pair-split kernels are built only with -DLWSB_EXPERIMENTS")
    ctx = api._context(0)
    po, pg = oracle.lws(fs, hop), gpu.lws(fs, hop)
    A = np.abs(po.stft(make_signal("tonal", 4, n)))
    try:
        ctx.set_tuning(0, cluster, 0)
        ctx.set_variant(lag, variant)
        for thr in (np.zeros(its), gpu.get_thresholds(its, 2.0, 0.1, 1)):
            Y = pg.batch_lws(A, thresholds=thr)
            plan = ctx.last_batch_plan()
            assert plan is not None and plan["cluster"] == cluster
            assert (plan["tensor_memory"] == 0) == (variant == 2), plan
            assert variant >= 10 or plan["tensor_memory"] == {2: 0, 3: 3}[variant], plan
            assert plan["threads"] <= (480 if variant == 3 else 256)
            assert lag == 0 or plan["sweep_lag"] == lag
            _close(Y, po.batch_lws(A, thresholds=thr), "variant %d %s" % (variant, plan))
    finally:
        ctx.set_tuning(0, 0, 0)
        ctx.set_variant(0, 0)


@pytest.mark.parametrize("variant", [2, 3, 13, 14])
def test_strip_kernel_variants_custom_window_and_complex_input(gpu, oracle, variant):
    """A window whose |W| > 1e-12 mask differs from the default pattern (run-time mask path) and a complex input."""
    from lws_b200 import api
    if variant >= 3 and not _experiments():
        pytest.skip("the two-lane and pair-split kernels are built only with -DLWSB_EXPERIMENTS")
    ctx = api._context(0)
    g = golden("custom_win")
    case = [c for c in CASES if c["name"] == "custom_win"][0]
    po, pg = _ctor(oracle, case), _ctor(gpu, case)
    rng = np.random.default_rng(11)
    X = g["X"]
    S = np.abs(X) * np.exp(1j * rng.uniform(-np.pi, np.pi, X.shape))
    try:
        ctx.set_variant(0, variant)
        for inp in (np.abs(X), S):
            for thr in (np.zeros(7), gpu.get_thresholds(7, 2.0, 0.2, 1)):
                _close(pg.batch_lws(inp, thresholds=thr), po.batch_lws(inp, thresholds=thr), "custom window, variant %d" % variant)
    finally:
        ctx.set_variant(0, 0)


def test_branch_free_sqrt_and_division_match_the_library(gpu):
    """The pair-split kernels' sqrt / division (MUFU seed + Newton steps, no slow-path branch) give the bits of
    __dsqrt_rn / __ddiv_rn on every sample inside their fast ranges (10^8 samples, all exponents)."""
    from lws_b200 import api
    ctx = api._context(0)
    chk_s, bad_s, chk_d, bad_d = ctx.debug_fast_math(100_000_000, 7)
    assert chk_s > 90_000_000 and chk_d > 50_000_000, (chk_s, chk_d)  # half the division samples have any exponent: many quotients leave the fast range
    assert bad_s == 0 and bad_d == 0, (bad_s, bad_d)


BLOCK4_CASES = [  # fsize, hop, samples, iterations, cluster, sweeps per pass, variant
    (512, 128, 6000, 6, 1, 0, 0), (512, 128, 6000, 6, 2, 0, 0), (512, 128, 6000, 7, 4, 2, 0), (512, 128, 700, 5, 2, 2, 0),
    (1024, 256, 30000, 12, 2, 0, 0), (1024, 256, 30000, 12, 4, 0, 0), (1024, 256, 30000, 12, 8, 0, 0), (1024, 256, 30000, 9, 8, 2, 0),
    (128, 64, 9000, 10, 1, 0, 0), (128, 64, 9000, 10, 2, 3, 0), (512, 128, 32000, 100, 0, 0, 0), (1024, 256, 30000, 12, 4, 0, 11),
    (512, 128, 9000, 9, 2, 0, 11),
]


@needs_experiments
@pytest.mark.parametrize("fs,hop,n,its,cluster,sweeps,variant", BLOCK4_CASES, ids=["%d-%d-n%d-it%d-C%d-G%d-V%d" % c for c in BLOCK4_CASES])
def test_strip_kernel_four_bin_blocks(gpu, oracle, fs, hop, n, its, cluster, sweeps, variant):
    """4 bins per block, frames 3 blocks apart (the L = 5 halo bins of a strip span two blocks): same bits."""
    from lws_b200 import api
    ctx = api._context(0)
    po, pg = oracle.lws(fs, hop), gpu.lws(fs, hop)
    A = np.abs(po.stft(make_signal("tonal", 3, n)))
    try:
        ctx.set_tuning(0, cluster, sweeps)
        ctx.set_block_bins(4)
        ctx.set_variant(0, variant)
        for thr in (np.zeros(its), gpu.get_thresholds(its, 100 if its > 50 else 2.0, 0.1, 1)):
            Y = pg.batch_lws(A, thresholds=thr)
            plan = ctx.last_batch_plan()
            assert plan is not None and plan["block_bins"] == 4, plan
            assert cluster == 0 or plan["cluster"] == cluster
            _close(Y, po.batch_lws(A, thresholds=thr), "4-bin blocks %s" % (plan,))
    finally:
        ctx.set_tuning(0, 0, 0)
        ctx.set_block_bins(0)
        ctx.set_variant(0, 0)


@needs_experiments
def test_strip_kernel_four_bin_blocks_ragged_batch(gpu, oracle):
    from lws_b200 import api
    ctx = api._context(0)
    po, pg = oracle.lws(512, 128), gpu.lws(512, 128)
    As = [np.abs(po.stft(make_signal("white" if i % 2 else "tonal", 50 + i, 3000 + 700 * (i % 9)))) for i in range(45)]
    thr = gpu.get_thresholds(9, 2.0, 0.2, 1)
    want = [po.batch_lws(A, thresholds=thr) for A in As]
    try:
        ctx.set_block_bins(4)
        for cl in (4, 2, 0):
            ctx.set_tuning(0, cl, 0)
            Ys = pg.batch_lws(As, thresholds=thr)
            assert ctx.last_batch_plan()["block_bins"] == 4
            for i, (Y, W) in enumerate(zip(Ys, want)):
                _close(Y, W, "ragged member %d, cluster %d, 4-bin blocks" % (i, cl))
    finally:
        ctx.set_tuning(0, 0, 0)
        ctx.set_block_bins(0)


# ------------------------------------------------------------------ fused calls
@pytest.mark.parametrize("fs,hop,mode,perfectrec", [(512, 128, "music", True), (512, 128, "speech", True), (256, 64, "music", False),
                                                     (128, 64, "music", True)])
def test_reconstruct_equals_the_chained_calls(gpu, oracle, fs, hop, mode, perfectrec):
    """lwsb_reconstruct (waveform -> waveform on the device) against the chain stft -> abs -> run_lws -> istft: bit-identical
    to the four separate calls of this library (same kernels, same intermediate values), and -- where the iteration does
    not amplify the 1e-16 difference between the two FFTs (no NoFuture_LWSQ4 stage) -- equal to the
    oracle's chain to round-off."""
    kw = dict(mode=mode, perfectrec=perfectrec, batch_iterations=12)
    po, pg = oracle.lws(fs, hop, **kw), gpu.lws(fs, hop, **kw)
    x = np.stack([make_signal("tonal", 21, 6000), make_signal("white", 22, 6000)])
    y, cons = pg.reconstruct(x, return_consistency=True)
    Yg = pg.run_lws(np.abs(pg.stft(x)))
    yg = pg.istft(Yg)
    assert y.shape == yg.shape and np.array_equal(y, yg)
    for b in range(2):
        assert abs(cons[b] - pg.get_consistency(Yg[b])) < 1e-9
        assert po.istft(po.stft(x[b])).shape == y[b].shape
    if mode == "speech":  # batch sweeps only
        for b in range(2):
            want = po.istft(po.run_lws(np.abs(po.stft(x[b]))))
            assert np.abs(y[b] - want).max() <= 1e-8 * max(1.0, np.abs(want).max())
    y1 = pg.reconstruct(x[0])
    assert y1.shape == y[0].shape and np.array_equal(y1, y[0])


def test_per_sweep_consistency_trace(gpu, oracle):
    """batch_lws_trace: the consistency after every k-th sweep, computed on the resident batch between staged lwsb_batch calls;
    the result equals batch_lws (ghost frames stay frozen across the calls), the trace equals the oracle's consistency of the
    oracle's partial results."""
    po, pg = oracle.lws(512, 128, batch_alpha=2), gpu.lws(512, 128, batch_alpha=2)
    As = [np.abs(po.stft(make_signal(k, 80 + i, 6000))) for i, k in enumerate(("tonal", "white"))]
    thr = gpu.get_thresholds(12, 2, 0.1, 1)
    Y, tr = pg.batch_lws_trace(As[0], thresholds=thr, every=4)
    _close(Y, po.batch_lws(As[0], thresholds=thr), "traced batch_lws")
    assert tr.shape == (3,) and tr[0] < tr[-1]
    for k in range(3):
        assert abs(tr[k] - po.get_consistency(po.batch_lws(As[0], thresholds=thr[:4 * (k + 1)]))) < 1e-7
    Yb, trb = pg.batch_lws_trace(np.stack([As[0], As[1]]), iterations=6, every=1)
    assert trb.shape == (6, 2) and np.array_equal(Yb, pg.batch_lws(np.stack([As[0], As[1]]), iterations=6))


def test_cuda_tensors_in_and_out(gpu, oracle):
    """torch CUDA tensors in -> torch CUDA tensors out (complex128, same GPU), nothing crosses PCIe; same bits as the host path."""
    import torch
    po, pg = oracle.lws(512, 128, mode="music", batch_iterations=10, batch_alpha=2), gpu.lws(512, 128, mode="music", batch_iterations=10, batch_alpha=2)
    A = np.abs(po.stft(make_signal("tonal", 90, 9000)))
    At = torch.from_numpy(A).cuda()
    Y = pg.batch_lws(At)
    assert torch.is_tensor(Y) and Y.is_cuda and Y.dtype == torch.complex128 and Y.shape == At.shape
    _close(Y.cpu().numpy(), po.batch_lws(A), "batch_lws on a CUDA tensor")
    _close(pg.run_lws(At).cpu().numpy(), po.run_lws(A), "run_lws on a CUDA tensor")
    _close(pg.online_lws(At).cpu().numpy(), po.online_lws(A), "online_lws on a CUDA tensor")
    _close(pg.nofuture_lws(At.to(torch.complex128)).cpu().numpy(), po.nofuture_lws(A), "nofuture_lws on a complex CUDA tensor")
    B3 = torch.stack([At[:30], At[30:60]])
    Y3 = pg.batch_lws(B3)
    assert Y3.shape == B3.shape and np.array_equal(Y3[1].cpu().numpy(), pg.batch_lws(A[30:60]))
    Yl = pg.batch_lws([At[:30], At[:55]])
    assert isinstance(Yl, list) and np.array_equal(Yl[1].cpu().numpy(), pg.batch_lws(A[:55]))
    assert pg.batch_lws(At, iterations=0) is At
    with pytest.raises(ValueError):
        pg.batch_lws(At[:, :-1])


def test_consistency_on_device(gpu, oracle):
    po, pg = oracle.lws(512, 128), gpu.lws(512, 128)
    x = make_signal("tonal", 5, 8000)
    X = po.stft(x)
    rng = np.random.default_rng(3)
    S = np.abs(X) * np.exp(1j * rng.uniform(-np.pi, np.pi, X.shape))
    for Z in (S, po.batch_lws(np.abs(X), iterations=5)):
        assert abs(pg.get_consistency(Z) - po.get_consistency(Z)) < 1e-8
    # a true STFT is consistent to round-off: both sides report ~300 dB, compare loosely
    assert pg.get_consistency(X) > 250 and po.get_consistency(X) > 250
    from lws_b200 import transforms
    both = transforms.get_consistency(np.stack([S, S[::-1]]), 512, 128, pg.awin, pg.swin, perfectrec=True)
    assert both.shape == (2,) and abs(both[0] - po.get_consistency(S)) < 1e-8


def test_batch_sharded_over_two_gpus_in_one_process(gpu, oracle):
    """device=[0, 1]: utterances split contiguously over the GPUs, one host thread and context per device, no collective
    (skipped on a single-GPU box)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    po = oracle.lws(512, 128)
    p2 = gpu.lws(512, 128, device=[0, 1])
    As = [np.abs(po.stft(make_signal("tonal" if i % 2 else "white", 70 + i, 4000 + 500 * (i % 5)))) for i in range(9)]
    thr = gpu.get_thresholds(8, 2.0, 0.2, 1)
    Ys = p2.batch_lws(As, thresholds=thr)
    for A, Y in zip(As, Ys):
        _close(Y, po.batch_lws(A, thresholds=thr), "two-GPU shard")
    Yr = p2.run_lws(As[:4])
    p1 = gpu.lws(512, 128)
    for a, b in zip(Yr, p1.run_lws(As[:4])):
        assert np.array_equal(a, b)


def test_many_short_utterances_more_work_items_than_clusters(gpu, oracle):
    """300 ragged utterances, 3 sweeps per pass: ~1200 work items dealt to the resident clusters in several rounds."""
    from lws_b200 import api
    ctx = api._context(0)
    po, pg = oracle.lws(512, 128), gpu.lws(512, 128)
    rng = np.random.default_rng(17)
    As = [np.abs(po.stft(make_signal("white" if i % 3 else "tonal", 300 + i, int(rng.integers(1500, 6000))))) for i in range(300)]
    thr = gpu.get_thresholds(12, 2.0, 0.15, 1)
    try:
        ctx.set_tuning(0, 2, 3)
        Ys = pg.batch_lws(As, thresholds=thr)
        plan = ctx.last_batch_plan()
        assert plan is not None and plan["sweeps_per_pass"] <= 3
    finally:
        ctx.set_tuning(0, 0, 0)
    for i in range(0, 300, 7):
        _close(Ys[i], po.batch_lws(As[i], thresholds=thr), "utterance %d of 300" % i)
    one = pg.batch_lws(As[123], thresholds=thr)
    assert np.array_equal(one, Ys[123])

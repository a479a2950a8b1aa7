"""Pin the oracle: it must reproduce the reference bit-for-bit.

(1) against the committed golden vectors (generated from the compiled reference by
    tools/make_golden.py) -- runs everywhere;
(2) against the compiled reference itself (oracle/_ref) when it is present -- every C entry
    point through ctypes and the whole Python API.
"""
import ctypes
import os

import numpy as np
import pytest

from conftest import CASES, SMALL_CASES, ROOT, golden

NAMES = [c["name"] for c in SMALL_CASES]


def _ctor(mod, case, **extra):
    kw = dict(case["kwargs"])
    kw.update(extra)
    if case["name"] == "custom_win":
        g = golden("custom_win")
        return mod.lws(g["awin_in"], case["args"][1], swin=g["swin_in"], **kw)
    return mod.lws(*case["args"], **kw)


@pytest.mark.parametrize("case", SMALL_CASES, ids=NAMES)
def test_helpers_match_golden(oracle, case, capsys):
    g = golden(case["name"])
    p = _ctor(oracle, case, mode="music")
    for k in ("awin", "swin", "W", "W_ai", "W_af"):
        assert np.array_equal(getattr(p, k), g[k]), k
    assert np.array_equal(p.stft(g["x"]), g["X"])
    assert np.array_equal(p.istft(g["X"]), g["xrec"])
    if "consistency" in g:
        assert p.get_consistency(g["Sc"]) == float(g["consistency"])


@pytest.mark.parametrize("case", SMALL_CASES, ids=NAMES)
def test_sweeps_match_golden(oracle, case):
    g = golden(case["name"])
    p = _ctor(oracle, case, mode="music")
    A = np.abs(g["X"])
    z = np.zeros
    checks = {
        "batch_zero": lambda: p.batch_lws(A, thresholds=z(5 if "Sc" in g else 4)),
        "nofuture_def": lambda: p.nofuture_lws(A),
        "online_def": lambda: p.online_lws(A, iterations=3 if "Sc" in g else 2),
    }
    if "Sc" in g:
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
        assert np.array_equal(fn(), g[k]), k
    nb = 8 if "Sc" in g else 6
    assert np.array_equal(_ctor(oracle, case, mode="music", batch_iterations=nb, batch_alpha=1.0).run_lws(A), g["run"])


def test_cfg1_short_matches_golden(oracle):
    g = golden("cfg1_short")
    A = np.abs(g["X"])
    assert np.array_equal(oracle.lws(512, 128).batch_lws(A), g["batch_def"])
    assert np.array_equal(oracle.lws(512, 128, mode="music").run_lws(A), g["run_music"])


def test_api_behaviours(oracle):
    """reference behaviours the oracle must share."""
    p = oracle.lws(32, 8)
    A = np.abs(p.stft(np.random.default_rng(3).standard_normal(300)))
    A0 = A.copy()
    Y = p.batch_lws(A, thresholds=np.zeros(2))
    assert Y.dtype == np.complex128 and Y.flags.c_contiguous and np.array_equal(A, A0)
    assert np.allclose(np.abs(Y), A, rtol=1e-12, atol=1e-14)
    Sc = A.astype(np.complex128)
    assert p.batch_lws(Sc, iterations=0) is Sc
    with pytest.raises(ValueError):
        p.batch_lws(A[:, :-1], iterations=1)
    x = np.random.default_rng(4).standard_normal(1000)
    assert np.abs(p.istft(p.stft(x))[:1000] - x).max() < 1e-13


# ------------------------------------------------------------------ live reference checks
def _ref_lib():
    path = os.path.join(ROOT, "oracle", "_ref", "liblws_ref.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/liblws_ref.so not built")
    return ctypes.CDLL(path)


def _rand_problem(rng, Nreal, T, L, Q, dense):
    Np, Tp = Nreal + 2 * L, T + 2 * (Q - 1)
    Sr, Si = rng.standard_normal((Tp, Np)), rng.standard_normal((Tp, Np))
    wr, wi = 0.1 * rng.standard_normal((Q, Q, L + 1)), 0.1 * rng.standard_normal((Q, Q, L + 1))
    wf = np.ones((Q, Q, L + 1), dtype=np.intc) if dense else (rng.uniform(size=(Q, Q, L + 1)) > 0.3).astype(np.intc)
    amp = np.abs(rng.standard_normal((Tp, Np)))
    return Sr, Si, wr, wi, wf, amp


@pytest.mark.parametrize("Q,fold", [(2, 2), (4, 4), (4, 0), (8, 0), (3, 0), (2, 0)])
@pytest.mark.parametrize("dense", [True, False])
def test_c_entry_points_bitexact(oracle, Q, fold, dense):
    """Arbitrary (non-symmetric) weights: exercises exactly the formula of each C variant."""
    ref, orc = _ref_lib(), oracle.lib()
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)
    d = lambda a: a.ctypes.data_as(dp)
    rng = np.random.default_rng(Q * 10 + fold)
    Nreal, T, L = 21, 9, 5
    prob = _rand_problem(rng, Nreal, T, L, Q, dense)
    suffix = {2: "Q2", 4: "Q4", 0: "anyQ"}[fold]
    qarg = [Q] if fold == 0 else []
    for thr in (0.0, 0.4):
        for kind in ("LWS", "NoFuture_LWS"):
            a = [x.copy() for x in prob]
            b = [x.copy() for x in prob]
            getattr(ref, "ref_" + kind + suffix)(d(a[0]), d(a[1]), d(a[2]), d(a[3]), a[4].ctypes.data_as(ip), d(a[5]),
                                                 Nreal, T, L, *qarg, ctypes.c_double(thr))
            fn = orc.orc_batch_sweep if kind == "LWS" else orc.orc_nofuture_sweep
            fn(fold, d(b[0]), d(b[1]), d(b[2]), d(b[3]), b[4].ctypes.data_as(ip), d(b[5]), Nreal, T, L, Q, thr)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (kind, thr)
            assert not np.array_equal(a[0], prob[0])
        for M, M0 in [(1, 0), (1, 1), (3, 4), (2, 3), (T, T + Q)]:
            for update in (2, 1):
                a = [x.copy() for x in prob]
                b = [x.copy() for x in prob]
                getattr(ref, "ref_Asym_UpdatePhase" + suffix)(
                    d(a[0]), d(a[1]), d(a[2]), d(a[3]), a[4].ctypes.data_as(ip), d(a[5]), Nreal, M, M0, L, *qarg,
                    ctypes.c_double(thr), update)
                orc.orc_asym_update(fold, d(b[0]), d(b[1]), d(b[2]), d(b[3]), b[4].ctypes.data_as(ip), d(b[5]),
                                    Nreal, M, M0, L, Q, thr, update)
                assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (M, M0, thr, update)


def test_c_helpers_bitexact(oracle):
    ref, orc = _ref_lib(), oracle.lib()
    dp = ctypes.POINTER(ctypes.c_double)
    d = lambda a: a.ctypes.data_as(dp)
    rng = np.random.default_rng(9)
    Nreal, M, L, Q = 17, 6, 5, 4
    Sr, Si = rng.standard_normal((M, Nreal)), rng.standard_normal((M, Nreal))
    shp = (M + 2 * (Q - 1), Nreal + 2 * L)
    Ea, Eb, Fa, Fb = (np.zeros(shp) for _ in range(4))
    ref.ref_ExtendSpec(d(Ea), d(Fa), d(Sr), d(Si), Nreal, M, L, Q)
    orc.orc_extend_spec(d(Eb), d(Fb), d(Sr), d(Si), Nreal, M, L, Q)
    assert np.array_equal(Ea, Eb) and np.array_equal(Fa, Fb)
    assert np.array_equal(Ea + 1j * Fa, oracle.extspec(Sr + 1j * Si, L, Q))
    A, B = np.zeros(shp), np.zeros(shp)
    ref.ref_ComputeAmpSpec(d(Ea), d(Fa), d(A), A.size)
    orc.orc_amp_spec(d(Eb), d(Fb), d(B), B.size)
    assert np.array_equal(A, B)
    o1, o2, o3, o4 = (np.zeros((M, Nreal)) for _ in range(4))
    ref.ref_CopySpec(d(Ea), d(Fa), d(o1), d(o2), Nreal, M, L, Q)
    orc.orc_copy_spec(d(Eb), d(Fb), d(o3), d(o4), Nreal, M, L, Q)
    assert np.array_equal(o1, Sr) and np.array_equal(o3, Sr) and np.array_equal(o2, o4)


@pytest.mark.parametrize("case", SMALL_CASES[:8], ids=NAMES[:8])
def test_python_api_vs_live_reference(oracle, ref_module, case, capsys):
    """Fresh random inputs (not the golden ones) through both Python APIs."""
    a, b = _ctor(ref_module, case, mode="music"), _ctor(oracle, case, mode="music")
    x = np.random.default_rng(1234).standard_normal(case["n"] + 37)
    X = a.stft(x)
    assert np.array_equal(X, b.stft(x))
    A = np.abs(X)
    assert np.array_equal(a.run_lws(A), b.run_lws(A))
    thr = np.array([0.7, 0.3, 0.0])
    assert np.array_equal(a.batch_lws(A, thresholds=thr), b.batch_lws(A, thresholds=thr))
    assert np.array_equal(a.online_lws(A, thresholds=thr), b.online_lws(A, thresholds=thr))
    assert np.array_equal(a.nofuture_lws(A, thresholds=thr), b.nofuture_lws(A, thresholds=thr))


# ---------------------------------------------------------------------------------- *fractionalQ paths
from conftest import FRAC_CASES, ref_fractional, make_signal  # noqa: E402

FRAC_NAMES = [c["name"] for c in FRAC_CASES]


def _frac_checks(p, g):
    A = np.abs(g["X"])
    return {
        "batch_zero": lambda: p.batch_lws(A, thresholds=np.zeros(5)),
        "batch_mid": lambda: p.batch_lws(A, thresholds=g["thr_mid"]),
        "batch_cplx": lambda: p.batch_lws(g["Sc"], thresholds=np.zeros(3)),
        "nofuture_def": lambda: p.nofuture_lws(A),
        "nofuture_zero": lambda: p.nofuture_lws(A, thresholds=np.zeros(2)),
        "online_def": lambda: p.online_lws(A, iterations=3),
        "online_zero": lambda: p.online_lws(g["Sc"], thresholds=np.zeros(2)),
    }


@pytest.mark.parametrize("case", FRAC_CASES, ids=FRAC_NAMES)
def test_fractional_paths_match_golden(oracle, case):
    """hop not dividing the frame size / use_simplifications=False: per-frequency weight rows (LWSfractionalQ & co.);
    golden vectors from the reference's C functions on tables with the zero row N the reference reads out of bounds"""
    g = golden(case["name"])
    p = oracle.lws(*case["args"], mode="music", **case["kwargs"])
    for k in ("awin", "swin", "W", "W_ai", "W_af"):
        assert np.array_equal(getattr(p, k), g[k]), k
    assert p.W.shape[0] == 2 * (g["X"].shape[1] - 1) and p.W.shape[0] != p.W.shape[1]
    for k, fn in _frac_checks(p, g).items():
        assert np.array_equal(fn(), g[k]), k
    pr = oracle.lws(*case["args"], mode="music", batch_iterations=8, batch_alpha=1.0, **case["kwargs"])
    assert np.array_equal(pr.run_lws(np.abs(g["X"])), g["run"])


@pytest.mark.parametrize("fs,hop,kw", [(64, 20, {}), (32, 8, {"use_simplifications": False}), (128, 48, {"look_ahead": 1}), (60, 12, {"use_simplifications": False})])
def test_fractional_oracle_vs_reference_functions(oracle, ref_module, fs, hop, kw):
    """fresh inputs: oracle == the reference's *fractionalQ functions (zero row N appended to the tables), bit for bit"""
    po, pr = oracle.lws(fs, hop, mode="music", **kw), ref_module.lws(fs, hop, mode="music", **kw)
    for k in ("W", "W_ai", "W_af"):
        assert np.array_equal(getattr(po, k), getattr(pr, k))
    A = np.abs(pr.stft(make_signal("tonal", 9, 2500)))
    thr = oracle.get_thresholds(4, 1.5, 0.3, 1)
    assert np.array_equal(po.batch_lws(A, thresholds=thr), ref_fractional(ref_module, "batch", A, pr.W, thr))
    assert np.array_equal(po.nofuture_lws(A, thresholds=thr[:2]), ref_fractional(ref_module, "nofuture", A, pr.W_ai, thr[:2]))
    assert np.array_equal(po.online_lws(A, thresholds=thr[:3]),
                          ref_fractional(ref_module, "online", A, pr.W, thr[:3], pr.W_ai, pr.W_af, pr.look_ahead))

"""CPU-side checks of the product's host logic and C-ABI (no GPU, no compute calls):

* liblws_b200.so loads and exports every symbol include/lws_b200.h declares;
* create_weights (host C++) and the window helpers against the golden vectors;
* the stencil tables + the wavefront / pipelined-sweep / online-chain schedules the kernels
  execute, replayed in numpy with *concurrent-step semantics* (all bins of a step read the
  state as it was when the step began), must reproduce the sequential oracle exactly;
* reference API behaviours that are decided before any kernel runs.
"""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, SMALL_CASES, golden, relF

import lws_b200
from lws_b200 import _native, dsp

NAMES = [c["name"] for c in SMALL_CASES]


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "lws_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(lwsb_\w+)\s*\(", hdr))
    assert len(declared) >= 25
    L = ctypes.CDLL(_native.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), "missing export: " + name
    assert declared == set(_native.SIGNATURES), declared ^ set(_native.SIGNATURES)
    assert _native.lib().lwsb_version() >= 100


def test_no_cpu_fallback_without_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        _native.Context(0)
    with pytest.raises(RuntimeError):
        lws_b200.lws(32, 8).batch_lws(np.ones((5, 17)), iterations=1)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "lws_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "lws_oracle" not in src and "oracle/" not in src, f


@pytest.mark.parametrize("case", SMALL_CASES, ids=NAMES)
def test_windows_and_weights_match_golden(case, capsys):
    g = golden(case["name"])
    kw = dict(case["kwargs"])
    if case["name"] == "custom_win":
        p = lws_b200.lws(g["awin_in"], case["args"][1], swin=g["swin_in"], mode="music", **kw)
    else:
        p = lws_b200.lws(*case["args"], mode="music", **kw)
    assert np.array_equal(p.awin, g["awin"]) and np.array_equal(p.swin, g["swin"])
    if "win_ai" in g:
        assert np.array_equal(p.win_ai, g["win_ai"]) and np.array_equal(p.win_af, g["win_af"])
    for k in ("W", "W_ai", "W_af"):
        assert np.array_equal(getattr(p, k), g[k]), k  # bit-identical tables: parity depends on it
    # the C-ABI's host C++ twin of create_weights agrees to rounding
    Wn = dsp.create_weights_native(p.awin, p.swin, p.fshift, p.L)
    assert Wn.shape == g["W"].shape and np.abs(Wn - g["W"]).max() < 1e-14
    assert p.nofuture_iterations == 1 and p.online_iterations == 10 and p.batch_iterations == 100


def test_stft_geometry_matches_oracle(oracle):
    lib = _native.lib()
    for fs, hop in ((32, 8), (64, 16), (48, 8), (36, 12), (50, 20), (512, 128)):
        awin = np.ones(fs)
        for n in (1, hop - 1, hop, fs, fs + 1, 1000, 1001, 4096):
            for pr in (True, False):
                if not pr and n < fs:
                    continue
                M = oracle.stft(np.zeros(n), fs, hop, awin, perfectrec=pr).shape[0]
                assert lib.lwsb_stft_frames(n, fs, hop, int(pr)) == M, (fs, hop, n, pr)


def test_api_checks_before_any_kernel():
    p = lws_b200.lws(32, 8)
    Sc = np.ones((6, 17), dtype=np.complex128)
    assert p.batch_lws(Sc, iterations=0) is Sc
    assert p.online_lws(Sc) is Sc and p.nofuture_lws(Sc) is Sc  # default mode: 0 iterations
    out = p.batch_lws(np.ones((6, 17), dtype=np.float32), iterations=0)
    assert out.dtype == np.complex128
    with pytest.raises(ValueError, match="non-negative frequencies"):
        p.batch_lws(np.ones((6, 16)), iterations=1)
    with pytest.raises(ValueError):
        lws_b200.stft(np.ones((2, 2, 2)), 4, 2, np.ones(4))
    with pytest.raises(ValueError, match="Odd ffts"):
        lws_b200.stft(np.ones(20), 5, 2, np.ones(5))
    with pytest.raises(ValueError):
        lws_b200.lws(32, 8, fftsize=35)
    with pytest.raises(TypeError):
        lws_b200.lws(np.ones((2, 32)), 8)
    # module-level istft / get_consistency with an fftsize that is not 2(Nreal-1): the reference's numpy expressions cannot
    # broadcast (lws.pyx:121-126, 143) and raise ValueError; checked against the compiled reference for these shapes
    S17 = np.ones((6, 33), dtype=np.complex128)
    for fft, wlen in ((32, 64), (128, 64), (128, 128), (66, 64), (64, 128)):
        with pytest.raises(ValueError, match="broadcast"):
            lws_b200.istft(S17, 16, np.ones(wlen), fftsize=fft)
    with pytest.raises(ValueError, match="broadcast"):
        lws_b200.get_consistency(S17, 128, 16, np.ones(128), np.ones(128))
    assert np.array_equal(lws_b200.get_thresholds(4, 100, 0.1, 1), 100 * np.exp(-0.1 * np.arange(4)))


# ------------------------------------------------------------------ schedule replays
FOLD = {2: 2, 4: 4}


def _tables(W, fold, rframe, cframe):
    Q = W.shape[1]
    return [_native.debug_terms(W, fold, rframe, cframe, p) for p in range(Q)]


def _update(E, A, row, c, L, Nreal, terms, thr, src=None):
    """one bin, the kernels' update_bin/commit_bin; reads from `src` (step snapshot)"""
    src = E if src is None else src
    a = A[row, c + L]
    if not (a > thr):
        return
    dr, dk, co = terms[c % len(terms)]
    t = np.sum(co * src[row + dr, c + L + dk])
    mag = abs(t)
    if mag > 0:
        v = t * a / mag
        E[row, c + L] = v
        if 1 <= c <= L:
            E[row, L - c] = np.conj(v)
        elif Nreal - 1 - L <= c <= Nreal - 2:
            E[row, L + 2 * (Nreal - 1) - c] = np.conj(v)


@pytest.mark.parametrize("name", ["q2", "q4", "q8b", "q3", "q4_L3"])
def test_pipelined_wavefront_equals_sequential_sweeps(oracle, name):
    """k_sweeps_generic's order: sweep i, frame m, bin c at step c + (L+1)*(m + Q*i)."""
    case = [c for c in SMALL_CASES if c["name"] == name][0]
    p = lws_b200.lws(*case["args"], **case["kwargs"])
    po = oracle.lws(*case["args"], **case["kwargs"])
    Q, L = p.W.shape[1], p.W.shape[2] - 1
    A0 = np.abs(golden(name)["X"])[:14]
    T, Nreal = A0.shape
    thr = np.array([0.9, 0.0, 0.3, 0.0])
    mean = np.mean(A0)
    E = dsp.extspec(A0.astype(np.complex128), L, Q)
    A = np.abs(E)
    terms = _tables(p.W, FOLD.get(Q, 0), Q, 1)
    S, iters = L + 1, len(thr)
    tmax = S * ((T - 1) + Q * (iters - 1)) + Nreal - 1
    for t in range(tmax + 1):
        snap = E.copy()
        for i in range(iters):
            for m in range(T):
                c = t - S * (m + Q * i)
                if 0 <= c < Nreal:
                    _update(E, A, m + Q - 1, c, L, Nreal, terms, thr[i] * mean, src=snap)
    got = E[Q - 1:Q - 1 + T, L:L + Nreal]
    want = po.batch_lws(A0, thresholds=thr)
    assert relF(got, want) < 1e-12


@pytest.mark.parametrize("name,LA", [("q4", 3), ("q2_la4", 4), ("q8_la2", 2), ("q4_la0", 0)])
def test_online_chain_equals_reference_schedule(oracle, name, LA):
    """k_online_generic's order: row update j of the TF-RTISI-LA chain, bin c at step c + (L+1)*j."""
    case = [c for c in SMALL_CASES if c["name"] == name][0]
    p = lws_b200.lws(*case["args"], **case["kwargs"])
    po = oracle.lws(*case["args"], **case["kwargs"])
    assert p.look_ahead == LA
    Q, L = p.W.shape[1], p.W.shape[2] - 1
    A0 = np.abs(golden(name)["X"])[:9]
    T, Nreal = A0.shape
    iters = 2
    thr = lws_b200.get_thresholds(iters, 1, 0.1, 1)
    mean = np.mean(A0)
    E = dsp.extspec(A0.astype(np.complex128), L, Q)
    A = np.abs(E)
    fold = FOLD.get(Q, 0)
    tabs = {("W", rf): _tables(p.W, fold, rf, 1) for rf in range(2, Q + 1)}
    tabs["ai"] = _tables(p.W_ai, fold, 1, 0)
    tabs["af"] = _tables(p.W_af, fold, 1, 1)
    chain = _native.debug_online_chain(T, iters, LA, Q)
    assert len(chain) == sum(1 + iters * (min(LA, m) + 1) for m in range(T))
    S = L + 1
    for t in range(S * (len(chain) - 1) + Nreal):
        snap = E.copy()
        for j in range(max(0, (t - Nreal) // S), min(len(chain) - 1, t // S) + 1):
            c = t - S * j
            if not (0 <= c < Nreal):
                continue
            row, which, rframe, cframe, ti = chain[j]
            terms = tabs[("W", rframe)] if which == 0 else (tabs["ai"] if which == 1 else tabs["af"])
            _update(E, A, row, c, L, Nreal, terms, 0.0 if ti < 0 else thr[ti] * mean, src=snap)
    got = E[Q - 1:Q - 1 + T, L:L + Nreal]
    want = po.online_lws(A0, thresholds=thr)
    assert relF(got, want) < 1e-12


@pytest.mark.parametrize("name,LA", [("q4", 3), ("q2_la4", 4), ("q4_la0", 0), ("q4_la5", 5)])
def test_online_chain_two_bins_per_step(oracle, name, LA):
    """k_online_ring2's order: row update j runs S = 8 bins behind row update j-1 and every task takes TWO bins per step
    (2t - S*j and the next one; S >= 2 + L).  Concurrent semantics: a task reads what the step began with plus its own
    writes of the step, and no task may read a cell another task writes in the same step."""
    case = [c for c in SMALL_CASES if c["name"] == name][0]
    p = lws_b200.lws(*case["args"], **case["kwargs"])
    po = oracle.lws(*case["args"], **case["kwargs"])
    Q, L = p.W.shape[1], p.W.shape[2] - 1
    A0 = np.abs(golden(name)["X"])[:9]
    T, Nreal = A0.shape
    iters = 2
    thr = lws_b200.get_thresholds(iters, 1, 0.1, 1)
    mean = np.mean(A0)
    E = dsp.extspec(A0.astype(np.complex128), L, Q)
    A = np.abs(E)
    fold = FOLD.get(Q, 0)
    tabs = {("W", rf): _tables(p.W, fold, rf, 1) for rf in range(2, Q + 1)}
    tabs["ai"] = _tables(p.W_ai, fold, 1, 0)
    tabs["af"] = _tables(p.W_af, fold, 1, 1)
    chain = _native.debug_online_chain(T, iters, LA, Q)
    S, BPS = 8, 2
    assert S % Q == 0 and S >= BPS + L
    for t in range((S * (len(chain) - 1) + Nreal + BPS - 1) // BPS + 1):
        snap = E.copy()
        written = {}
        reads = []
        for j in range(len(chain)):
            c0 = BPS * t - S * j
            if c0 + BPS - 1 < 0 or c0 >= Nreal:
                continue
            row, which, rframe, cframe, ti = chain[j]
            terms = tabs[("W", rframe)] if which == 0 else (tabs["ai"] if which == 1 else tabs["af"])
            th = 0.0 if ti < 0 else thr[ti] * mean
            own = {}
            for c in range(max(c0, 0), min(c0 + BPS, Nreal)):
                a = A[row, c + L]
                if not (a > th):
                    continue
                dr, dk, co = terms[c % len(terms)]
                vals = np.empty(len(co), dtype=np.complex128)
                for q in range(len(co)):
                    key = (row + dr[q], c + L + dk[q])
                    if key in own:
                        vals[q] = own[key]
                    else:
                        vals[q] = snap[key]
                        reads.append(key + (j,))
                tsum = np.sum(co * vals)
                if abs(tsum) > 0:
                    v = tsum * a / abs(tsum)
                    tg = [((row, c + L), v)]
                    if 1 <= c <= L:
                        tg.append(((row, L - c), np.conj(v)))
                    elif Nreal - 1 - L <= c <= Nreal - 2:
                        tg.append(((row, L + 2 * (Nreal - 1) - c), np.conj(v)))
                    for key, vv in tg:
                        E[key] = vv
                        own[key] = vv
                        written[key] = j
        for (r_, c_, j) in reads:
            assert written.get((r_, c_), j) == j, "cell read and written by different tasks in one step"
    got = E[Q - 1:Q - 1 + T, L:L + Nreal]
    want = po.online_lws(A0, thresholds=thr)
    assert relF(got, want) < 1e-12


def _flow_replay(p, po, A0, LA, iters, S, K):
    """k_online_flow's schedule: row update j is on bin b - S j at bin-step b; K warps per row update take the bin-steps in turn.  A
    warp forms the term values of the other frames for bin-step b any time after its own hand-over of bin-step b - K: modelled as
    "every such cell was last written at bin-step <= b - K" (then reading it at any moment of the window gives the same value);
    the centre-frame terms are read after the hand-over of b - 1.  No cell read at b may be written at b by another row update."""
    Q, L = p.W.shape[1], p.W.shape[2] - 1
    T, Nreal = A0.shape
    thr = lws_b200.get_thresholds(iters, 1, 0.1, 1)
    mean = np.mean(A0)
    E = dsp.extspec(A0.astype(np.complex128), L, Q)
    A = np.abs(E)
    fold = FOLD.get(Q, 0)
    tabs = {("W", rf): _tables(p.W, fold, rf, 1) for rf in range(2, Q + 1)}
    tabs["ai"] = _tables(p.W_ai, fold, 1, 0)
    tabs["af"] = _tables(p.W_af, fold, 1, 1)
    chain = _native.debug_online_chain(T, iters, LA, Q)
    last = np.full(E.shape, -10 ** 9, dtype=np.int64)   # bin-step of the last write of a cell
    for b in range(S * (len(chain) - 1) + Nreal):
        snap = E.copy()
        written, reads = {}, []
        for j in range(max(0, (b - Nreal) // S), min(len(chain) - 1, b // S) + 1):
            c = b - S * j
            if not (0 <= c < Nreal):
                continue
            row, which, rframe, cframe, ti = chain[j]
            terms = tabs[("W", rframe)] if which == 0 else (tabs["ai"] if which == 1 else tabs["af"])
            a = A[row, c + L]
            if not (a > (0.0 if ti < 0 else thr[ti] * mean)):
                continue
            dr, dk, co = terms[c % len(terms)]
            for q in range(len(co)):
                key = (row + dr[q], c + L + dk[q])
                reads.append(key + (j,))
                if dr[q] != 0 and last[key] > b - K:
                    return None, "row update %d, bin %d reads cell %s written at bin-step %d > %d - %d" % (j, c, key, last[key], b, K)
            tsum = np.sum(co * snap[row + dr, c + L + dk])
            if abs(tsum) > 0:
                v = tsum * a / abs(tsum)
                tg = [((row, c + L), v)]
                if 1 <= c <= L:
                    tg.append(((row, L - c), np.conj(v)))
                elif Nreal - 1 - L <= c <= Nreal - 2:
                    tg.append(((row, L + 2 * (Nreal - 1) - c), np.conj(v)))
                for key, val in tg:
                    E[key] = val
                    written[key] = j
        for (r_, c_, j) in reads:
            if written.get((r_, c_), j) != j:
                return None, "cell (%d, %d) read by row update %d and written by %d in bin-step %d" % (r_, c_, j, written[(r_, c_)], b)
        for key in written:
            last[key] = b
    return E[Q - 1:Q - 1 + T, L:L + Nreal], None


@pytest.mark.parametrize("name,LA", [("q4", 3), ("q2_la4", 4), ("q4_la0", 0), ("q4_la5", 5), ("q4_la1", 1)])
def test_online_flow_schedule(oracle, name, LA):
    """k_online_flow: lag S = K + L between row updates is sufficient for K warps per row update taking turns (K = 2, 3, 4), and
    S = K + L - 1 is not (the replay finds a cell read inside the window in which it is written)."""
    case = [c for c in SMALL_CASES if c["name"] == name][0]
    p = lws_b200.lws(*case["args"], **case["kwargs"])
    po = oracle.lws(*case["args"], **case["kwargs"])
    assert p.look_ahead == LA
    L = p.W.shape[2] - 1
    A0 = np.abs(golden(name)["X"])[:9]
    iters = 2
    want = po.online_lws(A0, thresholds=lws_b200.get_thresholds(iters, 1, 0.1, 1))
    for K in (2, 3, 4):
        got, err = _flow_replay(p, po, A0, LA, iters, K + L, K)
        assert err is None, (K, err)
        assert relF(got, want) < 1e-12
        got, err = _flow_replay(p, po, A0, LA, iters, K + L + 2, K)
        assert err is None and relF(got, want) < 1e-12
    _, err = _flow_replay(p, po, A0, LA, iters, 4 + L - 1, 4)
    assert err is not None



def test_nofuture_q4_table_reproduces_reference_indexing(oracle):
    """LWSB_FOLD_NF4 terms applied with the reference's flat offset (m+dr)*Np + 2e + dk, raster order."""
    p, po = lws_b200.lws(32, 8), oracle.lws(32, 8)
    Q, L = 4, 5
    A0 = np.abs(golden("q4")["X"])[:10]
    T, Nreal = A0.shape
    Np = Nreal + 2 * L
    E = dsp.extspec(A0.astype(np.complex128), L, Q)
    A = np.abs(E)
    terms = _tables(p.W_ai, 5, 1, 0)
    thr = 0.2 * np.mean(A0)
    flat = E.reshape(-1)
    for m in range(Q - 1, T + Q - 1):
        for c in range(Nreal):
            a = A[m, c + L]
            if not (a > thr):
                continue
            dr, dk, co = terms[c % Q]
            t = np.sum(co * flat[(m + dr) * Np + 2 * (c + L) + dk])
            if abs(t) > 0:
                v = t * a / abs(t)
                E[m, c + L] = v
                if 1 <= c <= L:
                    E[m, L - c] = np.conj(v)
                elif Nreal - 1 - L <= c <= Nreal - 2:
                    E[m, L + 2 * (Nreal - 1) - c] = np.conj(v)
    want = po.nofuture_lws(A0, thresholds=np.array([0.2]))
    assert relF(E[Q - 1:Q - 1 + T, L:L + Nreal], want) < 1e-12


def test_native_create_weights_is_parity_grade_for_batch_sweeps(oracle):
    """SURVEY.md row a15.  lwsb_create_weights (host C++) cannot reproduce numpy's BLAS `dot` bit for bit: its tables differ
    from the reference's by ~1e-16.  Weight perturbations do not touch the exact cancellations the batch iteration's
    stability rests on (those come from the mirrored DATA), so 100 batch sweeps with the native tables stay within 1e-10
    of the reference's -- five orders inside north_star's 1e-5.  (NoFuture_LWSQ4, numerically expanding because of its
    indexing slip, amplifies the same 1e-16 to O(0.1): a caller that needs run_lws parity in music mode passes the
    reference's W, as INTEGRATION.md says.)"""
    from lws_b200 import dsp
    from conftest import make_signal
    for fs, hop, kind in ((512, 128, "tonal"), (256, 32, "white")):
        po = oracle.lws(fs, hop)
        Wn = dsp.create_weights_native(po.awin, po.swin, hop, 5)
        assert np.abs(Wn - po.W).max() < 4e-16 and np.array_equal(np.abs(Wn) > 1e-12, np.abs(po.W) > 1e-12)
        A = np.abs(po.stft(make_signal(kind, 7, 16000)))
        thr = oracle.get_thresholds(100, 100, 0.1, 1)
        Y0, Y1 = oracle.batch_lws(A, po.W, thr), oracle.batch_lws(A, Wn, thr)
        assert np.linalg.norm(Y1 - Y0) / np.linalg.norm(Y0) < 1e-10


def _frame_base(m, it, LA):
    """lwsb_online_frame_base (lwsb_common.h): number of row updates before frame m"""
    if LA <= 0:
        return m * (1 + it)
    if m <= LA:
        return m + it * m * (m + 1) // 2
    return LA + it * LA * (LA + 1) // 2 + (m - LA) * (1 + it * (LA + 1))


def _frame_of(it, LA, j):
    m = 0
    while _frame_base(m + 1, it, LA) <= j:
        m += 1
    return m


@pytest.mark.parametrize("T,Nreal,iters,LA,Q", [(9, 33, 2, 3, 4), (1, 17, 3, 3, 4), (30, 129, 1, 0, 2), (12, 257, 2, 5, 4), (3, 513, 4, 1, 4),
                                                 (40, 65, 1, 2, 2)])
def test_online_flow_barrier_protocol_terminates(T, Nreal, iters, LA, Q):
    """The control flow of k_online_flow as a discrete simulation: K warps per group of 32 row updates, hand-over barriers 1 + b mod K
    shared by the warps that committed bin-step b and the warps about to run b + 1, CTA barriers around the ring residency changes
    (with the role of bin-step bs taking its hand-over BEFORE the drain).  Every warp must reach the end, every barrier must be
    met by exactly the warps it is sized for, for any shape -- no dead-lock, no stray arrival."""
    K, L = 4, 5
    S = K + L
    n = _frame_base(T, iters, LA)
    assert n == len(_native.debug_online_chain(T, iters, LA, Q))
    tasks = (Nreal + S - 1) // S + 1
    G = (tasks + 31) // 32
    bend = S * (n - 1) + (Nreal - 1) + 1
    Tp = T + 2 * (Q - 1)

    def warp(role, grp):
        hi, ev_j = -1, 0
        b, bmod, jhi = role, role, 0
        while b <= bend:
            waited = False
            if bmod < K and jhi >= ev_j:
                jh = min(jhi, n - 1)
                mfront = _frame_of(iters, LA, jh)
                ev_j = _frame_base(mfront + 1, iters, LA) if mfront + 1 < T else 1 << 62
                need_hi = min(Tp - 1, mfront + 2 * (Q - 1))
                if need_hi > hi:
                    if bmod == 0 and b > 0:
                        yield ("bar", 1 + (role + K - 1) % K); waited = True
                    yield ("cta",); yield ("cta",); yield ("cta",)
                    hi = need_hi
            if b > 0 and not waited:
                yield ("bar", 1 + (role + K - 1) % K)
            if b < bend:
                yield ("bar", 1 + role)
            if b + K > bend:
                break
            b += K; bmod += K
            if bmod >= S:
                bmod -= S; jhi += 1
        yield ("cta",)

    warps = {(r, g): warp(r, g) for r in range(K) for g in range(G)}
    waiting = {}
    for key, w in warps.items():
        waiting[key] = next(w, None)
    steps = 0
    while any(v is not None for v in waiting.values()):
        steps += 1
        progressed = False
        live = [k for k, v in waiting.items() if v is not None]
        if all(waiting[k] == ("cta",) for k in live) and len(live) == K * G:
            for k in live:
                waiting[k] = next(warps[k], None)
            progressed = True
        else:
            for bid in range(1, K + 1):
                at = [k for k in live if waiting[k] == ("bar", bid)]
                assert len(at) <= 2 * G, "barrier %d oversubscribed: %s" % (bid, at)
                if len(at) == 2 * G:
                    assert sorted(set(r for r, _ in at)) == sorted({(bid - 1) % K, bid % K}), at  # committers of b and waiters of b + 1
                    for k in at:
                        waiting[k] = next(warps[k], None)
                    progressed = True
        assert progressed, "dead-lock: %s" % {k: v for k, v in waiting.items() if v is not None}
        assert steps < 10 * (bend + 10) * K


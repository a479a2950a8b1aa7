"""Shared pytest plumbing.

* ``-m "not gpu"`` : oracle vs golden vectors, host logic, C-ABI symbol checks (CPU only).
* ``-m gpu``       : parity of the CUDA path against the oracle, through the C-ABI.

The oracle (oracle/) is the checker and is imported only here, never by lws_b200.
"""
import importlib.util
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_cases(key="cases"):
    with open(os.path.join(GOLDEN, "cases.json")) as f:
        return json.load(f)[key]


CASES = load_cases()
FRAC_CASES = load_cases("fractional_cases")  # the *fractionalQ paths: hop not dividing the frame size / use_simplifications=False
SMALL_CASES = [c for c in CASES if c["name"] not in ("cfg1_short",)]


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def load_ref_module():
    """The compiled reference (oracle/_ref/lws_ref*.so) or None when it was not built."""
    d = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isdir(d):
        return None
    for f in os.listdir(d):
        if f.startswith("lws_ref") and f.endswith(".so"):
            spec = importlib.util.spec_from_file_location("lws_ref", os.path.join(d, f))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod
    return None


@pytest.fixture(scope="session")
def oracle():
    import lws_oracle
    lws_oracle.lib()
    return lws_oracle


@pytest.fixture(scope="session")
def ref_module():
    m = load_ref_module()
    if m is None:
        pytest.skip("compiled reference (oracle/_ref) not present")
    return m


def relF(y, yref):
    """rel-Frobenius error, the parity metric of SURVEY.md section 8c."""
    return float(np.linalg.norm(y - yref) / max(np.linalg.norm(yref), 1e-300))


def make_signal(kind, seed, n):
    if kind == "white":
        return np.random.default_rng(seed).standard_normal(n)
    t = np.arange(n) / 16000.0
    f0 = 120 + 30 * np.sin(2 * np.pi * 3 * t)
    ph = 2 * np.pi * np.cumsum(f0) / 16000.0
    x = sum(np.sin(k * ph) / k for k in range(1, 30))
    x = x * (0.5 + 0.5 * np.sin(2 * np.pi * 2 * t)) ** 2
    return x + 0.01 * np.random.default_rng(seed).standard_normal(n)


# ---------------------------------------------------------------------------------- *fractionalQ reference driver
def ref_fractional(ref_module, stage, S, W, thresholds, W_ai=None, W_af=None, LA=3):
    """The reference's *fractionalQ C functions (lwslib.cpp:376-467, 693-764, 1276-1492) driven the way its binding
    drives them (lws.pyx:209-375), but on weight tables with a zero row N appended: the reference reads that row, one
    past its table, at the DC bin -- undefined behaviour that makes the compiled module's results depend on the heap.
    With the padded tables the reference's own code is deterministic; oracle and CUDA path are pinned to it."""
    import ctypes
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "liblws_ref.so"))
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)
    d = lambda a: a.ctypes.data_as(dp)
    S = np.asarray(S).astype(np.complex128)
    L, Q = W.shape[2] - 1, W.shape[1]
    T, Nreal = S.shape

    def split(Wx):
        Wp = np.concatenate([Wx, np.zeros_like(Wx[:1])], axis=0)
        return (np.ascontiguousarray(Wp.real), np.ascontiguousarray(Wp.imag),
                np.ascontiguousarray(np.abs(Wp) > 1.0e-12, dtype=np.intc))
    E = ref_module.extspec(S, L, Q)
    Er, Ei = np.ascontiguousarray(E.real), np.ascontiguousarray(E.imag)
    amp = np.ascontiguousarray(np.abs(E))
    mean = np.mean(np.abs(S))
    wr, wi, wf = split(W)
    if stage in ("batch", "nofuture"):
        fn = lib.ref_LWSfractionalQ if stage == "batch" else lib.ref_NoFuture_LWSfractionalQ
        for t in thresholds:
            fn(d(Er), d(Ei), d(wr), d(wi), wf.ctypes.data_as(ip), d(amp), Nreal, T, L, Q, ctypes.c_double(t * mean))
    else:
        ar, ai, af = split(W_ai)
        fr, fi, ff = split(W_af)
        thr = np.ascontiguousarray(np.asarray(thresholds, dtype=np.float64) * mean)
        lib.ref_TF_RTISI_LA(d(Er), d(Ei), d(wr), d(wi), d(ar), d(ai), d(fr), d(fi), wf.ctypes.data_as(ip), af.ctypes.data_as(ip),
                            ff.ctypes.data_as(ip), d(amp), len(thr), int(LA), Nreal, T, L, Q, ctypes.c_double(float(Q)),  # Qfloat: read only when update == 1
                            0, d(thr), 2)
    return Er[(Q - 1):(Q - 1 + T), L:(Nreal + L)] + 1j * Ei[(Q - 1):(Q - 1 + T), L:(Nreal + L)]

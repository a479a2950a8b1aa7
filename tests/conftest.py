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


def load_cases():
    with open(os.path.join(GOLDEN, "cases.json")) as f:
        return json.load(f)["cases"]


CASES = load_cases()
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

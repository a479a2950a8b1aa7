"""Host-side DSP helpers of the reference's Python layer (python/lws.pyx:10-206).

Window construction and threshold schedules are tiny host computations in the reference
and stay numpy here; `create_weights` runs in the library's host C++, and
`stft` / `istft` / `get_consistency` run on the GPU through the C-ABI.
"""
from __future__ import annotations

import numpy as np

from . import _native


def hann(n, symmetric=True, use_offset=False):
    """Hann window (lws.pyx:10-19): half-sample-offset ("symmetric") or periodic."""
    if symmetric:
        return 0.5 * (1 - np.cos(2 * np.pi * (np.arange(1, 2 * n, 2)) / (2 * n)))
    offset = 1 if use_offset else 0
    return 0.5 * (1 - np.cos(2 * np.pi * (np.arange(n) + offset) / n))


def synthwin(awin, fshift, swin=None):
    """Synthesis window normalised so that overlap-add of awin*swin is 1 (lws.pyx:22-40)."""
    fsize = len(awin)
    Q = int(np.ceil(float(fsize) / float(fshift)))
    if swin is None:
        swin = awin
    prod = np.hstack([awin * swin, np.zeros((Q * fshift - fsize,))])
    norm = np.sum(np.reshape(prod, (Q, fshift)), axis=0)
    norm = np.tile(norm, (1, Q))[0, :fsize]
    if min(norm) <= 0:
        raise ValueError('The normalizer is not strictly positive')
    return swin / norm


def extspec(S, L, Q):
    """Extended spectrogram (lws.pyx:146-157).  The CUDA path builds this on the device
    (k_extend); the host version exists because the reference module exports it."""
    T, Nreal = S.shape
    E = np.zeros((T + 2 * (Q - 1), Nreal + 2 * L), dtype=S.dtype)
    E[(Q - 1):(Q - 1 + T), L:(Nreal + L)] = S
    E[:, 0:L] = np.conjugate(E[:, (2 * L):L:-1])
    E[:, (Nreal + L):(Nreal + 2 * L)] = np.conjugate(E[:, (Nreal + L - 2):(Nreal - 2):-1])
    E[:(Q - 1)] = np.atleast_2d(E[Q - 1])
    E[(Q - 1 + T):] = np.atleast_2d(E[Q - 2 + T])
    return E


def create_weights(awin, swin, fshift, L, use_summarized_weights=True):
    """Complex LWS weights, shape (Qprime, Q, L+1) (lws.pyx:160-181).

    Evaluated with the same numpy operations as the reference (complex exponential tables, a
    BLAS ``dot`` with the window products, two broadcast multiplications): the LWS iteration
    is sensitive to the last bit of these 96..384 numbers ("why bit-exact"), and a
    BLAS dot cannot be reproduced natively.  ``lwsb_create_weights`` in the C-ABI is the host
    C++ twin (equal to ~1e-16); the class uses this one so that results match the reference
    bit for bit."""
    T = len(awin)
    Q = int(np.ceil(float(T) / float(fshift)))
    Qfloat = float(T) / float(fshift)
    Qprime = Q if (T % fshift == 0 and use_summarized_weights) else T
    kcol = np.atleast_2d(np.arange(L + 1)).T
    dft_rows = np.exp(-1j * 2 * np.pi * kcol * np.arange(T) / T)
    winprod = np.zeros((T, Q))
    for q in range(Q):
        t = np.arange(T - q * fshift)
        winprod[t, q] = awin[t] * swin[t + q * fshift] / T
    W = (dft_rows.dot(winprod)) * np.exp(-1j * 2 * np.pi * kcol * np.arange(Q) / Qfloat)
    W[0, 0] = W[0, 0] - 1
    phase = np.exp(1j * 2 * np.pi * np.atleast_2d(np.arange(Qprime)).T * np.arange(Q) / Qfloat)
    W = W[:, np.newaxis] * phase[np.newaxis, :]
    return W.transpose((1, 2, 0))


def create_weights_native(awin, swin, fshift, L, use_summarized_weights=True):
    """The C-ABI's host C++ version of the same table (lwsb_create_weights)."""
    return _native.create_weights(awin, swin, fshift, L, use_summarized_weights)


def build_asymmetric_windows(awin_swin, fshift):
    """Mirrored envelopes of RTISI-LA (lws.pyx:184-200); input is the product awin*swin."""
    T = len(awin_swin)
    Q = int(np.ceil(float(T) / float(fshift)))
    tails = np.zeros((T, Q))
    tails[:, 0] = awin_swin
    for q in range(Q):
        n = T - q * fshift
        tails[:n, q] = awin_swin[q * fshift:q * fshift + n]
    win_ai = np.sum(tails[:, 1:], axis=1)[::-1]
    win_af = np.sum(tails, axis=1)[::-1]
    if T % fshift == 2:  # reference behaviour (lws.pyx:198), kept as is
        win_ai = awin_swin
    return win_ai, win_af


def get_thresholds(iterations, alpha, beta, gamma):
    """Sparsity thresholds alpha*exp(-beta*i^gamma) (lws.pyx:203-206)."""
    return alpha * np.exp(- beta * np.arange(iterations) ** gamma)

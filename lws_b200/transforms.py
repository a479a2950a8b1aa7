"""stft / istft / get_consistency of the reference (python/lws.pyx:43-144) on the GPU.

Same signatures, argument checks and output shapes as the reference functions.  Extensions
that cannot collide with reference behaviour: a 2-D ``x`` of shape (B, nsamples) / a 3-D
spectrogram (B, M, Nreal) is processed as a batch of equal-length signals in one launch
(the reference raises ValueError for those), and ``device=`` selects the GPU.
"""
from __future__ import annotations

import numpy as np

from . import dsp


def _ctx(device):
    from .api import _context, _devices
    return _context(_devices(device)[0])


def stft(x, fsize, fshift, awin, fftsize=None, perfectrec=False, *, device=None):
    """STFT with a fixed frame shift (lws.pyx:43-90)."""
    x = np.asarray(x)
    batched = x.ndim == 2
    if x.ndim not in (1, 2):
        raise ValueError('We only deal with single channel signals here')
    if fftsize is None:
        fftsize = fsize
    if fftsize % 2 == 1:
        raise ValueError('Odd ffts not supported.')
    awin = np.squeeze(np.asarray(awin, dtype=np.float64))
    if awin.shape != (fsize,):
        raise ValueError('operands could not be broadcast together: frame (%d,) window %s' % (fsize, awin.shape))
    if np.iscomplexobj(x):
        raise TypeError('real signals only')
    xb = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64)
    ctx = _ctx(device)
    with ctx.lock:
        S = ctx.stft(xb, awin, int(fsize), int(fshift), int(fftsize), perfectrec is True)
    return S if batched else S[0]


def istft(spec, fshift, swin, awin=None, fftsize=None, perfectrec=False, *, device=None):
    """iSTFT with a fixed frame shift (lws.pyx:93-137)."""
    spec = np.asarray(spec)
    batched = spec.ndim == 3
    if spec.ndim not in (2, 3):
        raise ValueError('We only deal with single channel signals here')
    M, N = spec.shape[-2:]
    if N % 2 != 1:
        raise ValueError('We expect the spectrogram to only have non-negative frequencies')
    fsize = 2 * (N - 1)
    if awin is not None:
        swin = dsp.synthwin(awin, fshift, swin=swin)
    if fftsize is not None and fftsize != fsize:
        raise NotImplementedError('istft with fftsize != 2*(Nreal-1) is not supported by the CUDA implementation')
    swin = np.squeeze(np.asarray(swin, dtype=np.float64))
    if len(swin) > fsize:
        raise ValueError('operands could not be broadcast together: frame (%d,) window %s' % (fsize, swin.shape))
    Sb = np.ascontiguousarray(spec if batched else spec[None], dtype=np.complex128)
    ctx = _ctx(device)
    with ctx.lock:
        sig = ctx.istft(Sb, swin, int(fshift))
    if perfectrec is True:
        residual_size = fsize % fshift
        pre_pad_length = fsize - fshift if residual_size == 0 else fsize - residual_size
        sig = sig[:, pre_pad_length:(fshift - fsize)]  # same slice as lws.pyx:135 (empty when fshift == fsize)
    return sig if batched else sig[0]


def get_consistency(S, fsize, fshift, awin, swin, perfectrec=False, *, device=None):
    """Consistency in dB (lws.pyx:140-144): 20 log10(|S| / |stft(istft(S)) - S|), computed on the device in one call
    (istft, stft and the two Frobenius norms; the spectrogram crosses PCIe once).  A 3-D S gives one value per
    spectrogram."""
    S = np.asarray(S)
    batched = S.ndim == 3
    if S.ndim not in (2, 3):
        raise ValueError('We only deal with single channel signals here')
    if S.shape[-1] % 2 != 1:
        raise ValueError('We expect the spectrogram to only have non-negative frequencies')
    if 2 * (S.shape[-1] - 1) != fsize:
        raise NotImplementedError('get_consistency with fftsize != fsize is not supported by the CUDA implementation')
    awin = np.squeeze(np.asarray(awin, dtype=np.float64))
    swin = np.squeeze(np.asarray(swin, dtype=np.float64))
    Sb = np.ascontiguousarray(S if batched else S[None], dtype=np.complex128)
    ctx = _ctx(device)
    with ctx.lock:
        c = ctx.consistency(Sb, awin, swin, int(fshift), perfectrec is True)
    return c if batched else float(c[0])

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
    if fftsize is None:
        fftsize = fsize
    swin = np.squeeze(np.asarray(swin, dtype=np.float64))
    if swin.ndim == 0:
        swin = np.full(1, float(swin))
    if fftsize > len(swin):
        swin = np.hstack([swin, np.zeros((fftsize - len(swin),))])
    # lws.pyx:121-126: the frame is the first 2(Nreal-1) samples of an n=fftsize inverse transform, times the (padded)
    # window.  Unless fftsize == 2(Nreal-1) == len(window) numpy cannot broadcast those; the reference raises
    # ValueError there (checked against the compiled reference), and so does this.
    frame_len = min(fsize, fftsize)
    if len(swin) == 1:
        swin = np.full(frame_len, swin[0])
    if frame_len != len(swin) or frame_len != fsize:
        raise ValueError('operands could not be broadcast together with shapes (%d,) (%d,) ' % (frame_len, len(swin)))
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
        # lws.pyx:143: stft(istft(S)) has fsize/2+1 bins, S has others: numpy's subtraction raises there
        raise ValueError('operands could not be broadcast together with shapes (%d,) (%d,) ' % (fsize // 2 + 1, S.shape[-1]))
    awin = np.squeeze(np.asarray(awin, dtype=np.float64))
    swin = np.squeeze(np.asarray(swin, dtype=np.float64))
    Sb = np.ascontiguousarray(S if batched else S[None], dtype=np.complex128)
    ctx = _ctx(device)
    with ctx.lock:
        c = ctx.consistency(Sb, awin, swin, int(fshift), perfectrec is True)
    return c if batched else float(c[0])

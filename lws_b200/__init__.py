"""lws_b200 -- B200-native (sm_100a CUDA) implementation of the LWS phase-recovery hot path
of Jonathan-LeRoux/lws, behind the reference's own Python API:

    import lws_b200 as lws
    p = lws.lws(512, 128, mode="speech")
    X = p.stft(x); Y = p.run_lws(abs(X)); y = p.istft(Y)

Module attributes mirror python/lws.pyx:8-378.  There is no CPU compute path: the CUDA library
(lws_b200/liblws_b200.so, built by `python -m lws_b200.build`) and a GPU are required for
everything that touches a spectrogram.
"""
__version__ = "1.2.8"            # the reference's (lws.pyx:8): a drop-in answers the same
__b200_version__ = "2.0"           # this implementation's own
__reference_version__ = "1.2.8"

from .dsp import (hann, synthwin, extspec, create_weights, build_asymmetric_windows, get_thresholds)  # noqa: F401
from .api import batch_lws, nofuture_lws, online_lws, lws, OnlineStream  # noqa: F401


def stft(x, fsize, fshift, awin, fftsize=None, perfectrec=False, **kw):
    from . import transforms
    return transforms.stft(x, fsize, fshift, awin, fftsize=fftsize, perfectrec=perfectrec, **kw)


def istft(spec, fshift, swin, awin=None, fftsize=None, perfectrec=False, **kw):
    from . import transforms
    return transforms.istft(spec, fshift, swin, awin=awin, fftsize=fftsize, perfectrec=perfectrec, **kw)


def get_consistency(S, fsize, fshift, awin, swin, perfectrec=False, **kw):
    from . import transforms
    return transforms.get_consistency(S, fsize, fshift, awin, swin, perfectrec=perfectrec, **kw)

"""The reference's Python surface (python/lws.pyx:209-499) on top of the CUDA library.

Same names, positional order, defaults and error behaviour as the Cython module, so that
``import lws_b200 as lws`` is a drop-in for the hot path.  Extensions (keyword-only or
impossible in the reference, so they cannot collide):

* ``S`` may be a 3-D array ``(B, T, Nreal)`` or a list of 2-D arrays ``(T_i, Nreal)``: the
  whole batch is processed by one call on the GPU (the reference has no batch dimension).
* ``device=`` on the class / functions selects the GPU(s); a list shards the utterances over
  several GPUs (independent utterances, no collective).
"""
from __future__ import annotations

import os
import threading

import numpy as np

from . import _native, _pinned, dsp
from .dsp import get_thresholds

_EVEN = 'Please only include non-negative frequencies in the input spectrogram.'
_ctx_lock = threading.Lock()
_contexts = {}


def _context(device, slot=0):
    """One context (stream, device buffers) per (device, slot); slot 1 is the second half of an overlapped batch."""
    key = device if slot == 0 else (device, slot)
    with _ctx_lock:
        c = _contexts.get(key)
        if c is None:
            c = _contexts[key] = _native.Context(device)
        return c


# Optional (LWSB_OVERLAP_MIN=n, off by default): a batch of at least n utterances on ONE device is cut in two (58 % / 42 %)
# and handled by two host threads with a context each, so that the second half's host-to-device copy and pre-processing
# overlap the first half's sweeps and the first half's device-to-host copy overlaps the second half's sweeps (the strip
# kernels themselves run one after the other, lws_b200.h).  Measured on B200 at BASELINE configs[1]: 125 ms per call
# against 120 ms without it -- two launches of 37 / 27 utterances fill the 74 clusters' rounds of work items worse than
# one launch of 64, which costs more than the ~5 ms of copies it hides.
OVERLAP_MIN = int(os.environ.get("LWSB_OVERLAP_MIN", "0"))


def _devices(device):
    if device is None:
        return [0]
    if isinstance(device, (list, tuple)):
        devs = [int(d) for d in device]
        if len(set(devs)) != len(devs):
            # two shards on one device would share that device's context (resident batch, stream)
            raise ValueError('device list contains a GPU more than once: %r' % (devs,))
        if not devs:
            raise ValueError('empty device list')
        return devs
    return [int(device)]


def _as_batch(S):
    """-> (list of 2-D C-contiguous float64/complex128 arrays, kind, rebuild(outs))."""
    if isinstance(S, (list, tuple)):
        arrs, shape = [np.asarray(a) for a in S], "list"
    else:
        S = np.asarray(S)
        if S.ndim == 3:
            arrs, shape = [S[b] for b in range(S.shape[0])], "3d"
        else:
            arrs, shape = [S], "2d"
    if any(a.ndim != 2 for a in arrs):
        raise ValueError('expected (T, Nreal) spectrograms')
    cplx = any(np.iscomplexobj(a) for a in arrs)
    dt = np.complex128 if cplx else np.float64
    arrs = [np.ascontiguousarray(a, dtype=dt) for a in arrs]
    return arrs, (_native.C128 if cplx else _native.F64), shape


def _rebuild(outs, shape, out=None, whole=None):
    if out is not None:
        return out
    if shape == "2d":
        return outs[0]
    if shape == "3d":
        return whole if whole is not None else np.stack(outs)
    return outs


class _Outs(list):
    """per-utterance result arrays; `whole` is the (B, T, Nreal) array they are slices of, if any"""
    whole = None


def _alloc_outs(arrs, shape, out=None):
    if out is not None:
        # caller-provided result buffer(s) (e.g. pinned host memory): complex128, C-contiguous
        outs = list(out) if isinstance(out, (list, tuple)) else ([out[b] for b in range(len(arrs))] if shape == "3d" else [out])
        for a, o in zip(arrs, outs):
            if o.shape != a.shape or o.dtype != np.complex128 or not o.flags.c_contiguous:
                raise ValueError('out= must be C-contiguous complex128 of the input shape')
        if len(outs) != len(arrs):
            raise ValueError('out= does not match the batch')
        return outs
    # fresh result arrays, as the reference returns them -- in page-locked memory where that pays (_pinned.py)
    if shape == "3d":
        big = _pinned.empty((len(arrs),) + arrs[0].shape)
        outs = _Outs(big[b] for b in range(len(arrs)))
        outs.whole = big
        return outs
    return _Outs(_pinned.empty(a.shape) for a in arrs)


def _passthrough(S):
    # iterations == 0: the reference returns the (complex128-cast) input itself (lws.pyx:212-220)
    if isinstance(S, (list, tuple)):
        return [a if a.dtype == np.complex128 else a.astype(np.complex128) for a in map(np.asarray, S)]
    S = np.asarray(S) if not isinstance(S, np.ndarray) else S
    return S if S.dtype == np.complex128 else S.astype(np.complex128)


def _check_shapes(arrs):
    nreal = arrs[0].shape[1]
    if any(a.shape[1] != nreal for a in arrs):
        raise ValueError('all spectrograms of a batch must have the same number of bins')
    if nreal % 2 == 0:
        raise ValueError(_EVEN)
    if any(a.shape[0] < 1 for a in arrs):
        raise ValueError('empty spectrogram')


def _check_weights(W, use_simplifications):
    """Variant dispatch of lws.pyx:246-247 / 299-300 / lwslib.cpp:1441: per-frequency weight rows (frame shift not dividing
    the frame size, or use_simplifications=False) select the reference's *fractionalQ variants.  The library recognises
    them by Qprime != Q; what has no counterpart in the reference is refused here."""
    W = np.asarray(W)
    if W.ndim != 3:
        raise ValueError('weights must have shape (Qprime, Q, L+1)')
    if not use_simplifications and W.shape[0] == W.shape[1] and W.shape[0] > 1:
        # summarised weights with use_simplifications=False: the reference would index row n-L of a Q-row table (out of
        # bounds for every bin beyond Q); create_weights never produces this combination
        raise ValueError('use_simplifications=False needs per-frequency weights (create_weights(..., use_summarized_weights=False))')


def _shard(frames, k):
    """Utterance indices per device: longest-processing-time-first over the frame counts (the work of an utterance is
    proportional to its number of frames), SURVEY.md section 8e.  Equal lengths give the contiguous balanced split.
    Returns k' <= k non-empty index lists; the indices inside a list are in increasing order."""
    n = len(frames)
    k = max(1, min(k, n))
    if len(set(frames)) <= 1:
        b = [(n * i) // k for i in range(k + 1)]
        return [list(range(b[i], b[i + 1])) for i in range(k)]
    load, parts = [0] * k, [[] for _ in range(k)]
    for i in sorted(range(n), key=lambda i: (-frames[i], i)):
        d = min(range(k), key=lambda d: (load[d], d))
        parts[d].append(i)
        load[d] += frames[i]
    return [sorted(p) for p in parts if p]


def _run_sharded(devices, arrs, outs, fn, overlap=False):
    """fn(ctx, arrays, outs) on each device's share; one host thread per GPU (ctypes releases the GIL, so the GPUs run
    concurrently).  A context is single-threaded (one resident batch, one stream): every use holds its lock, so host
    threads that share a device are serialised like the reference's calls are by the GIL."""
    parts = _shard([a.shape[0] for a in arrs], len(devices))
    slots = [0] * len(parts)
    if len(parts) == 1 and overlap and OVERLAP_MIN > 0 and len(arrs) >= OVERLAP_MIN:
        cut = (len(arrs) * 58 + 99) // 100
        parts, devices, slots = [list(range(cut)), list(range(cut, len(arrs)))], [devices[0], devices[0]], [0, 1]
    if len(parts) == 1:
        ctx = _context(devices[0])
        with ctx.lock:
            fn(ctx, arrs, outs)
        return
    errs = []

    def work(dev, slot, idx):
        try:
            ctx = _context(dev, slot)
            with ctx.lock:
                fn(ctx, [arrs[i] for i in idx], [outs[i] for i in idx])
        except BaseException as e:  # re-raised in the caller
            errs.append(e)

    ts = [threading.Thread(target=work, args=(devices[i], slots[i], idx)) for i, idx in enumerate(parts)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    if errs:
        raise errs[0]


def _cuda_tensors(S):
    """torch CUDA tensor(s) in -> (list of 2-D tensors, kind, shape, device index) or None.  An extension:
    spectrograms that already live in HBM skip PCIe in both directions; the result is a torch tensor on the same GPU."""
    try:
        import torch
    except ImportError:
        return None
    items = list(S) if isinstance(S, (list, tuple)) else [S]
    if not items or not all(torch.is_tensor(t) and t.is_cuda for t in items):
        return None
    if isinstance(S, (list, tuple)):
        ts, shape = items, "list"
    elif S.dim() == 3:
        ts, shape = [S[b] for b in range(S.shape[0])], "3d"
    else:
        ts, shape = [S], "2d"
    if any(t.dim() != 2 for t in ts):
        raise ValueError('expected (T, Nreal) spectrograms')
    cplx = any(t.is_complex() for t in ts)
    dt = torch.complex128 if cplx else torch.float64
    ts = [t.to(dt).contiguous() for t in ts]
    if len({t.device.index for t in ts}) != 1:
        raise ValueError('all spectrograms of a batch must be on the same GPU')
    return ts, (_native.C128 if cplx else _native.F64), shape, ts[0].device.index


def _run_cuda(cu, weights, stage):
    """stage(ctx) on a device-resident batch: load from the tensors' pointers, compute, store into fresh tensors"""
    import torch
    ts, kind, shape, dev = cu
    nreal = ts[0].shape[1]
    if any(t.shape[1] != nreal for t in ts):
        raise ValueError('all spectrograms of a batch must have the same number of bins')
    if nreal % 2 == 0:
        raise ValueError(_EVEN)
    if shape == "3d":
        whole = torch.empty((len(ts),) + tuple(ts[0].shape), dtype=torch.complex128, device=ts[0].device)
        outs = [whole[b] for b in range(len(ts))]
    else:
        whole, outs = None, [torch.empty(tuple(t.shape), dtype=torch.complex128, device=t.device) for t in ts]
    torch.cuda.current_stream(ts[0].device).synchronize()  # the producers of the inputs have finished
    ctx = _context(dev)
    with ctx.lock:
        for which, Wx in weights:
            ctx.set_weights(which, Wx)
        ctx.load_device([t.data_ptr() for t in ts], [t.shape[0] for t in ts], nreal, kind)
        stage(ctx)
        ctx.store_device([o.data_ptr() for o in outs])
    return outs[0] if shape == "2d" else (whole if shape == "3d" else outs)


def batch_lws(S, W, thresholds, use_simplifications=True, *, device=None, flags=0, out=None):
    """Batch-mode LWS phase reconstruction (lws.pyx:209-258)."""
    cu = _cuda_tensors(S)
    if cu is not None:
        _check_weights(W, use_simplifications)
        return S if len(thresholds) == 0 else _run_cuda(cu, [(_native.W, W)], lambda ctx: ctx.batch(thresholds, flags))
    if len(thresholds) == 0:
        return _passthrough(S)
    arrs, kind, shape = _as_batch(S)
    _check_shapes(arrs)
    _check_weights(W, use_simplifications)
    outs = _alloc_outs(arrs, shape, out)

    def fn(ctx, a, o):
        ctx.set_weights(_native.W, W)
        ctx.batch_lws(a, kind, thresholds, flags, outs=o)

    _run_sharded(_devices(device), arrs, outs, fn, overlap=True)
    return _rebuild(outs, shape, out, getattr(outs, 'whole', None))


def nofuture_lws(S, W, thresholds, use_simplifications=True, *, device=None, flags=0, out=None):
    """LWS using past frames only, typically for initialisation (lws.pyx:261-311)."""
    cu = _cuda_tensors(S)
    if cu is not None:
        _check_weights(W, use_simplifications)
        return S if len(thresholds) == 0 else _run_cuda(cu, [(_native.W, W)], lambda ctx: ctx.nofuture(_native.W, thresholds, flags))
    if len(thresholds) == 0:
        return _passthrough(S)
    arrs, kind, shape = _as_batch(S)
    _check_shapes(arrs)
    _check_weights(W, use_simplifications)
    outs = _alloc_outs(arrs, shape, out)

    def fn(ctx, a, o):
        ctx.set_weights(_native.W, W)
        ctx.nofuture_lws(_native.W, a, kind, thresholds, flags, outs=o)

    _run_sharded(_devices(device), arrs, outs, fn)
    return _rebuild(outs, shape, out, getattr(outs, 'whole', None))


def online_lws(S, W, W_ai, W_af, thresholds, LA, fshift, use_simplifications=True, *, device=None, flags=0, out=None):
    """Online (TF-RTISI-LA) LWS phase reconstruction (lws.pyx:314-375)."""
    cu = _cuda_tensors(S)
    if cu is not None:
        _check_weights(W, use_simplifications)
        return S if len(thresholds) == 0 else _run_cuda(cu, [(_native.W, W), (_native.W_AI, W_ai), (_native.W_AF, W_af)],
                                                        lambda ctx: ctx.online(thresholds, int(LA), flags))
    if len(thresholds) == 0:
        return _passthrough(S)
    arrs, kind, shape = _as_batch(S)
    _check_shapes(arrs)
    _check_weights(W, use_simplifications)
    outs = _alloc_outs(arrs, shape, out)

    def fn(ctx, a, o):
        ctx.set_weights(_native.W, W)
        ctx.set_weights(_native.W_AI, W_ai)
        ctx.set_weights(_native.W_AF, W_af)
        ctx.online_lws(a, kind, thresholds, int(LA), flags, outs=o)

    _run_sharded(_devices(device), arrs, outs, fn)
    return _rebuild(outs, shape, out, getattr(outs, 'whole', None))


class OnlineStream(object):
    """Frame-in / frame-out ``online_lws`` (an extension).  TF_RTISI_LA (lwslib.cpp:1432-1491) is causal:
    the row updates of frame m read nothing beyond frame m, so they run as soon as the frame arrives and frame
    ``m - look_ahead`` is final once they have.  ``push(frames)`` returns the frames that became final, ``close()`` the rest.

    The reference scales its thresholds by the mean of |S| over the WHOLE utterance (lws.pyx:360-361), which a stream
    cannot know: ``mean_amp`` is the caller's value or estimate.  With the true mean the concatenated output is
    bit-identical to ``online_lws`` on the complete spectrogram.  Each stream owns a context (its frames stay in HBM)."""

    def __init__(self, plugin, n_bins, max_frames, mean_amp, iterations=None, thresholds=None, complex_input=False):
        if iterations is None:
            iterations = plugin.online_iterations
        if thresholds is None:
            thresholds = get_thresholds(iterations, plugin.online_alpha, plugin.online_beta, plugin.online_gamma)
        if len(thresholds) == 0:
            raise ValueError('a stream needs at least one online iteration')
        if n_bins % 2 == 0:
            raise ValueError(_EVEN)
        _check_weights(plugin.W, plugin.use_simplifications)
        self._kind = _native.C128 if complex_input else _native.F64
        self._dtype = np.complex128 if complex_input else np.float64
        self._n_bins = int(n_bins)
        self._ctx = _native.Context(_devices(plugin.device)[0])
        self._ctx.set_weights(_native.W, plugin.W)
        self._ctx.set_weights(_native.W_AI, plugin.W_ai)
        self._ctx.set_weights(_native.W_AF, plugin.W_af)
        self._ctx.stream_begin(n_bins, max_frames, self._kind, mean_amp, thresholds, plugin.look_ahead)
        self._given = 0

    def push(self, frames):
        """append frames (n, n_bins); returns the (k, n_bins) complex128 frames that are final now (k may be 0)"""
        f = np.ascontiguousarray(np.atleast_2d(frames), dtype=self._dtype)
        if f.ndim != 2 or f.shape[1] != self._n_bins:
            raise ValueError('expected frames of %d bins' % self._n_bins)
        if f.shape[0]:
            self._ctx.stream_push(f)
        return self._take()

    def _take(self):
        _, final = self._ctx.stream_frames()
        out = self._ctx.stream_read(self._given, final - self._given)
        self._given = final
        return out

    def close(self):
        """end of the utterance: the last look_ahead frames are final as they stand"""
        self._ctx.stream_end()
        out = self._take()
        self._ctx.close()
        return out


class lws(object):
    """Drop-in for ``lws.lws`` (lws.pyx:378-499).  Holds windows, weights and the iteration /
    threshold schedule; every spectrogram computation runs on the GPU."""

    def __init__(self, awin_or_fsize, fshift, L=5, swin=None, look_ahead=3,
                 nofuture_iterations=0, nofuture_alpha=1, nofuture_beta=0.1, nofuture_gamma=1,
                 online_iterations=0, online_alpha=1, online_beta=0.1, online_gamma=1,
                 batch_iterations=100, batch_alpha=100, batch_beta=0.1, batch_gamma=1,
                 symmetric_win=True, mode=None, fftsize=None, perfectrec=True, use_simplifications=True,
                 *, device=None):
        if isinstance(awin_or_fsize, int):
            # default perfect-reconstruction window: sqrt-Hann, renormalised (lws.pyx:384-387)
            awin = np.sqrt(dsp.hann(awin_or_fsize, symmetric=symmetric_win, use_offset=False))
            awin = np.sqrt(awin * dsp.synthwin(awin, fshift))
        else:
            awin = awin_or_fsize
        if awin.ndim > 1:
            # the reference compares a shape tuple with an int here (lws.pyx:391), which raises
            # TypeError for a 2-D window with more than one row; same expression, same behaviour
            if (awin.ndim > 2) or (awin.shape[0] > 1 and awin.shape > 1):
                raise ValueError('The analysis window should be flat')
            else:
                awin = awin.flatten()
        if fftsize is None:
            fftsize = len(awin)
        if fftsize > len(awin):
            if (fftsize - len(awin)) % 2 != 0:
                raise ValueError('The zero-padding should add even length to the original window.')
            pad_length = (fftsize - len(awin)) // 2
            print('Zero-padding symmetrically around the original windows.\n'
                  'WARNING: for code simplicity, a consequence is that the first/last '
                  '{} samples of the signal will not be '.format(pad_length) +
                  'in the perfect reconstruction region.')
            pad = np.zeros(pad_length)
            awin = np.hstack((pad, awin, pad))
            if swin is not None:
                swin = np.hstack((pad, swin, pad))

        self.awin = awin
        if swin is not None:
            print('Provided synthesis window is renormalized for perfect reconstruction.')
        self.swin = dsp.synthwin(awin, fshift, swin=swin)
        self.fshift = fshift
        self.fsize = len(awin)
        self.perfectrec = perfectrec
        self.L = L
        if self.fsize % self.fshift == 0:
            self.Q = int(self.fsize / self.fshift)
        else:
            self.Q = self.fsize / self.fshift
        self.use_simplifications = use_simplifications
        self.W = dsp.create_weights(self.awin, self.swin, self.fshift, self.L,
                                    use_summarized_weights=self.use_simplifications)
        self.win_ai, self.win_af = dsp.build_asymmetric_windows(self.awin * self.swin, self.fshift)
        self.W_ai = dsp.create_weights(self.win_ai, self.swin, self.fshift, self.L,
                                       use_summarized_weights=self.use_simplifications)
        self.W_af = dsp.create_weights(self.win_af, self.swin, self.fshift, self.L,
                                       use_summarized_weights=self.use_simplifications)
        self.look_ahead = look_ahead

        if mode == 'speech':
            nofuture_iterations = 0
            online_iterations = 0
        elif mode == 'music':
            nofuture_iterations = 1
            online_iterations = 10

        self.batch_iterations = batch_iterations
        self.batch_alpha = batch_alpha
        self.batch_beta = batch_beta
        self.batch_gamma = batch_gamma
        self.online_iterations = online_iterations
        self.online_alpha = online_alpha
        self.online_beta = online_beta
        self.online_gamma = online_gamma
        self.nofuture_iterations = nofuture_iterations
        self.nofuture_alpha = nofuture_alpha
        self.nofuture_beta = nofuture_beta
        self.nofuture_gamma = nofuture_gamma
        self.device = device

        if (not np.allclose(awin, awin[::-1])):
            print('WARNING: It appears you are using an analysis window that is not symmetric.\n'
                  'The current code uses simplifications that rely on such symmetry, so the code may not behave properly.')

    # ---- transforms (GPU) -------------------------------------------------------------------
    def reconstruct(self, x, *, return_consistency=False):
        """Waveform -> waveform on the device (an extension; the reference chains four calls through the host):
        ``istft(run_lws(abs(stft(x))))`` for a signal or a (B, nsamples) batch of equal-length signals, with the class's
        iteration settings.  With ``return_consistency`` also the consistency (dB) of the reconstructed spectrogram(s)."""
        x = np.asarray(x)
        batched = x.ndim == 2
        if x.ndim not in (1, 2):
            raise ValueError('We only deal with single channel signals here')
        if np.iscomplexobj(x):
            raise TypeError('real signals only')
        _check_weights(self.W, self.use_simplifications)
        nf = get_thresholds(self.nofuture_iterations, self.nofuture_alpha, self.nofuture_beta, self.nofuture_gamma)
        on = get_thresholds(self.online_iterations, self.online_alpha, self.online_beta, self.online_gamma)
        ba = get_thresholds(self.batch_iterations, self.batch_alpha, self.batch_beta, self.batch_gamma)
        xb = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64)
        ctx = _context(_devices(self.device)[0])
        with ctx.lock:
            ctx.set_weights(_native.W, self.W)
            ctx.set_weights(_native.W_AI, self.W_ai)
            ctx.set_weights(_native.W_AF, self.W_af)
            res = ctx.reconstruct(xb, self.awin, self.swin, int(self.fsize), int(self.fshift), self.perfectrec is True, nf, on,
                                  self.look_ahead, ba, consistency=return_consistency)
        if return_consistency:
            y, c = res
            return (y, c) if batched else (y[0], float(c[0]))
        return res if batched else res[0]

    def get_consistency(self, S):
        from . import transforms
        return transforms.get_consistency(S, self.fsize, self.fshift, self.awin, self.swin,
                                          perfectrec=self.perfectrec, device=self.device)

    def stft(self, S):
        from . import transforms
        return transforms.stft(S, self.fsize, self.fshift, self.awin, perfectrec=self.perfectrec, device=self.device)

    def istft(self, S):
        from . import transforms
        # awin is not passed: swin was renormalised at construction (lws.pyx:465-467)
        return transforms.istft(S, self.fshift, self.swin, perfectrec=self.perfectrec, device=self.device)

    # ---- phase reconstruction (GPU) -----------------------------------------------------------
    def nofuture_lws(self, S, iterations=None, thresholds=None, *, out=None):
        if iterations is None:
            iterations = self.nofuture_iterations
        if thresholds is None:
            thresholds = get_thresholds(iterations, self.nofuture_alpha, self.nofuture_beta, self.nofuture_gamma)
        return nofuture_lws(S, self.W_ai, thresholds, use_simplifications=self.use_simplifications, device=self.device,
                            out=out)

    def online_lws(self, S, iterations=None, thresholds=None, *, out=None):
        if iterations is None:
            iterations = self.online_iterations
        if thresholds is None:
            thresholds = get_thresholds(iterations, self.online_alpha, self.online_beta, self.online_gamma)
        return online_lws(S, self.W, self.W_ai, self.W_af, thresholds, self.look_ahead, self.fshift,
                          use_simplifications=self.use_simplifications, device=self.device, out=out)

    def online_stream(self, n_bins, max_frames, mean_amp, iterations=None, thresholds=None, complex_input=False):
        """frame-in / frame-out online_lws with this object's windows and schedule (see OnlineStream)"""
        return OnlineStream(self, n_bins, max_frames, mean_amp, iterations, thresholds, complex_input)

    def batch_lws(self, S, iterations=None, thresholds=None, *, out=None):
        if iterations is None:
            iterations = self.batch_iterations
        if thresholds is None:
            thresholds = get_thresholds(iterations, self.batch_alpha, self.batch_beta, self.batch_gamma)
        return batch_lws(S, self.W, thresholds, use_simplifications=self.use_simplifications, device=self.device, out=out)

    def batch_lws_trace(self, S, iterations=None, thresholds=None, every=1):
        """batch_lws with the consistency (dB, lws.pyx:140-144) recorded every `every` sweeps, computed on the device without
        moving the batch (an extension).  Returns (result, trace): the result is bit-identical to
        batch_lws(S), trace[k] holds the consistency of every utterance after sweep min((k + 1) * every, iterations), and
        trace[-1] that of the result -- shape (n_points,) for one spectrogram, (n_points, B) for a batch."""
        if iterations is None:
            iterations = self.batch_iterations
        if thresholds is None:
            thresholds = get_thresholds(iterations, self.batch_alpha, self.batch_beta, self.batch_gamma)
        thresholds = np.asarray(thresholds, dtype=np.float64)
        if len(thresholds) == 0:
            return _passthrough(S), np.zeros((0,))
        arrs, kind, shape = _as_batch(S)
        _check_shapes(arrs)
        _check_weights(self.W, self.use_simplifications)
        if 2 * (arrs[0].shape[1] - 1) != self.fsize:
            raise ValueError('the consistency needs spectrograms of this object\'s frame size')
        outs = _alloc_outs(arrs, shape)
        ctx = _context(_devices(self.device)[0])
        trace = []
        with ctx.lock:
            ctx.set_weights(_native.W, self.W)
            ctx.load(arrs, kind)
            for i in range(0, len(thresholds), max(1, int(every))):
                ctx.batch(thresholds[i:i + max(1, int(every))])
                trace.append(ctx.resident_consistency(self.awin, self.swin, self.fshift, self.perfectrec is True))
            ctx.store(outs)
        trace = np.array(trace)
        return _rebuild(outs, shape, None, getattr(outs, 'whole', None)), (trace[:, 0] if shape == "2d" else trace)

    def run_lws(self, S, *, out=None):
        """nofuture -> online -> batch (lws.pyx:495-499), fused on the device: the spectrograms
        cross PCIe once in each direction instead of three times."""
        nf = get_thresholds(self.nofuture_iterations, self.nofuture_alpha, self.nofuture_beta, self.nofuture_gamma)
        on = get_thresholds(self.online_iterations, self.online_alpha, self.online_beta, self.online_gamma)
        ba = get_thresholds(self.batch_iterations, self.batch_alpha, self.batch_beta, self.batch_gamma)
        cu = _cuda_tensors(S)
        if cu is not None:  # torch CUDA tensors: the three stages on the device-resident batch, no PCIe traffic
            _check_weights(self.W, self.use_simplifications)
            if len(nf) + len(on) + len(ba) == 0:
                return S

            def stages(ctx):
                dirty = False
                for n, fn in ((len(nf), lambda: ctx.nofuture(_native.W_AI, nf)), (len(on), lambda: ctx.online(on, self.look_ahead)),
                              (len(ba), lambda: ctx.batch(ba))):
                    if n:
                        if dirty:
                            ctx.restage()
                        fn()
                        dirty = True
            return _run_cuda(cu, [(_native.W, self.W), (_native.W_AI, self.W_ai), (_native.W_AF, self.W_af)], stages)
        if len(nf) + len(on) + len(ba) == 0:
            return _passthrough(S)
        arrs, kind, shape = _as_batch(S)
        _check_shapes(arrs)
        _check_weights(self.W, self.use_simplifications)
        outs = _alloc_outs(arrs, shape, out)

        def fn(ctx, a, o):
            ctx.set_weights(_native.W, self.W)
            ctx.set_weights(_native.W_AI, self.W_ai)
            ctx.set_weights(_native.W_AF, self.W_af)
            ctx.run_lws(a, kind, nf, on, self.look_ahead, ba, outs=o)

        _run_sharded(_devices(self.device), arrs, outs, fn)
        return _rebuild(outs, shape, out, getattr(outs, 'whole', None))

"""ctypes binding of liblws_b200.so (include/lws_b200.h).  Thin by design: argument
marshalling and error-code -> exception translation only; there is no Python/numpy compute
path behind it -- if the CUDA library is missing or no GPU is present, calls raise."""
from __future__ import annotations

import ctypes
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LWSB_LIB_PATH") or os.path.join(_HERE, "liblws_b200.so")  # the override is for timing experiments with instrumented builds

C128, F64 = 0, 1
W, W_AI, W_AF = 0, 1, 2
HOST, DEVICE = 0, 1
FORCE_GENERIC, FORCE_ANYQ = 1, 2

ERR_CUDA, ERR_ARG, ERR_EVEN_NREAL, ERR_UNSUPPORTED, ERR_STATE, ERR_NOMEM = -1, -2, -3, -4, -5, -6

_vp, _dp, _ip = ctypes.c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)
_vpp = ctypes.POINTER(ctypes.c_void_p)
_ci, _cd, _ll = ctypes.c_int, ctypes.c_double, ctypes.c_longlong

# name -> (restype, argtypes): every symbol include/lws_b200.h declares
SIGNATURES = {
    "lwsb_version": (_ci, []),
    "lwsb_has_experiments": (_ci, []),
    "lwsb_strip_launch_mode": (_ci, []),
    "lwsb_last_error": (ctypes.c_char_p, [_vp]),
    "lwsb_create": (_ci, [_ci, _vp, _vpp]),
    "lwsb_destroy": (_ci, [_vp]),
    "lwsb_sync": (_ci, [_vp]),
    "lwsb_host_alloc": (_ci, [ctypes.c_ulonglong, _vpp]),
    "lwsb_host_free": (_ci, [_vp]),
    "lwsb_set_weights": (_ci, [_vp, _ci, _dp, _dp, _ci, _ci, _ci]),
    "lwsb_create_weights": (_ci, [_dp, _dp, _ci, _ci, _ci, _ci, _dp, _dp, _ip, _ip]),
    "lwsb_load": (_ci, [_vp, _vpp, _ip, _ci, _ci, _ci, _ci]),
    "lwsb_batch": (_ci, [_vp, _dp, _ci, _ci]),
    "lwsb_nofuture": (_ci, [_vp, _ci, _dp, _ci, _ci]),
    "lwsb_online": (_ci, [_vp, _dp, _ci, _ci, _ci]),
    "lwsb_store": (_ci, [_vp, _vpp, _ci]),
    "lwsb_restage": (_ci, [_vp]),
    "lwsb_batch_lws": (_ci, [_vp, _vpp, _vpp, _ip, _ci, _ci, _ci, _ci, _dp, _ci, _ci]),
    "lwsb_nofuture_lws": (_ci, [_vp, _ci, _vpp, _vpp, _ip, _ci, _ci, _ci, _ci, _dp, _ci, _ci]),
    "lwsb_online_lws": (_ci, [_vp, _vpp, _vpp, _ip, _ci, _ci, _ci, _ci, _dp, _ci, _ci, _ci]),
    "lwsb_run_lws": (_ci, [_vp, _vpp, _vpp, _ip, _ci, _ci, _ci, _ci, _dp, _ci, _dp, _ci, _ci, _dp, _ci, _ci]),
    "lwsb_stream_begin": (_ci, [_vp, _ci, _ci, _ci, _cd, _dp, _ci, _ci, _ci]),
    "lwsb_stream_push": (_ci, [_vp, _vp, _ci, _ci]),
    "lwsb_stream_frames": (_ci, [_vp, _ip, _ip]),
    "lwsb_stream_read": (_ci, [_vp, _vp, _ci, _ci, _ci]),
    "lwsb_stream_end": (_ci, [_vp]),
    "lwsb_stft_frames": (_ci, [_ci, _ci, _ci, _ci]),
        "lwsb_stft_prepad": (_ci, [_ci, _ci, _ci]),
    "lwsb_stft": (_ci, [_vp, _vp, _ci, _ci, _dp, _ci, _ci, _ci, _ci, _ci, _vp, _ci]),
    "lwsb_istft": (_ci, [_vp, _vp, _ci, _ci, _ci, _dp, _ci, _ci, _vp, _ci]),
    "lwsb_reconstruct_length": (_ll, [_ci, _ci, _ci, _ci]),
    "lwsb_reconstruct": (_ci, [_vp, _vp, _ci, _ci, _dp, _dp, _ci, _ci, _ci, _dp, _ci, _dp, _ci, _ci, _dp, _ci, _ci, _vp, _ci, _dp]),
    "lwsb_consistency": (_ci, [_vp, _vp, _ci, _ci, _ci, _dp, _dp, _ci, _ci, _ci, _ci, _dp]),
    "lwsb_resident_consistency": (_ci, [_vp, _dp, _dp, _ci, _ci, _ci, _dp]),
    "lwsb_last_compute_ms": (_ci, [_vp, ctypes.POINTER(ctypes.c_float)]),
    "lwsb_launch_count": (_ll, [_vp]),
    "lwsb_last_stage_ms": (_ci, [_vp, ctypes.POINTER(ctypes.c_float)]),
    "lwsb_last_batch_work": (_ci, [_vp, ctypes.POINTER(_ll)]),
    "lwsb_last_online_kernel": (_ci, [_vp]),
    "lwsb_last_batch_plan": (_ci, [_vp, _ip]),
    "lwsb_device_info": (_ci, [_vp, _ip, _ip, _ip, ctypes.POINTER(_ll)]),
    "lwsb_get_stats": (_ci, [_vp, _dp, _dp]),
    "lwsb_debug_terms": (_ci, [_dp, _dp, _ci, _ci, _ci, _ci, _ci, _ci, _ci, _ip, _ip, _dp, _dp]),
    "lwsb_debug_plan_strips": (_ci, [_ci, _ci, _ci, _ci, _ci, _ci, _ll, _ci, _ci, _ci, _ci, _ip]),
    "lwsb_set_block_bins": (_ci, [_vp, _ci]),
    "lwsb_set_tuning": (_ci, [_vp, _ll, _ci, _ci]),
    "lwsb_set_variant": (_ci, [_vp, _ci, _ci]),
    "lwsb_last_batch_cycles": (_ci, [_vp, ctypes.POINTER(ctypes.c_ulonglong)]),
    "lwsb_last_batch_trace": (_ci, [_vp, _ci, _ci, _ip, ctypes.POINTER(ctypes.c_ulonglong)]),
    "lwsb_debug_fast_math": (_ci, [_vp, _ll, ctypes.c_ulonglong, ctypes.POINTER(ctypes.c_ulonglong)]),
    "lwsb_debug_work_items": (_ci, [_ip, _ci, _ci, _ci, _ip]),
    "lwsb_debug_online_chain_length": (_ll, [_ci, _ci, _ci]),
    "lwsb_debug_online_task": (_ci, [_ci, _ci, _ci, _ci, _ll, _ip, _ip, _ip, _ip, _ip]),
}

_lib = None
_lock = threading.Lock()


def lib():
    """Load the CUDA library; fail loudly when it has not been built (no fallback)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    "lws_b200: %s is missing -- build it with `python -m lws_b200.build` "
                    "(nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
            L = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(L, name)  # AttributeError if the .so does not export it
                fn.restype, fn.argtypes = res, args
            _lib = L
    return _lib


def _dptr(a):
    return a.ctypes.data_as(_dp)


def _check(code, handle=None):
    if code >= 0:
        return code
    msg = lib().lwsb_last_error(handle)
    msg = msg.decode() if msg else "error %d" % code
    if code == ERR_EVEN_NREAL:
        raise ValueError(msg)
    if code == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    if code == ERR_ARG:
        raise ValueError("lws_b200: " + msg)
    raise RuntimeError("lws_b200: " + msg)


def create_weights(awin, swin, fshift, L, use_summarized_weights=True):
    """lws.pyx:160-181 through the library's host-side C++ (no device needed)."""
    awin = np.ascontiguousarray(awin, dtype=np.float64)
    swin = np.ascontiguousarray(swin, dtype=np.float64)
    T = len(awin)
    qp, q = _ci(0), _ci(0)
    _check(lib().lwsb_create_weights(_dptr(awin), _dptr(swin), T, int(fshift), int(L), int(bool(use_summarized_weights)),
                                     None, None, ctypes.byref(qp), ctypes.byref(q)))
    wr = np.empty((qp.value, q.value, L + 1))
    wi = np.empty_like(wr)
    _check(lib().lwsb_create_weights(_dptr(awin), _dptr(swin), T, int(fshift), int(L), int(bool(use_summarized_weights)),
                                     _dptr(wr), _dptr(wi), None, None))
    return wr + 1j * wi


def _ptr_array(arrays):
    return (ctypes.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])


class Context(object):
    """One CUDA device + stream + resident batch (lwsb_ctx)."""

    def __init__(self, device=0, stream=None):
        h = ctypes.c_void_p()
        _check(lib().lwsb_create(int(device), ctypes.c_void_p(stream) if stream else None, ctypes.byref(h)))
        self._h = h
        self.device = int(device)
        self._wkeys = {}
        # a context is single-threaded (lws_b200.h): callers that share one (lws_b200.api keeps one per device) hold this
        # lock for the whole set_weights .. store sequence
        self.lock = threading.RLock()

    def close(self):
        if getattr(self, "_h", None):
            try:
                lib().lwsb_destroy(self._h)
            except TypeError:  # interpreter shutdown: the module globals are already gone, the driver reclaims the context
                pass
            self._h = None

    __del__ = close

    def _c(self, code):
        return _check(code, self._h)

    # -- weights ------------------------------------------------------------------------
    def set_weights(self, which, Wc):
        Wc = np.asarray(Wc)
        if Wc.ndim != 3:
            raise ValueError("weights must have shape (Qprime, Q, L+1)")
        key = (Wc.shape, Wc.tobytes())
        if self._wkeys.get(which) == key:
            return
        wr = np.ascontiguousarray(Wc.real, dtype=np.float64)
        wi = np.ascontiguousarray(Wc.imag, dtype=np.float64)
        self._c(lib().lwsb_set_weights(self._h, which, _dptr(wr), _dptr(wi), Wc.shape[0], Wc.shape[1], Wc.shape[2] - 1))
        self._wkeys[which] = key

    # -- staged interface -----------------------------------------------------------------
    def load(self, arrays, kind):
        T = np.ascontiguousarray([a.shape[0] for a in arrays], dtype=np.intc)
        self._T, self._Nreal = T, arrays[0].shape[1]
        self._c(lib().lwsb_load(self._h, _ptr_array(arrays), T.ctypes.data_as(_ip), len(arrays), self._Nreal, kind, HOST))

    def load_device(self, ptrs, T, Nreal, kind):
        T = np.ascontiguousarray(T, dtype=np.intc)
        self._T, self._Nreal = T, int(Nreal)
        arr = (ctypes.c_void_p * len(ptrs))(*ptrs)
        self._c(lib().lwsb_load(self._h, arr, T.ctypes.data_as(_ip), len(ptrs), int(Nreal), kind, DEVICE))

    @staticmethod
    def _thr(thresholds):
        t = np.ascontiguousarray(thresholds, dtype=np.float64)
        return t, (_dptr(t) if len(t) else None), len(t)

    def batch(self, thresholds, flags=0):
        t, p, n = self._thr(thresholds)
        self._c(lib().lwsb_batch(self._h, p, n, flags))

    def nofuture(self, which, thresholds, flags=0):
        t, p, n = self._thr(thresholds)
        self._c(lib().lwsb_nofuture(self._h, which, p, n, flags))

    def online(self, thresholds, look_ahead, flags=0):
        t, p, n = self._thr(thresholds)
        self._c(lib().lwsb_online(self._h, p, n, int(look_ahead), flags))

    def restage(self):
        self._c(lib().lwsb_restage(self._h))

    def store(self, outs=None):
        if outs is None:
            outs = [np.empty((int(t), self._Nreal), dtype=np.complex128) for t in self._T]
        self._c(lib().lwsb_store(self._h, _ptr_array(outs), HOST))
        return outs

    def store_device(self, ptrs):
        arr = (ctypes.c_void_p * len(ptrs))(*ptrs)
        self._c(lib().lwsb_store(self._h, arr, DEVICE))

    def sync(self):
        self._c(lib().lwsb_sync(self._h))

    # -- one-shot interface ---------------------------------------------------------------
    def _io(self, arrays, outs):
        T = np.ascontiguousarray([a.shape[0] for a in arrays], dtype=np.intc)
        if outs is None:
            outs = [np.empty(a.shape, dtype=np.complex128) for a in arrays]
        return T, outs

    def batch_lws(self, arrays, kind, thresholds, flags=0, outs=None):
        T, outs = self._io(arrays, outs)
        t, p, n = self._thr(thresholds)
        self._c(lib().lwsb_batch_lws(self._h, _ptr_array(arrays), _ptr_array(outs), T.ctypes.data_as(_ip), len(arrays),
                                     arrays[0].shape[1], kind, HOST, p, n, flags))
        return outs

    def nofuture_lws(self, which, arrays, kind, thresholds, flags=0, outs=None):
        T, outs = self._io(arrays, outs)
        t, p, n = self._thr(thresholds)
        self._c(lib().lwsb_nofuture_lws(self._h, which, _ptr_array(arrays), _ptr_array(outs), T.ctypes.data_as(_ip),
                                        len(arrays), arrays[0].shape[1], kind, HOST, p, n, flags))
        return outs

    def online_lws(self, arrays, kind, thresholds, look_ahead, flags=0, outs=None):
        T, outs = self._io(arrays, outs)
        t, p, n = self._thr(thresholds)
        self._c(lib().lwsb_online_lws(self._h, _ptr_array(arrays), _ptr_array(outs), T.ctypes.data_as(_ip), len(arrays),
                                      arrays[0].shape[1], kind, HOST, p, n, int(look_ahead), flags))
        return outs

    def run_lws(self, arrays, kind, nf_thr, on_thr, look_ahead, b_thr, flags=0, outs=None):
        T, outs = self._io(arrays, outs)
        t1, p1, n1 = self._thr(nf_thr)
        t2, p2, n2 = self._thr(on_thr)
        t3, p3, n3 = self._thr(b_thr)
        self._c(lib().lwsb_run_lws(self._h, _ptr_array(arrays), _ptr_array(outs), T.ctypes.data_as(_ip), len(arrays),
                                   arrays[0].shape[1], kind, HOST, p1, n1, p2, n2, int(look_ahead), p3, n3, flags))
        return outs

    def run_lws_device(self, in_ptrs, out_ptrs, T, Nreal, kind, nf_thr, on_thr, look_ahead, b_thr, flags=0):
        """lwsb_run_lws on device buffers (CUDA pointers as ints): nothing crosses PCIe."""
        T = np.ascontiguousarray(T, dtype=np.intc)
        self._T, self._Nreal = T, int(Nreal)
        t1, p1, n1 = self._thr(nf_thr)
        t2, p2, n2 = self._thr(on_thr)
        t3, p3, n3 = self._thr(b_thr)
        a_in = (ctypes.c_void_p * len(in_ptrs))(*in_ptrs)
        a_out = (ctypes.c_void_p * len(out_ptrs))(*out_ptrs)
        self._c(lib().lwsb_run_lws(self._h, a_in, a_out, T.ctypes.data_as(_ip), len(in_ptrs), int(Nreal), kind, DEVICE,
                                   p1, n1, p2, n2, int(look_ahead), p3, n3, flags))

    # -- streaming online_lws ---------------------------------------------------------------
    def stream_begin(self, Nreal, max_frames, kind, mean_amp, thresholds, look_ahead, flags=0):
        t, p, n = self._thr(thresholds)
        self._c(lib().lwsb_stream_begin(self._h, int(Nreal), int(max_frames), kind, float(mean_amp), p, n, int(look_ahead), flags))
        self._Nreal = int(Nreal)

    def stream_push(self, frames):
        """frames: (n, Nreal) C-contiguous float64 / complex128 (the kind the stream was opened with)"""
        self._c(lib().lwsb_stream_push(self._h, frames.ctypes.data, frames.shape[0], HOST))

    def stream_frames(self):
        a, b = _ci(0), _ci(0)
        self._c(lib().lwsb_stream_frames(self._h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def stream_read(self, first, n):
        out = np.empty((int(n), self._Nreal), dtype=np.complex128)
        self._c(lib().lwsb_stream_read(self._h, out.ctypes.data, int(first), int(n), HOST))
        return out

    def stream_end(self):
        return self._c(lib().lwsb_stream_end(self._h))

    # -- transforms -------------------------------------------------------------------------
    def stft(self, x, awin, fsize, fshift, fftsize, perfectrec):
        """x: (B, nsamples) float64 C-contiguous -> (B, M, fftsize//2+1) complex128."""
        B, n = x.shape
        M = _check(lib().lwsb_stft_frames(n, fsize, fshift, int(bool(perfectrec))))
        pre = _check(lib().lwsb_stft_prepad(fsize, fshift, int(bool(perfectrec))))
        S = np.empty((B, M, fftsize // 2 + 1), dtype=np.complex128)
        awin = np.ascontiguousarray(awin, dtype=np.float64)
        self._c(lib().lwsb_stft(self._h, x.ctypes.data, B, n, _dptr(awin), fsize, fshift, fftsize, pre, M,
                                S.ctypes.data, HOST))
        return S

    def istft(self, S, swin, fshift):
        """S: (B, M, Nreal) complex128 C-contiguous -> (B, fshift*(M-1) + 2*(Nreal-1)) float64 (uncropped)."""
        B, M, Nreal = S.shape
        swin = np.ascontiguousarray(swin, dtype=np.float64)
        out = np.empty((B, fshift * (M - 1) + 2 * (Nreal - 1)))
        self._c(lib().lwsb_istft(self._h, S.ctypes.data, B, M, Nreal, _dptr(swin), len(swin), fshift,
                                 out.ctypes.data, HOST))
        return out

    def reconstruct(self, x, awin, swin, fsize, fshift, perfectrec, nf_thr, on_thr, look_ahead, ba_thr, flags=0,
                    consistency=False):
        """x: (B, nsamples) float64 -> y = istft(run_lws(|stft(x)|)), (B, n_out) float64, in one call on the device
        (the three weight sets must have been set); with `consistency`, also the consistency in dB of each result."""
        B, n = x.shape
        pr = int(bool(perfectrec))
        ylen = _check(lib().lwsb_reconstruct_length(n, fsize, fshift, pr))
        y = np.empty((B, ylen))
        awin = np.ascontiguousarray(awin, dtype=np.float64)
        swin = np.ascontiguousarray(swin, dtype=np.float64)
        thr = [np.ascontiguousarray(t, dtype=np.float64) for t in (nf_thr, on_thr, ba_thr)]
        cons = np.empty(B) if consistency else None
        # keep a 1-element buffer alive for empty threshold lists (the pointer is not read when the count is 0)
        ptr = [_dptr(t) if len(t) else _dptr(np.zeros(1)) for t in thr]
        self._c(lib().lwsb_reconstruct(self._h, x.ctypes.data, B, n, _dptr(awin), _dptr(swin), fsize, fshift, pr,
                                       ptr[0], len(thr[0]), ptr[1], len(thr[1]), int(look_ahead), ptr[2], len(thr[2]), int(flags),
                                       y.ctypes.data, HOST, _dptr(cons) if consistency else None))
        return (y, cons) if consistency else y

    def consistency(self, S, awin, swin, fshift, perfectrec):
        """S: (B, M, Nreal) complex128 -> (B,) consistency in dB (lws.pyx:140-144), computed on the device."""
        B, M, Nreal = S.shape
        awin = np.ascontiguousarray(awin, dtype=np.float64)
        swin = np.ascontiguousarray(swin, dtype=np.float64)
        out = np.empty(B)
        self._c(lib().lwsb_consistency(self._h, S.ctypes.data, B, M, Nreal, _dptr(awin), _dptr(swin), len(swin), fshift,
                                       int(bool(perfectrec)), HOST, _dptr(out)))
        return out

    def resident_consistency(self, awin, swin, fshift, perfectrec):
        """(B,) consistency in dB of the resident batch as it stands (between batch() calls: a per-sweep trace)"""
        awin = np.ascontiguousarray(awin, dtype=np.float64)
        swin = np.ascontiguousarray(swin, dtype=np.float64)
        out = np.empty(len(self._T))
        self._c(lib().lwsb_resident_consistency(self._h, _dptr(awin), _dptr(swin), len(swin), int(fshift), int(bool(perfectrec)), _dptr(out)))
        return out

    # -- introspection ----------------------------------------------------------------------
    def last_compute_ms(self):
        ms = ctypes.c_float(0)
        self._c(lib().lwsb_last_compute_ms(self._h, ctypes.byref(ms)))
        return float(ms.value)

    def last_stage_ms(self):
        """device ms of the stages run since the last load: dict(nofuture=, online=, batch=), None for a stage not run"""
        ms = (ctypes.c_float * 3)()
        self._c(lib().lwsb_last_stage_ms(self._h, ms))
        return {k: (float(v) if v >= 0 else None) for k, v in zip(("nofuture", "online", "batch"), ms)}

    def last_online_kernel(self):
        return int(lib().lwsb_last_online_kernel(self._h))

    def last_batch_work(self):
        out = (_ll * 4)()
        self._c(lib().lwsb_last_batch_work(self._h, out))
        return dict(zip(("bin_iters_nominal", "bin_iters_executed", "work_items", "passes"), [int(x) for x in out]))

    def set_tuning(self, smem_limit=0, cluster=0, sweeps_per_pass=0):
        self._c(lib().lwsb_set_tuning(self._h, int(smem_limit), int(cluster), int(sweeps_per_pass)))

    def set_block_bins(self, bins=0):
        self._c(lib().lwsb_set_block_bins(self._h, int(bins)))

    def set_variant(self, sweep_lag=0, tensor_memory=0):
        self._c(lib().lwsb_set_variant(self._h, int(sweep_lag), int(tensor_memory)))

    def batch_trace(self, enable=True, max_items=4096):
        """Switch the strip kernel's work-item time line on / off and return the one of the last batch() call: list of
        (utterance, pass, taken, primed, computed, written [ns from the first stamp], cycles of strip 0: control lane
        waiting for rows / polling neighbours, compute warp 0 working / waiting for the control warp)."""
        up = (ctypes.c_int * (2 * max_items))()
        st = (ctypes.c_ulonglong * (8 * max_items))()
        n = self._c(lib().lwsb_last_batch_trace(self._h, 1 if enable else 0, max_items, up, st))
        if n <= 0:
            return []
        t0 = min(st[8 * i] for i in range(n))
        return [(up[2 * i], up[2 * i + 1]) + tuple(int(st[8 * i + k]) - t0 for k in range(4)) + tuple(int(st[8 * i + k]) for k in range(4, 8))
                for i in range(n)]

    def debug_fast_math(self, n, seed=1):
        """(sqrt samples checked, differing, division samples checked, differing) of the kernels' branch-free sqrt / division."""
        out = (ctypes.c_ulonglong * 4)()
        self._c(lib().lwsb_debug_fast_math(self._h, int(n), int(seed), out))
        return tuple(int(x) for x in out)

    def last_batch_cycles(self):
        out = (ctypes.c_ulonglong * 13)()
        if self._c(lib().lwsb_last_batch_cycles(self._h, out)) != 1:
            return None
        keys = ("ctrl_publish", "ctrl_poll", "ctrl_tma", "warp_work", "warp_wait_strip", "warp_wait_neighbours", "warps",
                "c_setup", "c_own_terms", "c_wait_a", "c_chain_a", "c_wait_b", "c_chain_b")
        return dict(zip(keys, [int(x) for x in out]))

    def last_batch_plan(self):
        """dict describing the cluster strip plan of the last batch() call, or None (generic kernel)."""
        out = (ctypes.c_int * 16)()
        if self._c(lib().lwsb_last_batch_plan(self._h, out)) != 1:
            return None
        return dict(zip(PLAN_KEYS, list(out)))

    def launch_count(self):
        return int(lib().lwsb_launch_count(self._h))

    def device_info(self):
        sm, ma, mi, mem = _ci(0), _ci(0), _ci(0), _ll(0)
        self._c(lib().lwsb_device_info(self._h, ctypes.byref(sm), ctypes.byref(ma), ctypes.byref(mi), ctypes.byref(mem)))
        return dict(sm_count=sm.value, cc=(ma.value, mi.value), hbm_bytes=mem.value)

    def stats(self):
        B = len(self._T)
        mean, mx = np.empty(B), np.empty(B)
        self._c(lib().lwsb_get_stats(self._h, _dptr(mean), _dptr(mx)))
        return mean, mx


# ---- host-only mirrors of the device schedule (tests) ----------------------------------------
def debug_terms(Wc, fold, rframe, cframe, p):
    Wc = np.asarray(Wc)
    Q, L = Wc.shape[1], Wc.shape[2] - 1
    wr = np.ascontiguousarray(Wc.real)
    wi = np.ascontiguousarray(Wc.imag)
    mx = (2 * Q - 1) * (2 * L + 1)
    dr, dk = np.zeros(mx, dtype=np.intc), np.zeros(mx, dtype=np.intc)
    cr, ci = np.zeros(mx), np.zeros(mx)
    n = _check(lib().lwsb_debug_terms(_dptr(wr), _dptr(wi), Q, L, fold, rframe, cframe, p, mx, dr.ctypes.data_as(_ip),
                                      dk.ctypes.data_as(_ip), _dptr(cr), _dptr(ci)))
    return dr[:n], dk[:n], cr[:n] + 1j * ci[:n]


PLAN_KEYS = ("cluster", "blocks_per_strip", "virtual_blocks", "frame_slots", "sweeps_per_pass", "ring_rows",
             "ring_pitch", "threads", "smem_bytes", "sweep_lag", "sweep_fastest", "tensor_memory", "block_bins", "sweep_extra_from",
             "load_lead")


def debug_work_items(active_sweeps, sweeps_per_pass):
    """The strip kernel's work list [(utterance, pass), ...] for the given numbers of active sweeps per utterance."""
    a = (ctypes.c_int * len(active_sweeps))(*[int(x) for x in active_sweeps])
    n = _check(lib().lwsb_debug_work_items(a, len(active_sweeps), int(sweeps_per_pass), 0, None))
    out = (ctypes.c_int * (2 * max(n, 1)))()
    _check(lib().lwsb_debug_work_items(a, len(active_sweeps), int(sweeps_per_pass), n, out))
    return [(out[2 * i], out[2 * i + 1]) for i in range(n)]


def debug_plan_strips(Nreal, Q, L, iterations, maxT, B, smem_limit=232448, sm_count=148, cluster=0, sweeps=0, block=0):
    out = (ctypes.c_int * 16)()
    if _check(lib().lwsb_debug_plan_strips(Nreal, Q, L, iterations, maxT, B, smem_limit, sm_count, cluster, sweeps, block,
                                           out)) != 1:
        return None
    return dict(zip(PLAN_KEYS, list(out)))


def debug_online_chain(T, iterations, look_ahead, Q):
    n = int(lib().lwsb_debug_online_chain_length(T, iterations, look_ahead))
    out = []
    v = [_ci(0) for _ in range(5)]
    for j in range(n):
        _check(lib().lwsb_debug_online_task(T, iterations, look_ahead, Q, j, *[ctypes.byref(x) for x in v]))
        out.append(tuple(x.value for x in v))
    return out

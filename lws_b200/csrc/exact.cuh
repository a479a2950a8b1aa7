// exact.cuh -- the reference's per-bin arithmetic, operation for operation.
//
// Why bit-exactness: the LWS iteration keeps symmetric configurations (the real-valued DC and
// Nyquist bins, the zero-phase start) alive only through *exact* cancellations of its paired
// terms  ar*(bi+ci)+ai*(br-cr)  (lwslib.cpp:98-99); those configurations are unstable, so an
// implementation whose rounding differs by one ulp anywhere drifts to a different solution
// within ~50 sweeps (measured: 1e-2 rel-Frobenius on BASELINE configs).  Parity
// with the CPU reference therefore needs the same IEEE operations in the same order:
//   * every product and sum rounded separately -- the reference's x86-64 build has no FMA --
//     hence the explicit __dadd_rn / __dmul_rn intrinsics, which nvcc never contracts;
//   * terms accumulated in the order of the reference's loops for each variant
//     (LWSQ2 / LWSQ4 / LWSanyQ, lwslib.cpp:72-373; NoFuture_*, 473-690; Asym_UpdatePhase*,
//     776-1273), selected by (fold, rframe, cframe) exactly as the C variants differ;
//   * |t| = sqrt(t.r*t.r + t.i*t.i) and the normalisation (t * a) / |t| (lwslib.cpp:355-360);
//     sqrt and division are correctly rounded on both machines.
// The accessor E(dr, dk) returns cell (frame m + dr, bin n + dk) of the extended spectrogram,
// wherever the calling kernel keeps it (global memory or a shared-memory ring).
#pragma once
#include <cuda_runtime.h>
#include "fast_math.cuh"
#include "lwsb_common.h"

namespace lwsb {

struct LwsbW {        // one weight set in the reference's layout (Qprime, Q, L+1)
    const double *wr, *wi;
    const int *wf;    // |W| > 1e-12 (lws.pyx:231-232)
    int frac_rows;    // 0: summarised weights (Qprime = Q, row = bin mod Q); N = 2(Nreal-1): one row per FFT bin (the
                      // reference's *fractionalQ variants), the device table has N + 1 rows, row N zero / mask clear
};

// acc += w*b + conj(w)*c  (lwslib.cpp:98-99)
__device__ __forceinline__ void x_pair(double &tr, double &ti, double ar, double ai, double br, double bi, double cr, double ci)
{
    tr = __dadd_rn(tr, __dsub_rn(__dmul_rn(ar, __dadd_rn(br, cr)), __dmul_rn(ai, __dsub_rn(bi, ci))));
    ti = __dadd_rn(ti, __dadd_rn(__dmul_rn(ar, __dadd_rn(bi, ci)), __dmul_rn(ai, __dsub_rn(br, cr))));
}
// acc += w*b  (lwslib.cpp:500-501)
__device__ __forceinline__ void x_one(double &tr, double &ti, double ar, double ai, double br, double bi)
{
    tr = __dadd_rn(tr, __dsub_rn(__dmul_rn(ar, br), __dmul_rn(ai, bi)));
    ti = __dadd_rn(ti, __dadd_rn(__dmul_rn(ar, bi), __dmul_rn(ai, br)));
}
// acc += conj(w)*c  (lwslib.cpp:667-668)
__device__ __forceinline__ void x_one_conj(double &tr, double &ti, double ar, double ai, double cr, double ci)
{
    tr = __dadd_rn(tr, __dadd_rn(__dmul_rn(ar, cr), __dmul_rn(ai, ci)));
    ti = __dadd_rn(ti, __dsub_rn(__dmul_rn(ar, ci), __dmul_rn(ai, cr)));
}

// frames m-r and m+r, both sides (lwslib.cpp:105-132, 187-259, 316-353)
template <class Acc>
__device__ __forceinline__ void x_both(const Acc &E, const LwsbW &w, int L, int r, int wp, int wpn, int fold, bool minus,
                                       double &tr, double &ti)
{
    const int u = wp + r * (L + 1);
    if (w.wf[u]) {
        const double2 b = E(-r, 0), c = E(+r, 0);
        x_pair(tr, ti, w.wr[u], w.wi[u], b.x, b.y, c.x, c.y);
    }
    if (fold == LWSB_FOLD_ANY) {
        const int un = wpn + r * (L + 1);
        for (int k = 1; k <= L; ++k) {
            if (w.wf[u + k]) {
                const double2 b = E(-r, -k), c = E(+r, -k);
                x_pair(tr, ti, w.wr[u + k], w.wi[u + k], b.x, b.y, c.x, c.y);
            }
            if (w.wf[un + k]) {
                const double2 b = E(+r, +k), c = E(-r, +k);
                x_pair(tr, ti, w.wr[un + k], w.wi[un + k], b.x, b.y, c.x, c.y);
            }
        }
    } else {
        for (int k = 1; k <= L; ++k)
            if (w.wf[u + k]) {
                const double2 e1 = E(-r, -k), e2 = E(+r, +k), e3 = E(+r, -k), e4 = E(-r, +k);
                double br, bi, cr, ci;
                if (minus) { // lwslib.cpp:204-207
                    br = __dsub_rn(e1.x, e2.x); bi = __dsub_rn(e1.y, e2.y);
                    cr = __dsub_rn(e3.x, e4.x); ci = __dsub_rn(e3.y, e4.y);
                } else {     // lwslib.cpp:123-126
                    br = __dadd_rn(e1.x, e2.x); bi = __dadd_rn(e1.y, e2.y);
                    cr = __dadd_rn(e3.x, e4.x); ci = __dadd_rn(e3.y, e4.y);
                }
                x_pair(tr, ti, w.wr[u + k], w.wi[u + k], br, bi, cr, ci);
            }
    }
}

// frame m-r only (lwslib.cpp:491-516, 643-671, 862-878, 986-1001)
template <class Acc>
__device__ __forceinline__ void x_left(const Acc &E, const LwsbW &w, int L, int r, int wp, int wpn, int fold, bool minus,
                                       double &tr, double &ti)
{
    const int u = wp + r * (L + 1);
    if (w.wf[u]) {
        const double2 b = E(-r, 0);
        x_one(tr, ti, w.wr[u], w.wi[u], b.x, b.y);
    }
    if (fold == LWSB_FOLD_ANY) {
        const int un = wpn + r * (L + 1);
        for (int k = 1; k <= L; ++k) {
            if (w.wf[u + k]) {
                const double2 b = E(-r, -k);
                x_one(tr, ti, w.wr[u + k], w.wi[u + k], b.x, b.y);
            }
            if (w.wf[un + k]) {
                const double2 c = E(-r, +k);
                x_one_conj(tr, ti, w.wr[un + k], w.wi[un + k], c.x, c.y);
            }
        }
    } else {
        for (int k = 1; k <= L; ++k)
            if (w.wf[u + k]) {
                const double2 b = E(-r, -k);
                double2 c = E(-r, +k);
                if (minus) { c.x = -c.x; c.y = -c.y; } // lwslib.cpp:994-997
                x_pair(tr, ti, w.wr[u + k], w.wi[u + k], b.x, b.y, c.x, c.y);
            }
    }
}

// The weighted sum of one bin: p = bin mod Q.  (rframe, cframe): batch sweep = (Q, 1), no-future
// sweep = (1, 0), online row updates as Asym_UpdatePhase* derives them (lwslib.cpp:1141-1151).
// The `update == 1` branch of Asym_UpdatePhase* (centre-bin term) is not restated: both
// reference bindings pass update = 2 (lws.pyx:363, online_lws.cpp:160).
// `bin` is the bin index n - L: the weight rows are bin mod Q and (Q - bin mod Q) mod Q, or -- per-frequency tables,
// lwslib.cpp:393, 408 -- bin and N - bin (row N, which the reference reads one past its table at the DC bin, is the
// zero row the host appends: those terms are skipped).
template <class Acc>
__device__ __forceinline__ void x_weighted_sum(const Acc &E, const LwsbW &w, int Q, int L, int bin, int fold, int rframe,
                                               int cframe, double &tr, double &ti)
{
    tr = 0.0; ti = 0.0;
    const int p = bin % Q;
    const int wp = (w.frac_rows ? bin : p) * Q * (L + 1);
    const int wpn = (w.frac_rows ? w.frac_rows - bin : (Q - p) % Q) * Q * (L + 1);
    if (cframe)
        for (int k = 1; k <= L; ++k)
            if (w.wf[wp + k]) {
                const double2 b = E(0, -k), c = E(0, +k);
                x_pair(tr, ti, w.wr[wp + k], w.wi[wp + k], b.x, b.y, c.x, c.y);
            }
    if (fold == LWSB_FOLD_Q4 && (p & 1)) {
        // odd bins: odd frames first with the sign-flipped folding, then r = 2 (lwslib.cpp:186-235, 953-1052)
        for (int r = 1; r < Q; r += 2) {
            if (r < rframe) x_both(E, w, L, r, wp, wpn, fold, true, tr, ti);
            else x_left(E, w, L, r, wp, wpn, fold, true, tr, ti);
        }
        if (2 < rframe) x_both(E, w, L, 2, wp, wpn, fold, false, tr, ti);
        else x_left(E, w, L, 2, wp, wpn, fold, false, tr, ti);
    } else {
        for (int r = 1; r < rframe; ++r) x_both(E, w, L, r, wp, wpn, fold, false, tr, ti);
        for (int r = rframe; r < Q; ++r) x_left(E, w, L, r, wp, wpn, fold, false, tr, ti);
    }
}

// lwslib.cpp:355-360: returns false when |t| == 0 (the bin keeps its value).  sqrt and the two divisions are
// correctly rounded, like the reference's; they are computed without the library functions' slow-path branches
// (fast_math.cuh: one basic block, one reciprocal for both divisions), the rare operand outside the fast ranges
// falls back to the library functions.
__device__ __forceinline__ bool x_project(double tr, double ti, double a, double2 &out)
{
    const double x = __dadd_rn(__dmul_rn(tr, tr), __dmul_rn(ti, ti));
    const double nr = __dmul_rn(tr, a), ni = __dmul_rn(ti, a);
    bool sok, rok, dok1, dok2;
    double mag = fm_sqrt(x, sok);
    const double rcp = fm_rcp(mag, rok);
    out.x = fm_div(nr, mag, rcp, dok1);
    out.y = fm_div(ni, mag, rcp, dok2);
    if (!(x == 0.0) && !(sok && rok && dok1 && dok2)) {
        mag = __dsqrt_rn(x);
        out.x = __ddiv_rn(nr, mag); out.y = __ddiv_rn(ni, mag);
    }
    return x > 0.0; // sqrt(x) > 0 iff x > 0
}

// x_project with the range checks of the fast paths replaced by a SUFFICIENT condition on the inputs, known before the square
// root starts (so that it is not between the last division and the store): x in [2^-200, 2^900) and each numerator zero or
// in [2^-500, 2^900) in magnitude => |temp| in [2^-100, 2^450], quotients in [2^-950, 2^1000]: every check of fm_sqrt /
// fm_rcp / fm_div passes.  Anything else (denormal-scale or huge spectra) takes the library functions.  Same bits either way.
__device__ __forceinline__ bool x_project_early(double tr, double ti, double a, double2 &out)
{
    const double x = __dadd_rn(__dmul_rn(tr, tr), __dmul_rn(ti, ti));
    const double nr = __dmul_rn(tr, a), ni = __dmul_rn(ti, a);
    constexpr unsigned LO_X = (unsigned)(1023 - 200) << 20, HI = (unsigned)(1023 + 900) << 20, LO_N = (unsigned)(1023 - 500) << 20;
    const unsigned hx = (unsigned)__double2hiint(x), hr = (unsigned)__double2hiint(nr) & 0x7fffffffu, hi_ = (unsigned)__double2hiint(ni) & 0x7fffffffu;
    const bool fast = (hx - LO_X) < (HI - LO_X) && ((hr - LO_N) < (HI - LO_N) || nr == 0.0) && ((hi_ - LO_N) < (HI - LO_N) || ni == 0.0);
    if (fast) {
        bool o1, o2, o3, o4;
        const double mag = fm_sqrt(x, o1);
        const double rcp = fm_rcp(mag, o2);
        out.x = fm_div(nr, mag, rcp, o3);
        out.y = fm_div(ni, mag, rcp, o4);
        return true; // x >= 2^-200 > 0
    }
    const double mag = __dsqrt_rn(x);
    out.x = __ddiv_rn(nr, mag); out.y = __ddiv_rn(ni, mag);
    return x > 0.0;
}

// |z| as numpy computes it for a contiguous complex128 array on an FMA-capable x86-64
// (numpy >= 1.25 CDOUBLE_absolute, SIMD path): max * sqrt(fma(q, q, 1)), q = min / max.
// np.abs(ExtS) / np.abs(S) of lws.pyx:239-240 go through that loop; the oracle (numpy on the
// same host) is what tests compare against.
__device__ __forceinline__ double x_cabs(double re, double im)
{
    re = fabs(re); im = fabs(im);
    const double mx = fmax(re, im), mn = fmin(re, im);
    if (mx == 0.0) return 0.0;
    if (isinf(mx)) return mx;
    const double q = __ddiv_rn(mn, mx);
    return __dmul_rn(mx, __dsqrt_rn(__fma_rn(q, q, 1.0)));
}

} // namespace lwsb

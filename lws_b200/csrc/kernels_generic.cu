// kernels_generic.cu -- layout kernels (extend / amplitude / crop) and the generic wavefront
// kernels that execute ANY (Q, L, weight pattern) exactly in the reference's sequential order.
//
// Why a wavefront: the reference updates bins in place in raster order (frame-major,
// bin-minor; lwslib.cpp:291-292), i.e. Gauss-Seidel.  A bin (m, c) reads frames m-r at bins
// up to c+L (already updated) and frames m+r from c-L on (not yet updated).  Running row
// update number j of a chain at "time" c + (L+1)*j therefore reproduces the sequential result
// exactly; successive sweeps may follow each other Q frames apart.
// The generic kernels run one CTA per utterance in lock step (one bin per row update per
// step, __syncthreads between steps) with the state in global memory.  Every bin is computed
// with the reference's own operation sequence (exact.cuh), so results are bit-identical to
// the CPU reference.  They serve every configuration; the tuned batch kernel lives in
// kernels_batch.cu.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdlib>
#include <vector>
#include "lwsb_common.h"
#include "kernels.h"
#include "exact.cuh"

namespace lwsb {

// ------------------------------------------------------------------------------------------
// extend + amplitude + per-frame statistics   (lws.pyx:146-157, 235-240; lwslib.cpp:15-65)
// grid (max Tp, B), one CTA per extended row.
template <int KIND>
__global__ void k_extend(LwsbView v, const void *const *src, double *row_max)
{
    const int u = blockIdx.y;
    const int T = v.T[u];
    const int m = blockIdx.x;
    const int Tp = T + 2 * (v.Q - 1);
    if (m >= Tp) return;
    const int Nreal = v.Nreal, L = v.L;
    int p = m - (v.Q - 1);
    p = p < 0 ? 0 : (p > T - 1 ? T - 1 : p);
    const long long row = v.rowbase[u] + m;
    double2 *E = v.E + row * v.P;
    double *A = v.A + row * v.P;
    const int e0 = v.c0 - L; // physical column of extended column 0
    double mx = 0.0;
    for (int x = threadIdx.x; x < v.P; x += blockDim.x) {
        const int c = x - v.c0; // bin index, may be a mirrored one
        double2 val = make_double2(0.0, 0.0);
        double a = 0.0;
        if (x >= e0 && c < Nreal + L) {
            int cs = c;
            bool cj = false;
            if (c < 0) { cs = -c; cj = true; }
            else if (c >= Nreal) { cs = 2 * (Nreal - 1) - c; cj = true; }
            if (KIND == 0) {
                const double2 *S = reinterpret_cast<const double2 *>(src[u]);
                val = S[(long long)p * Nreal + cs];
                a = x_cabs(val.x, val.y); // np.abs(ExtS), lws.pyx:239
            } else {
                const double *S = reinterpret_cast<const double *>(src[u]);
                val = make_double2(S[(long long)p * Nreal + cs], 0.0);
                a = fabs(val.x);
            }
            if (cj) val.y = -val.y;
            if (c >= 0 && c < Nreal) mx = fmax(mx, a);
        }
        E[x] = val;
        A[x] = a;
    }
    // per-frame maximum of |S| (the mean is summed separately, in numpy's order: k_stats)
    __shared__ double sh_m[32];
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
    const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if ((threadIdx.x & 31) == 0) sh_m[w] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tm = 0.0;
        for (int i = 0; i < nw; ++i) tm = fmax(tm, sh_m[i]);
        row_max[row] = tm;
    }
}

// Streaming: the same rows for frames [m0, m0 + n) of utterance 0 as they arrive -- src holds the n new
// frames -- plus, with the first frame, the Q-1 frozen ghost rows above it (copies of frame 0, lws.pyx:155).  The ghost rows
// below the last frame are never read by the online chain (every row update uses frames up to the newest one only).
// grid n (+ Q-1 when m0 == 0)
template <int KIND>
__global__ void k_stream_extend(LwsbView v, const void *src, int m0, int n)
{
    const int ghosts = m0 == 0 ? v.Q - 1 : 0;
    const int idx = blockIdx.x;
    const int f = idx < ghosts ? 0 : idx - ghosts;              // frame inside src
    const long long row = v.rowbase[0] + (idx < ghosts ? idx : m0 + (v.Q - 1) + f);
    const int Nreal = v.Nreal, L = v.L;
    double2 *E = v.E + row * v.P;
    double *A = v.A + row * v.P;
    const int e0 = v.c0 - L;
    for (int x = threadIdx.x; x < v.P; x += blockDim.x) {
        const int c = x - v.c0;
        double2 val = make_double2(0.0, 0.0);
        double a = 0.0;
        if (x >= e0 && c < Nreal + L) {
            int cs = c;
            bool cj = false;
            if (c < 0) { cs = -c; cj = true; }
            else if (c >= Nreal) { cs = 2 * (Nreal - 1) - c; cj = true; }
            if (KIND == 0) {
                val = reinterpret_cast<const double2 *>(src)[(long long)f * Nreal + cs];
                a = x_cabs(val.x, val.y);
            } else {
                val = make_double2(reinterpret_cast<const double *>(src)[(long long)f * Nreal + cs], 0.0);
                a = fabs(val.x);
            }
            if (cj) val.y = -val.y;
        }
        E[x] = val;
        A[x] = a;
    }
}

// mean / max of |S| over the un-extended spectrogram (lws.pyx:240); grid B.
// `mean_amp = np.mean(np.abs(S))` is numpy's *pairwise* summation of the row-major flattened
// array (numpy/_core/src/umath/loops_utils.h.src, DOUBLE_pairwise_sum: blocks of <= 128
// elements summed with 8 interleaved accumulators, halves split at a multiple of 8) divided by
// the element count.  The thresholds are multiples of this number and the comparison
// `absspec > threshold` (lwslib.cpp:296) decides which bins move, so the sum is reproduced
// in the same order: the host walks the recursion once per array length (stat_tree) and emits the
// leaf blocks plus, per leaf, how many pending partial sums to combine after it; all threads sum
// leaves; thread 0 combines.
__global__ void __launch_bounds__(1024)
k_stats(LwsbView v, const double *row_max, double *mean_amp, double *max_amp, const int *leaf_tab, const int2 *tab_of, double *leaf_sum,
        long long scratch_stride)
{
    const int u = blockIdx.x;
    const int T = v.T[u];
    const long long row0 = v.rowbase[u] + (v.Q - 1);
    const long long n = (long long)T * v.Nreal;
    const int *tab = leaf_tab + 3 * (size_t)tab_of[u].x; // (offset, length, adds-after) per leaf: built on the host (stat_tree)
    const int n_leaves = tab_of[u].y;
    double *ls = leaf_sum + (size_t)u * scratch_stride;
    __shared__ double sh_m[32];
    for (int l = threadIdx.x; l < n_leaves; l += blockDim.x) {
        const long long lo = tab[3 * l];
        const int m = tab[3 * l + 1];
        // walk the leaf with a running (frame, bin) position instead of dividing per element
        long long fr = lo / v.Nreal;
        int bin = (int)(lo - fr * v.Nreal);
        const double *Arow = v.A + (row0 + fr) * v.P + v.c0;
        auto next = [&]() {
            const double a = Arow[bin];
            if (++bin == v.Nreal) { bin = 0; Arow += v.P; }
            return a;
        };
        double res;
        if (m < 8) {
            res = 0.0;
            for (int i = 0; i < m; ++i) res = __dadd_rn(res, next());
        } else {
            double r[8];
            for (int j = 0; j < 8; ++j) r[j] = next();
            int i = 8;
            for (; i < m - (m % 8); i += 8)
                for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], next());
            res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                            __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
            for (; i < m; ++i) res = __dadd_rn(res, next());
        }
        ls[l] = res;
    }
    double mx = 0.0;
    for (int m = threadIdx.x; m < T; m += blockDim.x) mx = fmax(mx, row_max[row0 + m]);
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
    const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if ((threadIdx.x & 31) == 0) sh_m[w] = mx;
    __syncthreads();
    // the leaf sums are combined in the order of numpy's recursion by ONE thread (a dependent chain); the leaf sums and the
    // additions-after counts reach it through shared memory in chunks (from global memory the walk cost an L2 round trip per leaf:
    // 0.3 ms at 2 500 leaves, 6 ms at the 45 000 of a 30 s utterance at 2048 / 256)
    constexpr int CH = 2048;
    __shared__ double sh_leaf[CH];
    __shared__ int sh_adds[CH];
    __shared__ double sh_st[64];
    __shared__ int sh_sp;
    if (threadIdx.x == 0) sh_sp = 0;
    for (int l0 = 0; l0 < n_leaves; l0 += CH) {
        const int cnt = min(CH, n_leaves - l0);
        __syncthreads(); // the leaf sums of this CTA are written (first pass) / the previous chunk is consumed
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) { sh_leaf[i] = ls[l0 + i]; sh_adds[i] = tab[3 * (l0 + i) + 2]; }
        __syncthreads();
        if (threadIdx.x == 0) {
            int sp = sh_sp;
            for (int i = 0; i < cnt; ++i) {
                sh_st[sp++] = sh_leaf[i];
                for (int a = sh_adds[i]; a > 0; --a) { sh_st[sp - 2] = __dadd_rn(sh_st[sp - 2], sh_st[sp - 1]); --sp; }
            }
            sh_sp = sp;
        }
    }
    if (threadIdx.x == 0) {
        mean_amp[u] = __ddiv_rn(sh_st[0], (double)n); // umr_sum(...) / count  (numpy _methods._mean)
        double tm = 0.0;
        for (int i = 0; i < nw; ++i) tm = fmax(tm, sh_m[i]);
        max_amp[u] = tm;
    }
}

// Re-derive the frozen ghost frames from the current first / last frame: what the next
// reference call's extspec() does when run_lws chains the three stages (lws.pyx:155-156).
// grid (2*(Q-1), B)
__global__ void k_refresh_ghosts(LwsbView v)
{
    const int u = blockIdx.y;
    const int T = v.T[u], Q = v.Q;
    const int g = blockIdx.x; // 0..Q-2 top ghosts, Q-1..2Q-3 bottom ghosts
    const long long base = v.rowbase[u];
    const long long dst = g < Q - 1 ? base + g : base + (T + Q - 1) + (g - (Q - 1));
    const long long src = g < Q - 1 ? base + (Q - 1) : base + (T + Q - 2);
    for (int x = threadIdx.x; x < v.P; x += blockDim.x) {
        v.E[dst * v.P + x] = v.E[src * v.P + x];
        v.A[dst * v.P + x] = v.A[src * v.P + x];
    }
}

// Recompute |E| and the per-frame statistics of the real frames from the CURRENT values: the
// `AmpSpec = np.abs(ExtS)` / `mean_amp` of the next chained reference call (lws.pyx:239-240).
// grid (max Tp, B)
__global__ void k_reamp(LwsbView v, double *row_max)
{
    const int u = blockIdx.y;
    const int T = v.T[u];
    const int m = blockIdx.x;
    if (m < v.Q - 1 || m >= T + v.Q - 1) return;
    const long long row = v.rowbase[u] + m;
    const double2 *E = v.E + row * v.P;
    double *A = v.A + row * v.P;
    const int e0 = v.c0 - v.L, Np = v.Nreal + 2 * v.L;
    double mx = 0.0;
    for (int e = threadIdx.x; e < Np; e += blockDim.x) {
        const double2 val = E[e0 + e];
        const double a = x_cabs(val.x, val.y);
        A[e0 + e] = a;
        const int c = e - v.L;
        if (c >= 0 && c < v.Nreal) mx = fmax(mx, a);
    }
    __shared__ double sh_m[32];
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
    const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if ((threadIdx.x & 31) == 0) sh_m[w] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tm = 0.0;
        for (int i = 0; i < nw; ++i) tm = fmax(tm, sh_m[i]);
        row_max[row] = tm;
    }
}

// crop + recombine (lws.pyx:256; lwslib.cpp:47-57); grid (max T, B)
__global__ void k_crop(LwsbView v, void *const *dst)
{
    const int u = blockIdx.y;
    const int m = blockIdx.x;
    if (m >= v.T[u]) return;
    const double2 *E = v.E + (v.rowbase[u] + m + v.Q - 1) * v.P + v.c0;
    double2 *out = reinterpret_cast<double2 *>(dst[u]) + (long long)m * v.Nreal;
    for (int c = threadIdx.x; c < v.Nreal; c += blockDim.x) out[c] = E[c];
}

// ------------------------------------------------------------------------------------------
// one bin update on the global-memory state, reference arithmetic (exact.cuh)
struct GlobalCell {
    const double2 *ctr; // &E[m][n]
    int P;
    __device__ __forceinline__ double2 operator()(int dr, int dk) const { return ctr[(long long)dr * P + dk]; }
};

__device__ __forceinline__ void update_bin(const LwsbView &v, const LwsbW &w, int fold, int rframe, int cframe,
                                           long long row, int c, double thr)
{
    double2 *Erow = v.E + row * v.P;
    const double a = v.A[row * v.P + v.c0 + c];
    if (!(a > thr)) return; // lwslib.cpp:295-296
    const GlobalCell cell{Erow + v.c0 + c, v.P};
    double tr, ti;
    x_weighted_sum(cell, w, v.Q, v.L, c, fold, rframe, cframe, tr, ti);
    double2 val;
    if (x_project(tr, ti, a, val)) {
        // lwslib.cpp:356-368: store, then refresh the mirrored copies at once
        Erow[v.c0 + c] = val;
        const int Nreal = v.Nreal;
        if (c >= 1 && c <= v.L) Erow[v.c0 - c] = make_double2(val.x, -val.y);
        else if (c >= Nreal - 1 - v.L && c <= Nreal - 2) Erow[v.c0 + 2 * (Nreal - 1) - c] = make_double2(val.x, -val.y);
    }
}

// ------------------------------------------------------------------------------------------
// `iters` pipelined sweeps (batch: rframe Q / cframe 1; no-future: rframe 1 / cframe 0).
// Sweep i, frame m, bin c runs at step  c + (L+1)*(m + Q*i).  One CTA per utterance.
__global__ void __launch_bounds__(1024)
k_sweeps_generic(LwsbView v, LwsbW w, int fold, int rframe, int cframe, const double *thresholds, int iters)
{
    const int u = blockIdx.x;
    const int T = v.T[u], Q = v.Q, Nreal = v.Nreal;
    const int S = v.L + 1;
    const long long row0 = v.rowbase[u] + (Q - 1);
    const double mean = v.mean_amp[u];
    const long long tmax = (long long)S * ((T - 1) + (long long)Q * (iters - 1)) + (Nreal - 1);
    for (long long t = 0; t <= tmax; ++t) {
        const long long vhi = t / S;
        const long long vlo = t < Nreal ? 0 : (t - (Nreal - 1) + S - 1) / S;
        const int nv = (int)(vhi - vlo + 1);
        long long imin = vlo - (T - 1);
        imin = imin <= 0 ? 0 : (imin + Q - 1) / Q;
        long long imax = vhi / Q;
        if (imax > iters - 1) imax = iters - 1;
        const int ni = (int)(imax - imin + 1);
        for (int idx = threadIdx.x; idx < nv * ni; idx += blockDim.x) {
            const long long vv = vlo + idx % nv;
            const int i = (int)imin + idx / nv;
            const long long m = vv - (long long)Q * i;
            if (m < 0 || m >= T) continue;
            const int c = (int)(t - S * vv);
            update_bin(v, w, fold, rframe, cframe, row0 + m, c, __dmul_rn(thresholds[i], mean)); // lws.pyx:245
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// online (TF-RTISI-LA) chain: row update j of the chain, bin c runs at step c + (L+1)*j.
// One CTA per utterance; thread k serves the chain positions j == k (mod blockDim).
// [j0, j1) restricts the launch to a range of chain positions (streaming: the positions of the frames just pushed; the
// earlier ones are complete, the later ones have not begun); j1 < 0 = the whole chain of the utterance's T frames.
__global__ void __launch_bounds__(1024)
k_online_generic(LwsbView v, LwsbW w0, LwsbW w_ai, LwsbW w_af, int fold, const double *thresholds, int iters, int LA,
                 long long j0, long long j1)
{
    const int u = blockIdx.x;
    const int T = v.T[u], Q = v.Q, Nreal = v.Nreal;
    const int S = v.L + 1;
    const long long base = v.rowbase[u];
    const double mean = v.mean_amp[u];
    const long long n = j1 >= 0 ? j1 : lwsb_online_chain_len(T, iters, LA);
    const long long tmax = S * (n - 1) + (Nreal - 1);
    const int nt = blockDim.x; // >= ceil(Nreal / S) + 1 (launch_online_generic guarantees)
    long long jc = -1;
    LwsbOnlineTask task;
    double thr = 0.0;
    for (long long t = S * j0; t <= tmax; ++t) {
        const long long jhi = t / S;
        // the unique j <= jhi with j == tid (mod nt) and j > jhi - nt
        const long long d = (jhi - threadIdx.x) % nt;
        const long long j = jhi - (d < 0 ? d + nt : d);
        if (j >= j0 && j < n) {
            const long long c = t - S * j;
            if (c < Nreal) {
                if (j != jc) {
                    jc = j;
                    task = lwsb_online_decode(T, iters, LA, Q, j);
                    thr = task.thr < 0 ? 0.0 : __dmul_rn(thresholds[task.thr], mean); // lws.pyx:361, lwslib.cpp:1467
                }
                update_bin(v, task.which == 0 ? w0 : (task.which == 1 ? w_ai : w_af), fold, task.rframe, task.cframe,
                           base + task.row, (int)c, thr);
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// NoFuture_LWSQ4 as the reference computes it (lwslib.cpp:538-617).  `im` already contains the
// bin index and the reads add it again, so bin n of frame m reads the flat offsets
// (m-r)*Np + 2n +- k: frames m-3 .. m-1 at doubled bin positions and, for the upper half of the
// spectrum, the already-updated part of frame m itself (mirror cells included).
//
// This map is numerically *expanding* (a 1e-16 perturbation grows by ~2x per frame: measured,
// DESIGN.md), so parity needs the reference's arithmetic bit for bit: every product and sum is
// rounded separately (no FMA contraction -- the reference's x86-64 build has none), terms are
// added in the reference's order (r = 3,2,1; k = 1..L, then k = 0), |.| is
// sqrt(x*x + y*y) and the normalisation is (t * a) / |t|.
//
// Schedule: one CTA per utterance, frames in order.  Inside a frame bin n depends on the
// current frame only through columns <= 2n + L - Np (and, through the mirror cells, on bins
// <= 2L), so the frame is processed in waves: first every bin below (Np-L)/2 at once, then the
// remaining distance to the end of the frame halves with every wave (~log2 Np waves per frame).
__device__ __forceinline__ double2 nf4_load(const double2 *E0, int P, int Np, int row, int off)
{
    // flat offset row*Np + off of the reference layout, 0 <= off < 2*Np
    if (off >= Np) { off -= Np; row += 1; }
    return E0[(long long)row * P + off];
}

__global__ void __launch_bounds__(512)
k_nofuture_q4(LwsbView v, const double *wr, const double *wi, const int *wf, const double *thresholds, int iters)
{
    const int u = blockIdx.x;
    const int T = v.T[u], Nreal = v.Nreal, L = v.L;
    constexpr int Q = 4;
    const int Np = Nreal + 2 * L, Naux = Nreal + L - 1, P = v.P;
    const double mean = v.mean_amp[u];
    double2 *E0 = v.E + v.rowbase[u] * (long long)P + (v.c0 - L); // extended (row 0, column 0)
    const double *A0 = v.A + v.rowbase[u] * (long long)P + (v.c0 - L);
    const int lim1 = (Np - L + 1) / 2; // first n whose reads reach the current frame
    for (int it = 0; it < iters; ++it) {
        const double thr = __dmul_rn(thresholds[it], mean); // lws.pyx:298
        for (int m = Q - 1; m < T + Q - 1; ++m) {
            int done = L;
            while (done <= Naux) {
                int hi = (2 * L < done) ? (done + Np - L + 1) / 2 : lim1;
                if (hi < done + 1) hi = done + 1;
                if (hi > Naux + 1) hi = Naux + 1;
                for (int n = done + threadIdx.x; n < hi; n += blockDim.x) {
                    const double a = A0[(long long)m * P + n];
                    if (!(a > thr)) continue;
                    double tr = 0.0, ti = 0.0;
                    const int wp = ((n - L) % Q) * Q * (L + 1);
                    for (int r = Q - 1; r > 0; --r) {
                        const int wu = wp + r * (L + 1);
                        const bool minus = ((n - L) & 1) && (r & 1);
                        for (int k = 1; k <= L; ++k) {
                            if (!wf[wu + k]) continue;
                            const double ar = wr[wu + k], ai = wi[wu + k];
                            const double2 b = nf4_load(E0, P, Np, m - r, 2 * n - k);
                            double2 c = nf4_load(E0, P, Np, m - r, 2 * n + k);
                            if (minus) { c.x = -c.x; c.y = -c.y; }
                            tr = __dadd_rn(tr, __dsub_rn(__dmul_rn(ar, __dadd_rn(b.x, c.x)), __dmul_rn(ai, __dsub_rn(b.y, c.y))));
                            ti = __dadd_rn(ti, __dadd_rn(__dmul_rn(ar, __dadd_rn(b.y, c.y)), __dmul_rn(ai, __dsub_rn(b.x, c.x))));
                        }
                        if (wf[wu]) {
                            const double ar = wr[wu], ai = wi[wu];
                            const double2 b = nf4_load(E0, P, Np, m - r, 2 * n);
                            tr = __dadd_rn(tr, __dsub_rn(__dmul_rn(ar, b.x), __dmul_rn(ai, b.y)));
                            ti = __dadd_rn(ti, __dadd_rn(__dmul_rn(ar, b.y), __dmul_rn(ai, b.x)));
                        }
                    }
                    const double mag = __dsqrt_rn(__dadd_rn(__dmul_rn(tr, tr), __dmul_rn(ti, ti)));
                    if (mag > 0.0) {
                        const double2 val = make_double2(__ddiv_rn(__dmul_rn(tr, a), mag), __ddiv_rn(__dmul_rn(ti, a), mag));
                        double2 *Erow = E0 + (long long)m * P;
                        Erow[n] = val;
                        if (n >= L + 1 && n < 2 * L + 1) Erow[2 * L - n] = make_double2(val.x, -val.y);
                        else if (n >= Nreal - 1 && n < Naux) Erow[2 * Naux - n] = make_double2(val.x, -val.y);
                    }
                }
                __syncthreads();
                done = hi;
            }
        }
    }
}

// The same sweep with its working set in shared memory (L = 5).  The global-memory kernel above issues the ~35 cell loads of
// a bin one after the other behind its `continue`s (each an L2 round trip): 21 ms for one sweep over 64 x 628 frames, ~65 k
// cycles per frame.  Here frames m-3 .. m live in a shared-memory window laid out exactly like the reference's flat buffer --
// consecutive rows Np cells apart, so that "row m-r, offset 2n +- k" is one address plus an immediate whether or not the
// offset runs into the next row: 8 row slots (rows m-3 .. m+1 are live) + a ninth that mirrors slot 0 (the row after slot 7).  Row m+1 and its amplitudes
// are fetched with cp.async while frame m is processed; dropped terms (|W| <= 1e-12) are added as -0.0 (the running sum starts
// at +0.0 and can never be -0.0: adding a zero of either sign leaves its bits), so a bin is straight-line code.
template <int L>
__global__ void __launch_bounds__(512)
k_nofuture_q4_ring(LwsbView v, const double *wr, const double *wi, const int *wf, const double *thresholds, int iters)
{
    constexpr int Q = 4;
    const int u = blockIdx.x;
    const int T = v.T[u], Nreal = v.Nreal;
    const int Np = Nreal + 2 * L, Naux = Nreal + L - 1, P = v.P;
    extern __shared__ __align__(16) unsigned char nf_smem[];
    constexpr int NS = 8;
    double2 *win = reinterpret_cast<double2 *>(nf_smem);            // [NS + 1][Np]
    double2 *w2 = win + (NS + 1) * Np;                                      // (wr, wi)[Q][Q][L + 1]
    double *amp = reinterpret_cast<double *>(w2 + Q * Q * (L + 1)); // [2][Np + 1]
    unsigned *keep = reinterpret_cast<unsigned *>(amp + 2 * (Np + 1)); // [Q][Q] bit k
    for (int i = threadIdx.x; i < Q * Q * (L + 1); i += blockDim.x) w2[i] = make_double2(wr[i], wi[i]);
    for (int i = threadIdx.x; i < Q * Q; i += blockDim.x) {
        unsigned f = 0;
        for (int k = 0; k <= L; ++k) f |= wf[i * (L + 1) + k] ? 1u << k : 0u;
        keep[i] = f;
    }
    const double mean = v.mean_amp[u];
    double2 *E0 = v.E + v.rowbase[u] * (long long)P + (v.c0 - L); // extended (row 0, column 0)
    const double *A0 = v.A + v.rowbase[u] * (long long)P + (v.c0 - L);
    const int lim1 = (Np - L + 1) / 2; // first n whose reads reach the current frame
    // extended row -> its slot (and the mirror of slot 0); its amplitudes -> amp[row & 1] (not for the ghost rows, which are never
    // updated: two copies in flight to the same amplitude buffer would not be ordered)
    auto fetch_row = [&](int row, bool with_amp) {
        const unsigned slot = (unsigned)(row & (NS - 1));
        for (int x = threadIdx.x; x < Np; x += blockDim.x) {
            const double2 *src = E0 + (long long)row * P + x;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(win + slot * Np + x)), "l"(src) : "memory");
            if (slot == 0)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(win + NS * Np + x)), "l"(src) : "memory");
            if (with_amp)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(amp + (row & 1) * (Np + 1) + x)),
                             "l"(A0 + (long long)row * P + x) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (int it = 0; it < iters; ++it) {
        const double thr = __dmul_rn(thresholds[it], mean); // lws.pyx:298
        __syncthreads();
        for (int row = 0; row < Q; ++row) fetch_row(row, row == Q - 1); // the ghost rows 0 .. Q-2 and the first frame
        for (int m = Q - 1; m < T + Q - 1; ++m) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads(); // row m and its amplitudes are in; everybody is done with frame m-1
            if (m + 1 < T + Q - 1) fetch_row(m + 1, true); // into the slot of row m-7
            const double *arow = amp + (m & 1) * (Np + 1);
            double2 *cur = win + (m & (NS - 1)) * Np, *cur_dup = (m & (NS - 1)) == 0 ? win + NS * Np : nullptr;
            int done = L;
            while (done <= Naux) {
                int hi = (2 * L < done) ? (done + Np - L + 1) / 2 : lim1;
                if (hi < done + 1) hi = done + 1;
                if (hi > Naux + 1) hi = Naux + 1;
                for (int n = done + threadIdx.x; n < hi; n += blockDim.x) {
                    const double a = arow[n];
                    if (!(a > thr)) continue;
                    double tr = 0.0, ti = 0.0;
                    const int pr = (n - L) & (Q - 1);
                    const bool oddn = pr & 1;
#pragma unroll
                    for (int r = Q - 1; r > 0; --r) {
                        const double2 *wrow = w2 + (pr * Q + r) * (L + 1);
                        const unsigned kb = keep[pr * Q + r];
                        const bool minus = oddn && (r & 1);
                        const double2 *base = win + ((m - r) & (NS - 1)) * Np + 2 * n; // flat offset (m-r)*Np + 2n: may run into the next row
#pragma unroll
                        for (int k = 1; k <= L; ++k) {
                            const double2 ww = wrow[k], b = base[-k];
                            double2 c = base[k];
                            if (minus) { c.x = -c.x; c.y = -c.y; }
                            const double vr = __dsub_rn(__dmul_rn(ww.x, __dadd_rn(b.x, c.x)), __dmul_rn(ww.y, __dsub_rn(b.y, c.y)));
                            const double vi = __dadd_rn(__dmul_rn(ww.x, __dadd_rn(b.y, c.y)), __dmul_rn(ww.y, __dsub_rn(b.x, c.x)));
                            const bool kp = (kb >> k) & 1u;
                            tr = __dadd_rn(tr, kp ? vr : -0.0); ti = __dadd_rn(ti, kp ? vi : -0.0);
                        }
                        {
                            const double2 ww = wrow[0], b = base[0];
                            const double vr = __dsub_rn(__dmul_rn(ww.x, b.x), __dmul_rn(ww.y, b.y));
                            const double vi = __dadd_rn(__dmul_rn(ww.x, b.y), __dmul_rn(ww.y, b.x));
                            const bool kp = kb & 1u;
                            tr = __dadd_rn(tr, kp ? vr : -0.0); ti = __dadd_rn(ti, kp ? vi : -0.0);
                        }
                    }
                    double2 val;
                    if (x_project(tr, ti, a, val)) {
                        cur[n] = val;
                        if (cur_dup) cur_dup[n] = val;
                        int mir = -1;
                        if (n >= L + 1 && n < 2 * L + 1) mir = 2 * L - n;
                        else if (n >= Nreal - 1 && n < Naux) mir = 2 * Naux - n;
                        if (mir >= 0) {
                            cur[mir] = make_double2(val.x, -val.y);
                            if (cur_dup) cur_dup[mir] = make_double2(val.x, -val.y);
                        }
                    }
                }
                __syncthreads();
                done = hi;
            }
            for (int x = threadIdx.x; x < Np; x += blockDim.x) E0[(long long)m * P + x] = cur[x]; // frame m is final for this sweep
        }
    }
}

// ------------------------------------------------------------------------------------------
// post-order walk of numpy's pairwise(a, n) = n <= 128 ? leaf : pairwise(a, n2) + pairwise(a + n2, n - n2),
// n2 = (n/2) rounded down to a multiple of 8: appends (offset, length, additions after this leaf) triples
void stat_tree(long long n, std::vector<int> &tab)
{
    struct Node { long long off, len; int state; };
    std::vector<Node> st;
    st.push_back(Node{0, n, 0});
    const size_t first = tab.size();
    while (!st.empty()) {
        Node &top = st.back();
        if (top.len <= 128) {
            tab.push_back((int)top.off); tab.push_back((int)top.len); tab.push_back(0);
            st.pop_back();
        } else if (top.state == 0) {
            long long n2 = top.len / 2; n2 -= n2 % 8;
            top.state = 1;
            const Node right{top.off + n2, top.len - n2, 0}, left{top.off, n2, 0};
            st.push_back(right); // the left half is processed first
            st.push_back(left);
        } else {
            tab[tab.size() - 1] += 1; // both halves done: one addition after the last leaf emitted
            st.pop_back();
        }
    }
    (void)first;
}

// ------------------------------------------------------------------------------------------
// launch wrappers
void launch_extend(const LwsbView &v, int kind, const void *const *src, const StatScratch &sc, double *mean_amp,
                   double *max_amp, int maxTp, cudaStream_t s)
{
    dim3 grid(maxTp, v.B);
    if (kind == 0) k_extend<0><<<grid, 256, 0, s>>>(v, src, sc.row_max);
    else k_extend<1><<<grid, 256, 0, s>>>(v, src, sc.row_max);
    k_stats<<<v.B, 1024, 0, s>>>(v, sc.row_max, mean_amp, max_amp, sc.leaf_tab, sc.tab_of, sc.leaf_sum, sc.stride);
}

void launch_refresh_ghosts(const LwsbView &v, cudaStream_t s)
{
    if (v.Q < 2) return;
    dim3 grid(2 * (v.Q - 1), v.B);
    k_refresh_ghosts<<<grid, 256, 0, s>>>(v);
}

void launch_reextend(const LwsbView &v, const StatScratch &sc, double *mean_amp, double *max_amp, int maxTp,
                     cudaStream_t s)
{
    dim3 grid(maxTp, v.B);
    k_reamp<<<grid, 256, 0, s>>>(v, sc.row_max);
    launch_refresh_ghosts(v, s);
    k_stats<<<v.B, 1024, 0, s>>>(v, sc.row_max, mean_amp, max_amp, sc.leaf_tab, sc.tab_of, sc.leaf_sum, sc.stride);
}

void launch_crop(const LwsbView &v, void *const *dst, int maxT, cudaStream_t s)
{
    dim3 grid(maxT, v.B);
    k_crop<<<grid, 256, 0, s>>>(v, dst);
}

void launch_sweeps_generic(const LwsbView &v, const LwsbW &w, int fold, int rframe, int cframe, const double *thr,
                           int iters, cudaStream_t s)
{
    k_sweeps_generic<<<v.B, 1024, 0, s>>>(v, w, fold, rframe, cframe, thr, iters);
}

void launch_online_generic(const LwsbView &v, const LwsbW *w3, int fold, const double *thr, int iters, int LA,
                           cudaStream_t s, long long j0, long long j1)
{
    int nt = (v.Nreal + v.L) / (v.L + 1) + 1;
    nt = (nt + 31) / 32 * 32;
    k_online_generic<<<v.B, nt, 0, s>>>(v, w3[0], w3[1], w3[2], fold, thr, iters, LA, j0, j1);
}

void launch_stream_extend(const LwsbView &v, int kind, const void *src, int m0, int n, cudaStream_t s)
{
    const int grid = n + (m0 == 0 ? v.Q - 1 : 0);
    if (kind == 0) k_stream_extend<0><<<grid, 256, 0, s>>>(v, src, m0, n);
    else k_stream_extend<1><<<grid, 256, 0, s>>>(v, src, m0, n);
}

void launch_nofuture_q4(const LwsbView &v, const LwsbW &w, const double *thr, int iters, cudaStream_t s)
{
    const double *wr = w.wr, *wi = w.wi;
    const int *wf = w.wf;
    int nt = (v.Nreal + 2 * v.L) / 2 + 32;
    nt = nt > 512 ? 512 : (nt + 31) / 32 * 32;
    const int Np = v.Nreal + 2 * v.L;
    const size_t bytes = (size_t)9 * Np * 16 + (size_t)16 * 6 * 16 + (size_t)2 * (Np + 1) * 8 + 16 * 4 + 16;
    const char *e = getenv("LWSB_NOFUTURE_RING");
    if (v.L == 5 && bytes <= 200 * 1024 && !(e && atoi(e) == 0)) { // working set in shared memory
        auto kern = k_nofuture_q4_ring<5>;
        if (bytes > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        kern<<<v.B, nt, bytes, s>>>(v, wr, wi, wf, thr, iters);
        return;
    }
    k_nofuture_q4<<<v.B, nt, 0, s>>>(v, wr, wi, wf, thr, iters);
}

} // namespace lwsb

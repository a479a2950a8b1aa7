// kernels_generic.cu -- layout kernels (extend / amplitude / crop) and the generic wavefront
// kernels that execute ANY (Q, L, weight pattern) exactly in the reference's sequential order.
//
// Why a wavefront: the reference updates bins in place in raster order (frame-major,
// bin-minor; lwslib.cpp:291-292), i.e. Gauss-Seidel.  A bin (m, c) reads frames m-r at bins
// up to c+L (already updated) and frames m+r from c-L on (not yet updated).  Running row
// update number j of a chain at "time" c + (L+1)*j therefore reproduces the sequential result
// exactly; successive sweeps may follow each other Q frames apart.
// The generic kernels run one CTA per utterance in lock step (one bin per row update per
// step, __syncthreads between steps) with the state in global memory.  They are the
// reference-exact fallback for every configuration; the tuned batch kernel lives in
// kernels_batch.cu.
#include <cuda_runtime.h>
#include "lwsb_common.h"
#include "kernels.h"

namespace lwsb {

// ------------------------------------------------------------------------------------------
// extend + amplitude + per-frame statistics   (lws.pyx:146-157, 235-240; lwslib.cpp:15-65)
// grid (max Tp, B), one CTA per extended row.
template <int KIND>
__global__ void k_extend(LwsbView v, const void *const *src, double *row_sum, double *row_max)
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
    double s = 0.0, mx = 0.0;
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
                a = hypot(val.x, val.y);
            } else {
                const double *S = reinterpret_cast<const double *>(src[u]);
                val = make_double2(S[(long long)p * Nreal + cs], 0.0);
                a = fabs(val.x);
            }
            if (cj) val.y = -val.y;
            if (c >= 0 && c < Nreal) { s += a; mx = fmax(mx, a); }
        }
        E[x] = val;
        A[x] = a;
    }
    // deterministic block reduction (fixed tree): the mean must not depend on scheduling
    __shared__ double sh_s[32], sh_m[32];
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_down_sync(0xffffffffu, s, o);
        mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
    }
    const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if ((threadIdx.x & 31) == 0) { sh_s[w] = s; sh_m[w] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ts = 0.0, tm = 0.0;
        for (int i = 0; i < nw; ++i) { ts += sh_s[i]; tm = fmax(tm, sh_m[i]); }
        row_sum[row] = ts;
        row_max[row] = tm;
    }
}

// mean / max of |S| over the un-extended spectrogram (lws.pyx:240); grid B
__global__ void k_stats(LwsbView v, const double *row_sum, const double *row_max, double *mean_amp, double *max_amp)
{
    const int u = blockIdx.x;
    const int T = v.T[u];
    const long long r0 = v.rowbase[u] + (v.Q - 1);
    double s = 0.0, mx = 0.0;
    for (int m = threadIdx.x; m < T; m += blockDim.x) { s += row_sum[r0 + m]; mx = fmax(mx, row_max[r0 + m]); }
    __shared__ double sh_s[32], sh_m[32];
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_down_sync(0xffffffffu, s, o);
        mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
    }
    const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if ((threadIdx.x & 31) == 0) { sh_s[w] = s; sh_m[w] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ts = 0.0, tm = 0.0;
        for (int i = 0; i < nw; ++i) { ts += sh_s[i]; tm = fmax(tm, sh_m[i]); }
        mean_amp[u] = ts / ((double)T * (double)v.Nreal);
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
__global__ void k_reamp(LwsbView v, double *row_sum, double *row_max)
{
    const int u = blockIdx.y;
    const int T = v.T[u];
    const int m = blockIdx.x;
    if (m < v.Q - 1 || m >= T + v.Q - 1) return;
    const long long row = v.rowbase[u] + m;
    const double2 *E = v.E + row * v.P;
    double *A = v.A + row * v.P;
    const int e0 = v.c0 - v.L, Np = v.Nreal + 2 * v.L;
    double s = 0.0, mx = 0.0;
    for (int e = threadIdx.x; e < Np; e += blockDim.x) {
        const double2 val = E[e0 + e];
        const double a = hypot(val.x, val.y);
        A[e0 + e] = a;
        const int c = e - v.L;
        if (c >= 0 && c < v.Nreal) { s += a; mx = fmax(mx, a); }
    }
    __shared__ double sh_s[32], sh_m[32];
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_down_sync(0xffffffffu, s, o);
        mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
    }
    const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if ((threadIdx.x & 31) == 0) { sh_s[w] = s; sh_m[w] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ts = 0.0, tm = 0.0;
        for (int i = 0; i < nw; ++i) { ts += sh_s[i]; tm = fmax(tm, sh_m[i]); }
        row_sum[row] = ts;
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
// one bin update, generic stencil
__device__ __forceinline__ void commit_bin(const LwsbView &v, double2 *Erow, int c, double tr, double ti, double a)
{
    // lwslib.cpp:356-368: normalise to the stored magnitude, refresh the mirrored copies at once
    const double mag = sqrt(tr * tr + ti * ti);
    if (mag > 0.0) {
        const double2 val = make_double2(tr * a / mag, ti * a / mag);
        Erow[v.c0 + c] = val;
        const int Nreal = v.Nreal;
        if (c >= 1 && c <= v.L) Erow[v.c0 - c] = make_double2(val.x, -val.y);
        else if (c >= Nreal - 1 - v.L && c <= Nreal - 2) Erow[v.c0 + 2 * (Nreal - 1) - c] = make_double2(val.x, -val.y);
    }
}

__device__ __forceinline__ void update_bin(const LwsbView &v, const LwsbStencil &st, long long row, int c, double thr)
{
    double2 *Erow = v.E + row * v.P;
    const double a = v.A[row * v.P + v.c0 + c];
    if (!(a > thr)) return; // lwslib.cpp:295-296
    const int p = c % v.Q;
    const LwsbTerm *tm = st.terms + (size_t)p * st.maxt;
    const int cnt = st.count[p];
    double tr = 0.0, ti = 0.0;
    const double2 *ctr = Erow + v.c0 + c;
    for (int e = 0; e < cnt; ++e) {
        const LwsbTerm t = tm[e];
        const double2 x = ctr[(long long)t.dr * v.P + t.dk];
        tr = fma(t.cr, x.x, tr); tr = fma(-t.ci, x.y, tr);
        ti = fma(t.cr, x.y, ti); ti = fma(t.ci, x.x, ti);
    }
    commit_bin(v, Erow, c, tr, ti, a);
}

// ------------------------------------------------------------------------------------------
// `iters` pipelined sweeps (batch: rframe Q / cframe 1; no-future: rframe 1 / cframe 0).
// Sweep i, frame m, bin c runs at step  c + (L+1)*(m + Q*i).  One CTA per utterance.
__global__ void __launch_bounds__(1024)
k_sweeps_generic(LwsbView v, LwsbStencil st, const double *thresholds, int iters)
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
            update_bin(v, st, row0 + m, c, thresholds[i] * mean); // lws.pyx:245
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// online (TF-RTISI-LA) chain: row update j of the chain, bin c runs at step c + (L+1)*j.
// One CTA per utterance; thread k serves the chain positions j == k (mod blockDim).
__global__ void __launch_bounds__(1024)
k_online_generic(LwsbView v, const LwsbStencil *sts, const double *thresholds, int iters, int LA)
{
    const int u = blockIdx.x;
    const int T = v.T[u], Q = v.Q, Nreal = v.Nreal;
    const int S = v.L + 1;
    const long long base = v.rowbase[u];
    const double mean = v.mean_amp[u];
    const long long n = lwsb_online_chain_len(T, iters, LA);
    const long long tmax = S * (n - 1) + (Nreal - 1);
    const int nt = blockDim.x; // >= ceil(Nreal / S) + 1 (launch_online_generic guarantees)
    long long jc = -1;
    LwsbOnlineTask task;
    LwsbStencil st;
    double thr = 0.0;
    for (long long t = 0; t <= tmax; ++t) {
        const long long jhi = t / S;
        // the unique j <= jhi with j == tid (mod nt) and j > jhi - nt
        const long long d = (jhi - threadIdx.x) % nt;
        const long long j = jhi - (d < 0 ? d + nt : d);
        if (j >= 0 && j < n) {
            const long long c = t - S * j;
            if (c < Nreal) {
                if (j != jc) {
                    jc = j;
                    task = lwsb_online_decode(T, iters, LA, Q, j);
                    st = sts[task.which == 0 ? task.rframe - 1 : (task.which == 1 ? Q : Q + 1)];
                    thr = task.thr < 0 ? 0.0 : thresholds[task.thr] * mean; // lws.pyx:361, lwslib.cpp:1467
                }
                update_bin(v, st, base + task.row, (int)c, thr);
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// NoFuture_LWSQ4 as the reference computes it (lwslib.cpp:538-617): the doubled bin offset
// makes frame m read up to bin 2c+L of frame m-1 and even the already-updated part of frame
// m itself, so frames cannot be skewed by a constant; the (single) sweep is run in raster
// order by one warp per utterance with the <= 3*(2L+1) stencil terms spread over the lanes.
__global__ void __launch_bounds__(32)
k_nofuture_q4(LwsbView v, LwsbStencil st, const double *thresholds, int iters)
{
    const int u = blockIdx.x;
    const int T = v.T[u], Q = v.Q, Nreal = v.Nreal, L = v.L;
    const int Np = Nreal + 2 * L;
    const int lane = threadIdx.x;
    const double mean = v.mean_amp[u];
    double2 *E0 = v.E + v.rowbase[u] * v.P + (v.c0 - L); // extended (row 0, column 0)
    for (int it = 0; it < iters; ++it) {
        const double thr = thresholds[it] * mean;
        for (int m = Q - 1; m < T + Q - 1; ++m) {
            double2 *Erow = v.E + (v.rowbase[u] + m) * v.P;
            const double *Arow = v.A + (v.rowbase[u] + m) * v.P;
            for (int c = 0; c < Nreal; ++c) {
                const double a = Arow[v.c0 + c];
                if (!(a > thr)) continue;
                const int p = c % Q;
                const LwsbTerm *tm = st.terms + (size_t)p * st.maxt;
                const int cnt = st.count[p];
                double tr = 0.0, ti = 0.0;
                for (int e = lane; e < cnt; e += 32) {
                    const LwsbTerm t = tm[e];
                    const long long f = (long long)(m + t.dr) * Np + 2 * (c + L) + t.dk; // reference flat offset
                    const long long fr = f / Np, fc = f % Np;
                    const double2 x = E0[fr * v.P + fc];
                    tr = fma(t.cr, x.x, tr); tr = fma(-t.ci, x.y, tr);
                    ti = fma(t.cr, x.y, ti); ti = fma(t.ci, x.x, ti);
                }
                for (int o = 16; o > 0; o >>= 1) {
                    tr += __shfl_xor_sync(0xffffffffu, tr, o);
                    ti += __shfl_xor_sync(0xffffffffu, ti, o);
                }
                if (lane == 0) commit_bin(v, Erow, c, tr, ti, a);
                __syncwarp();
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// launch wrappers
void launch_extend(const LwsbView &v, int kind, const void *const *src, double *row_sum, double *row_max,
                   double *mean_amp, double *max_amp, int maxTp, cudaStream_t s)
{
    dim3 grid(maxTp, v.B);
    if (kind == 0) k_extend<0><<<grid, 256, 0, s>>>(v, src, row_sum, row_max);
    else k_extend<1><<<grid, 256, 0, s>>>(v, src, row_sum, row_max);
    k_stats<<<v.B, 256, 0, s>>>(v, row_sum, row_max, mean_amp, max_amp);
}

void launch_refresh_ghosts(const LwsbView &v, cudaStream_t s)
{
    if (v.Q < 2) return;
    dim3 grid(2 * (v.Q - 1), v.B);
    k_refresh_ghosts<<<grid, 256, 0, s>>>(v);
}

void launch_reextend(const LwsbView &v, double *row_sum, double *row_max, double *mean_amp, double *max_amp,
                     int maxTp, cudaStream_t s)
{
    dim3 grid(maxTp, v.B);
    k_reamp<<<grid, 256, 0, s>>>(v, row_sum, row_max);
    launch_refresh_ghosts(v, s);
    k_stats<<<v.B, 256, 0, s>>>(v, row_sum, row_max, mean_amp, max_amp);
}

void launch_crop(const LwsbView &v, void *const *dst, int maxT, cudaStream_t s)
{
    dim3 grid(maxT, v.B);
    k_crop<<<grid, 256, 0, s>>>(v, dst);
}

void launch_sweeps_generic(const LwsbView &v, const LwsbStencil &st, const double *thr, int iters, cudaStream_t s)
{
    k_sweeps_generic<<<v.B, 1024, 0, s>>>(v, st, thr, iters);
}

void launch_online_generic(const LwsbView &v, const LwsbStencil *sts, const double *thr, int iters, int LA,
                           cudaStream_t s)
{
    int nt = (v.Nreal + v.L) / (v.L + 1) + 1;
    nt = (nt + 31) / 32 * 32;
    k_online_generic<<<v.B, nt, 0, s>>>(v, sts, thr, iters, LA);
}

void launch_nofuture_q4(const LwsbView &v, const LwsbStencil &st, const double *thr, int iters, cudaStream_t s)
{
    k_nofuture_q4<<<v.B, 32, 0, s>>>(v, st, thr, iters);
}

} // namespace lwsb

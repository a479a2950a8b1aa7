// kernels_online.cu -- online_lws (TF-RTISI-LA, lwslib.cpp:1424-1492) with the working set in shared memory.
//
// The whole schedule is one chain of row updates (lwsb_common.h: lwsb_online_decode); row update j runs S bins behind row
// update j-1.  The ~Nreal/S row updates in flight touch only the extended rows [m_lo - LA, m_hi + 2(Q-1)]; a ring of R rows (R a
// power of two, sized on the host from the exact maximum of the schedule) keeps them in shared memory, rows are copied in when
// the front of the chain first needs them and written back when the tail has passed.  One CTA per utterance.  Arithmetic: the
// reference's, operation for operation (exact.cuh).  Kernels, in the order launch_online_ring tries them:
//   k_online_flow  (Q <= 4, shipping)  K = 4 warps per row update taking turns, so that only the order-bound chain of a bin is on
//                                      the critical path; per-lane data instead of per-residue code; S = K + L = 9
//   k_online_duo   (Q <= 4)            two bins per step on two lanes (spectra too wide for four warps per row update)
//   k_online_ring2 / k_online_ring     two bins / one bin per step in one lane; S a multiple of Q so that every thread of a step is on
//                                      the same residue and the per-residue weights come from the parameter bank (Q = 8)
//   k_online_rail  (experiments build) value warps + chain warps
// (k_online_generic in kernels_generic.cu serves every other shape from global memory.)
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdlib>
#include <type_traits>
#include "exact.cuh"
#include "kernels.h"
#include "lwsb_common.h"

namespace lwsb {

namespace {

constexpr int OL = 5; // stencil reach this kernel is specialised for

template <int Q>
struct OnlineW { // the three weight sets (W, W_ai, W_af) in the reference layout, in the kernel parameter bank
    double wr[3][Q][Q][OL + 1];
    double wi[3][Q][Q][OL + 1];
    unsigned flag[3][Q][Q]; // bit k: |W[p][r][k]| > 1e-12
};

struct OnlineRingCell {
    const double2 *ring;
    int rmask, pitch, row, col;
    __device__ __forceinline__ double2 operator()(int dr, int dk) const
    {
        return ring[(size_t)((row + dr) & rmask) * pitch + col + dk];
    }
};

// Asym_UpdatePhase{Q2,Q4,anyQ} for one bin of residue P (lwslib.cpp:776-1273): frame pairs r < rframe on both
// sides, r >= rframe on the left only, centre-frame terms when cframe; weight set `ws` chosen at run time.
//
// The lanes of a warp are on different row updates (init / look-ahead / newest frame: different rframe, cframe
// and weight set), so the rule is written WITHOUT data-dependent branches:
//   * a "left only" term is the two-sided term with zeros in place of the frame m+r values -- bit-identical:
//     ar*(br+0) - ai*(bi-0) = ar*br - ai*bi (lwslib.cpp:500-501), and for the folded forms b = e1 -+ 0,
//     c = 0 -+ e4 reproduces c = -+E[m-r][n+k] (lwslib.cpp:994-997) exactly;
//   * a term whose |W| <= 1e-12 flag is clear, or a centre term without cframe, is computed and dropped by a
//     select (the reference skips it; adding nothing is the same).
template <int Q, int P, int FOLD>
__device__ __forceinline__ void online_weighted_sum(const OnlineRingCell &E, const OnlineW<Q> &w, int ws, int rframe, int cframe,
                                                    double &tr, double &ti)
{
    constexpr int PN = (Q - P) % Q;
    tr = 0.0; ti = 0.0;
    const double2 zero = make_double2(0.0, 0.0);
    auto add_if = [&](bool f, double ar, double ai, double br, double bi, double cr, double ci) {
        const double nr = __dadd_rn(tr, __dsub_rn(__dmul_rn(ar, __dadd_rn(br, cr)), __dmul_rn(ai, __dsub_rn(bi, ci))));
        const double ni = __dadd_rn(ti, __dadd_rn(__dmul_rn(ar, __dadd_rn(bi, ci)), __dmul_rn(ai, __dsub_rn(br, cr))));
        tr = f ? nr : tr; ti = f ? ni : ti;
    };
    {
        const unsigned f0 = cframe ? w.flag[ws][P][0] : 0u;
#pragma unroll
        for (int k = 1; k <= OL; ++k) {
            const double2 b = E(0, -k), c = E(0, +k);
            add_if((f0 >> k) & 1u, w.wr[ws][P][0][k], w.wi[ws][P][0][k], b.x, b.y, c.x, c.y);
        }
    }
    auto pair = [&](auto rc, auto minusc) {
        constexpr int r = decltype(rc)::value;
        constexpr bool minus = decltype(minusc)::value;
        const unsigned f = w.flag[ws][P][r], fn = w.flag[ws][PN][r];
        const bool both = r < rframe;
        {
            const double2 b = E(-r, 0), c0 = E(+r, 0);
            const double2 c = both ? c0 : zero;
            add_if(f & 1u, w.wr[ws][P][r][0], w.wi[ws][P][r][0], b.x, b.y, c.x, c.y);
        }
#pragma unroll
        for (int k = 1; k <= OL; ++k) {
            const double2 e1 = E(-r, -k), e4 = E(-r, +k), e2l = E(+r, +k), e3l = E(+r, -k);
            const double2 e2 = both ? e2l : zero, e3 = both ? e3l : zero;
            if (FOLD == LWSB_FOLD_ANY) {
                add_if((f >> k) & 1u, w.wr[ws][P][r][k], w.wi[ws][P][r][k], e1.x, e1.y, e3.x, e3.y);
                add_if((fn >> k) & 1u, w.wr[ws][PN][r][k], w.wi[ws][PN][r][k], e2.x, e2.y, e4.x, e4.y);
            } else {
                double br, bi, cr, ci;
                if (minus) {
                    br = __dsub_rn(e1.x, e2.x); bi = __dsub_rn(e1.y, e2.y);
                    cr = __dsub_rn(e3.x, e4.x); ci = __dsub_rn(e3.y, e4.y);
                } else {
                    br = __dadd_rn(e1.x, e2.x); bi = __dadd_rn(e1.y, e2.y);
                    cr = __dadd_rn(e3.x, e4.x); ci = __dadd_rn(e3.y, e4.y);
                }
                add_if((f >> k) & 1u, w.wr[ws][P][r][k], w.wi[ws][P][r][k], br, bi, cr, ci);
            }
        }
    };
    using T_ = std::true_type;
    using F_ = std::false_type;
    if constexpr (FOLD == LWSB_FOLD_Q4 && (P & 1)) { // odd bins: r = 1, 3 sign-flipped, then r = 2 (lwslib.cpp:953-1052)
        pair(std::integral_constant<int, 1>{}, T_{});
        pair(std::integral_constant<int, 3>{}, T_{});
        pair(std::integral_constant<int, 2>{}, F_{});
    } else {
        if constexpr (Q > 1) pair(std::integral_constant<int, 1>{}, F_{});
        if constexpr (Q > 2) pair(std::integral_constant<int, 2>{}, F_{});
        if constexpr (Q > 3) pair(std::integral_constant<int, 3>{}, F_{});
        if constexpr (Q > 4) pair(std::integral_constant<int, 4>{}, F_{});
        if constexpr (Q > 5) pair(std::integral_constant<int, 5>{}, F_{});
        if constexpr (Q > 6) pair(std::integral_constant<int, 6>{}, F_{});
        if constexpr (Q > 7) pair(std::integral_constant<int, 7>{}, F_{});
    }
}

// ---------------------------------------------------------------- two bins per step (Q <= 4)
// The same rule split in two: the values of the inter-frame terms (frames m -+ r, r >= 1: they do not depend on the bin
// updated just before) and the order-bound part (centre-frame terms, the sum in the reference's order, projection).
// A task takes two consecutive bins per step: the term values of both are formed first -- independent work that
// overlaps -- then the two chains run one after the other.  Half the CTA barriers per bin as well.
template <int Q, int FOLD>
struct OnlineVals {
    static constexpr int PER_R = FOLD == LWSB_FOLD_ANY ? 1 + 2 * OL : 1 + OL;
    static constexpr int N = (Q - 1) * PER_R;
    double r[N > 0 ? N : 1], i[N > 0 ? N : 1];
};

__device__ __forceinline__ void online_value(double ar, double ai, double br, double bi, double cr, double ci, double &vr, double &vi)
{
    vr = __dsub_rn(__dmul_rn(ar, __dadd_rn(br, cr)), __dmul_rn(ai, __dsub_rn(bi, ci)));
    vi = __dadd_rn(__dmul_rn(ar, __dadd_rn(bi, ci)), __dmul_rn(ai, __dsub_rn(br, cr)));
}

// the frame pairs in the order the reference adds them: odd bins of the Q4 folding take r = 1, 3 (sign-flipped) then 2
template <int Q, int P, int FOLD, class F>
__device__ __forceinline__ void online_for_each_pair(F &&f)
{
    using T_ = std::true_type;
    using F_ = std::false_type;
    constexpr int PER_R = OnlineVals<Q, FOLD>::PER_R;
    if constexpr (FOLD == LWSB_FOLD_Q4 && (P & 1)) {
        f(std::integral_constant<int, 1>{}, T_{}, std::integral_constant<int, 0>{});
        f(std::integral_constant<int, 3>{}, T_{}, std::integral_constant<int, PER_R>{});
        f(std::integral_constant<int, 2>{}, F_{}, std::integral_constant<int, 2 * PER_R>{});
    } else {
        if constexpr (Q > 1) f(std::integral_constant<int, 1>{}, F_{}, std::integral_constant<int, 0>{});
        if constexpr (Q > 2) f(std::integral_constant<int, 2>{}, F_{}, std::integral_constant<int, PER_R>{});
        if constexpr (Q > 3) f(std::integral_constant<int, 3>{}, F_{}, std::integral_constant<int, 2 * PER_R>{});
    }
}

template <int Q, int P, int FOLD>
__device__ __forceinline__ void online_values(const OnlineRingCell &E, const OnlineW<Q> &w, int ws, int rframe, OnlineVals<Q, FOLD> &v)
{
    constexpr int PN = (Q - P) % Q;
    const double2 zero = make_double2(0.0, 0.0);
    online_for_each_pair<Q, P, FOLD>([&](auto rc, auto minusc, auto basec) {
        constexpr int r = decltype(rc)::value;
        constexpr bool minus = decltype(minusc)::value;
        constexpr int base = decltype(basec)::value;
        const bool both = r < rframe;
        {
            const double2 b = E(-r, 0), c0 = E(+r, 0);
            const double2 c = both ? c0 : zero;
            online_value(w.wr[ws][P][r][0], w.wi[ws][P][r][0], b.x, b.y, c.x, c.y, v.r[base], v.i[base]);
        }
#pragma unroll
        for (int k = 1; k <= OL; ++k) {
            const double2 e1 = E(-r, -k), e4 = E(-r, +k), e2l = E(+r, +k), e3l = E(+r, -k);
            const double2 e2 = both ? e2l : zero, e3 = both ? e3l : zero;
            if (FOLD == LWSB_FOLD_ANY) {
                online_value(w.wr[ws][P][r][k], w.wi[ws][P][r][k], e1.x, e1.y, e3.x, e3.y, v.r[base + 2 * k - 1], v.i[base + 2 * k - 1]);
                online_value(w.wr[ws][PN][r][k], w.wi[ws][PN][r][k], e2.x, e2.y, e4.x, e4.y, v.r[base + 2 * k], v.i[base + 2 * k]);
            } else {
                double br, bi, cr, ci;
                if (minus) {
                    br = __dsub_rn(e1.x, e2.x); bi = __dsub_rn(e1.y, e2.y);
                    cr = __dsub_rn(e3.x, e4.x); ci = __dsub_rn(e3.y, e4.y);
                } else {
                    br = __dadd_rn(e1.x, e2.x); bi = __dadd_rn(e1.y, e2.y);
                    cr = __dadd_rn(e3.x, e4.x); ci = __dadd_rn(e3.y, e4.y);
                }
                online_value(w.wr[ws][P][r][k], w.wi[ws][P][r][k], br, bi, cr, ci, v.r[base + k], v.i[base + k]);
            }
        }
    });
}

// centre-frame terms (they see the bin updated just before), then the term values in the reference's order
template <int Q, int P, int FOLD>
__device__ __forceinline__ void online_accumulate(const OnlineRingCell &E, const OnlineW<Q> &w, int ws, int cframe,
                                                  const OnlineVals<Q, FOLD> &v, double &tr, double &ti)
{
    constexpr int PN = (Q - P) % Q;
    tr = 0.0; ti = 0.0;
    auto add_if = [&](bool f, double vr, double vi) {
        const double nr = __dadd_rn(tr, vr), ni = __dadd_rn(ti, vi);
        tr = f ? nr : tr; ti = f ? ni : ti;
    };
    {
        const unsigned f0 = cframe ? w.flag[ws][P][0] : 0u;
#pragma unroll
        for (int k = 1; k <= OL; ++k) {
            const double2 b = E(0, -k), c = E(0, +k);
            double vr, vi;
            online_value(w.wr[ws][P][0][k], w.wi[ws][P][0][k], b.x, b.y, c.x, c.y, vr, vi);
            add_if((f0 >> k) & 1u, vr, vi);
        }
    }
    online_for_each_pair<Q, P, FOLD>([&](auto rc, auto, auto basec) {
        constexpr int r = decltype(rc)::value;
        constexpr int base = decltype(basec)::value;
        const unsigned f = w.flag[ws][P][r], fn = w.flag[ws][PN][r];
        add_if(f & 1u, v.r[base], v.i[base]);
#pragma unroll
        for (int k = 1; k <= OL; ++k) {
            if (FOLD == LWSB_FOLD_ANY) {
                add_if((f >> k) & 1u, v.r[base + 2 * k - 1], v.i[base + 2 * k - 1]);
                add_if((fn >> k) & 1u, v.r[base + 2 * k], v.i[base + 2 * k]);
            } else add_if((f >> k) & 1u, v.r[base + k], v.i[base + k]);
        }
    });
}

// bins c0 (residue P0) and c0 + 1 of one row update
template <int Q, int FOLD, int P0>
__device__ __forceinline__ void online_two_bins(double2 *ring, int rmask, int pitch, const OnlineW<Q> &w, const LwsbOnlineTask &task,
                                                int c0, int Nreal, bool act0, bool act1, double a0, double a1)
{
    constexpr int L = OL;
    OnlineVals<Q, FOLD> v0, v1;
    const OnlineRingCell cell0{ring, rmask, pitch, task.row, L + c0}, cell1{ring, rmask, pitch, task.row, L + c0 + 1};
    if (act0) online_values<Q, P0, FOLD>(cell0, w, task.which, task.rframe, v0);
    if (act1) online_values<Q, (P0 + 1) % Q, FOLD>(cell1, w, task.which, task.rframe, v1);
    double2 *Rrow = ring + (size_t)(task.row & rmask) * pitch;
    auto commit = [&](int c, double tr, double ti, double a) {
        double2 val;
        if (x_project(tr, ti, a, val)) {
            Rrow[L + c] = val;
            if (c >= 1 && c <= L) Rrow[L - c] = make_double2(val.x, -val.y);
            else if (c >= Nreal - 1 - L && c <= Nreal - 2) Rrow[L + 2 * (Nreal - 1) - c] = make_double2(val.x, -val.y);
        }
    };
    if (act0) {
        double tr, ti;
        online_accumulate<Q, P0, FOLD>(cell0, w, task.which, task.cframe, v0, tr, ti);
        commit(c0, tr, ti, a0);
    }
    if (act1) {
        double tr, ti;
        online_accumulate<Q, (P0 + 1) % Q, FOLD>(cell1, w, task.which, task.cframe, v1, tr, ti);
        commit(c0 + 1, tr, ti, a1);
    }
}

template <int Q, int FOLD>
__global__ void __launch_bounds__(256)
k_online_ring2(LwsbView v, const __grid_constant__ OnlineW<Q> w, const double *thresholds, int iters, int LA, int R, int pitch,
               int S, unsigned *status)
{
    static_assert(Q <= 4 && Q % 2 == 0, "two bins per step: the first bin of a step has an even residue");
    extern __shared__ __align__(16) unsigned char online_smem[];
    double2 *ring = reinterpret_cast<double2 *>(online_smem);
    __shared__ OnlineW<Q> wsm; // indexed per lane (each lane its own row update): shared memory serves divergent addresses
    for (int i = threadIdx.x; i < (int)(sizeof(OnlineW<Q>) / 4); i += blockDim.x)
        reinterpret_cast<unsigned *>(&wsm)[i] = reinterpret_cast<const unsigned *>(&w)[i];
    __syncthreads();
    const int u = blockIdx.x;
    const int T = v.T[u], Nreal = v.Nreal, P = v.P;
    constexpr int L = OL;
    const int Np = Nreal + 2 * L, Tp = T + 2 * (Q - 1), rmask = R - 1;
    double2 *E0 = v.E + v.rowbase[u] * (long long)P + (v.c0 - L); // extended (row 0, column 0)
    const double *A0 = v.A + v.rowbase[u] * (long long)P + (v.c0 - L);
    const double mean = v.mean_amp[u];
    const long long n = lwsb_online_chain_len(T, iters, LA);
    const long long bmax = (long long)S * (n - 1) + (Nreal - 1); // last "bin time": row update j is at bin (bin time - S*j)
    const int nt = blockDim.x;
    int lo = 0, hi = -1; // extended rows [lo, hi] are resident
    long long jc = -1;
    LwsbOnlineTask task;
    double thr = 0.0;
    for (long long bt = 0; bt <= bmax + 1; bt += 2) { // bin time of the first of the step's two bins (S is even: bt % S == 0 is hit)
        if (bt % S == 0) { // the front of the chain moves to a new row update: residency check (uniform across the CTA)
            const long long jhi = min(bt / S, n - 1);
            long long jlo = bt < Nreal ? 0 : (bt - (Nreal - 1) + S - 1) / S;
            if (jlo > n - 1) jlo = n - 1;
            const int need_hi = min(Tp - 1, lwsb_online_frame(iters, LA, jhi) + 2 * (Q - 1));
            const int need_lo = max(0, lwsb_online_frame(iters, LA, jlo) - LA);
            if (need_hi > hi) {
                if (need_hi - need_lo + 1 > R && threadIdx.x == 0) atomicCAS(status, 0u, 0xE1000000u | (unsigned)u);
                for (int e = lo; e < need_lo; ++e) // rows the chain has left: back to global memory
                    if (e >= Q - 1 && e < T + Q - 1)
                        for (int x = threadIdx.x; x < Np; x += nt) E0[(long long)e * P + x] = ring[(size_t)(e & rmask) * pitch + x];
                lo = need_lo;
                __syncthreads();
                for (int e = hi + 1; e <= need_hi; ++e) // rows the chain is about to reach
                    for (int x = threadIdx.x; x < Np; x += nt) ring[(size_t)(e & rmask) * pitch + x] = E0[(long long)e * P + x];
                hi = need_hi;
                __syncthreads();
            }
        }
        const long long jhi = bt / S;
        const long long d = (jhi - threadIdx.x) % nt;
        const long long j = jhi - (d < 0 ? d + nt : d);
        if (j >= 0 && j < n) {
            const int c0 = (int)(bt - (long long)S * j); // even, >= 0
            if (c0 < Nreal) {
                if (j != jc) {
                    jc = j;
                    task = lwsb_online_decode(T, iters, LA, Q, j);
                    thr = task.thr < 0 ? 0.0 : __dmul_rn(thresholds[task.thr], mean); // lws.pyx:361, lwslib.cpp:1467
                }
                const double *ap = A0 + (long long)task.row * P + L + c0;
                const double a0 = __ldg(ap), a1 = c0 + 1 < Nreal ? __ldg(ap + 1) : 0.0;
                const bool act0 = a0 > thr, act1 = c0 + 1 < Nreal && a1 > thr; // lwslib.cpp:295-296
                if (act0 || act1) {
                    if (Q == 2 || (c0 & 2) == 0) online_two_bins<Q, FOLD, 0>(ring, rmask, pitch, wsm, task, c0, Nreal, act0, act1, a0, a1);
                    else online_two_bins<Q, FOLD, 2 % Q>(ring, rmask, pitch, wsm, task, c0, Nreal, act0, act1, a0, a1);
                }
            }
        }
        __syncthreads();
    }
    for (int e = lo; e <= hi; ++e)
        if (e >= Q - 1 && e < T + Q - 1)
            for (int x = threadIdx.x; x < Np; x += nt) E0[(long long)e * P + x] = ring[(size_t)(e & rmask) * pitch + x];
}

// ---------------------------------------------------------------- two bins per step on two lanes (Q <= 4)
// k_online_ring2's step is the in-order stream of one thread: term values of bin c0, term values of bin c0 + 1, the
// order-bound chain of c0 (centre terms, ordered sum, sqrt, division, commit), the chain of c0 + 1 -- ~2 400 instructions at
// the issue rate of a lone warp (ncu: 26 % issue, fp64 pipe 12 %, shared-memory pipe 2 %: the SM is idle).  Here the two
// bins of a task go to lane l of warp w (bin c0) and lane l of warp w + NW (bin c0 + 1): the term values of both bins are
// formed at the same time on different schedulers' slots, then the chain of c0 runs, a named barrier hands the committed
// bin over (bar.arrive / bar.sync, the producer-consumer idiom of the PTX manual), and the chain of c0 + 1 follows.  Same
// operations in the same order per bin: same bits.
__device__ __forceinline__ void online_pair_arrive(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void online_pair_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

// values of the centre-frame terms k = K0 .. L of a bin (K0 = 1 for the first bin of a step: every neighbour it reads was
// committed in an earlier step; K0 = 2 for the second bin, whose k = 1 term reads the bin its partner lane is about to commit)
struct OnlineCentre { double r[OL], i[OL]; };

template <int Q, int P, int K0>
__device__ __forceinline__ void online_centre_values(const OnlineRingCell &E, const OnlineW<Q> &w, int ws, OnlineCentre &cv)
{
#pragma unroll
    for (int k = K0; k <= OL; ++k) {
        const double2 b = E(0, -k), c = E(0, +k);
        online_value(w.wr[ws][P][0][k], w.wi[ws][P][0][k], b.x, b.y, c.x, c.y, cv.r[k - 1], cv.i[k - 1]);
    }
}

// the order-bound part with the centre values of k >= K0 already formed: same additions in the same order as online_accumulate
template <int Q, int P, int FOLD, int K0>
__device__ __forceinline__ void online_accumulate_pre(const OnlineRingCell &E, const OnlineW<Q> &w, int ws, int cframe,
                                                      const OnlineCentre &cv, const OnlineVals<Q, FOLD> &v, double &tr, double &ti)
{
    constexpr int PN = (Q - P) % Q;
    tr = 0.0; ti = 0.0;
    auto add_if = [&](bool f, double vr, double vi) {
        const double nr = __dadd_rn(tr, vr), ni = __dadd_rn(ti, vi);
        tr = f ? nr : tr; ti = f ? ni : ti;
    };
    {
        const unsigned f0 = cframe ? w.flag[ws][P][0] : 0u;
#pragma unroll
        for (int k = 1; k <= OL; ++k) {
            double vr, vi;
            if (k < K0) {
                const double2 b = E(0, -k), c = E(0, +k);
                online_value(w.wr[ws][P][0][k], w.wi[ws][P][0][k], b.x, b.y, c.x, c.y, vr, vi);
            } else { vr = cv.r[k - 1]; vi = cv.i[k - 1]; }
            add_if((f0 >> k) & 1u, vr, vi);
        }
    }
    online_for_each_pair<Q, P, FOLD>([&](auto rc, auto, auto basec) {
        constexpr int r = decltype(rc)::value;
        constexpr int base = decltype(basec)::value;
        const unsigned f = w.flag[ws][P][r], fn = w.flag[ws][PN][r];
        add_if(f & 1u, v.r[base], v.i[base]);
#pragma unroll
        for (int k = 1; k <= OL; ++k) {
            if (FOLD == LWSB_FOLD_ANY) {
                add_if((f >> k) & 1u, v.r[base + 2 * k - 1], v.i[base + 2 * k - 1]);
                add_if((fn >> k) & 1u, v.r[base + 2 * k], v.i[base + 2 * k]);
            } else add_if((f >> k) & 1u, v.r[base + k], v.i[base + k]);
        }
    });
}

template <int Q, int FOLD, int P, int K0>
__device__ __forceinline__ void online_commit_bin(double2 *ring, int rmask, int pitch, const OnlineW<Q> &w, const LwsbOnlineTask &task,
                                                  const OnlineRingCell &cell, const OnlineCentre &cv, const OnlineVals<Q, FOLD> &v, int c,
                                                  int Nreal, double a)
{
    constexpr int L = OL;
    double tr, ti;
    online_accumulate_pre<Q, P, FOLD, K0>(cell, w, task.which, task.cframe, cv, v, tr, ti);
    double2 val;
    if (x_project(tr, ti, a, val)) {
        double2 *Rrow = ring + (size_t)(task.row & rmask) * pitch;
        Rrow[L + c] = val;
        if (c >= 1 && c <= L) Rrow[L - c] = make_double2(val.x, -val.y);
        else if (c >= Nreal - 1 - L && c <= Nreal - 2) Rrow[L + 2 * (Nreal - 1) - c] = make_double2(val.x, -val.y);
    }
}

template <int Q, int FOLD>
__global__ void __launch_bounds__(256)
k_online_duo(LwsbView v, const __grid_constant__ OnlineW<Q> w, const double *thresholds, int iters, int LA, int R, int pitch,
             int S, unsigned *status)
{
    static_assert(Q <= 4 && Q % 2 == 0, "two bins per step: the first bin of a step has an even residue");
    extern __shared__ __align__(16) unsigned char online_smem[];
    double2 *ring = reinterpret_cast<double2 *>(online_smem);
    __shared__ OnlineW<Q> wsm; // indexed per lane (each lane its own row update): shared memory serves divergent addresses
    for (int i = threadIdx.x; i < (int)(sizeof(OnlineW<Q>) / 4); i += blockDim.x)
        reinterpret_cast<unsigned *>(&wsm)[i] = reinterpret_cast<const unsigned *>(&w)[i];
    __syncthreads();
    const int u = blockIdx.x;
    const int T = v.T[u], Nreal = v.Nreal, P = v.P;
    constexpr int L = OL;
    const int Np = Nreal + 2 * L, Tp = T + 2 * (Q - 1), rmask = R - 1;
    double2 *E0 = v.E + v.rowbase[u] * (long long)P + (v.c0 - L); // extended (row 0, column 0)
    const double *A0 = v.A + v.rowbase[u] * (long long)P + (v.c0 - L);
    const double mean = v.mean_amp[u];
    const long long n = lwsb_online_chain_len(T, iters, LA);
    const long long bmax = (long long)S * (n - 1) + (Nreal - 1);
    const int nall = blockDim.x, nt = nall >> 1;          // nt tasks in flight, two lanes each
    const int half = threadIdx.x >= nt ? 1 : 0;           // 0: bin c0, 1: bin c0 + 1 (warp-uniform: nt is a multiple of 32)
    const int tix = threadIdx.x - half * nt;
    const int bar = 1 + (tix >> 5);                        // named barrier of the warp pair (0 is __syncthreads)
    int lo = 0, hi = -1; // extended rows [lo, hi] are resident
    long long jc = -1;
    LwsbOnlineTask task;
    task.row = Q - 1; task.which = 0; task.rframe = 1; task.cframe = 0; task.thr = -1;
    double thr = 0.0, a_next = 0.0;
    long long d = (nt - tix % nt) % nt;                   // (jhi - tix) mod nt at jhi = 0
    for (long long bt = 0; bt <= bmax + 1; bt += 2) {
        if (bt % S == 0) { // the front of the chain moves to a new row update: residency check (uniform across the CTA)
            const long long jhi = min(bt / S, n - 1);
            long long jlo = bt < Nreal ? 0 : (bt - (Nreal - 1) + S - 1) / S;
            if (jlo > n - 1) jlo = n - 1;
            const int need_hi = min(Tp - 1, lwsb_online_frame(iters, LA, jhi) + 2 * (Q - 1));
            const int need_lo = max(0, lwsb_online_frame(iters, LA, jlo) - LA);
            if (need_hi > hi) {
                if (need_hi - need_lo + 1 > R && threadIdx.x == 0) atomicCAS(status, 0u, 0xE1000000u | (unsigned)u);
                for (int e = lo; e < need_lo; ++e) // rows the chain has left: back to global memory
                    if (e >= Q - 1 && e < T + Q - 1)
                        for (int x = threadIdx.x; x < Np; x += nall) E0[(long long)e * P + x] = ring[(size_t)(e & rmask) * pitch + x];
                lo = need_lo;
                __syncthreads();
                for (int e = hi + 1; e <= need_hi; ++e) // rows the chain is about to reach
                    for (int x = threadIdx.x; x < Np; x += nall) ring[(size_t)(e & rmask) * pitch + x] = E0[(long long)e * P + x];
                hi = need_hi;
                __syncthreads();
            }
        }
        // the row update this lane works on: j = jhi - d with d = (jhi - tix) mod nt, kept incrementally (jhi grows by one
        // every S / 2 steps)
        if (bt % S == 0 && bt > 0) d = d + 1 == nt ? 0 : d + 1;
        const long long jhi = bt / S;
        const long long j = jhi - d;
        bool act = false;
        int c = 0;
        double a = 0.0;
        if (j >= 0 && j < n) {
            const int c0 = (int)(bt - (long long)S * j); // even, >= 0
            c = c0 + half;
            if (c < Nreal) {
                if (j != jc) {
                    jc = j;
                    task = lwsb_online_decode(T, iters, LA, Q, j);
                    thr = task.thr < 0 ? 0.0 : __dmul_rn(thresholds[task.thr], mean); // lws.pyx:361, lwslib.cpp:1467
                    a = __ldg(A0 + (long long)task.row * P + L + c);
                } else a = a_next;
                if (c + 2 < Nreal) a_next = __ldg(A0 + (long long)task.row * P + L + c + 2); // this lane's bin of the next step
                act = a > thr; // lwslib.cpp:295-296
            }
        }
        // this lane's bin: term values first (both lanes of a task at the same time), then the chains in bin order
        const bool p2 = Q > 2 && (c & 2);                 // residue of the bin: (c mod 4) = half + (p2 ? 2 : 0); Q = 2: half
        const OnlineRingCell cell{ring, rmask, pitch, task.row, L + c};
        OnlineVals<Q, FOLD> vals;
        OnlineCentre cv;
        if (act) {
            if (half == 0) {
                if (!p2) { online_values<Q, 0, FOLD>(cell, wsm, task.which, task.rframe, vals); online_centre_values<Q, 0, 1>(cell, wsm, task.which, cv); }
                else { online_values<Q, 2 % Q, FOLD>(cell, wsm, task.which, task.rframe, vals); online_centre_values<Q, 2 % Q, 1>(cell, wsm, task.which, cv); }
            } else {
                if (!p2) { online_values<Q, 1, FOLD>(cell, wsm, task.which, task.rframe, vals); online_centre_values<Q, 1, 2>(cell, wsm, task.which, cv); }
                else { online_values<Q, 3 % Q, FOLD>(cell, wsm, task.which, task.rframe, vals); online_centre_values<Q, 3 % Q, 2>(cell, wsm, task.which, cv); }
            }
        }
        if (half == 0) {
            if (act) {
                if (!p2) online_commit_bin<Q, FOLD, 0, 1>(ring, rmask, pitch, wsm, task, cell, cv, vals, c, Nreal, a);
                else online_commit_bin<Q, FOLD, 2 % Q, 1>(ring, rmask, pitch, wsm, task, cell, cv, vals, c, Nreal, a);
            }
            __syncwarp();
            online_pair_arrive(bar);
        } else {
            __syncwarp();
            online_pair_sync(bar);
            // near the ends of the spectrum a centre neighbour n -+ k, k >= 2, can be the MIRROR cell of the bin the partner
            // has just committed (bin b's mirror sits at -b, resp. 2 (Nreal - 1) - b): those values are formed again now
            if (act && (c - 1 <= L || c - 1 >= Nreal - 2 - L)) {
                if (!p2) online_centre_values<Q, 1, 2>(cell, wsm, task.which, cv);
                else online_centre_values<Q, 3 % Q, 2>(cell, wsm, task.which, cv);
            }
            if (act) {
                if (!p2) online_commit_bin<Q, FOLD, 1, 2>(ring, rmask, pitch, wsm, task, cell, cv, vals, c, Nreal, a);
                else online_commit_bin<Q, FOLD, 3 % Q, 2>(ring, rmask, pitch, wsm, task, cell, cv, vals, c, Nreal, a);
            }
        }
        __syncthreads();
    }
    for (int e = lo; e <= hi; ++e)
        if (e >= Q - 1 && e < T + Q - 1)
            for (int x = threadIdx.x; x < Np; x += nall) E0[(long long)e * P + x] = ring[(size_t)(e & rmask) * pitch + x];
}

// ---------------------------------------------------------------- one bin per step, K warps per task taking turns (Q <= 4)
// The chain of a bin (centre-frame terms that see the bin committed just before, the ordered sum, sqrt, two divisions) is a
// dependent sequence no schedule can shorten; the term values of the other frames are not on it.  k_online_duo still runs
// them one after the other in every lane.  Here a task (row update) is served by K lanes, lane l of K different warps
// ("roles"): role rho takes the bins of bin-steps b = rho, rho + K, ...  While role rho runs the chain of bin-step b, roles
// rho + 1 .. rho + K - 1 are forming the term values of bin-steps b + 1 .. b + K - 1, so that per bin-step only a chain is on
// the critical path.  Hand-over is a named barrier per bin-step: the warps that committed bin-step b and the warps about to run
// the chain of b + 1 meet on barrier 1 + b mod K (all groups of 32 tasks: consecutive tasks sit in neighbouring lanes and
// groups).
//
// Dependencies (bin time b: task j is on bin b - S j).  The term values of task j, bin c read columns c - L .. c + L of other
// rows; task j - 1 must have committed through column c + L and task j + 1 must not have reached column c - L.  A role starts
// the values of bin-step b after its own hand-over of bin-step b - K, when all bin-steps <= b - K are complete: task j - 1 is
// then through column c - K + S >= c + L iff S >= K + L; the chains of bin-steps > b cannot start before b's, so task j + 1
// is at most on column c + K - 1 - S < c - L.  The centre-frame terms read bins (and mirror cells) of the task's own row
// committed up to one bin-step ago: all of them are formed after the hand-over; only k = 1 is on the critical path, the
// others fill the bubbles of the dependent sum.
//
// Everything that depends on the residue of the bin, on the kind of row update or on the parity rule of the Q4 folding is
// per-lane DATA (weight rows, row offsets with the all-zero row standing in for frames that are not used, a sign mask), not
// code: one body for every bin, S need not be a multiple of Q.  A term the reference skips (|W| <= 1e-12, no centre frame)
// gets the value -0.0: the running sum starts at +0.0 and can never be -0.0, so adding a zero of either sign leaves its bits.
// Shared memory is addressed with 32-bit shared-window addresses and immediate offsets (ld.shared / st.shared).
__device__ __forceinline__ void flow_bar(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ double flow_flip(double x, unsigned m) { return __hiloint2double(__double2hiint(x) ^ (int)m, __double2loint(x)); }
__device__ __forceinline__ double flow_keep(double x, bool keep) { return keep ? x : -0.0; }

template <int OFF> // cell at byte address a + 16 OFF of the shared window
__device__ __forceinline__ double2 flow_ld(unsigned a)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v.x), "=d"(v.y) : "r"(a), "n"(16 * OFF) : "memory");
    return v;
}
__device__ __forceinline__ void flow_st(unsigned a, double x, double y)
{
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(x), "d"(y) : "memory");
}
template <int I, int N, class F>
__device__ __forceinline__ void flow_for(F &&f)
{
    if constexpr (I < N) { f(std::integral_constant<int, I>{}); flow_for<I + 1, N>(f); }
}

template <int Q, int FOLD, int K>
__global__ void __launch_bounds__(256)
k_online_flow(LwsbView v, const __grid_constant__ OnlineW<Q> w, const double *thresholds, int iters, int LA, int R, int pitch,
              int S, unsigned *status)
{
    static_assert(Q <= 4 && Q >= 2, "frame pairs held in registers");
    constexpr int L = OL, NP = Q - 1, PER_R = OnlineVals<Q, FOLD>::PER_R;
    extern __shared__ __align__(16) unsigned char online_smem[];
    double2 *ring = reinterpret_cast<double2 *>(online_smem);
    double2 *zrow = ring + (size_t)R * pitch;                       // a row of zeros: the frames m + r a row update does not use
    // (wr, wi)[3][Q] blocks of Q (L + 1) cells + 1 of padding: the lanes of a warp differ in weight set and residue, the odd
    // block stride spreads them over the bank groups (without it all of them collide on one: 3 Q-way conflicts per load)
    constexpr int WB = Q * (L + 1) + 1;
    double2 *w2 = zrow + pitch;
    unsigned *wf = reinterpret_cast<unsigned *>(w2 + 3 * Q * WB); // flags [3][Q][Q]
    for (int i = threadIdx.x; i < 3 * Q * Q * (L + 1); i += blockDim.x)
        w2[i / (Q * (L + 1)) * WB + i % (Q * (L + 1))] = make_double2((&w.wr[0][0][0][0])[i], (&w.wi[0][0][0][0])[i]);
    for (int i = threadIdx.x; i < 3 * Q * Q; i += blockDim.x) wf[i] = (&w.flag[0][0][0])[i];
    for (int i = threadIdx.x; i < pitch; i += blockDim.x) zrow[i] = make_double2(0.0, 0.0);
    __syncthreads();
    const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring), w2_s = (unsigned)__cvta_generic_to_shared(w2);
    const int u = blockIdx.x;
    const int T = v.T[u], Nreal = v.Nreal, P = v.P;
    const int Np = Nreal + 2 * L, Tp = T + 2 * (Q - 1), rmask = R - 1;
    double2 *E0 = v.E + v.rowbase[u] * (long long)P + (v.c0 - L); // extended (row 0, column 0)
    const double *A0 = v.A + v.rowbase[u] * (long long)P + (v.c0 - L);
    const double mean = v.mean_amp[u];
    const long long n = lwsb_online_chain_len(T, iters, LA);
    const int wmap = (S >> 16) & 3;
    S &= 0xffff;
    long long bend = (long long)S * (n - 1) + (Nreal - 1) + 1; // one empty bin-step closes the last hand-over
    const int nall = blockDim.x, G = nall / (32 * K), nt = 32 * G;   // nt tasks in flight, K lanes each
    // warp w issues on scheduler w mod 4.  With two groups the warps of roles rho and rho + 1 share a pair of schedulers: the chain
    // of bin-step b then runs next to a warp that is waiting for its own hand-over (b + 1) half of the time, instead of always next
    // to one in the middle of its term values
    const int wid = threadIdx.x >> 5;
    int role = wid / G, grp = wid % G;
    if (K == 4 && G == 2 && wmap == 1) { role = (wid >> 1) == 1 ? 2 : ((wid >> 1) == 2 ? 1 : (wid >> 1)); grp = wid & 1; }
    if (K == 4 && wmap == 2) { role = wid & 3; grp = wid >> 2; } // both groups of a role on one scheduler
    const int tix = grp * 32 + (threadIdx.x & 31);
    int handover = 2 * nt;                                           // threads on a hand-over barrier
    int bar_mine = 1 + role, bar_prev = 1 + (role + K - 1) % K;      // b mod K == role for every bin-step of this warp
    // loop invariants that sit between a commit and its hand-over: kept in registers (opaque to rematerialisation)
    asm volatile("" : "+r"(handover), "+r"(bar_mine), "+r"(bar_prev), "+l"(bend));
    int lo = 0, hi = -1;                                             // extended rows [lo, hi] are resident
    long long jc = -1, jhi = 0, ev_j = 0;                            // ev_j: chain position at which the front reaches the next frame
    int bmod = role;                                                 // b mod S (role < K <= S)
    int d = (nt - tix) % nt;                                         // (jhi - tix) mod nt at jhi = 0
    // per-lane constants of the task (and of the residue of this role's bins): byte offsets into the ring / weight tables
    unsigned off_m[NP], off_p[NP], off_own = 0, wofs[NP], wofs_n[NP], wofs_c = 0;
    unsigned flg[NP], flg_n[NP], flg_c = 0, sgn[NP];
    int row = Q - 1;
#pragma unroll
    for (int s = 0; s < NP; ++s) { off_m[s] = off_p[s] = wofs[s] = wofs_n[s] = 0; flg[s] = flg_n[s] = sgn[s] = 0; }
    int pc = -1;
    double thr = 0.0, a_next = 0.0;
    for (long long b = role; b <= bend; b += K) {
        bool waited = false;
        // a multiple of S in (b - K, b]: the front of the chain reaches a new row update (each role sees each once); rows have
        // to come in when that row update is the first of a frame
        if (bmod < K && jhi >= ev_j) {
            const long long bs = b - bmod;
            const long long jh = min(jhi, n - 1);
            long long jlo = bs < Nreal ? 0 : (bs - (Nreal - 1) + S - 1) / S;
            if (jlo > n - 1) jlo = n - 1;
            const int mfront = lwsb_online_frame(iters, LA, jh);
            ev_j = mfront + 1 < T ? lwsb_online_frame_base(mfront + 1, iters, LA) : 0x7fffffffffffffffLL;
            const int need_hi = min(Tp - 1, mfront + 2 * (Q - 1));
            if (need_hi > hi) {
                const int need_lo = max(0, lwsb_online_frame(iters, LA, jlo) - LA);
                // drain: the role of bin-step bs takes its hand-over first (the committers of bs - 1 wait for it), then
                // everybody meets: all bin-steps < bs are complete and none >= bs has begun
                if (bmod == 0 && b > 0) { flow_bar(bar_prev, handover); waited = true; }
                __syncthreads();
                if (need_hi - need_lo + 1 > R && threadIdx.x == 0) atomicCAS(status, 0u, 0xE1000000u | (unsigned)u);
                for (int e = lo; e < need_lo; ++e) // rows the chain has left: back to global memory
                    if (e >= Q - 1 && e < T + Q - 1)
                        for (int x = threadIdx.x; x < Np; x += nall) E0[(long long)e * P + x] = ring[(size_t)(e & rmask) * pitch + x];
                __syncthreads();
                for (int e = hi + 1; e <= need_hi; ++e) // rows the chain is about to reach
                    for (int x = threadIdx.x; x < Np; x += nall) ring[(size_t)(e & rmask) * pitch + x] = E0[(long long)e * P + x];
                __syncthreads();
                lo = need_lo; hi = need_hi;
            }
        }
        // the row update this lane works on: j = jhi - d, d = (jhi - tix) mod nt, on bin c = b - S j = bmod + S d
        const long long j = jhi - d;
        const int c = bmod + S * d;
        int ctl = (b + K <= bend ? 1 : 0) | (b < bend ? 2 : 0) | (b > 0 && !waited ? 4 : 0); // decided before the chain, not after it
        asm volatile("" : "+r"(ctl));
        const bool more = ctl & 1, handoff = ctl & 2, take = ctl & 4;
        bool act = false;
        double a = 0.0;
        if (j >= 0 && j < n && c < Nreal) {
            const int p = c % Q;
            if (j != jc || p != pc) {
                LwsbOnlineTask task = lwsb_online_decode(T, iters, LA, Q, j);
                if (j != jc) {
                    thr = task.thr < 0 ? 0.0 : __dmul_rn(thresholds[task.thr], mean); // lws.pyx:361, lwslib.cpp:1467
                    a = __ldg(A0 + (long long)task.row * P + L + c);
                } else a = a_next;
                jc = j; pc = p;
                row = task.row;
                const int pn = (Q - p) % Q;
                const bool odd = FOLD == LWSB_FOLD_Q4 && (p & 1);
                off_own = (unsigned)((row & rmask) * pitch) * 16u;
                wofs_c = (unsigned)((task.which * Q + p) * WB) * 16u;
                flg_c = task.cframe ? wf[(task.which * Q + p) * Q + 0] : 0u;
#pragma unroll
                for (int s = 0; s < NP; ++s) {
                    // the frame pairs in the order the reference adds them: odd bins of the Q4 folding take r = 1, 3
                    // (sign-flipped) then 2 (lwslib.cpp:953-1052)
                    const int r = odd ? (s == 0 ? 1 : (s == 1 ? 3 : 2)) : s + 1;
                    sgn[s] = (odd && s < 2) ? 0x80000000u : 0u;
                    off_m[s] = (unsigned)(((row - r) & rmask) * pitch) * 16u;
                    off_p[s] = (unsigned)(r < task.rframe ? ((row + r) & rmask) * pitch : R * pitch) * 16u; // R * pitch: the zero row
                    wofs[s] = (unsigned)((task.which * Q + p) * WB + r * (L + 1)) * 16u;
                    wofs_n[s] = (unsigned)((task.which * Q + pn) * WB + r * (L + 1)) * 16u;
                    flg[s] = wf[(task.which * Q + p) * Q + r];
                    flg_n[s] = wf[(task.which * Q + pn) * Q + r];
                }
            } else a = a_next;
            if (c + K < Nreal) a_next = __ldg(A0 + (long long)row * P + L + c + K); // this lane's next bin
            act = a > thr; // lwslib.cpp:295-296
        }
        // ---- term values of the other frames
        const unsigned colb = (unsigned)(L + c) * 16u;
        const unsigned a_own = ring_s + off_own + colb, a_wc = w2_s + wofs_c;
        OnlineVals<Q, FOLD> vals;
        double2 wc[L]; // centre weights, fetched before the hand-over
        if (act) {
#pragma unroll
            for (int s = 0; s < NP; ++s) {
                const unsigned am = ring_s + off_m[s] + colb, ap = ring_s + off_p[s] + colb, aw = w2_s + wofs[s], awn = w2_s + wofs_n[s];
                const int base = s * PER_R;
                {
                    const double2 bb = flow_ld<0>(am), cc = flow_ld<0>(ap), ww = flow_ld<0>(aw);
                    double vr, vi;
                    online_value(ww.x, ww.y, bb.x, bb.y, cc.x, cc.y, vr, vi);
                    vals.r[base] = flow_keep(vr, flg[s] & 1u); vals.i[base] = flow_keep(vi, flg[s] & 1u);
                }
                flow_for<1, L + 1>([&](auto kc) {
                    constexpr int k = decltype(kc)::value;
                    const double2 e1 = flow_ld<-k>(am), e4 = flow_ld<k>(am), e2 = flow_ld<k>(ap), e3 = flow_ld<-k>(ap), ww = flow_ld<k>(aw);
                    const bool keep = (flg[s] >> k) & 1u;
                    if constexpr (FOLD == LWSB_FOLD_ANY) {
                        const double2 wn = flow_ld<k>(awn);
                        const bool keepn = (flg_n[s] >> k) & 1u;
                        double vr, vi;
                        online_value(ww.x, ww.y, e1.x, e1.y, e3.x, e3.y, vr, vi);
                        vals.r[base + 2 * k - 1] = flow_keep(vr, keep); vals.i[base + 2 * k - 1] = flow_keep(vi, keep);
                        online_value(wn.x, wn.y, e2.x, e2.y, e4.x, e4.y, vr, vi);
                        vals.r[base + 2 * k] = flow_keep(vr, keepn); vals.i[base + 2 * k] = flow_keep(vi, keepn);
                    } else {
                        // b = e1 -+ e2, c = e3 -+ e4 (lwslib.cpp:204-207, 228-231): x - y == x + (-y) bit for bit
                        const double br = __dadd_rn(e1.x, flow_flip(e2.x, sgn[s])), bi = __dadd_rn(e1.y, flow_flip(e2.y, sgn[s]));
                        const double cr = __dadd_rn(e3.x, flow_flip(e4.x, sgn[s])), ci = __dadd_rn(e3.y, flow_flip(e4.y, sgn[s]));
                        double vr, vi;
                        online_value(ww.x, ww.y, br, bi, cr, ci, vr, vi);
                        vals.r[base + k] = flow_keep(vr, keep); vals.i[base + k] = flow_keep(vi, keep);
                    }
                });
            }
            flow_for<1, L + 1>([&](auto kc) { constexpr int k = decltype(kc)::value; wc[k - 1] = flow_ld<k>(a_wc); });
        }
        // addresses of the commit, formed before the hand-over (and kept: opaque to rematerialisation)
        unsigned a_cell = a_own, a_mir = a_own;
        if (c >= 1 && c <= L) a_mir = a_own - 32u * (unsigned)c;                                      // column L - c
        else if (c >= Nreal - 1 - L && c <= Nreal - 2) a_mir = a_own + 32u * (unsigned)(Nreal - 1 - c); // column L + 2 (Nreal - 1) - c
        asm volatile("" : "+r"(a_cell), "+r"(a_mir));
        if (take) flow_bar(bar_prev, handover); // every bin of bin-step b - 1 is committed
        // ---- the chain: centre-frame values, the sum in the reference's order, projection, commit
        if (act) {
            double cvr[L], cvi[L];
            flow_for<1, L + 1>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                const double2 bb = flow_ld<-k>(a_cell), cc = flow_ld<k>(a_cell);
                double vr, vi;
                online_value(wc[k - 1].x, wc[k - 1].y, bb.x, bb.y, cc.x, cc.y, vr, vi);
                cvr[k - 1] = flow_keep(vr, (flg_c >> k) & 1u); cvi[k - 1] = flow_keep(vi, (flg_c >> k) & 1u);
            });
            double tr = 0.0, ti = 0.0;
#pragma unroll
            for (int k = 0; k < L; ++k) { tr = __dadd_rn(tr, cvr[k]); ti = __dadd_rn(ti, cvi[k]); }
#pragma unroll
            for (int i = 0; i < OnlineVals<Q, FOLD>::N; ++i) { tr = __dadd_rn(tr, vals.r[i]); ti = __dadd_rn(ti, vals.i[i]); }
            double2 val;
            if (x_project_early(tr, ti, a, val)) {
                flow_st(a_cell, val.x, val.y);
                if (a_mir != a_cell) flow_st(a_mir, val.x, -val.y);
            }
        }
        if (handoff) flow_bar(bar_mine, handover);
        if (!more) break;
        bmod += K;
        if (bmod >= S) { bmod -= S; ++jhi; if (++d == nt) d = 0; }
    }
    __syncthreads();
    for (int e = lo; e <= hi; ++e)
        if (e >= Q - 1 && e < T + Q - 1)
            for (int x = threadIdx.x; x < Np; x += nall) E0[(long long)e * P + x] = ring[(size_t)(e & rmask) * pitch + x];
}

#ifdef LWSB_EXPERIMENTS
// ---------------------------------------------------------------- value warps and chain warps (Q <= 4, folded rules)
// EXPERIMENT (round 2, -DLWSB_EXPERIMENTS and LWSB_ONLINE_RAIL=1): bit-exact, but 238 ms at configs[2] against 137 ms for
// k_online_flow -- every instruction of the chain warp is on the critical path (1 330 cycles per bin with its bookkeeping, the
// loads of 18 staged values and all five centre-frame values), and the value warps pay 80 register moves per bin to slide their
// windows (profiles/r2_online_experiments.txt).
// k_online_flow hands every bin of a task from one warp to the next: the bin committed one bin-step ago travels through shared
// memory and a barrier on the critical path, and each of the K warps loads the whole 6 x 11 neighbourhood of its bin (ncu: the
// shared-memory pipe of the busy SMs is ~75 % occupied, the loads of the chain queue behind them).  Here the work of a task
// is split by KIND instead: per group of 32 tasks, warp r = 1 .. Q - 1 forms the term values of frame pair (m - r, m + r) for
// EVERY bin of the task, and one chain warp runs the order-bound part of every bin.
//   * A value warp keeps the 11 columns of its two rows in registers and slides them: two loads per bin instead of 22.  It
//     runs one bin ahead of the chain warp and parks the 6 values of a bin in a staging area (two buffers, by parity of time).
//   * The chain warp reads the 3 x 6 values, forms the centre-frame values, adds in the reference's order, projects, commits.
//     The bin committed just before is the lane's own: it stays in a register, nothing is handed over.
//   * One CTA barrier per bin-step; within a bin-step nobody reads a cell that somebody writes (see below).
// Time t: value warps are on bin t - S j of task j, the chain warp on bin t - 1 - S j.  The values of bin c read columns
// c - L .. c + L of the rows of earlier tasks: task j - 1 has committed through column c - 2 + S at the end of iteration t - 1,
// so S >= L + 2; the new column of the sliding window, c + L, is one the chain warp of task j - 1 committed at iteration
// t - 1 at the latest.  Cells written in iteration t (column c - 1 of the task's own row and its mirror cell) are read by no
// value warp in iteration t: task j + 1 is S columns behind, task j - 1 S columns ahead, the own row is only read by the
// chain lane itself.  Residue, kind of row update and the parity rule of the Q4 folding are per-lane data as in k_online_flow.
template <int Q, int FOLD>
__global__ void __launch_bounds__(256)
k_online_rail(LwsbView v, const __grid_constant__ OnlineW<Q> w, const double *thresholds, int iters, int LA, int R, int pitch,
              int S, unsigned *status)
{
    static_assert((Q == 4 && FOLD == LWSB_FOLD_Q4) || (Q == 2 && FOLD == LWSB_FOLD_Q2), "folded rules: 1 + L values per frame pair");
    constexpr int L = OL, NP = Q - 1, PER_R = 1 + L, WB = Q * (L + 1) + 1, W11 = 2 * L + 1;
    constexpr int SLOT = NP * PER_R + 1; // cells per task and buffer in the staging area (odd: consecutive lanes in different bank groups)
    extern __shared__ __align__(16) unsigned char online_smem[];
    double2 *ring = reinterpret_cast<double2 *>(online_smem);
    double2 *zrow = ring + (size_t)R * pitch;
    double2 *w2 = zrow + pitch;
    const int nall = blockDim.x, G = nall / (32 * Q), nt = 32 * G;
    double2 *stage = w2 + 3 * Q * WB;                                  // [2][nt][SLOT]
    unsigned *wf = reinterpret_cast<unsigned *>(stage + 2 * nt * SLOT); // flags [3][Q][Q]
    for (int i = threadIdx.x; i < 3 * Q * Q * (L + 1); i += blockDim.x)
        w2[i / (Q * (L + 1)) * WB + i % (Q * (L + 1))] = make_double2((&w.wr[0][0][0][0])[i], (&w.wi[0][0][0][0])[i]);
    for (int i = threadIdx.x; i < 3 * Q * Q; i += blockDim.x) wf[i] = (&w.flag[0][0][0])[i];
    for (int i = threadIdx.x; i < pitch; i += blockDim.x) zrow[i] = make_double2(0.0, 0.0);
    for (int i = threadIdx.x; i < 2 * nt * SLOT; i += blockDim.x) stage[i] = make_double2(0.0, 0.0);
    __syncthreads();
    const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring), w2_s = (unsigned)__cvta_generic_to_shared(w2);
    const unsigned stage_s = (unsigned)__cvta_generic_to_shared(stage);
    const int u = blockIdx.x;
    const int T = v.T[u], Nreal = v.Nreal, P = v.P;
    const int Np = Nreal + 2 * L, Tp = T + 2 * (Q - 1), rmask = R - 1;
    double2 *E0 = v.E + v.rowbase[u] * (long long)P + (v.c0 - L);
    const double *A0 = v.A + v.rowbase[u] * (long long)P + (v.c0 - L);
    const double mean = v.mean_amp[u];
    const long long n = lwsb_online_chain_len(T, iters, LA);
    const long long tend = (long long)S * (n - 1) + (Nreal - 1) + 1;   // the chain warp is one iteration behind
    const int wid = threadIdx.x >> 5, role = wid / G, tix = (wid % G) * 32 + (threadIdx.x & 31);
    const bool chain = role == NP;
    const int r = role + 1;                                            // frame pair of a value warp
    const int lag = chain ? 1 : 0;                                     // local time tl = t - lag; the lane is on bin tl - S j
    int lo = 0, hi = -1;
    long long jhi = 0, ev_j = 0;                                       // jhi = tl / S (>= 0), ev_j: see k_online_flow
    long long jf = 0;                                                  // t / S: the front of the value warps (residency)
    int tmod = 0, fmod = 0;                                            // tl mod S, t mod S
    int d = (nt - tix) % nt;
    // value warps: the two rows of the frame pair, columns c - L .. c + L
    double2 wm[W11], wp[W11];
#pragma unroll
    for (int i = 0; i < W11; ++i) wm[i] = wp[i] = make_double2(0.0, 0.0);
    unsigned off_m = 0, off_p = 0, wtask = 0, flags4 = 0;              // flags4: the 1 + L flag bits of residues 0 .. Q - 1, 8 bits each
    // chain warp
    unsigned off_own = 0, wtask_c = 0, flags_c4 = 0;
    int row = Q - 1;
    double thr = 0.0, a_next = 0.0;
    for (long long t = 0; t <= tend; ++t) {
        if (fmod == 0 && jf >= ev_j) { // the front reaches a new row update, the first of a frame: rows come in, rows the tail has left go out
            const long long jh = min(jf, n - 1);
            const long long tc = t - 1; // time of the chain warp, which is still to run this iteration
            long long jlo = tc < Nreal ? 0 : (tc - (Nreal - 1) + S - 1) / S;
            if (jlo > n - 1) jlo = n - 1;
            const int mfront = lwsb_online_frame(iters, LA, jh);
            ev_j = mfront + 1 < T ? lwsb_online_frame_base(mfront + 1, iters, LA) : 0x7fffffffffffffffLL;
            const int need_hi = min(Tp - 1, mfront + 2 * (Q - 1));
            if (need_hi > hi) {
                const int need_lo = max(0, lwsb_online_frame(iters, LA, jlo) - LA);
                if (need_hi - need_lo + 1 > R && threadIdx.x == 0) atomicCAS(status, 0u, 0xE1000000u | (unsigned)u);
                for (int e = lo; e < need_lo; ++e)
                    if (e >= Q - 1 && e < T + Q - 1)
                        for (int x = threadIdx.x; x < Np; x += nall) E0[(long long)e * P + x] = ring[(size_t)(e & rmask) * pitch + x];
                __syncthreads();
                for (int e = hi + 1; e <= need_hi; ++e)
                    for (int x = threadIdx.x; x < Np; x += nall) ring[(size_t)(e & rmask) * pitch + x] = E0[(long long)e * P + x];
                __syncthreads();
                lo = need_lo; hi = need_hi;
            }
        }
        const bool started = t >= lag;
        const long long j = jhi - d;
        const int c = tmod + S * d;
        const bool on = started && j >= 0 && j < n && c < Nreal;
        const unsigned buf_v = (unsigned)(t & 1), buf_c = buf_v ^ 1u;
        if (!chain) {
            if (on) {
                if (c == 0) { // a new task: its rows, weights and flags; the whole window
                    const LwsbOnlineTask task = lwsb_online_decode(T, iters, LA, Q, j);
                    off_m = (unsigned)(((task.row - r) & rmask) * pitch) * 16u;
                    off_p = (unsigned)(r < task.rframe ? ((task.row + r) & rmask) * pitch : R * pitch) * 16u;
                    wtask = (unsigned)(task.which * Q * WB + r * (L + 1)) * 16u;
                    flags4 = 0;
#pragma unroll
                    for (int pp = 0; pp < Q; ++pp) flags4 |= (wf[(task.which * Q + pp) * Q + r] & 0xffu) << (8 * pp);
                    const unsigned am = ring_s + off_m + (unsigned)L * 16u, ap = ring_s + off_p + (unsigned)L * 16u;
                    flow_for<0, W11>([&](auto ic) { constexpr int i = decltype(ic)::value; wm[i] = flow_ld<i - L>(am); wp[i] = flow_ld<i - L>(ap); });
                } else {
#pragma unroll
                    for (int i = 0; i + 1 < W11; ++i) { wm[i] = wm[i + 1]; wp[i] = wp[i + 1]; }
                    const unsigned colb = (unsigned)(L + c + L) * 16u;
                    wm[W11 - 1] = flow_ld<0>(ring_s + off_m + colb); wp[W11 - 1] = flow_ld<0>(ring_s + off_p + colb);
                }
                const int p = c % Q;
                const unsigned aw = w2_s + wtask + (unsigned)(p * WB) * 16u;
                const unsigned flg = (flags4 >> (8 * p)) & 0xffu;
                const unsigned sgn = (FOLD == LWSB_FOLD_Q4 && (p & 1) && r != 2) ? 0x80000000u : 0u;
                const unsigned ast = stage_s + (unsigned)(((int)buf_v * nt + tix) * SLOT + (r - 1) * PER_R) * 16u;
                {
                    const double2 ww = flow_ld<0>(aw);
                    double vr, vi;
                    online_value(ww.x, ww.y, wm[L].x, wm[L].y, wp[L].x, wp[L].y, vr, vi);
                    flow_st(ast, flow_keep(vr, flg & 1u), flow_keep(vi, flg & 1u));
                }
                flow_for<1, L + 1>([&](auto kc) {
                    constexpr int k = decltype(kc)::value;
                    const double2 e1 = wm[L - k], e4 = wm[L + k], e2 = wp[L + k], e3 = wp[L - k], ww = flow_ld<k>(aw);
                    const bool keep = (flg >> k) & 1u;
                    // b = e1 -+ e2, c = e3 -+ e4 (lwslib.cpp:204-207, 228-231): x - y == x + (-y) bit for bit
                    const double br = __dadd_rn(e1.x, flow_flip(e2.x, sgn)), bi = __dadd_rn(e1.y, flow_flip(e2.y, sgn));
                    const double cr = __dadd_rn(e3.x, flow_flip(e4.x, sgn)), ci = __dadd_rn(e3.y, flow_flip(e4.y, sgn));
                    double vr, vi;
                    online_value(ww.x, ww.y, br, bi, cr, ci, vr, vi);
                    flow_st(ast + 16u * k, flow_keep(vr, keep), flow_keep(vi, keep));
                });
            }
        } else if (on) {
            double a;
            if (c == 0) {
                const LwsbOnlineTask task = lwsb_online_decode(T, iters, LA, Q, j);
                thr = task.thr < 0 ? 0.0 : __dmul_rn(thresholds[task.thr], mean); // lws.pyx:361, lwslib.cpp:1467
                row = task.row;
                off_own = (unsigned)((row & rmask) * pitch) * 16u;
                wtask_c = (unsigned)(task.which * Q * WB) * 16u;
                flags_c4 = 0;
                if (task.cframe) {
#pragma unroll
                    for (int pp = 0; pp < Q; ++pp) flags_c4 |= (wf[(task.which * Q + pp) * Q + 0] & 0xffu) << (8 * pp);
                }
                a = __ldg(A0 + (long long)row * P + L + c);
            } else a = a_next;
            if (c + 1 < Nreal) a_next = __ldg(A0 + (long long)row * P + L + c + 1);
            if (a > thr) { // lwslib.cpp:295-296
                const int p = c % Q;
                const bool odd = FOLD == LWSB_FOLD_Q4 && (p & 1);
                const unsigned a_cell = ring_s + off_own + (unsigned)(L + c) * 16u, a_wc = w2_s + wtask_c + (unsigned)(p * WB) * 16u;
                const unsigned flg_c = (flags_c4 >> (8 * p)) & 0xffu;
                const unsigned ast = stage_s + (unsigned)(((int)buf_c * nt + tix) * SLOT) * 16u;
                // the frame pairs in the order the reference adds them: odd bins of the Q4 folding take r = 1, 3 then 2 (lwslib.cpp:953-1052)
                const unsigned as1 = ast + (unsigned)((odd ? 2 : 1) * PER_R) * 16u, as2 = ast + (unsigned)((odd ? 1 : 2) * PER_R) * 16u;
                double tr = 0.0, ti = 0.0;
                flow_for<1, L + 1>([&](auto kc) {
                    constexpr int k = decltype(kc)::value;
                    const double2 bb = flow_ld<-k>(a_cell), cc = flow_ld<k>(a_cell), ww = flow_ld<k>(a_wc);
                    double vr, vi;
                    online_value(ww.x, ww.y, bb.x, bb.y, cc.x, cc.y, vr, vi);
                    tr = __dadd_rn(tr, flow_keep(vr, (flg_c >> k) & 1u)); ti = __dadd_rn(ti, flow_keep(vi, (flg_c >> k) & 1u));
                });
                flow_for<0, PER_R>([&](auto kc) { const double2 x = flow_ld<decltype(kc)::value>(ast); tr = __dadd_rn(tr, x.x); ti = __dadd_rn(ti, x.y); });
                if constexpr (NP > 1) flow_for<0, PER_R>([&](auto kc) { const double2 x = flow_ld<decltype(kc)::value>(as1); tr = __dadd_rn(tr, x.x); ti = __dadd_rn(ti, x.y); });
                if constexpr (NP > 2) flow_for<0, PER_R>([&](auto kc) { const double2 x = flow_ld<decltype(kc)::value>(as2); tr = __dadd_rn(tr, x.x); ti = __dadd_rn(ti, x.y); });
                double2 val;
                if (x_project(tr, ti, a, val)) {
                    flow_st(a_cell, val.x, val.y);
                    if (c >= 1 && c <= L) flow_st(a_cell - 32u * (unsigned)c, val.x, -val.y);                                            // column L - c
                    else if (c >= Nreal - 1 - L && c <= Nreal - 2) flow_st(a_cell + 32u * (unsigned)(Nreal - 1 - c), val.x, -val.y); // L + 2 (Nreal - 1) - c
                }
            }
        }
        __syncthreads();
        if (started) { if (++tmod == S) { tmod = 0; ++jhi; if (++d == nt) d = 0; } }
        if (++fmod == S) { fmod = 0; ++jf; }
    }
    for (int e = lo; e <= hi; ++e)
        if (e >= Q - 1 && e < T + Q - 1)
            for (int x = threadIdx.x; x < Np; x += nall) E0[(long long)e * P + x] = ring[(size_t)(e & rmask) * pitch + x];
}

#endif // LWSB_EXPERIMENTS

template <int Q, int FOLD, int P>
__device__ __forceinline__ void online_sum_for_residue(int p, const OnlineRingCell &E, const OnlineW<Q> &w, int ws, int rframe,
                                                       int cframe, double &tr, double &ti)
{
    if constexpr (P < Q) {
        if (p == P) online_weighted_sum<Q, P, FOLD>(E, w, ws, rframe, cframe, tr, ti);
        else online_sum_for_residue<Q, FOLD, P + 1>(p, E, w, ws, rframe, cframe, tr, ti);
    }
}

template <int Q, int FOLD>
__global__ void __launch_bounds__(256)
k_online_ring(LwsbView v, const __grid_constant__ OnlineW<Q> w, const double *thresholds, int iters, int LA, int R, int pitch,
              int S, unsigned *status)
{
    extern __shared__ __align__(16) unsigned char online_smem[];
    double2 *ring = reinterpret_cast<double2 *>(online_smem);
    // the weight sets are indexed per lane (each lane its own row update): shared memory serves divergent
    // addresses at full rate, the constant bank would replay them
    __shared__ OnlineW<Q> wsm;
    for (int i = threadIdx.x; i < (int)(sizeof(OnlineW<Q>) / 4); i += blockDim.x)
        reinterpret_cast<unsigned *>(&wsm)[i] = reinterpret_cast<const unsigned *>(&w)[i];
    __syncthreads();
    const int u = blockIdx.x;
    const int T = v.T[u], Nreal = v.Nreal, P = v.P;
    constexpr int L = OL;
    const int Np = Nreal + 2 * L, Tp = T + 2 * (Q - 1), rmask = R - 1;
    double2 *E0 = v.E + v.rowbase[u] * (long long)P + (v.c0 - L); // extended (row 0, column 0)
    const double *A0 = v.A + v.rowbase[u] * (long long)P + (v.c0 - L);
    const double mean = v.mean_amp[u];
    const long long n = lwsb_online_chain_len(T, iters, LA);
    const long long tmax = (long long)S * (n - 1) + (Nreal - 1);
    const int nt = blockDim.x;
    int lo = 0, hi = -1; // extended rows [lo, hi] are resident
    long long jc = -1;
    LwsbOnlineTask task;
    double thr = 0.0, a_next = 0.0;
    for (long long t = 0; t <= tmax; ++t) {
        if (t % S == 0) { // the front of the chain moves to a new row update: residency check (uniform across the CTA)
            const long long jhi = min(t / S, n - 1);
            long long jlo = t < Nreal ? 0 : (t - (Nreal - 1) + S - 1) / S;
            if (jlo > n - 1) jlo = n - 1;
            const int need_hi = min(Tp - 1, lwsb_online_frame(iters, LA, jhi) + 2 * (Q - 1));
            const int need_lo = max(0, lwsb_online_frame(iters, LA, jlo) - LA);
            if (need_hi > hi) {
                if (need_hi - need_lo + 1 > R && threadIdx.x == 0) atomicCAS(status, 0u, 0xE1000000u | (unsigned)u);
                for (int e = lo; e < need_lo; ++e) // rows the chain has left: back to global memory
                    if (e >= Q - 1 && e < T + Q - 1)
                        for (int x = threadIdx.x; x < Np; x += nt) E0[(long long)e * P + x] = ring[(size_t)(e & rmask) * pitch + x];
                lo = need_lo;
                __syncthreads();
                for (int e = hi + 1; e <= need_hi; ++e) // rows the chain is about to reach
                    for (int x = threadIdx.x; x < Np; x += nt) ring[(size_t)(e & rmask) * pitch + x] = E0[(long long)e * P + x];
                hi = need_hi;
                __syncthreads();
            }
        }
        const long long jhi = t / S;
        const long long d = (jhi - threadIdx.x) % nt;
        const long long j = jhi - (d < 0 ? d + nt : d);
        if (j >= 0 && j < n) {
            const int c = (int)(t - (long long)S * j);
            if (c < Nreal) {
                double a;
                if (j != jc) {
                    jc = j;
                    task = lwsb_online_decode(T, iters, LA, Q, j);
                    thr = task.thr < 0 ? 0.0 : __dmul_rn(thresholds[task.thr], mean); // lws.pyx:361, lwslib.cpp:1467
                    a = A0[(long long)task.row * P + L + c];
                } else a = a_next;
                if (c + 1 < Nreal) a_next = __ldg(A0 + (long long)task.row * P + L + c + 1); // amplitude of the next step
                if (a > thr) { // lwslib.cpp:295-296
                    double2 *Rrow = ring + (size_t)(task.row & rmask) * pitch;
                    const OnlineRingCell cell{ring, rmask, pitch, task.row, L + c};
                    double tr = 0.0, ti = 0.0;
                    online_sum_for_residue<Q, FOLD, 0>(c % Q, cell, wsm, task.which, task.rframe, task.cframe, tr, ti);
                    double2 val;
                    if (x_project(tr, ti, a, val)) {
                        Rrow[L + c] = val;
                        if (c >= 1 && c <= L) Rrow[L - c] = make_double2(val.x, -val.y);
                        else if (c >= Nreal - 1 - L && c <= Nreal - 2) Rrow[L + 2 * (Nreal - 1) - c] = make_double2(val.x, -val.y);
                    }
                }
            }
        }
        __syncthreads();
    }
    for (int e = lo; e <= hi; ++e)
        if (e >= Q - 1 && e < T + Q - 1)
            for (int x = threadIdx.x; x < Np; x += nt) E0[(long long)e * P + x] = ring[(size_t)(e & rmask) * pitch + x];
}

// Largest number of extended rows resident at once: rows are loaded when the front of the chain first needs
// them (at steps t = k*S) and rows behind the tail are dropped at the same moment -- same formulas as the kernel.
int online_max_span(int T, int Nreal, int S, int Q, int iters, int LA, int lag = 0)
{
    const long long n = lwsb_online_chain_len(T, iters, LA);
    const int Tp = T + 2 * (Q - 1);
    int span = 0;
    for (long long k = 0; k < n; ++k) {
        const long long t = k * S - lag; // the tail (k_online_rail: its chain warps) is `lag` iterations behind the front
        long long jl = t < Nreal ? 0 : (t - (Nreal - 1) + S - 1) / S;
        if (jl > n - 1) jl = n - 1;
        const int hi = std::min(Tp - 1, lwsb_online_frame(iters, LA, k) + 2 * (Q - 1));
        const int lo = std::max(0, lwsb_online_frame(iters, LA, jl) - LA);
        span = std::max(span, hi - lo + 1);
    }
    return span;
}

template <int Q, int FOLD>
cudaError_t launch_t(const LwsbView &v, const OnlineW<Q> &w, const double *thr, int iters, int LA, int R, int pitch, int S, int nt,
                     size_t bytes, size_t smem_limit, unsigned *status, int *which_kernel, int flowK, cudaStream_t s)
{
    *which_kernel = 1;
#ifdef LWSB_EXPERIMENTS
    if constexpr ((Q == 4 && FOLD == LWSB_FOLD_Q4) || (Q == 2 && FOLD == LWSB_FOLD_Q2)) {
        if (flowK < 0) { // value warps and chain warps; `nt` is the CTA size, `bytes` includes zero row, weights and staging area
            auto kern5 = k_online_rail<Q, FOLD>;
            if (bytes > 48 * 1024) {
                cudaError_t e = cudaFuncSetAttribute(kern5, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
                if (e != cudaSuccess) return e;
            }
            kern5<<<v.B, nt, bytes, s>>>(v, w, thr, iters, LA, R, pitch, S, status);
            *which_kernel = 5;
            return cudaGetLastError();
        }
    }
#endif
    if constexpr (Q <= 4) {
        if (flowK > 0) { // K warps per task taking turns; `nt` is the CTA size, `bytes` includes the zero row and the weights
            auto kern4 = flowK == 4 ? k_online_flow<Q, FOLD, 4> : (flowK == 3 ? k_online_flow<Q, FOLD, 3> : k_online_flow<Q, FOLD, 2>);
            if (bytes > 48 * 1024) {
                cudaError_t e = cudaFuncSetAttribute(kern4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
                if (e != cudaSuccess) return e;
            }
            kern4<<<v.B, nt, bytes, s>>>(v, w, thr, iters, LA, R, pitch, S, status);
            *which_kernel = 4;
            return cudaGetLastError();
        }
        const char *e_duo = getenv("LWSB_ONLINE_DUO");
        if (S >= 2 + OL && S % 2 == 0 && 2 * nt <= 256 && !(e_duo && atoi(e_duo) == 0)) { // two bins per step on two lanes
            auto kern3 = k_online_duo<Q, FOLD>;
            if (bytes > 48 * 1024) {
                cudaError_t e = cudaFuncSetAttribute(kern3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
                if (e != cudaSuccess) return e;
            }
            kern3<<<v.B, 2 * nt, bytes, s>>>(v, w, thr, iters, LA, R, pitch, S, status);
            *which_kernel = 3;
            return cudaGetLastError();
        }
        if (S >= 2 + OL && S % 2 == 0) { // two bins per step
            *which_kernel = 2;
            auto kern2 = k_online_ring2<Q, FOLD>;
            if (bytes > 48 * 1024) {
                cudaError_t e = cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
                if (e != cudaSuccess) return e;
            }
            kern2<<<v.B, nt, bytes, s>>>(v, w, thr, iters, LA, R, pitch, S, status);
            return cudaGetLastError();
        }
    }
    auto kern = k_online_ring<Q, FOLD>;
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return e;
    }
    kern<<<v.B, nt, bytes, s>>>(v, w, thr, iters, LA, R, pitch, S, status);
    return cudaGetLastError();
}

template <int Q>
cudaError_t launch_q(const LwsbView &v, const double *const *wr, const double *const *wi, int fold, const double *thr, int iters,
                     int LA, int R, int pitch, int S, int nt, size_t bytes, size_t smem_limit, unsigned *status, int *which_kernel, int flowK,
                     cudaStream_t s)
{
    OnlineW<Q> w;
    for (int ws = 0; ws < 3; ++ws)
        for (int p = 0; p < Q; ++p)
            for (int r = 0; r < Q; ++r) {
                unsigned f = 0;
                for (int k = 0; k <= OL; ++k) {
                    const size_t i = ((size_t)p * Q + r) * (OL + 1) + k;
                    w.wr[ws][p][r][k] = wr[ws][i]; w.wi[ws][p][r][k] = wi[ws][i];
                    if (std::hypot(wr[ws][i], wi[ws][i]) > 1.0e-12) f |= 1u << k; // lws.pyx:347-352
                }
                w.flag[ws][p][r] = f;
            }
    if (fold == LWSB_FOLD_ANY) return launch_t<Q, LWSB_FOLD_ANY>(v, w, thr, iters, LA, R, pitch, S, nt, bytes, smem_limit, status, which_kernel, flowK, s);
    if constexpr (Q == 4) { if (fold == LWSB_FOLD_Q4) return launch_t<4, LWSB_FOLD_Q4>(v, w, thr, iters, LA, R, pitch, S, nt, bytes, smem_limit, status, which_kernel, flowK, s); }
    if constexpr (Q == 2) { if (fold == LWSB_FOLD_Q2) return launch_t<2, LWSB_FOLD_Q2>(v, w, thr, iters, LA, R, pitch, S, nt, bytes, smem_limit, status, which_kernel, flowK, s); }
    return cudaErrorInvalidValue;
}

} // namespace

// returns false when this kernel does not serve the shape (the global-memory kernel takes over)
bool launch_online_ring(const LwsbView &v, const double *const *wr_host, const double *const *wi_host, int fold,
                        const double *thr, int iters, int LA, const int *T_host, size_t smem_limit, unsigned *status,
                        cudaStream_t s, cudaError_t *err, int *which_kernel)
{
    *err = cudaSuccess;
    *which_kernel = 0;
    const int Q = v.Q;
    if (v.L != OL || !(Q == 2 || Q == 4 || Q == 8)) return false;
    // smallest multiple of Q >= L + 1 (Q > 4: one bin per step) or >= L + 2 (Q <= 4: two bins per step): every thread of a
    // step on the same residue, and the bins a task takes in one step stay clear of its neighbours' in the chain
    int S = Q <= 4 ? (OL + 2 + Q - 1) / Q * Q : (OL + 1 + Q - 1) / Q * Q;
    int flowK = 0;
#ifdef LWSB_EXPERIMENTS
    if (Q <= 4 && (fold == LWSB_FOLD_Q4 || fold == LWSB_FOLD_Q2)) { // k_online_rail: value warps + chain warps (LWSB_ONLINE_RAIL=1 enables, LWSB_ONLINE_RAIL_S: the lag)
        const char *e_rail = getenv("LWSB_ONLINE_RAIL"), *e_s = getenv("LWSB_ONLINE_RAIL_S"), *e_pm = getenv("LWSB_ONLINE_FLOW_PITCH");
        if (e_rail && atoi(e_rail) == 1) {
            const int Sr = std::max(OL + 2, e_s ? atoi(e_s) : 9);
            const int tasks = (v.Nreal + Sr - 1) / Sr + 2, G = (tasks + 31) / 32;
            if (G * Q * 32 <= 256) {
                int spanr = 0, lastTr = -1;
                for (int b = 0; b < v.B; ++b)
                    if (T_host[b] != lastTr) { lastTr = T_host[b]; spanr = std::max(spanr, online_max_span(lastTr, v.Nreal, Sr, Q, iters, LA, 1)); }
                int Rr = 8;
                while (Rr < spanr) Rr *= 2;
                const int pm = e_pm ? atoi(e_pm) & 7 : 2; // see k_online_flow below: consecutive lanes in consecutive bank groups
                int pitchr = v.Nreal + 2 * OL;
                while ((pitchr & 7) != pm) ++pitchr;
                const size_t bytesr = (size_t)(Rr + 1) * pitchr * sizeof(double2) + (size_t)3 * Q * (Q * (OL + 1) + 1) * 16 +
                                      (size_t)2 * (32 * G) * ((Q - 1) * (OL + 1) + 1) * 16 + (size_t)3 * Q * Q * 4;
                if (bytesr + 1024 <= smem_limit) {
                    switch (Q) {
                    case 2: *err = launch_q<2>(v, wr_host, wi_host, fold, thr, iters, LA, Rr, pitchr, Sr, G * Q * 32, bytesr, smem_limit, status, which_kernel, -1, s); break;
                    case 4: *err = launch_q<4>(v, wr_host, wi_host, fold, thr, iters, LA, Rr, pitchr, Sr, G * Q * 32, bytesr, smem_limit, status, which_kernel, -1, s); break;
                    }
                    return true;
                }
            }
        }
    }
#endif
    if (Q <= 4) { // k_online_flow: K warps per task, any lag S >= K + L (LWSB_ONLINE_FLOW=0 disables, =2/3 sets K; LWSB_ONLINE_FLOW_S the lag)
        const char *e_flow = getenv("LWSB_ONLINE_FLOW"), *e_s = getenv("LWSB_ONLINE_FLOW_S");
        int K = e_flow ? atoi(e_flow) : 4;
        if (K == 1 || K > 4) K = 4;
        if (K >= 2) {
            const int Sf = std::max(K + OL, e_s ? atoi(e_s) : 0);
            const int tasks = (v.Nreal + Sf - 1) / Sf + 1, G = (tasks + 31) / 32;
            if (G * K * 32 <= 256) {
                int spanf = 0, lastTf = -1;
                for (int b = 0; b < v.B; ++b)
                    if (T_host[b] != lastTf) { lastTf = T_host[b]; spanf = std::max(spanf, online_max_span(lastTf, v.Nreal, Sf, Q, iters, LA)); }
                int Rf = 8;
                while (Rf < spanf) Rf *= 2;
                // the lanes of a warp are on consecutive row updates: columns S apart, frames cycling through the look-ahead
                // window.  With pitch = 2 (mod 8) cells and S = 9 the bank group advances by one from lane to lane whatever
                // the look-ahead (2 frame - column), i.e. a half-warp's 16 cells fall two per bank group: no replays
                const char *e_pm = getenv("LWSB_ONLINE_FLOW_PITCH");
                const int pm = e_pm ? atoi(e_pm) & 7 : 2;
                int pitchf = v.Nreal + 2 * OL;
                while ((pitchf & 7) != pm) ++pitchf;
                const size_t bytesf = (size_t)(Rf + 1) * pitchf * sizeof(double2) + (size_t)3 * Q * ((Q * (OL + 1) + 1) * 16 + Q * 4);
                if (bytesf + 1024 <= smem_limit) {
                    const char *e_map = getenv("LWSB_ONLINE_FLOW_MAP");
                    flowK = K; S = Sf | (((e_map ? atoi(e_map) : 1) & 3) << 16);
                    switch (Q) {
                    case 2: *err = launch_q<2>(v, wr_host, wi_host, fold, thr, iters, LA, Rf, pitchf, S, G * K * 32, bytesf, smem_limit, status, which_kernel, flowK, s); break;
                    case 4: *err = launch_q<4>(v, wr_host, wi_host, fold, thr, iters, LA, Rf, pitchf, S, G * K * 32, bytesf, smem_limit, status, which_kernel, flowK, s); break;
                    }
                    return true;
                }
            }
        }
    }
    int span = 0, lastT = -1;
    for (int b = 0; b < v.B; ++b) // exact span for every distinct length of the batch
        if (T_host[b] != lastT) { lastT = T_host[b]; span = std::max(span, online_max_span(lastT, v.Nreal, S, Q, iters, LA)); }
    int R = 8;
    while (R < span) R *= 2;
    int pitch = v.Nreal + 2 * OL;
    if ((pitch & 1) == 0) ++pitch;
    const size_t bytes = (size_t)R * pitch * sizeof(double2);
    if (bytes + (size_t)Q * Q * 3 * ((OL + 1) * 16 + 4) + 1024 > smem_limit) return false; // ring + the weight sets
    int nt = (v.Nreal + S - 1) / S + 1;
    nt = (nt + 31) / 32 * 32;
    if (nt > 256) return false;
    switch (Q) {
    case 2: *err = launch_q<2>(v, wr_host, wi_host, fold, thr, iters, LA, R, pitch, S, nt, bytes, smem_limit, status, which_kernel, flowK, s); break;
    case 4: *err = launch_q<4>(v, wr_host, wi_host, fold, thr, iters, LA, R, pitch, S, nt, bytes, smem_limit, status, which_kernel, flowK, s); break;
    case 8: *err = launch_q<8>(v, wr_host, wi_host, fold, thr, iters, LA, R, pitch, S, nt, bytes, smem_limit, status, which_kernel, flowK, s); break;
    }
    return true;
}

} // namespace lwsb

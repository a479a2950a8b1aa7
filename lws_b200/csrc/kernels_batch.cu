// kernels_batch.cu -- the tuned batch_lws kernel: column strips on a thread-block cluster.
//
// Work decomposition (DESIGN.md has the derivation and the dependency proofs)
//   * One CLUSTER of C CTAs per utterance; CTA c ("strip c") owns the bins [c*W, (c+1)*W),
//     W = 8*NBr, of EVERY frame and streams the frames through a ring of R rows in its own
//     shared memory (row = W + 2L complex128: own bins plus an L-bin halo on either side).
//     Rows enter by TMA bulk copies (cp.async.bulk, mbarrier-signalled) LEAD frames ahead of
//     their first use and leave by TMA bulk stores when their last sweep of the pass is done.
//   * Bins are processed in BLOCKS of 8 consecutive bins ("macro-step").  Within a strip the
//     thread of frame m runs 2 blocks behind the thread of frame m-1 (the stencil reaches
//     5 bins, i.e. into the next block, so 2 blocks is the smallest safe lag) and sweep g+1
//     runs Q frames behind sweep g: thread (j, g), j in [0, NS), g in [0, G), handles block
//     xb = (t - 2j) mod NBV of frame  j + NS*floor((t - 2j)/NBV) - Q*g  at macro-step t.
//     All of this is the reference's raster order relaxed only where the data dependences
//     allow it, so every bin sees exactly the neighbour values the sequential code sees.
//   * G sweeps are in flight per pass (as many as the ring has room for); a call of `iters`
//     sweeps takes ceil(active/G) passes over the utterance, where sweeps whose threshold is
//     not below max|S| are dropped up front (they cannot move any bin).
//   * Strips run in lock step, NBr macro-steps apart (strip c+1 behind strip c).  A bin block
//     on a strip edge is written into the neighbour's halo through distributed shared memory;
//     after every macro-step the control warp publishes the strip's progress to both
//     neighbours (st.release.cluster) and polls theirs (ld.acquire.cluster) while the compute
//     warps work on the next macro-step.
//   * Arithmetic is the reference's, operation for operation (exact.cuh): results are
//     bit-identical to lwslib's LWSQ2 / LWSQ4 / LWSanyQ.
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <type_traits>
#include "exact.cuh"
#include "fast_math.cuh"
#include "kernels.h"
#include "lwsb_common.h"

namespace cg = cooperative_groups;

namespace lwsb {

namespace {

constexpr int SL = 5;          // stencil reach in bins this kernel is specialised for (L)
constexpr int SBK = 8;         // bins per block
constexpr int SLEAD = 2;       // frames of TMA look-ahead
constexpr unsigned SPIN_LIMIT = 1u << 24; // polls before a wait is declared dead (seconds)
constexpr unsigned PASS_SPIN_LIMIT = 1u << 27; // waits for another cluster's pass: it may still be busy with earlier work items
#ifdef LWSB_PAIR_EXPERIMENTS
constexpr int pair_thread_cap(int max_sweeps) { return max_sweeps < 0 ? 512 : 256; }
#else
constexpr int pair_thread_cap(int) { return 256; }
#endif
// planner cost model of the pair-split kernels (cycles per macro-step: fixed + per warp; weight of the bank-conflict factor)
constexpr double PAIR_T0 = 8000.0, PAIR_T1 = 500.0, PAIR_TF = 0.3;
constexpr int PAIR_THREADS_MAX = 512;  // pair-split kernels: 15 task warps (240 tasks) + the control warp at 128 registers
constexpr int PAIR_THREADS_PLAN = 256; // what the planner uses: 7 task warps + the control warp keep 255 registers (measured fastest)
#ifndef LWSB_PAIR_DEFAULT_MODE
#define LWSB_PAIR_DEFAULT_MODE 1 // odd frame pairs (r = 1, 3) in register windows, no explicit pipelining
#endif // pair-split kernels: 15 task warps (240 tasks) + the control warp at 128 registers

template <int Q>
struct StripW {                // one weight set, reference layout, in the kernel parameter bank
    double wr[Q][Q][SL + 1];
    double wi[Q][Q][SL + 1];
    unsigned flag[Q][Q];       // bit k: |W[p][r][k]| > 1e-12
    int fold;                  // LWSB_FOLD_*
};

struct StripParams {
    LwsbView v;
    const double *thr;         // [iters] unscaled thresholds
    const double *max_amp;     // [B]
    int iters;
    int C, NBr, NBV, NS, G, R, pitch, QS, GFAST;
    unsigned *status;          // [0]: 0 ok, else first watchdog code
    // work list: one item per (utterance, pass), pass-major; cluster k takes items k, k + #clusters, ...  A pass reads
    // each frame after the previous pass of the same utterance -- possibly running on another cluster at the
    // same time -- has written it back: done[((u * max_pass) + pass) * 8 + strip] counts the frames that pass has
    // written so far (one writer per counter: the counters only grow)
    const int2 *items;
    int n_items, max_pass;
    unsigned *done;
    unsigned long long *trace; // optional [n_items][8]: globaltimer stamps (ns) item taken, ring primed, last macro-step done, written back;
                               // cycles of strip 0: control lane in row waits / neighbour polls, compute warp 0 at work / waiting for the control warp
};

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_inval(uint64_t *bar)
{
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, unsigned parity)
{
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// global -> shared bulk copy (TMA), completion counted in bytes on an mbarrier of this CTA
__device__ __forceinline__ void tma_load_row(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global bulk copy (TMA), tracked by the bulk async-group of the issuing thread
__device__ __forceinline__ void tma_store_row(void *dst, const void *src, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ unsigned ld_acquire_cluster(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.cluster.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_cluster(unsigned *p, unsigned v)
{
    asm volatile("st.release.cluster.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned *p, unsigned v)
{
    asm volatile("st.release.gpu.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// all but the most recent bulk store of this thread have been written to global memory
__device__ __forceinline__ void tma_store_wait_all_but_one() { asm volatile("cp.async.bulk.wait_group 1;" ::: "memory"); }

// CTA-wide barrier usable from the role-split (control / compute) code paths
__device__ __forceinline__ void cta_sync() { asm volatile("bar.sync 1;" ::: "memory"); }

// Bounded waiting: a protocol error must end the kernel, not hang the GPU.  Returns false when
// the wait has to be abandoned (this or another CTA recorded a time-out in *status).
__device__ __forceinline__ bool keep_waiting(unsigned &spins, unsigned *status, unsigned code)
{
    if ((++spins & 1023u) == 0) {
        if (*reinterpret_cast<volatile unsigned *>(status) != 0u) return false;
        if (spins > SPIN_LIMIT) { atomicCAS(status, 0u, code); return false; }
    }
    return true;
}

// ---------------------------------------------------------------- one block of 8 bins
// Accessor over the shared-memory ring: rowoff[dr + Q - 1] is the byte offset of ring row
// (frame + dr) and `col` the ring column of the bin being updated.
template <int Q>
struct RingCell {
    const unsigned char *ring;
    unsigned rowoff[2 * Q - 1];
    int col;
    __device__ __forceinline__ double2 operator()(int dr, int dk) const
    {
        return *reinterpret_cast<const double2 *>(ring + rowoff[dr + Q - 1] + (unsigned)(col + dk) * 16u);
    }
};

// the reference's weighted sum with everything but the data resolved at compile time:
// P = bin mod Q, weights / flags straight from the parameter bank
template <int Q, int P, int FOLD>
__device__ __forceinline__ void strip_weighted_sum(const RingCell<Q> &E, const StripW<Q> &w, double &tr, double &ti)
{
    constexpr int PN = (Q - P) % Q;
    tr = 0.0; ti = 0.0;
    // centre frame, bins n -+ k (lwslib.cpp:88-101, 169-182, 299-312)
#pragma unroll
    for (int k = 1; k <= SL; ++k)
        if (w.flag[P][0] & (1u << k)) {
            const double2 b = E(0, -k), c = E(0, +k);
            x_pair(tr, ti, w.wr[P][0][k], w.wi[P][0][k], b.x, b.y, c.x, c.y);
        }
    auto both = [&](auto rc, auto minusc) {
        constexpr int r = decltype(rc)::value;
        constexpr bool minus = decltype(minusc)::value;
        if (w.flag[P][r] & 1u) {
            const double2 b = E(-r, 0), c = E(+r, 0);
            x_pair(tr, ti, w.wr[P][r][0], w.wi[P][r][0], b.x, b.y, c.x, c.y);
        }
#pragma unroll
        for (int k = 1; k <= SL; ++k) {
            if (FOLD == LWSB_FOLD_ANY) {
                if (w.flag[P][r] & (1u << k)) {
                    const double2 b = E(-r, -k), c = E(+r, -k);
                    x_pair(tr, ti, w.wr[P][r][k], w.wi[P][r][k], b.x, b.y, c.x, c.y);
                }
                if (w.flag[PN][r] & (1u << k)) {
                    const double2 b = E(+r, +k), c = E(-r, +k);
                    x_pair(tr, ti, w.wr[PN][r][k], w.wi[PN][r][k], b.x, b.y, c.x, c.y);
                }
            } else if (w.flag[P][r] & (1u << k)) {
                const double2 e1 = E(-r, -k), e2 = E(+r, +k), e3 = E(+r, -k), e4 = E(-r, +k);
                double br, bi, cr, ci;
                if (minus) {
                    br = __dsub_rn(e1.x, e2.x); bi = __dsub_rn(e1.y, e2.y);
                    cr = __dsub_rn(e3.x, e4.x); ci = __dsub_rn(e3.y, e4.y);
                } else {
                    br = __dadd_rn(e1.x, e2.x); bi = __dadd_rn(e1.y, e2.y);
                    cr = __dadd_rn(e3.x, e4.x); ci = __dadd_rn(e3.y, e4.y);
                }
                x_pair(tr, ti, w.wr[P][r][k], w.wi[P][r][k], br, bi, cr, ci);
            }
        }
    };
    using T_ = std::true_type;
    using F_ = std::false_type;
    if constexpr (FOLD == LWSB_FOLD_Q4 && (P & 1)) {
        // odd bins: r = 1, 3 with the sign-flipped folding, then r = 2 (lwslib.cpp:186-235)
        static_assert(FOLD != LWSB_FOLD_Q4 || Q == 4, "the Q4 folding is defined for Q = 4");
        both(std::integral_constant<int, 1>{}, T_{});
        both(std::integral_constant<int, 3>{}, T_{});
        both(std::integral_constant<int, 2>{}, F_{});
    } else {
        if constexpr (Q > 1) both(std::integral_constant<int, 1>{}, F_{});
        if constexpr (Q > 2) both(std::integral_constant<int, 2>{}, F_{});
        if constexpr (Q > 3) both(std::integral_constant<int, 3>{}, F_{});
        if constexpr (Q > 4) both(std::integral_constant<int, 4>{}, F_{});
        if constexpr (Q > 5) both(std::integral_constant<int, 5>{}, F_{});
        if constexpr (Q > 6) both(std::integral_constant<int, 6>{}, F_{});
        if constexpr (Q > 7) both(std::integral_constant<int, 7>{}, F_{});
    }
}

struct BlockCtx {
    unsigned char *ring;        // this CTA's ring
    unsigned char *ring_left;   // left / right neighbour's ring through DSMEM (nullptr at the ends)
    unsigned char *ring_right;
    unsigned ownoff;            // byte offset of the frame's ring row
    int xb;                     // block index inside the strip
    int n0;                     // first bin of the block (global bin index)
    int b0;                     // first bin of the strip
    int Nreal, NBr;
    bool first_strip;
};

template <int Q, int P, int FOLD>
__device__ __forceinline__ void strip_update_bin(RingCell<Q> &cell, const StripW<Q> &w, const BlockCtx &bc, int i, double a)
{
    cell.col = SL + SBK * bc.xb + i;
    double tr, ti;
    strip_weighted_sum<Q, P, FOLD>(cell, w, tr, ti);
    double2 val;
    if (!x_project(tr, ti, a, val)) return;
    const int n = bc.n0 + i;
    double2 *own = reinterpret_cast<double2 *>(bc.ring + bc.ownoff);
    own[cell.col] = val;
    const double2 cj = make_double2(val.x, -val.y);
    // mirrored copies, refreshed at once (lwslib.cpp:362-368); ring column of bin q is SL + q - b0
    if (n >= 1 && n <= SL) { if (bc.first_strip) own[SL - n] = cj; }
    else if (n >= bc.Nreal - 1 - SL && n <= bc.Nreal - 2) own[SL + 2 * (bc.Nreal - 1) - n - bc.b0] = cj;
    // halo copies in the neighbouring strips (distributed shared memory)
    if (bc.xb == 0 && i < SL && bc.ring_left)
        reinterpret_cast<double2 *>(bc.ring_left + bc.ownoff)[SL + SBK * bc.NBr + i] = val;
    if (bc.xb == bc.NBr - 1 && i >= SBK - SL && bc.ring_right)
        reinterpret_cast<double2 *>(bc.ring_right + bc.ownoff)[i - (SBK - SL)] = val;
}

template <int Q, int FOLD, int I>
__device__ __forceinline__ void strip_update_block(RingCell<Q> &cell, const StripW<Q> &w, const BlockCtx &bc,
                                                   const double *amp, unsigned active)
{
    // bins of a block in order; the block starts at a multiple of 8 bins, so bin I has residue I mod Q
    if constexpr (I < SBK) {
        if (active & (1u << I)) strip_update_bin<Q, I % Q, FOLD>(cell, w, bc, I, amp[I]);
        strip_update_block<Q, FOLD, I + 1>(cell, w, bc, amp, active);
    }
}

// ---------------------------------------------------------------- software-pipelined block update (Q <= 4)
// The reference fixes the ORDER in which the terms of a bin are added, not when their values
// are computed: every inter-frame term  ar*(br+cr) - ai*(bi-ci)  depends only on neighbour frames,
// so the values for bin i+1 are formed (loads included) while the additions, the square root and
// the division of bin i -- one long dependent chain -- are in flight.  The code below is branch
// free inside a block so that the instruction scheduler can interleave the two streams.
//
// PAT selects how the |W| > 1e-12 mask is applied: 1 = the pattern of the default sqrt-Hann
// windows, known at compile time (r = 0: k = 1; r = Q/2: k in {0,1,2,4}; all other k set; the
// host checks the actual mask against it), 0 = any mask, applied with selects at run time.
template <int Q, int PAT>
__device__ __forceinline__ constexpr bool pat_has(int r, int k)
{
    if (PAT == 0) return true;
    if (r == 0) return k == 1;
    if (2 * r == Q) return k == 0 || k == 1 || k == 2 || k == 4;
    return true;
}

template <int FOLD>
struct TermCount { static constexpr int per_r = FOLD == LWSB_FOLD_ANY ? 1 + 2 * SL : 1 + SL; };

template <int Q, int FOLD>
struct BinTerms { // values of the inter-frame terms of one bin, in the reference's order of addition
    static constexpr int N = (Q - 1) * TermCount<FOLD>::per_r;
    double r[N], i[N];
};

__device__ __forceinline__ void pair_value(double ar, double ai, double br, double bi, double cr, double ci, double &vr, double &vi)
{
    vr = __dsub_rn(__dmul_rn(ar, __dadd_rn(br, cr)), __dmul_rn(ai, __dsub_rn(bi, ci)));
    vi = __dadd_rn(__dmul_rn(ar, __dadd_rn(bi, ci)), __dmul_rn(ai, __dsub_rn(br, cr)));
}

// term values of frame pair (m - R_, m + R_) into slots [BASE, BASE + per_r)
template <int Q, int P, int FOLD, int PAT, int R_, bool MINUS, int BASE>
__device__ __forceinline__ void term_values_r(const RingCell<Q> &E, const StripW<Q> &w, BinTerms<Q, FOLD> &tv)
{
    constexpr int PN = (Q - P) % Q;
    if (pat_has<Q, PAT>(R_, 0)) {
        const double2 b = E(-R_, 0), c = E(+R_, 0);
        pair_value(w.wr[P][R_][0], w.wi[P][R_][0], b.x, b.y, c.x, c.y, tv.r[BASE], tv.i[BASE]);
    }
#pragma unroll
    for (int k = 1; k <= SL; ++k) {
        if (FOLD == LWSB_FOLD_ANY) {
            if (pat_has<Q, PAT>(R_, k)) {
                const double2 b = E(-R_, -k), c = E(+R_, -k);
                pair_value(w.wr[P][R_][k], w.wi[P][R_][k], b.x, b.y, c.x, c.y, tv.r[BASE + 2 * k - 1], tv.i[BASE + 2 * k - 1]);
                const double2 b2 = E(+R_, +k), c2 = E(-R_, +k);
                pair_value(w.wr[PN][R_][k], w.wi[PN][R_][k], b2.x, b2.y, c2.x, c2.y, tv.r[BASE + 2 * k], tv.i[BASE + 2 * k]);
            }
        } else if (pat_has<Q, PAT>(R_, k)) {
            const double2 e1 = E(-R_, -k), e2 = E(+R_, +k), e3 = E(+R_, -k), e4 = E(-R_, +k);
            double br, bi, cr, ci;
            if (MINUS) {
                br = __dsub_rn(e1.x, e2.x); bi = __dsub_rn(e1.y, e2.y);
                cr = __dsub_rn(e3.x, e4.x); ci = __dsub_rn(e3.y, e4.y);
            } else {
                br = __dadd_rn(e1.x, e2.x); bi = __dadd_rn(e1.y, e2.y);
                cr = __dadd_rn(e3.x, e4.x); ci = __dadd_rn(e3.y, e4.y);
            }
            pair_value(w.wr[P][R_][k], w.wi[P][R_][k], br, bi, cr, ci, tv.r[BASE + k], tv.i[BASE + k]);
        }
    }
}

// add the values of frame pair R_ in the reference's order
template <int Q, int P, int FOLD, int PAT, int R_, int BASE>
__device__ __forceinline__ void term_accumulate_r(const StripW<Q> &w, const BinTerms<Q, FOLD> &tv, double &tr, double &ti)
{
    constexpr int PN = (Q - P) % Q;
    auto add = [&](int slot, unsigned flagword, int k) {
        if (PAT == 1) { tr = __dadd_rn(tr, tv.r[slot]); ti = __dadd_rn(ti, tv.i[slot]); }
        else {
            const bool f = (flagword >> k) & 1u;
            const double nr = __dadd_rn(tr, tv.r[slot]), ni = __dadd_rn(ti, tv.i[slot]);
            tr = f ? nr : tr; ti = f ? ni : ti;
        }
    };
    if (pat_has<Q, PAT>(R_, 0)) add(BASE, w.flag[P][R_], 0);
#pragma unroll
    for (int k = 1; k <= SL; ++k) {
        if (!pat_has<Q, PAT>(R_, k)) continue;
        if (FOLD == LWSB_FOLD_ANY) {
            add(BASE + 2 * k - 1, w.flag[P][R_], k);
            add(BASE + 2 * k, w.flag[PN][R_], k);
        } else add(BASE + k, w.flag[P][R_], k);
    }
}

template <int Q, int P, int FOLD, int PAT>
__device__ __forceinline__ void bin_term_values(const RingCell<Q> &E, const StripW<Q> &w, BinTerms<Q, FOLD> &tv)
{
    constexpr int TPR = TermCount<FOLD>::per_r;
    if constexpr (FOLD == LWSB_FOLD_Q4 && (P & 1)) { // odd bins: r = 1, 3 sign-flipped, then r = 2 (lwslib.cpp:186-235)
        term_values_r<Q, P, FOLD, PAT, 1, true, 0>(E, w, tv);
        term_values_r<Q, P, FOLD, PAT, 3, true, TPR>(E, w, tv);
        term_values_r<Q, P, FOLD, PAT, 2, false, 2 * TPR>(E, w, tv);
    } else {
        if constexpr (Q > 1) term_values_r<Q, P, FOLD, PAT, 1, false, 0>(E, w, tv);
        if constexpr (Q > 2) term_values_r<Q, P, FOLD, PAT, 2, false, TPR>(E, w, tv);
        if constexpr (Q > 3) term_values_r<Q, P, FOLD, PAT, 3, false, 2 * TPR>(E, w, tv);
    }
}

template <int Q, int P, int FOLD, int PAT>
__device__ __forceinline__ void bin_accumulate(const StripW<Q> &w, const BinTerms<Q, FOLD> &tv, double &tr, double &ti)
{
    constexpr int TPR = TermCount<FOLD>::per_r;
    if constexpr (FOLD == LWSB_FOLD_Q4 && (P & 1)) {
        term_accumulate_r<Q, P, FOLD, PAT, 1, 0>(w, tv, tr, ti);
        term_accumulate_r<Q, P, FOLD, PAT, 3, TPR>(w, tv, tr, ti);
        term_accumulate_r<Q, P, FOLD, PAT, 2, 2 * TPR>(w, tv, tr, ti);
    } else {
        if constexpr (Q > 1) term_accumulate_r<Q, P, FOLD, PAT, 1, 0>(w, tv, tr, ti);
        if constexpr (Q > 2) term_accumulate_r<Q, P, FOLD, PAT, 2, TPR>(w, tv, tr, ti);
        if constexpr (Q > 3) term_accumulate_r<Q, P, FOLD, PAT, 3, 2 * TPR>(w, tv, tr, ti);
    }
}

// bins I .. 7 of a block; `tv` holds the inter-frame term values of bin I on entry
template <int Q, int FOLD, int PAT, int I>
__device__ __forceinline__ void pipelined_block(RingCell<Q> &cell, const StripW<Q> &w, const BlockCtx &bc, const double *amp,
                                                unsigned active, BinTerms<Q, FOLD> &tv, double2 *newv, unsigned &committed)
{
    if constexpr (I < SBK) {
        constexpr int P = I % Q;
        const int col = SL + SBK * bc.xb + I;
        // (1) the next bin's inter-frame terms: independent of everything below
        BinTerms<Q, FOLD> tvn;
        if constexpr (I + 1 < SBK) {
            cell.col = col + 1;
            bin_term_values<Q, (I + 1) % Q, FOLD, PAT>(cell, w, tvn);
        }
        // (2) this bin: centre-frame terms (they see the bins just updated), then the ordered sum
        cell.col = col;
        double tr = 0.0, ti = 0.0;
#pragma unroll
        for (int k = 1; k <= SL; ++k)
            if (pat_has<Q, PAT>(0, k)) {
                const double2 b = cell(0, -k), c = cell(0, +k);
                double vr, vi;
                pair_value(w.wr[P][0][k], w.wi[P][0][k], b.x, b.y, c.x, c.y, vr, vi);
                if (PAT == 1) { tr = __dadd_rn(tr, vr); ti = __dadd_rn(ti, vi); }
                else {
                    const bool f = (w.flag[P][0] >> k) & 1u;
                    const double nr = __dadd_rn(tr, vr), ni = __dadd_rn(ti, vi);
                    tr = f ? nr : tr; ti = f ? ni : ti;
                }
            }
        bin_accumulate<Q, P, FOLD, PAT>(w, tv, tr, ti);
        // |t| = sqrt(tr*tr + ti*ti), new value (t * a) / |t| (lwslib.cpp:355-360), branch free (fast_math.cuh): one
        // reciprocal serves both divisions; the rare bin outside the fast ranges goes through the library functions
        double2 val;
        const double x = __dadd_rn(__dmul_rn(tr, tr), __dmul_rn(ti, ti));
        const double nr = __dmul_rn(tr, amp[I]), ni = __dmul_rn(ti, amp[I]);
        const bool act = (active >> I) & 1u;
        bool sok, rok, dok1, dok2;
        double mag = fm_sqrt(x, sok);
        const double rcp = fm_rcp(mag, rok);
        val.x = fm_div(nr, mag, rcp, dok1);
        val.y = fm_div(ni, mag, rcp, dok2);
        if (act && !(x == 0.0) && !(sok && rok && dok1 && dok2)) {
            mag = __dsqrt_rn(x);
            val.x = __ddiv_rn(nr, mag); val.y = __ddiv_rn(ni, mag);
        }
        const bool ok = act && x > 0.0;
        // (3) commit: own cell and its mirrored copy (lwslib.cpp:356-368)
        double2 *own = reinterpret_cast<double2 *>(bc.ring + bc.ownoff);
        const int n = bc.n0 + I;
        int mcol = col;
        if (bc.first_strip && n >= 1 && n <= SL) mcol = SL - n;
        else if (n >= bc.Nreal - 1 - SL && n <= bc.Nreal - 2) mcol = SL + 2 * (bc.Nreal - 1) - n - bc.b0;
        if (ok) {
            own[col] = val;
            own[mcol] = make_double2(val.x, mcol == col ? val.y : -val.y);
            committed |= 1u << I;
        }
        newv[I] = val;
        if constexpr (I + 1 < SBK) pipelined_block<Q, FOLD, PAT, I + 1>(cell, w, bc, amp, active, tvn, newv, committed);
    }
}

template <int Q, int FOLD, int PAT>
__device__ __forceinline__ void strip_update_block_pipelined(RingCell<Q> &cell, const StripW<Q> &w, const BlockCtx &bc,
                                                             const double *amp, unsigned active)
{
    BinTerms<Q, FOLD> tv;
    cell.col = SL + SBK * bc.xb;
    bin_term_values<Q, 0, FOLD, PAT>(cell, w, tv);
    double2 newv[SBK];
    unsigned committed = 0;
    pipelined_block<Q, FOLD, PAT, 0>(cell, w, bc, amp, active, tv, newv, committed);
    // halo copies in the neighbouring strips (distributed shared memory), edge blocks only
    if (bc.xb == 0 && bc.ring_left) {
        double2 *dst = reinterpret_cast<double2 *>(bc.ring_left + bc.ownoff) + SL + SBK * bc.NBr;
#pragma unroll
        for (int i = 0; i < SL; ++i)
            if ((committed >> i) & 1u) dst[i] = newv[i];
    }
    if (bc.xb == bc.NBr - 1 && bc.ring_right) {
        double2 *dst = reinterpret_cast<double2 *>(bc.ring_right + bc.ownoff);
#pragma unroll
        for (int i = SBK - SL; i < SBK; ++i)
            if ((committed >> i) & 1u) dst[i - (SBK - SL)] = newv[i];
    }
}

#include "strip_pair.cuh"

// ---------------------------------------------------------------- tensor memory as a term-value scratch pad (TM kernels)
// The register file cannot hold the 59 neighbour values of a bin for several bins at once, and shared memory is
// full of ring rows, but the SM's 256 KB of tensor memory is idle in this (tensor-core free) kernel.  TMEM is
// private to a lane quarter, and warps w and w+4 share a quarter: a PRODUCER warp (4..7) loads the neighbour
// frames once per half block (two 14-bin windows per frame pair, kept in registers and reused by four bins),
// forms the inter-frame term values of four bins and parks them in TMEM (tcgen05.st); its CONSUMER twin (0..3)
// fetches them (tcgen05.ld) and runs the order-bound part: the additions in the reference's order, sqrt, divide,
// commit.  The two warps sit on the same scheduler, so the producer's issue-bound stream fills the consumer's
// dependency stalls.  Layout per lane: buffer h (half block) at columns [256h, 256h+256), bin b of the half at
// +64b, term number n (in order of addition) at +4n: {re.lo, re.hi, im.lo, im.hi}.
__device__ __forceinline__ void tm_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tm_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void pair_arrive(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void pair_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

__device__ __forceinline__ void tm_st8(uint32_t taddr, const uint32_t *v)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tm_ld64(uint32_t taddr, uint32_t *v)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
        "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]),
          "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]),
          "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]),
          "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
        : "r"(taddr)
        : "memory");
}

// position (in order of addition) of the first term of frame pair r for a bin of residue p, default-window mask
template <int Q, int FOLD>
__device__ __forceinline__ constexpr int tm_slot_base(int p, int r)
{
    // terms per frame pair: 1 + 5, except r = Q/2 with k in {0,1,2,4}: 4
    int base = 0;
    if (FOLD == LWSB_FOLD_Q4 && (p & 1)) { // order 1, 3, 2
        if (r == 1) return 0;
        if (r == 3) return 6;
        return 12;
    }
    for (int q = 1; q < r; ++q) base += (2 * q == Q) ? 4 : 6;
    return base;
}
template <int Q>
__device__ __forceinline__ constexpr int tm_terms_per_bin()
{
    int n = 0;
    for (int q = 1; q < Q; ++q) n += (2 * q == Q) ? 4 : 6;
    return n; // 16 for Q = 4, 4 for Q = 2
}

// PRODUCER: term values of the four bins [4H, 4H+4) of the block for frame pair R_ (default mask), into TMEM
template <int Q, int FOLD, int H, int R_>
__device__ __forceinline__ void tm_produce_r(const unsigned char *ring, const unsigned *rowoff, int colbase, const StripW<Q> &w,
                                             uint32_t tbuf)
{
    // windows: ring columns colbase - 5 .. colbase + 8 of frames m - R_ and m + R_
    double2 wm[14], wp[14];
    const double2 *rm = reinterpret_cast<const double2 *>(ring + rowoff[Q - 1 - R_]) + (colbase - SL);
    const double2 *rp = reinterpret_cast<const double2 *>(ring + rowoff[Q - 1 + R_]) + (colbase - SL);
#pragma unroll
    for (int q = 0; q < 14; ++q) { wm[q] = rm[q]; wp[q] = rp[q]; }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        constexpr int dummy = 0; (void)dummy;
        const int p = (4 * H + b) % Q;                  // compile-time after unrolling
        const bool minus = FOLD == LWSB_FOLD_Q4 && (p & 1) && (R_ & 1);
        const int slot0 = tm_slot_base<Q, FOLD>(p, R_);
        double vr[6], vi[6];
        int n = 0;
        pair_value(w.wr[p][R_][0], w.wi[p][R_][0], wm[b + 5].x, wm[b + 5].y, wp[b + 5].x, wp[b + 5].y, vr[n], vi[n]);
        ++n;
#pragma unroll
        for (int k = 1; k <= SL; ++k) {
            if (!pat_has<Q, 1>(R_, k)) continue;
            const double2 e1 = wm[b + 5 - k], e2 = wp[b + 5 + k], e3 = wp[b + 5 - k], e4 = wm[b + 5 + k];
            double br, bi, cr, ci;
            if (minus) {
                br = __dsub_rn(e1.x, e2.x); bi = __dsub_rn(e1.y, e2.y);
                cr = __dsub_rn(e3.x, e4.x); ci = __dsub_rn(e3.y, e4.y);
            } else {
                br = __dadd_rn(e1.x, e2.x); bi = __dadd_rn(e1.y, e2.y);
                cr = __dadd_rn(e3.x, e4.x); ci = __dadd_rn(e3.y, e4.y);
            }
            pair_value(w.wr[p][R_][k], w.wi[p][R_][k], br, bi, cr, ci, vr[n], vi[n]);
            ++n;
        }
        // two terms (8 words) per store
        const uint32_t t0 = tbuf + (uint32_t)(64 * b + 4 * slot0);
#pragma unroll
        for (int q = 0; q + 1 < 6; q += 2) {
            if (q >= n) break;
            uint32_t x[8];
            x[0] = (uint32_t)__double2loint(vr[q]); x[1] = (uint32_t)__double2hiint(vr[q]);
            x[2] = (uint32_t)__double2loint(vi[q]); x[3] = (uint32_t)__double2hiint(vi[q]);
            x[4] = (uint32_t)__double2loint(vr[q + 1]); x[5] = (uint32_t)__double2hiint(vr[q + 1]);
            x[6] = (uint32_t)__double2loint(vi[q + 1]); x[7] = (uint32_t)__double2hiint(vi[q + 1]);
            tm_st8(t0 + 4 * q, x);
        }
    }
}

// Work split of a task between its two warps (Q = 4): the PRODUCER forms the terms of the frame pairs r = 1 and 3
// (12 of 16 terms), the CONSUMER those of r = 2 (4 terms) before it starts on the ordered sums -- roughly equal
// instruction counts, so both warps of a scheduler stay busy.  Q = 2 has a single frame pair: producer only.
template <int Q, int FOLD, int H, bool CONSUMER_SHARE>
__device__ __forceinline__ void tm_produce_half(const unsigned char *ring, const unsigned *rowoff, int xb, const StripW<Q> &w,
                                                uint32_t tlane)
{
    const int colbase = SL + SBK * xb + 4 * H;
    const uint32_t tbuf = tlane + 256u * H;
    if constexpr (Q == 4) {
        if constexpr (CONSUMER_SHARE) tm_produce_r<Q, FOLD, H, 2>(ring, rowoff, colbase, w, tbuf);
        else {
            tm_produce_r<Q, FOLD, H, 1>(ring, rowoff, colbase, w, tbuf);
            tm_produce_r<Q, FOLD, H, 3>(ring, rowoff, colbase, w, tbuf);
        }
    } else {
        if constexpr (!CONSUMER_SHARE) {
            if constexpr (Q > 1) tm_produce_r<Q, FOLD, H, 1>(ring, rowoff, colbase, w, tbuf);
            if constexpr (Q > 2) tm_produce_r<Q, FOLD, H, 2>(ring, rowoff, colbase, w, tbuf);
            if constexpr (Q > 3) tm_produce_r<Q, FOLD, H, 3>(ring, rowoff, colbase, w, tbuf);
        }
    }
}

// CONSUMER: the order-bound part of the four bins [4H, 4H+4): centre-frame term, the ordered sum, projection, commit
template <int Q, int FOLD, int H, int B_>
__device__ __forceinline__ void tm_consume_bins(const StripW<Q> &w, const BlockCtx &bc, const double *amp, unsigned active,
                                                uint32_t tbuf, double2 *newv, unsigned &committed)
{
    if constexpr (B_ < 4) {
        constexpr int I = 4 * H + B_;
        constexpr int P = I % Q;
        constexpr int NT = tm_terms_per_bin<Q>();
        uint32_t x[64];
        tm_ld64(tbuf + 64u * B_, x);
        double2 *own = reinterpret_cast<double2 *>(bc.ring + bc.ownoff);
        const int col = SL + SBK * bc.xb + I;
        // centre-frame term (default mask: k = 1 only); sees the bin updated just before (lwslib.cpp:169-182)
        const double2 b = own[col - 1], c = own[col + 1];
        double tr, ti;
        pair_value(w.wr[P][0][1], w.wi[P][0][1], b.x, b.y, c.x, c.y, tr, ti);
        tr = __dadd_rn(0.0, tr); ti = __dadd_rn(0.0, ti);
        tm_wait_ld();
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            tr = __dadd_rn(tr, __hiloint2double((int)x[4 * n + 1], (int)x[4 * n]));
            ti = __dadd_rn(ti, __hiloint2double((int)x[4 * n + 3], (int)x[4 * n + 2]));
        }
        double2 val;
        const bool ok = x_project(tr, ti, amp[I], val) && ((active >> I) & 1u);
        const int n = bc.n0 + I;
        int mcol = col;
        if (bc.first_strip && n >= 1 && n <= SL) mcol = SL - n;
        else if (n >= bc.Nreal - 1 - SL && n <= bc.Nreal - 2) mcol = SL + 2 * (bc.Nreal - 1) - n - bc.b0;
        if (ok) {
            own[col] = val;
            own[mcol] = make_double2(val.x, mcol == col ? val.y : -val.y);
            committed |= 1u << I;
        }
        newv[I] = val;
        tm_consume_bins<Q, FOLD, H, B_ + 1>(w, bc, amp, active, tbuf, newv, committed);
    }
}

// ---------------------------------------------------------------- the kernel
// TM = true: 8 warps, 0..3 consume and 4..7 produce (through tensor memory); needs the default mask
// PAIR > 0: two lanes per task (strip_pair.cuh); PAIR - 1 = window mode (0..2) + 3 * explicit pipelining (0/1); NREG: register cap
template <int Q, int FOLD, int PAT, bool TM, int PAIR, int NREG>
__global__ void __maxnreg__(NREG)
k_batch_strips(const __grid_constant__ StripParams prm, const __grid_constant__ StripW<Q> w)
{
    static_assert(!TM || (PAT == 1 && FOLD != LWSB_FOLD_ANY && Q <= 4), "the TMEM layout is laid out for the folded default-mask terms");
    static_assert(!(TM && PAIR), "one variant at a time");
    static_assert(!PAIR || Q <= 4, "the pair-split update is unrolled for Q <= 4");
    cg::cluster_group cluster = cg::this_cluster();
    const int C = prm.C;
    const int c = (int)cluster.block_rank();
    const int cid = blockIdx.x / C, ncl = gridDim.x / C;
    const int tid = threadIdx.x;
    // control duties (neighbour hand-shake, TMA traffic): a warp of its own, or -- TM kernels, which need two
    // warps per scheduler and all their registers -- lane 0 of the last producer warp
    const int nct = TM ? (int)blockDim.x : (int)blockDim.x - 32; // task threads
    const bool is_ctrl = !TM && tid >= nct;
    const bool ctl = TM ? (tid == (int)blockDim.x - 32) : (is_ctrl && (tid & 31) == 0); // the one thread doing them
    const int lane = tid & 31;
    const LwsbView &v = prm.v;
    const int NBr = prm.NBr, NBV = prm.NBV, NS = prm.NS, G = prm.G, R = prm.R, pitch = prm.pitch;
    const int QS = prm.QS; // frames between consecutive sweeps (>= Q; odd so that a warp's rows spread over all banks)
    const int Nreal = v.Nreal, P = v.P;
    const unsigned rowbytes = (unsigned)pitch * 16u;

    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *ring = smem;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + (size_t)R * rowbytes);
    unsigned *flags = reinterpret_cast<unsigned *>(mbar + R); // [0]: progress of the left neighbour, [1]: of the right one
    int *nact = reinterpret_cast<int *>(flags + 4);
    int *act = nact + 4;                                      // indices of the sweeps that can move a bin

    unsigned char *ring_left = c > 0 ? reinterpret_cast<unsigned char *>(cluster.map_shared_rank(ring, c - 1)) : nullptr;
    unsigned char *ring_right = c < C - 1 ? reinterpret_cast<unsigned char *>(cluster.map_shared_rank(ring, c + 1)) : nullptr;
    unsigned *flag_at_left = c > 0 ? cluster.map_shared_rank(flags, c - 1) + 1 : nullptr;       // I am its right neighbour
    unsigned *flag_at_right = c < C - 1 ? cluster.map_shared_rank(flags, c + 1) + 0 : nullptr;  // I am its left neighbour

    const int b0 = c * NBr * SBK;                                     // first bin of the strip
    int nb_my = (Nreal - b0 + SBK - 1) / SBK;                         // blocks holding real bins
    nb_my = nb_my < 0 ? 0 : (nb_my > NBr ? NBr : nb_my);
    const int gcol0 = v.c0 + b0 - SL;                                 // global column of ring column 0
    const unsigned load_bytes = (unsigned)(SBK * NBr + 2 * SL) * 16u; // a full ring row
    const int wb_lo = c == 0 ? 0 : SL;                                // ring columns written back (mirrors included at the ends)
    const int wb_hi = c == C - 1 ? SL + (Nreal - b0) + SL : SL + SBK * NBr;

    // per-thread slot: frame residue j, sweep slot g
    // thread order: sweep slot fastest (lanes of a quarter-warp sit QS rows apart: conflict free for odd QS) or
    // frame slot fastest (lanes on consecutive frames); the planner picks the one with fewer bank conflicts
    // task index: TM: producer lane l of warp w+4 serves consumer lane l of warp w; PAIR: lanes 2p, 2p+1 share task p
    const int tix = TM ? (tid & 127) : (PAIR ? (tid >> 1) : tid);
    const bool is_producer = TM && tid >= 128 && !is_ctrl;
    const int j = prm.GFAST ? tix / G : tix % NS, g = prm.GFAST ? tix % G : tix / NS;
    const bool has_slot = !is_ctrl && j < NS && g < G;

    // tensor memory: all 512 columns of this SM (one CTA per SM), base address through shared memory
    __shared__ uint32_t tm_base_smem;
    uint32_t tlane = 0;
    if constexpr (TM) {
        if (tid < 32) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tm_base_smem)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        tm_fence_before();
        __syncthreads();
        tm_fence_after();
        tlane = tm_base_smem + ((uint32_t)((tid >> 5) & 3) << 21); // lane field (bits 31:16) = 32 * quarter
    }

    bool mbar_live = false;
    long long tm_publish = 0, tm_poll = 0, tm_house = 0, tm_work = 0, tm_waitA = 0, tm_waitB = 0; // cycle counters (status[2..])
    long long tm_rows = 0;                                                                        // control lane: waiting for another pass's rows
    long long ph[6] = {0, 0, 0, 0, 0, 0}; // consumer phases of the TM kernels: set-up, own terms, wait A, chain A, wait B, chain B
    for (int item = cid; item < prm.n_items; item += ncl) {
        const int u = prm.items[item].x, pass = prm.items[item].y;
        const bool tracer = prm.trace != nullptr && c == 0 && tid == 0;
        if (tracer) prm.trace[8 * item + 0] = global_ns();
        const long long it_rows0 = tm_rows, it_poll0 = tm_poll, it_work0 = tm_work, it_wait0 = tm_waitB;
        const int T = v.T[u];
        const int Tp = T + 2 * (Q - 1);
        const long long grow0 = v.rowbase[u];
        const double mean = v.mean_amp[u];
        __syncthreads();
        if (tid == 0) {
            const double mx = prm.max_amp[u];
            int n = 0;
            for (int i = 0; i < prm.iters; ++i)
                if (__dmul_rn(prm.thr[i], mean) < mx) act[n++] = i; // a sweep with threshold >= max|S| moves nothing
            *nact = n;
        }
        __syncthreads();
        const int n_act = *nact;
        {
            const int Gp = min(G, n_act - pass * G);
            const int nsteps = 2 * (T - 1 + QS * (Gp - 1)) + NBV;
            const double thr = (has_slot && g < Gp) ? __dmul_rn(prm.thr[act[pass * G + g]], mean) : 0.0; // lws.pyx:245

            // Row e may be loaded once the previous pass of this utterance -- on this or on another cluster -- has
            // written it back in the three strips the load spans (ghost frames are never rewritten).
            auto wait_rows = [&](int e) {
                const int mf = e - (Q - 1);
                if (pass == 0 || mf < 0 || mf >= T) return;
                const long long w0 = clock64();
                const unsigned need = (unsigned)(mf + 1);
                const unsigned *prev = prm.done + ((size_t)u * prm.max_pass + (pass - 1)) * 8;
                for (int cc = max(c - 1, 0); cc <= min(c + 1, C - 1); ++cc) {
                    unsigned spins = 0;
                    while (ld_acquire_gpu(prev + cc) < need) {
                        __nanosleep(64);
                        if ((++spins & 1023u) == 0) {
                            if (*reinterpret_cast<volatile unsigned *>(prm.status) != 0u) break;
                            if (spins > PASS_SPIN_LIMIT) { atomicCAS(prm.status, 0u, 0x40000000u | (c << 24) | ((pass & 0xff) << 16) | (e & 0xffff)); break; }
                        }
                    }
                }
                fence_proxy_async();
                tm_rows += clock64() - w0;
            };
            // ---- pass prologue: this cluster's previous work item fully written back, ring (re)initialised
            if (ctl) { tma_store_wait_all(); fence_proxy_async(); __threadfence(); }
            cluster.sync();
            if (ctl) {
                if (!mbar_live) {
                    for (int s = 0; s < R; ++s) mbar_init(&mbar[s], 1);
                    fence_mbar_init();
                    mbar_live = true;
                }
                flags[0] = 0; flags[1] = 0;
                fence_proxy_async();
                const int npre = min(Tp, 2 * (Q - 1) + SLEAD + 1);
                for (int e = 0; e < npre; ++e) {
                    wait_rows(e);
                    mbar_expect_tx(&mbar[e % R], load_bytes);
                    tma_load_row(ring + (size_t)(e % R) * rowbytes, v.E + (grow0 + e) * P + gcol0, load_bytes, &mbar[e % R]);
                }
            }
            cluster.sync();
            if (tracer) prm.trace[8 * item + 1] = global_ns();

            // ---- control actions (executed by the thread `ctl` only)
            auto poll = [&](int t) { // conditions for macro-step t (DESIGN.md "strip hand-shake")
                if (c > 0) {
                    const unsigned need = (unsigned)min(t + NBr, nsteps);
                    unsigned spins = 0;
                    while (ld_acquire_cluster(&flags[0]) < need) {
                        __nanosleep(32);
                        if (!keep_waiting(spins, prm.status, 0x10000000u | (c << 24) | ((pass & 0xff) << 16) | (t & 0xffff))) break;
                    }
                }
                if (c < C - 1 && t - NBr > 0) {
                    const unsigned need = (unsigned)(t - NBr);
                    unsigned spins = 0;
                    while (ld_acquire_cluster(&flags[1]) < need) {
                        __nanosleep(32);
                        if (!keep_waiting(spins, prm.status, 0x20000000u | (c << 24) | ((pass & 0xff) << 16) | (t & 0xffff))) break;
                    }
                }
            };
            // publish first: the neighbours' next macro-step waits for this one.  (Polling for step t+1 before
            // publishing step t would dead-lock: in lock step the neighbour's matching step finishes only after
            // it has seen this strip's step t.)
            auto publish = [&](int done) {
                if (flag_at_left) st_release_cluster(flag_at_left, (unsigned)done);
                if (flag_at_right) st_release_cluster(flag_at_right, (unsigned)done);
            };
            // TMA traffic after macro-step t, overlapped with macro-step t + 1: nothing here touches a row in use
            auto housekeeping = [&](int t) {
                if (((t + 1) & 1) == 0) { // next frame into the slot freed longest ago
                    const int e = (t + 1) / 2 + 2 * (Q - 1) + SLEAD;
                    if (e < Tp) {
                        tma_store_wait_read(); // bulk stores issued a macro-step or more ago: long finished reading
                        wait_rows(e);
                        fence_proxy_async();
                        mbar_expect_tx(&mbar[e % R], load_bytes);
                        tma_load_row(ring + (size_t)(e % R) * rowbytes, v.E + (grow0 + e) * P + gcol0, load_bytes, &mbar[e % R]);
                    }
                }
                const int tf = t - (nb_my - 1); // frame whose last sweep of this pass finished its last real block in step t
                if (nb_my > 0 && tf >= 0 && (tf & 1) == 0) {
                    const int m = tf / 2 - QS * (Gp - 1);
                    if (m >= 0 && m < T) {
                        const int e = m + Q - 1;
                        fence_proxy_async();
                        tma_store_row(v.E + (grow0 + e) * P + gcol0 + wb_lo,
                                      ring + (size_t)(e % R) * rowbytes + (size_t)wb_lo * 16u, (unsigned)(wb_hi - wb_lo) * 16u);
                        // frames 0 .. m-1 of this pass are in global memory: tell the next pass of this utterance
                        tma_store_wait_all_but_one();
                        if (m >= 1) { fence_proxy_async(); __threadfence(); st_release_gpu(prm.done + ((size_t)u * prm.max_pass + pass) * 8 + c, (unsigned)m); }
                    }
                }
            };
            // Slot s received Tp/R (+1) rows in this pass, one mbarrier phase each.  The waits address phases by
            // parity counted from the start of the pass, so slots that saw an odd number of phases get one empty
            // phase: every barrier starts the next pass at parity 0 again.
            auto fix_parity = [&]() {
                for (int s = 0; s < R; ++s)
                    if ((Tp / R + (s < Tp % R ? 1 : 0)) & 1) mbar_arrive(&mbar[s]);
            };

            if (is_ctrl) {
                // ================= control warp (kernels without tensor memory) =================
                if (ctl) poll(0);
                __syncwarp();
                cta_sync(); // releases the compute warps into macro-step 0
                for (int t = 0; t < nsteps; ++t) {
                    cta_sync(); // macro-step t computed by every thread of the strip
                    const long long c0 = clock64();
                    if (ctl) publish(t + 1);
                    const long long c1 = clock64();
                    if (ctl && t + 1 < nsteps) poll(t + 1);
                    __syncwarp();
                    const long long c2 = clock64();
                    cta_sync(); // releases the compute warps into macro-step t + 1
                    if (ctl) housekeeping(t);
                    __syncwarp();
                    tm_publish += c1 - c0; tm_poll += c2 - c1; tm_house += clock64() - c2;
                }
                if (ctl) fix_parity();
            } else {
                // ================= task warps =================
                if constexpr (TM) {
                    int xb = -2 * j;
                    int m = j - QS * g;
                    if (ctl) poll(0);
                    __syncwarp();
                    cta_sync(); // macro-step 0 verified
                    for (int t = 0; t < nsteps; ++t) {
                        const long long k0 = clock64();
                        if ((t & 1) == 0) {
                            const int k = t >> 1;
                            const int e_lo = k == 0 ? 0 : k + 2 * (Q - 1), e_hi = k + 2 * (Q - 1);
                            for (int e = e_lo; e <= e_hi && e < Tp; ++e) {
                                unsigned spins = 0;
                                while (!mbar_try_wait(&mbar[e % R], (unsigned)((e / R) & 1)))
                                    if (!keep_waiting(spins, prm.status, 0x30000000u | (c << 24) | ((pass & 0xff) << 16) | (t & 0xffff))) break;
                            }
                        }
                        const bool valid = has_slot && g < Gp && xb >= 0 && xb < nb_my && m >= 0 && m < T;
                        const int e = valid ? m + Q - 1 : Q - 1;
                        const int xbv = valid ? xb : 0;
                        const int n0 = b0 + SBK * xbv;
                        double amp[SBK];
                        unsigned active = 0;
                        if (valid) {
                            const double2 *ap = reinterpret_cast<const double2 *>(v.A + (grow0 + e) * P + v.c0 + n0);
#pragma unroll
                            for (int q = 0; q < SBK / 2; ++q) {
                                const double2 a2 = __ldg(ap + q);
                                amp[2 * q] = a2.x; amp[2 * q + 1] = a2.y;
                            }
#pragma unroll
                            for (int i = 0; i < SBK; ++i)
                                if (n0 + i < Nreal && amp[i] > thr) active |= 1u << i;
                        } else {
#pragma unroll
                            for (int i = 0; i < SBK; ++i) amp[i] = 0.0;
                        }
                        // the tcgen05 operations are warp collectives: the producer / consumer pair takes the same
                        // decision (same task indices, same amplitudes), lanes without work run on harmless data
                        const long long q0 = clock64();
                        ph[0] += q0 - k0;
                        if (__any_sync(0xffffffffu, active != 0)) {
                            unsigned rowoff[2 * Q - 1];
                            const int es = e % R;
#pragma unroll
                            for (int d = 0; d < 2 * Q - 1; ++d) {
                                int sl = es + d - (Q - 1);
                                sl = sl < 0 ? sl + R : (sl >= R ? sl - R : sl);
                                rowoff[d] = (unsigned)sl * rowbytes;
                            }
                            const int bar0 = 2 + 2 * ((tid >> 5) & 3);
                            if (is_producer) {
                                tm_produce_half<Q, FOLD, 0, false>(ring, rowoff, xbv, w, tlane);
                                tm_wait_st(); tm_fence_before(); pair_arrive(bar0);
                                tm_produce_half<Q, FOLD, 1, false>(ring, rowoff, xbv, w, tlane);
                                tm_wait_st(); tm_fence_before(); pair_arrive(bar0 + 1);
                            } else {
                                BlockCtx bc;
                                bc.ring = ring; bc.ring_left = ring_left; bc.ring_right = ring_right;
                                bc.ownoff = rowoff[Q - 1]; bc.xb = xbv; bc.n0 = n0; bc.b0 = b0;
                                bc.Nreal = Nreal; bc.NBr = NBr; bc.first_strip = (c == 0);
                                double2 newv[SBK];
                                unsigned committed = 0;
                                // the consumer's share of the term values, then the order-bound part
                                tm_produce_half<Q, FOLD, 0, true>(ring, rowoff, xbv, w, tlane);
                                tm_produce_half<Q, FOLD, 1, true>(ring, rowoff, xbv, w, tlane);
                                tm_wait_st();
                                const long long q1 = clock64();
                                pair_sync(bar0); tm_fence_after();
                                const long long q2 = clock64();
                                tm_consume_bins<Q, FOLD, 0, 0>(w, bc, amp, active, tlane, newv, committed);
                                const long long q3 = clock64();
                                pair_sync(bar0 + 1); tm_fence_after();
                                const long long q4 = clock64();
                                tm_consume_bins<Q, FOLD, 1, 0>(w, bc, amp, active, tlane + 256u, newv, committed);
                                const long long q5 = clock64();
                                ph[1] += q1 - q0; ph[2] += q2 - q1; ph[3] += q3 - q2; ph[4] += q4 - q3; ph[5] += q5 - q4;
                                if (bc.xb == 0 && ring_left) {
                                    double2 *dst = reinterpret_cast<double2 *>(ring_left + bc.ownoff) + SL + SBK * NBr;
#pragma unroll
                                    for (int i = 0; i < SL; ++i)
                                        if ((committed >> i) & 1u) dst[i] = newv[i];
                                }
                                if (bc.xb == NBr - 1 && ring_right) {
                                    double2 *dst = reinterpret_cast<double2 *>(ring_right + bc.ownoff);
#pragma unroll
                                    for (int i = SBK - SL; i < SBK; ++i)
                                        if ((committed >> i) & 1u) dst[i - (SBK - SL)] = newv[i];
                                }
                            }
                        }
                        const long long k1 = clock64();
                        cta_sync(); // macro-step t done
                        if (++xb == NBV) { xb = 0; m += NS; }
                        const long long k2 = clock64();
                        if (ctl) { publish(t + 1); if (t + 1 < nsteps) poll(t + 1); }
                        __syncwarp();
                        cta_sync(); // neighbours ready for macro-step t + 1
                        const long long k3 = clock64();
                        if (ctl) housekeeping(t);
                        __syncwarp();
                        if (is_producer) { tm_publish += k1 - k0; tm_poll += k2 - k1; tm_house += k3 - k2; }
                        else { tm_work += k1 - k0; tm_waitA += k2 - k1; tm_waitB += k3 - k2; }
                    }
                    if (ctl) fix_parity();
                } else if constexpr (PAIR != 0) {
                    // ---- two lanes per task: every lane of a warp runs the block update when any task of the
                    // warp has work (the lane pairs exchange values by shuffles); lanes without work compute on
                    // harmless cells and commit nothing
                    const int h = tid & 1;
                    int xb = -2 * j;
                    int m = j - QS * g;
                    // amplitudes of the block (row-major plane in global memory, read-only), fetched one macro-step ahead
                    double ampn[SBK];
                    auto fetch_amp = [&](int xb_, int m_) {
                        const bool valid_ = has_slot && g < Gp && xb_ >= 0 && xb_ < nb_my && m_ >= 0 && m_ < T;
                        if (valid_) {
                            const double2 *ap = reinterpret_cast<const double2 *>(v.A + (grow0 + m_ + Q - 1) * P + v.c0 + b0 + SBK * xb_);
#pragma unroll
                            for (int q = 0; q < SBK / 2; ++q) {
                                const double2 a2 = __ldg(ap + q);
                                ampn[2 * q] = a2.x; ampn[2 * q + 1] = a2.y;
                            }
                        }
                    };
                    fetch_amp(xb, m);
                    cta_sync();
                    for (int t = 0; t < nsteps; ++t) {
                        const long long k0 = clock64();
                        if ((t & 1) == 0) {
                            const int k = t >> 1;
                            const int e_lo = k == 0 ? 0 : k + 2 * (Q - 1), e_hi = k + 2 * (Q - 1);
                            for (int e = e_lo; e <= e_hi && e < Tp; ++e) {
                                unsigned spins = 0;
                                while (!mbar_try_wait(&mbar[e % R], (unsigned)((e / R) & 1)))
                                    if (!keep_waiting(spins, prm.status, 0x30000000u | (c << 24) | ((pass & 0xff) << 16) | (t & 0xffff))) break;
                            }
                        }
                        const bool valid = has_slot && g < Gp && xb >= 0 && xb < nb_my && m >= 0 && m < T;
                        const int e = valid ? m + Q - 1 : Q - 1;
                        const int xbv = valid ? xb : 0;
                        const int n0 = b0 + SBK * xbv;
                        double amp[SBK];
                        unsigned active = 0;
#pragma unroll
                        for (int i = 0; i < SBK; ++i) {
                            amp[i] = valid ? ampn[i] : 0.0;
                            if (valid && n0 + i < Nreal && amp[i] > thr) active |= 1u << i; // lwslib.cpp:295-296
                        }
                        // next macro-step's task of this slot
                        int xb1 = xb + 1, m1 = m;
                        if (xb1 == NBV) { xb1 = 0; m1 += NS; }
                        fetch_amp(xb1, m1);
                        if (__any_sync(0xffffffffu, active != 0)) {
                            PairCell<Q> cell;
                            cell.base = ring + 8 * h;
                            const int es = e % R;
#pragma unroll
                            for (int d = 0; d < 2 * Q - 1; ++d) {
                                int sl = es + d - (Q - 1);
                                sl = sl < 0 ? sl + R : (sl >= R ? sl - R : sl);
                                cell.rowoff[d] = (unsigned)sl * rowbytes;
                            }
                            cell.col0 = SL + SBK * xbv;
                            BlockCtx bc;
                            bc.ring = ring; bc.ring_left = ring_left; bc.ring_right = ring_right;
                            bc.ownoff = cell.rowoff[Q - 1]; bc.xb = xbv; bc.n0 = n0; bc.b0 = b0;
                            bc.Nreal = Nreal; bc.NBr = NBr; bc.first_strip = (c == 0);
                            pair_update_block<Q, FOLD, PAT, (PAIR - 1) % 3, (PAIR - 1) / 3>(cell, w, bc, amp, active, h);
                        }
                        const long long k1 = clock64();
                        cta_sync(); // macro-step t done
                        xb = xb1; m = m1;
                        const long long k2 = clock64();
                        cta_sync(); // neighbours ready for macro-step t + 1
                        tm_work += k1 - k0; tm_waitA += k2 - k1; tm_waitB += clock64() - k2;
                    }
                } else {
                int xb = -2 * j;          // block index; negative while the slot has not started
                int m = j - QS * g;       // frame of the slot
                // amplitudes of the block (row-major plane in global memory, read-only), fetched one macro-step ahead:
                // their L2 / HBM latency would otherwise sit in front of every block
                double ampn[SBK];
                auto fetch_amp = [&](int xb_, int m_) {
                    if (has_slot && g < Gp && xb_ >= 0 && xb_ < nb_my && m_ >= 0 && m_ < T) {
                        const double2 *ap = reinterpret_cast<const double2 *>(v.A + (grow0 + m_ + Q - 1) * P + v.c0 + b0 + SBK * xb_);
#pragma unroll
                        for (int q = 0; q < SBK / 2; ++q) {
                            const double2 a2 = __ldg(ap + q);
                            ampn[2 * q] = a2.x; ampn[2 * q + 1] = a2.y;
                        }
                    }
                };
                fetch_amp(xb, m);
                cta_sync();               // control warp has verified macro-step 0
                for (int t = 0; t < nsteps; ++t) {
                    const long long k0 = clock64();
                    if ((t & 1) == 0) {
                        // rows entering use at this frame clock must have landed
                        const int k = t >> 1;
                        const int e_lo = k == 0 ? 0 : k + 2 * (Q - 1), e_hi = k + 2 * (Q - 1);
                        for (int e = e_lo; e <= e_hi && e < Tp; ++e) {
                            unsigned spins = 0;
                            while (!mbar_try_wait(&mbar[e % R], (unsigned)((e / R) & 1)))
                                if (!keep_waiting(spins, prm.status, 0x30000000u | (c << 24) | ((pass & 0xff) << 16) | (t & 0xffff))) break;
                        }
                    }
                    int xb1 = xb + 1, m1 = m; // next macro-step's task of this slot
                    if (xb1 == NBV) { xb1 = 0; m1 += NS; }
                    const bool valid = has_slot && g < Gp && xb >= 0 && xb < nb_my && m >= 0 && m < T;
                    double amp[SBK];
#pragma unroll
                    for (int i = 0; i < SBK; ++i) amp[i] = ampn[i];
                    fetch_amp(xb1, m1);
                    if (valid) {
                        const int e = m + Q - 1;
                        const int n0 = b0 + SBK * xb;
                        unsigned active = 0;
#pragma unroll
                        for (int i = 0; i < SBK; ++i)
                            if (n0 + i < Nreal && amp[i] > thr) active |= 1u << i; // lwslib.cpp:295-296
                        if (active) {
                            RingCell<Q> cell;
                            cell.ring = ring;
                            const int es = e % R;
#pragma unroll
                            for (int d = 0; d < 2 * Q - 1; ++d) {
                                int s = es + d - (Q - 1);
                                s = s < 0 ? s + R : (s >= R ? s - R : s);
                                cell.rowoff[d] = (unsigned)s * rowbytes;
                            }
                            BlockCtx bc;
                            bc.ring = ring; bc.ring_left = ring_left; bc.ring_right = ring_right;
                            bc.ownoff = cell.rowoff[Q - 1]; bc.xb = xb; bc.n0 = n0; bc.b0 = b0;
                            bc.Nreal = Nreal; bc.NBr = NBr; bc.first_strip = (c == 0);
                            if constexpr (Q <= 4) strip_update_block_pipelined<Q, FOLD, PAT>(cell, w, bc, amp, active);
                            else strip_update_block<Q, FOLD, 0>(cell, w, bc, amp, active);
                        }
                    }
                    const long long k1 = clock64();
                    cta_sync(); // macro-step t done
                    xb = xb1; m = m1;
                    const long long k2 = clock64();
                    cta_sync(); // neighbours ready for macro-step t + 1
                    tm_work += k1 - k0; tm_waitA += k2 - k1; tm_waitB += clock64() - k2;
                }
                }
            }
        }
        // ---- item epilogue: everything written back before the ring is reused; the pass is complete
        if (tracer) prm.trace[8 * item + 2] = global_ns();
        if (ctl) { tma_store_wait_all(); fence_proxy_async(); __threadfence(); st_release_gpu(prm.done + ((size_t)u * prm.max_pass + pass) * 8 + c, (unsigned)T); }
        if (prm.trace != nullptr && c == 0 && ctl) {
            prm.trace[8 * item + 3] = global_ns();
            prm.trace[8 * item + 4] = (unsigned long long)(tm_rows - it_rows0);
            prm.trace[8 * item + 5] = (unsigned long long)(tm_poll - it_poll0);
        }
        if (tracer) {
            prm.trace[8 * item + 6] = (unsigned long long)(tm_work - it_work0);
            prm.trace[8 * item + 7] = (unsigned long long)(tm_waitB - it_wait0);
        }
    }
    // cycle accounting of cluster 0 (introspection: lwsb_last_batch_cycles): control lane and one lane per compute warp
    if (cid == 0 && lane == 0) {
        unsigned long long *acc = reinterpret_cast<unsigned long long *>(prm.status + 2);
        if (is_ctrl || is_producer) { // control warp, or (TM kernels) producer warps: work / wait strip / wait neighbours
            atomicAdd(acc + 0, (unsigned long long)tm_publish); atomicAdd(acc + 1, (unsigned long long)tm_poll);
            atomicAdd(acc + 2, (unsigned long long)tm_house);
        } else {
            atomicAdd(acc + 3, (unsigned long long)tm_work); atomicAdd(acc + 4, (unsigned long long)tm_waitA);
            atomicAdd(acc + 5, (unsigned long long)tm_waitB); atomicAdd(acc + 6, 1ull);
            for (int q = 0; q < 6; ++q) atomicAdd(acc + 7 + q, (unsigned long long)ph[q]);
        }
    }
    cluster.sync(); // no CTA leaves while a neighbour may still address its shared memory
    if constexpr (TM) {
        if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm_base_smem) : "memory");
    }
}

// ---------------------------------------------------------------- self-check of the branch-free sqrt / division
// Inputs: a 64-bit mix of the index (every exponent from 2^-1022 to 2^1023 and signs on the numerator); counts
// the samples inside the fast ranges and those among them whose bits differ from __dsqrt_rn / __ddiv_rn.
__global__ void k_debug_fast_math(long long n, unsigned long long seed, unsigned long long *out)
{
    unsigned long long chk_s = 0, bad_s = 0, chk_d = 0, bad_d = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        auto mix = [](unsigned long long z) {
            z += 0x9e3779b97f4a7c15ull; z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
            z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31);
        };
        const unsigned long long a = mix(seed + 2 * (unsigned long long)i), b = mix(seed + 2 * (unsigned long long)i + 1);
        // half the samples: exponents near 1 (the values the sweeps see), half: any exponent
        const bool wide = (i & 1) != 0;
        auto make = [&](unsigned long long bits, bool neg_ok) {
            unsigned long long e = (bits >> 52) & 0x7ff;
            if (!wide) e = 1023 - 40 + e % 80;
            if (e == 0x7ff) e = 0x7fe;
            unsigned long long v = (bits & 0x000fffffffffffffull) | (e << 52);
            if (neg_ok && (bits >> 63)) v |= 1ull << 63;
            return __longlong_as_double((long long)v);
        };
        const double x = make(a, false), num = (i % 97 == 0) ? 0.0 : make(b, true);
        bool sok, rok, dok;
        const double s = fm_sqrt(x, sok);
        if (sok) { ++chk_s; if (__double_as_longlong(s) != __double_as_longlong(__dsqrt_rn(x))) ++bad_s; }
        const double q = fm_div(num, x, fm_rcp(x, rok), dok);
        if (rok && dok) { ++chk_d; if (__double_as_longlong(q) != __double_as_longlong(__ddiv_rn(num, x))) ++bad_d; }
    }
    atomicAdd(out + 0, chk_s); atomicAdd(out + 1, bad_s); atomicAdd(out + 2, chk_d); atomicAdd(out + 3, bad_d);
}

} // namespace

// ---------------------------------------------------------------- host side
namespace {

template <int Q, int FOLD, int PAT, bool TM, int PAIR = 0, int NREG = 255>
cudaError_t launch_strips_t(const StripParams &prm, const StripW<Q> &w, const StripPlan &pl, int B, cudaStream_t s)
{
    auto kern = k_batch_strips<Q, FOLD, PAT, TM, PAIR, NREG>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem_bytes);
    if (e != cudaSuccess) return e;
    if (pl.C > 8) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
    }
    // one CTA per SM: small plans (short rings) would otherwise share an SM while other SMs idle, and the passes of
    // one utterance -- a dependency chain across clusters -- would slow each other down
    int dev = 0, smem_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    const int smem_launch = std::max(pl.smem_bytes, std::min(smem_sm / 2 + 1024, pl.smem_limit));
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_launch);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(TM ? 256 : pl.nthreads);
    cfg.dynamicSmemBytes = smem_launch;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = pl.C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.gridDim = dim3(pl.C); // placeholder for the occupancy query
    int ncl = 0;
    e = cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg);
    if (e != cudaSuccess) return e;
    if (ncl < 1) return cudaErrorLaunchOutOfResources;
    cfg.gridDim = dim3((unsigned)(std::min(ncl, prm.n_items) * pl.C)); // all clusters resident: a pass may wait for another cluster's
    return cudaLaunchKernelEx(&cfg, kern, prm, w);
}

// pair-split kernels.  The register file is 16 K registers per SM sub-partition and warps are dealt to the four
// sub-partitions in turn: up to 8 warps keep 255 registers, 9-12 warps 168, 13-16 warps 128.  Measured on B200 the
// spills of the smaller caps cost more than the extra warps bring, so the planner stays at
// 7 task warps + the control warp and only that tier is built (all three with -DLWSB_PAIR_EXPERIMENTS).
template <int Q, int FOLD, int PAT>
cudaError_t launch_pair(const StripParams &prm, const StripW<Q> &w, const StripPlan &pl, int B, cudaStream_t s)
{
    const int mode = pl.TM - LWSB_VARIANT_PAIR; // window mode + 3 * explicit pipelining
    const int nt = pl.nthreads;
#ifdef LWSB_PAIR_EXPERIMENTS
    if (nt > PAIR_THREADS_MAX) return cudaErrorInvalidValue;
#define LWSB_PAIR_CASE(M_)                                                                                   \
    case M_:                                                                                                 \
        if (nt <= 256) return launch_strips_t<Q, FOLD, PAT, false, M_ + 1, 255>(prm, w, pl, B, s);           \
        if (nt <= 384) return launch_strips_t<Q, FOLD, PAT, false, M_ + 1, 168>(prm, w, pl, B, s);           \
        return launch_strips_t<Q, FOLD, PAT, false, M_ + 1, 128>(prm, w, pl, B, s);
#else
    if (nt > PAIR_THREADS_PLAN) return cudaErrorInvalidValue;
#define LWSB_PAIR_CASE(M_)                                                                                   \
    case M_: return launch_strips_t<Q, FOLD, PAT, false, M_ + 1, 255>(prm, w, pl, B, s);
#endif
    switch (mode) {
        LWSB_PAIR_CASE(LWSB_PAIR_DEFAULT_MODE)
#ifdef LWSB_PAIR_EXPERIMENTS
    default:
        if constexpr (Q == 4 && PAT == 1) {
            switch (mode) {
                LWSB_PAIR_CASE((LWSB_PAIR_DEFAULT_MODE + 1) % 6)
                LWSB_PAIR_CASE((LWSB_PAIR_DEFAULT_MODE + 2) % 6)
                LWSB_PAIR_CASE((LWSB_PAIR_DEFAULT_MODE + 3) % 6)
                LWSB_PAIR_CASE((LWSB_PAIR_DEFAULT_MODE + 4) % 6)
                LWSB_PAIR_CASE((LWSB_PAIR_DEFAULT_MODE + 5) % 6)
            }
        }
#endif
    }
#undef LWSB_PAIR_CASE
    if (mode != LWSB_PAIR_DEFAULT_MODE) { // a mode this build (or this Q / mask) has no kernel for: the default one
        StripPlan pd = pl;
        pd.TM = LWSB_VARIANT_PAIR + LWSB_PAIR_DEFAULT_MODE;
        return launch_pair<Q, FOLD, PAT>(prm, w, pd, B, s);
    }
    return cudaErrorInvalidValue;
}

template <int Q>
cudaError_t launch_strips_q(const StripParams &prm, const double *wr, const double *wi, int fold, const StripPlan &pl,
                            int B, cudaStream_t s)
{
    StripW<Q> w;
    for (int p = 0; p < Q; ++p)
        for (int r = 0; r < Q; ++r) {
            unsigned f = 0;
            for (int k = 0; k <= SL; ++k) {
                const size_t i = ((size_t)p * Q + r) * (SL + 1) + k;
                w.wr[p][r][k] = wr[i]; w.wi[p][r][k] = wi[i];
                if (std::hypot(wr[i], wi[i]) > 1.0e-12) f |= 1u << k; // lws.pyx:231-232
            }
            w.flag[p][r] = f;
        }
    w.fold = fold;
    // does the mask equal the default-window pattern the PAT = 1 kernels have compiled in?
    bool def = Q <= 4;
    for (int p = 0; p < Q && def; ++p)
        for (int r = 0; r < Q && def; ++r)
            for (int k = (r == 0 ? 1 : 0); k <= SL; ++k)
                if (pat_has<Q, 1>(r, k) != (((w.flag[p][r] >> k) & 1u) != 0)) { def = false; break; }
    if constexpr (Q <= 4) {
        if (fold == LWSB_FOLD_ANY) return launch_strips_t<Q, LWSB_FOLD_ANY, 0, false>(prm, w, pl, B, s);
        const bool tm = def && pl.TM == LWSB_VARIANT_TM && pl.NS * pl.G <= 128;
        const bool pair = pl.TM >= LWSB_VARIANT_PAIR;
        if constexpr (Q == 4) {
            if (fold == LWSB_FOLD_Q4) {
                if (pair) return def ? launch_pair<4, LWSB_FOLD_Q4, 1>(prm, w, pl, B, s) : launch_pair<4, LWSB_FOLD_Q4, 0>(prm, w, pl, B, s);
                if (tm) return launch_strips_t<4, LWSB_FOLD_Q4, 1, true>(prm, w, pl, B, s);
                return def ? launch_strips_t<4, LWSB_FOLD_Q4, 1, false>(prm, w, pl, B, s)
                           : launch_strips_t<4, LWSB_FOLD_Q4, 0, false>(prm, w, pl, B, s);
            }
        }
        if constexpr (Q == 2) {
            if (fold == LWSB_FOLD_Q2) {
                if (pair) return def ? launch_pair<2, LWSB_FOLD_Q2, 1>(prm, w, pl, B, s) : launch_pair<2, LWSB_FOLD_Q2, 0>(prm, w, pl, B, s);
                if (tm) return launch_strips_t<2, LWSB_FOLD_Q2, 1, true>(prm, w, pl, B, s);
                return def ? launch_strips_t<2, LWSB_FOLD_Q2, 1, false>(prm, w, pl, B, s)
                           : launch_strips_t<2, LWSB_FOLD_Q2, 0, false>(prm, w, pl, B, s);
            }
        }
    } else {
        if (fold == LWSB_FOLD_ANY) return launch_strips_t<Q, LWSB_FOLD_ANY, 0, false>(prm, w, pl, B, s);
    }
    return cudaErrorInvalidValue;
}

} // namespace

// Chooses cluster size, strip width and sweeps per pass.  Returns false when the shape is not
// served by this kernel (the generic kernel takes over).
bool plan_strips(int Nreal, int Q, int L, int iters, int maxT, int B, size_t smem_limit, int sm_count, StripPlan *out,
                 int force_cluster, int max_sweeps, int force_lag, int variant, int fold)
{
    if (L != SL || !(Q == 2 || Q == 4 || Q == 8) || iters < 1) return false;
    // variant: the pair-split kernel serves the folded Q = 2 / Q = 4 updates and is the default there
    const bool pair_ok = (Q == 2 && fold == LWSB_FOLD_Q2) || (Q == 4 && fold == LWSB_FOLD_Q4);
    int var = variant;
    // automatic = one thread per task: with the branch-free projection it is as fast as or faster than the pair-split
    // kernels on every plan measured; those stay selectable
    if (var == LWSB_VARIANT_AUTO) var = LWSB_VARIANT_SCALAR;
    if (var >= LWSB_VARIANT_PAIR && (!pair_ok || var > LWSB_VARIANT_PAIR + 5)) var = LWSB_VARIANT_SCALAR;
#ifndef LWSB_PAIR_EXPERIMENTS
    if (var >= LWSB_VARIANT_PAIR) var = LWSB_VARIANT_PAIR + LWSB_PAIR_DEFAULT_MODE;
#endif
    if (var == LWSB_VARIANT_TM && Q > 4) var = LWSB_VARIANT_SCALAR;
    const bool tm = var == LWSB_VARIANT_TM;
    const bool pair = var >= LWSB_VARIANT_PAIR;
    const int task_cap = tm ? 128 : (pair ? (pair_thread_cap(max_sweeps) - 32) / 2 : 256 - 32);
    const int nbt = (Nreal + SBK - 1) / SBK; // blocks holding real bins
    bool found = false;
    double best = 0.0;
    for (int C = 1; C <= 8; C *= 2) {
        if (force_cluster > 0 && C != force_cluster) continue;
        const int NBr = (nbt + C - 1) / C;
        const int NBV = NBr + (NBr & 1);
        const int NS = NBV / 2;
        if (NBr < 2) continue;
        // every strip needs real bins and the last one the whole upper mirror zone
        if ((C - 1) * NBr * SBK > Nreal - 1 - SL) continue;
        int pitch = SBK * NBr + 2 * SL;
        if ((pitch & 1) == 0) ++pitch; // odd pitch: conflict-free 128-bit accesses across a warp's rows
        const size_t rowbytes = (size_t)pitch * 16;
        const size_t fixed = 64 + 16 + (size_t)(iters + 8) * sizeof(int) + 256 + 256; // flags, sweep list, alignment, static shared memory of the kernel
        if (smem_limit < fixed + rowbytes * 8) continue;
        const int Rmax = (int)((smem_limit - fixed) / (rowbytes + 8));
        const int ncl = std::max(1, (sm_count * 9 / 10) / C); // GPC packing loses a few SMs to clusters
        // sweep lag: Q frames is the minimum; an odd lag lets the sweep-fastest thread order be bank-conflict free
        for (int QS = Q; QS <= Q + 1; ++QS) {
            if (Rmax < 2 * Q + SLEAD + NS) continue;
            int Gmax = (Rmax - 2 * Q - SLEAD - NS) / QS + 1;
            Gmax = std::min(Gmax, task_cap / NS);
            Gmax = std::min(Gmax, iters);
            if (max_sweeps > 0) Gmax = std::min(Gmax, max_sweeps);
            if (max_sweeps < -1) Gmax = std::min(Gmax, -max_sweeps); // experiments: negative = sweeps per pass without the pair kernels' thread cap
            for (int G = Gmax; G >= 1; --G) {
                // shared-memory wavefronts per warp access: 8 tasks (a quarter-warp of 128-bit accesses, or a
                // half-warp of lane pairs reading 64 bits each) hit 16-byte bank groups (j - QS*g) mod 8 (odd
                // pitch); the busiest group sets the count
                double f = 0.0; int gfast = 0;
                for (int order = 0; order < 2; ++order) {
                    double waves = 0.0; int quarters = 0;
                    for (int q0 = 0; q0 < NS * G; q0 += 8) {
                        int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, mx = 0;
                        for (int tid = q0; tid < q0 + 8 && tid < NS * G; ++tid) {
                            const int jj = order ? tid / G : tid % NS, gg = order ? tid % G : tid / NS;
                            mx = std::max(mx, ++cnt[((jj - QS * gg) % 8 + 8) % 8]);
                        }
                        waves += mx; ++quarters;
                    }
                    if (order == 0 || waves / quarters < f) { f = waves / quarters; gfast = order; }
                }
                if (force_lag > 0 && QS != force_lag) continue;
                const int npass = (iters + G - 1) / G;
                const double steps = 2.0 * (maxT + QS * G) + NBV + (C - 1) * NBr;
                // Cost models fitted on B200.  One thread per task, branch-free projection: a
                // macro-step costs ~6k cycles plus ~1.3k per compute warp and shared-memory wavefront factor
                // (measured over cluster sizes 2-8, 3-6 warps, conflict factors 1.0-1.75; 11.5k-19.5k cycles).
                // Pair-split: half the instructions per warp and twice the warps.  A pass adds a fixed prologue.
                const int cwarps = ((pair ? 2 : 1) * NS * G + 31) / 32;
                // per-warp cost scales with the terms per bin: 6 (Q = 2), 17 (Q = 4, folded), 74 (Q = 8)
                const double tscale = Q == 2 ? 0.4 : (Q == 4 ? 1.0 : 4.3);
                const double t_step = pair ? PAIR_T0 + PAIR_T1 * cwarps * (1.0 + PAIR_TF * (f - 1.0)) + (C > 2 ? 800.0 : 0.0)
                                           : 6000.0 + 1300.0 * tscale * cwarps * f + (C > 2 ? 300.0 : 0.0) + (C > 4 ? 1700.0 : 0.0);
                // work items = (utterance, pass) pairs dealt to the resident clusters in turn
                // throughput bound, and the critical path of one utterance: its passes run concurrently on different
                // clusters, each `lag` macro-steps behind the previous one (it reads what that one has written back)
                const double lag = 2.0 * (Q + SLEAD + QS * (G - 1)) + NBV + 2 + (C - 1) * NBr;
                const double cost = std::max(std::ceil((double)B * npass / ncl) * (steps * t_step + 60000.0),
                                             (steps + (npass - 1) * lag) * t_step + 60000.0);
                if (!found || cost < best) {
                    found = true; best = cost;
                    out->C = C; out->NBr = NBr; out->NBV = NBV; out->NS = NS; out->G = G; out->pitch = pitch; out->QS = QS; out->GFAST = gfast;
                    out->TM = tm ? LWSB_VARIANT_TM : (pair ? var : 0);
                    out->R = QS * (G - 1) + 2 * Q + SLEAD + NS;
                    out->nthreads = ((pair ? 2 : 1) * NS * G + 31) / 32 * 32 + 32;
                    out->smem_bytes = (int)(fixed + (size_t)out->R * (rowbytes + 8));
                    out->smem_limit = (int)smem_limit;
                }
            }
        }
    }
    return found;
}

int strips_min_pitch(int Nreal, int c0)
{
    // the widest plan reads up to 8 blocks past the last real block plus the right halo
    return c0 + ((Nreal + SBK - 1) / SBK + 8) * SBK + SL;
}

cudaError_t launch_batch_strips(const LwsbView &v, const double *wr_host, const double *wi_host, int fold,
                                const double *thr, const double *max_amp, int iters, const StripPlan &pl,
                                unsigned *status, const int *items, int n_items, int max_pass, unsigned *done, unsigned long long *trace,
                                cudaStream_t s)
{
    StripParams prm;
    prm.items = reinterpret_cast<const int2 *>(items); prm.n_items = n_items; prm.max_pass = max_pass; prm.done = done; prm.trace = trace;
    prm.v = v; prm.thr = thr; prm.max_amp = max_amp; prm.iters = iters;
    prm.C = pl.C; prm.NBr = pl.NBr; prm.NBV = pl.NBV; prm.NS = pl.NS; prm.G = pl.G; prm.R = pl.R; prm.pitch = pl.pitch; prm.QS = pl.QS; prm.GFAST = pl.GFAST;
    prm.status = status;
    switch (v.Q) {
    case 2: return launch_strips_q<2>(prm, wr_host, wi_host, fold, pl, v.B, s);
    case 4: return launch_strips_q<4>(prm, wr_host, wi_host, fold, pl, v.B, s);
    case 8: return launch_strips_q<8>(prm, wr_host, wi_host, fold, pl, v.B, s);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_debug_fast_math(long long n, unsigned long long seed, unsigned long long *out4, cudaStream_t s)
{
    k_debug_fast_math<<<296, 256, 0, s>>>(n, seed, out4);
    return cudaGetLastError();
}

} // namespace lwsb

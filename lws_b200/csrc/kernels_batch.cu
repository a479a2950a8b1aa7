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
//   * The passes of an utterance are WORK ITEMS dealt to the resident clusters in turn: a pass reads each frame from
//     global memory after the previous pass of the same utterance -- usually running on another cluster at the same
//     time -- has written it back (per-strip progress counters, st.release.gpu / ld.acquire.gpu around the TMA
//     traffic), so one utterance keeps ceil(active/G) clusters busy and a batch keeps every SM busy.
//   * Arithmetic is the reference's, operation for operation (exact.cuh): results are
//     bit-identical to lwslib's LWSQ2 / LWSQ4 / LWSanyQ.  sqrt and the divisions are the correctly rounded fast
//     paths without their slow-path branches (fast_math.cuh).
// Files: this one holds the PTX helpers, the planner and the dispatch; strip_body.inc the block update, the kernel and
// its launch code, compiled for 8- and for 4-bin blocks; strip_pair.cuh the two-lanes-per-task variant.
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <algorithm>
#include <cstdlib>
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
constexpr int SBK_MAX = 8;     // bins per block: 8 or 4 (namespaces bk8 / bk4 below)
constexpr int SLEAD = 2;       // frames of TMA look-ahead
constexpr int PUBLISH_EVERY = 4; // a pass publishes its progress to the next pass of the utterance every so many frames
constexpr unsigned SPIN_LIMIT = 1u << 24; // polls before a wait is declared dead (seconds)
constexpr unsigned PASS_SPIN_LIMIT = 1u << 27; // waits for another cluster's pass: it may still be busy with earlier work items
#ifdef LWSB_PAIR_EXPERIMENTS
constexpr int pair_thread_cap(int max_sweeps) { return max_sweeps < 0 ? 512 : 256; }
#else
constexpr int pair_thread_cap(int) { return 256; }
#endif
// planner cost model of the pair-split kernels (cycles per macro-step: fixed + per warp; weight of the bank-conflict factor)
constexpr double PAIR_T0 = 8000.0, PAIR_T1 = 500.0, PAIR_TF = 0.3;
#ifdef LWSB_PAIR_EXPERIMENTS
constexpr int PAIR_THREADS_MAX = 512;  // pair-split kernels: 15 task warps (240 tasks) + the control warp at 128 registers
#endif
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
    int C, NBr, NBV, NS, G, R, pitch, QS, GFAST, GX, LEAD;
    int poll_sleep; // LWSB_STRIP_POLL_SLEEP=1: the neighbour-flag polls back off with __nanosleep(32) (default: tight spin, the compute
                    // warps of the CTA are waiting for the control warp anyway: 0.5 ms of 98 at BASELINE configs[1])
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
// release pattern with ONE fence for several flags: fence.acq_rel.cluster, then relaxed stores
__device__ __forceinline__ void fence_release_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
__device__ __forceinline__ void st_relaxed_cluster(unsigned *p, unsigned v)
{
    asm volatile("st.relaxed.cluster.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
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

// The block update, the kernel and its launch code exist twice: 8 bins per block with frames 2 blocks apart (Q = 8, the
// tensor-memory variant, wide strips) and 4 bins per block with frames 3 blocks apart (12 instead of 16 bins of ring per
// frame in flight: more tasks fit the shared-memory ring of a narrow strip).
int g_strip_launch_mode = -1; // last launch: 1 cooperative, 0 plain (launch_strips_t)
namespace bk8 {
constexpr int SBK = 8, LAGB = 2;
constexpr bool HAS_TM = true;
#include "strip_body.inc"
} // namespace bk8
#ifdef LWSB_EXPERIMENTS
namespace bk4 {
constexpr int SBK = 4, LAGB = 3;
constexpr bool HAS_TM = false;
#include "strip_body.inc"
} // namespace bk4
#endif

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


// Chooses cluster size, strip width and sweeps per pass.  Returns false when the shape is not
// served by this kernel (the generic kernel takes over).
int strip_launch_mode() { return g_strip_launch_mode; }

bool plan_strips(int Nreal, int Q, int L, int iters, int maxT, int B, size_t smem_limit, int sm_count, StripPlan *out,
                 int force_cluster, int max_sweeps, int force_lag, int variant, int fold, int force_block, double avg_iters)
{
    if (avg_iters <= 0.0 || avg_iters > iters) avg_iters = iters; // sweeps per utterance that can move a bin: mean over the batch (iters = the largest)
    if (L != SL || !(Q == 2 || Q == 4 || Q == 8) || iters < 1) return false;
    // variant: the pair-split kernel serves the folded Q = 2 / Q = 4 updates and is the default there
    const bool pair_ok = (Q == 2 && fold == LWSB_FOLD_Q2) || (Q == 4 && fold == LWSB_FOLD_Q4);
    int var = variant;
    // automatic = one thread per task: with the branch-free projection it is as fast as or faster than the pair-split
    // kernels on every plan measured; those stay selectable
    if (var == LWSB_VARIANT_AUTO) var = LWSB_VARIANT_SCALAR;
    if (var == LWSB_VARIANT_DUO && !pair_ok) var = LWSB_VARIANT_SCALAR; // two lanes per task: the folded Q = 2 / Q = 4 updates
#ifndef LWSB_EXPERIMENTS
    if (var == LWSB_VARIANT_TM || var == LWSB_VARIANT_DUO || var >= LWSB_VARIANT_PAIR) var = LWSB_VARIANT_SCALAR;
#endif
    if (var >= LWSB_VARIANT_PAIR && (!pair_ok || var > LWSB_VARIANT_PAIR + 5)) var = LWSB_VARIANT_SCALAR;
#ifndef LWSB_PAIR_EXPERIMENTS
    if (var >= LWSB_VARIANT_PAIR) var = LWSB_VARIANT_PAIR + LWSB_PAIR_DEFAULT_MODE;
#endif
    if (var == LWSB_VARIANT_TM && Q > 4) var = LWSB_VARIANT_SCALAR;
    const bool tm = var == LWSB_VARIANT_TM;
    const bool pair = var >= LWSB_VARIANT_PAIR;
    const bool duo = var == LWSB_VARIANT_DUO;
    const int task_cap = tm ? 128 : (duo ? 7 * 32 : (pair ? (pair_thread_cap(max_sweeps) - 32) / 2 : 256 - 32)); // DUO: 14 named barriers = 7 warp pairs
    bool found = false;
    double best = 0.0;
    // block size: 8 bins with frames 2 blocks apart, or (Q <= 4, not the tensor-memory variant) 4 bins with frames 3
    // blocks apart -- 12 instead of 16 bins of ring per frame in flight, i.e. more tasks in a narrow strip's ring.
    // Measured on B200 (profiles/r1b_block_size_experiment.txt) the extra tasks do not pay: the shared-memory pipe, not
    // the number of resident tasks, bounds a busy SM, and twice the macro-steps mean twice the barriers.  4-bin blocks
    // are therefore only planned on request (lwsb_set_block_bins / LWSB_STRIP_BLOCK).
    for (int SBK = 8; SBK >= 4; SBK /= 2) {
    if (SBK == 4 && (Q > 4 || tm || force_block != 4)) continue;
    if (force_block > 0 && SBK != force_block) continue;
    const int LAGB = (SBK + SL + SBK - 1) / SBK;
    const int nbt = (Nreal + SBK - 1) / SBK; // blocks holding real bins
    for (int C = 1; C <= 8; C *= 2) {
        if (force_cluster > 0 && C != force_cluster) continue;
        const int NBr = (nbt + C - 1) / C;
        const int NBV = (NBr + LAGB - 1) / LAGB * LAGB;
        const int NS = NBV / LAGB;
        if (NBr < 2 || SBK * NBr < 2 * SL) continue;
        // every strip needs real bins and the last one the whole upper mirror zone
        if ((C - 1) * NBr * SBK > Nreal - 1 - SL) continue;
        // ring row: the strip's bins, an L-bin halo on either side and 7 cells of slack (the row starts (frame - slot *
        // pitch) mod 8 cells into its slot, see ring_off in strip_body.inc)
        const int pitch = SBK * NBr + 2 * SL + 7;
        const size_t rowbytes = (size_t)pitch * 16;
        // flags, sweep list, alignment, static shared memory of the kernel; Q > 4: the weight table of the bin-loop update
        const size_t fixed = 64 + 16 + (size_t)(iters + 8) * sizeof(int) + 256 + 256 + (Q > 4 ? (size_t)Q * Q * (SL + 1) * 16 + 16 : 0);
        if (smem_limit < fixed + rowbytes * 8) continue;
        const int Rmax = (int)((smem_limit - fixed) / (rowbytes + 8));
        // resident clusters (cudaOccupancyMaxActiveClusters on B200, one CTA per SM): 148 / 74 / 36-37 / 18
        const int ncl = std::max(1, C <= 2 ? sm_count / C : (C == 4 ? sm_count / C - 1 : sm_count / C));
        // sweep lag: Q frames is the minimum; an odd lag makes the sweep-fastest thread order bank-conflict free
        for (int QS = Q; QS <= Q + 1; ++QS) {
            if (Rmax < 2 * Q + SLEAD + NS) continue;
            int Gmax = (Rmax - 2 * Q - SLEAD - NS) / QS + 1;
            Gmax = std::min(Gmax, iters);
            if (max_sweeps > 0) Gmax = std::min(Gmax, max_sweeps);
            if (max_sweeps < -1) Gmax = std::min(Gmax, -max_sweeps); // experiments: negative = sweeps per pass without the pair kernels' thread cap
            for (int G = Gmax; G >= 1; --G) {
                // Shared-memory cycles per 128-bit warp access (measured, tools/ubench/lds.cu): the 16 lanes of a half-warp
                // are served together, one cycle per round of distinct 16-byte cells in different bank groups -- 2 cycles
                // when the 16 cells spread evenly over the 8 groups, more when a group is hit 3 times or more.  A lane's
                // bank group is (frame + column) mod 8 (ring_off) and the lanes of a warp are on the same column mod 8, so
                // what counts is the frames' residues.  Two thread orders:
                //   frame-fastest: lanes on consecutive frame slots (compact, but a half-warp that straddles the newest /
                //     oldest frame in flight or two sweeps hits a group three times);
                //   sweep-fastest: the sweeps of a frame slot padded to groups of 8 lanes, QS frames apart: with an odd
                //     QS each group covers 8 different residues whatever the slots' frames are -- conflict free.
                //   rotating (NS = 17 = 16 + 1 only): half-warp h holds sweep slot h and 16 of its 17 frame slots -- all but the one
                //     that wrapped last (slot ph), whose frame residue doubles its neighbour's; the G left-over tasks share a last
                //     half-warp.  The lane -> frame slot map changes with ph (a task keeps nothing in registers between macro-steps):
                //     every full half-warp is conflict free at every macro-step.
                // Rotating order with an even lag: the G left-over tasks (one per sweep slot, frame slot ph) have frame residues
                // ph + 1 - QS g, i.e. only two classes for QS = 4: a 4-way conflict in their half-warp.  One extra frame of lag from
                // sweep slot GX = (G + 1) / 2 on spreads them over four classes (2-way); its ring row is taken from the TMA look-ahead
                // (1 frame instead of 2: a frame clock is two macro-steps, ~17 us, a row load ~1 us).  LWSB_STRIP_NO_EXTRA disables.
                double f = 0.0; int gfast = 0, lanes_best = 0, gx_best = G;
                for (int order = 0; order < 4; ++order) { // order 3: rotating with the extra frame of lag
                    const int gx = order == 3 ? (G + 1) / 2 : G;
                    if (order == 3 && (G < 3 || (QS & 1) || getenv("LWSB_STRIP_NO_EXTRA"))) continue;
                    const int GP8 = (G + 7) & ~7;
                    const int lanes = order == 1 ? NS * GP8 : NS * G;
                    if (lanes > task_cap) continue;
                    if (order >= 2 && !(NS == 17 && LAGB == 2 && !pair && !duo && !tm)) continue;
                    if (order >= 2 && getenv("LWSB_STRIP_NO_ROTATE")) continue;
                    double cyc = 0.0, ideal = 0.0;
                    for (int ph = 0; ph < NS; ++ph)       // slots 0 .. ph have wrapped to their next frame (+NS)
                        for (int h0 = 0; h0 < lanes; h0 += 16) {
                            int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, mx = 0, act = 0;
                            for (int l = h0; l < h0 + 16 && l < lanes; ++l) {
                                int jj = order ? l / GP8 : l % NS, gg = order ? l % GP8 : l / NS;
                                if (order >= 2) {
                                    if (l < 16 * G) { gg = l / 16; jj = l % 16 < ph ? l % 16 : l % 16 + 1; }
                                    else { gg = l - 16 * G; jj = ph; }
                                }
                                if (gg >= G) continue;
                                const int fr = jj + (jj <= ph ? NS : 0) - QS * gg - (gg >= gx ? 1 : 0);
                                mx = std::max(mx, ++cnt[((fr % 8) + 8) % 8]); ++act;
                            }
                            cyc += mx; ideal += (act + 7) / 8;
                        }
                    const double ff = ideal > 0.0 ? cyc / ideal : 1.0;
                    if (lanes_best == 0 || ff < f) { f = ff; gfast = order == 3 ? 2 : order; lanes_best = lanes; gx_best = gx; }
                }
                if (lanes_best == 0) continue;
                if (force_lag > 0 && QS != force_lag) continue;
                const int npass = (iters + G - 1) / G;
                const double steps = (double)LAGB * (maxT + QS * G) + NBV + (C - 1) * NBr;
                // Cost models fitted on B200.  One thread per task, branch-free projection: a
                // macro-step of 8-bin blocks costs ~10.5k cycles (the in-order instruction stream of one warp) plus ~0.45k
                // per compute warp and shared-memory wavefront factor, and 1.4k more for each beyond 5.6 (the shared-memory
                // pipe saturates); measured over cluster sizes 2-8, 2-6 warps, conflict factors 1.0-1.75: 11.5k-19.5k cycles.
                // The stream scales with the bins of a block, the two CTA barriers do not.
                // Pair-split: half the instructions per warp and twice the warps.  A pass adds a fixed prologue.
                const int cwarps = ((pair || duo ? 2 : 1) * lanes_best + 31) / 32;
                // per-warp cost scales with the terms per bin: 6 (Q = 2), 17 (Q = 4, folded), 74 (Q = 8)
                // (Q = 8: bin-loop block update with the default-window mask compiled in, measured 46k cycles per macro-step at
                // 3 warps on BASELINE configs[4]; 84k before, with one branch per term)
                const double tscale = Q == 2 ? 0.4 : (Q == 4 ? 1.0 : 4.0);
                const double bscale = SBK / 8.0;
                // Bank conflicts enter the time weakly: measured on B200 (profiles/r2_strip_experiments.txt) the conflict-free
                // order (sweep-fastest, odd sweep lag, rotated ring rows: 1.0 instead of ~1.5 cycles per access) is 15 % faster
                // at cluster 4 with the same sweeps per pass (128 vs 151 ms on BASELINE configs[1]); at cluster 2 it costs two
                // sweeps per pass and loses.  The stream of a warp is bound by fp64 issue and load latency first.
                // Effective cycles per macro-step of a launch (kernel time / rounds of work items / steps per item, so the waits
                // for neighbour strips and for other passes are in it), fitted on BASELINE configs[1] at default and zero
                // thresholds: cluster 2, 4 warps: 17.2-17.8k; cluster 4, 5 warps: 19.5-20.0k; cluster 8, 4 warps: 19-24k, 6 warps: 36k
                // (a fifth warp shares a scheduler with another one: the step lasts as long as that scheduler's two streams;
                // eight strips in lock step wait for one another a lot: profiles/r2_strip_experiments.txt).
                const double t_step = pair ? (PAIR_T0 - 800.0 + PAIR_T1 * cwarps * (1.0 + PAIR_TF * (f - 1.0))) * bscale + 800.0 + (C > 2 ? 800.0 : 0.0)
                                           : (14700.0 + 500.0 * std::min(cwarps, 4) + 4500.0 * std::max(0, cwarps - 4)) * (1.0 + 0.3 * (f - 1.0)) * tscale * bscale +
                                                 800.0 + (C > 2 ? 1700.0 : 0.0) + (C > 4 ? 6000.0 : 0.0);
                // throughput bound, and the critical path of one utterance: its passes run concurrently on different
                // clusters, each `lag` macro-steps behind the previous one (it reads what that one has written back)
                const double lag = (double)LAGB * (Q + SLEAD + QS * (G - 1) + PUBLISH_EVERY) + NBV + LAGB + (C - 1) * NBr;
                const double items = (double)B * std::ceil(avg_iters / G); // work items of the launch (one per utterance and pass)
                const double cost = std::max(std::ceil(items / ncl) * (steps * t_step + 60000.0),
                                             (steps + (npass - 1) * lag) * t_step + 60000.0);
                if (!found || cost < best) {
                    found = true; best = cost;
                    out->SBK = SBK; out->LAGB = LAGB;
                    out->C = C; out->NBr = NBr; out->NBV = NBV; out->NS = NS; out->G = G; out->pitch = pitch; out->QS = QS; out->GFAST = gfast;
                    out->TM = tm ? LWSB_VARIANT_TM : (pair ? var : (duo ? LWSB_VARIANT_DUO : 0));
                    out->GX = gx_best; out->LEAD = gx_best < G ? SLEAD - 1 : SLEAD;
                    out->R = QS * (G - 1) + (gx_best < G ? 1 : 0) + 2 * Q + out->LEAD + NS; // the extra frame takes the row the shorter look-ahead frees
                    out->nthreads = duo ? 2 * ((lanes_best + 31) / 32 * 32) : ((pair ? 2 : 1) * lanes_best + 31) / 32 * 32 + 32;
                    out->smem_bytes = (int)(fixed + (size_t)out->R * (rowbytes + 8));
                    out->smem_limit = (int)smem_limit;
                }
            }
        }
    }
    }
    return found;
}

int strips_min_pitch(int Nreal, int c0)
{
    // the widest plan reads up to 8 blocks past the last real block plus the right halo
    return c0 + ((Nreal + SBK_MAX - 1) / SBK_MAX + 8) * SBK_MAX + SL;
}

cudaError_t launch_batch_strips(const LwsbView &v, const double *wr_host, const double *wi_host, int fold,
                                const double *thr, const double *max_amp, int iters, const StripPlan &pl,
                                unsigned *status, const int *items, int n_items, int max_pass, unsigned *done, unsigned long long *trace,
                                cudaStream_t s)
{
    StripParams prm;
    prm.items = reinterpret_cast<const int2 *>(items); prm.n_items = n_items; prm.max_pass = max_pass; prm.done = done; prm.trace = trace;
    prm.v = v; prm.thr = thr; prm.max_amp = max_amp; prm.iters = iters;
    prm.C = pl.C; prm.NBr = pl.NBr; prm.NBV = pl.NBV; prm.NS = pl.NS; prm.G = pl.G; prm.R = pl.R; prm.pitch = pl.pitch; prm.QS = pl.QS; prm.GFAST = pl.GFAST; prm.GX = pl.GX; prm.LEAD = pl.LEAD;
    { const char *e = getenv("LWSB_STRIP_POLL_SLEEP"); prm.poll_sleep = e ? atoi(e) : 0; }
    prm.status = status;
    if (pl.SBK == 4) {
#ifdef LWSB_EXPERIMENTS
        switch (v.Q) {
        case 2: return bk4::launch_strips_q<2>(prm, wr_host, wi_host, fold, pl, v.B, s);
        case 4: return bk4::launch_strips_q<4>(prm, wr_host, wi_host, fold, pl, v.B, s);
        }
#endif
        return cudaErrorInvalidValue;
    }
    switch (v.Q) {
    case 2: return bk8::launch_strips_q<2>(prm, wr_host, wi_host, fold, pl, v.B, s);
    case 4: return bk8::launch_strips_q<4>(prm, wr_host, wi_host, fold, pl, v.B, s);
    case 8: return bk8::launch_strips_q<8>(prm, wr_host, wi_host, fold, pl, v.B, s);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_debug_fast_math(long long n, unsigned long long seed, unsigned long long *out4, cudaStream_t s)
{
    k_debug_fast_math<<<296, 256, 0, s>>>(n, seed, out4);
    return cudaGetLastError();
}

} // namespace lwsb

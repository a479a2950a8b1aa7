// kernels.h -- launch wrappers implemented in the .cu files (host-callable).
#pragma once
#include <cuda_runtime.h>
#include <vector>
#include "lwsb_common.h"
#include "exact.cuh"

namespace lwsb {

// kernels_generic.cu
struct StatScratch {     // work space of the numpy-ordered mean (k_stats)
    double *row_max;     // [total rows]
    const int *leaf_tab; // summation trees (offset, length, additions-after) of the array lengths in the batch
    const int2 *tab_of;  // [B] (first leaf, leaf count) of utterance b's tree in leaf_tab
    double *leaf_sum;    // [B][stride]
    long long stride;    // leaves reserved per utterance
};
void stat_tree(long long n, std::vector<int> &tab); // host: numpy's pairwise-summation tree for n elements

void launch_extend(const LwsbView &v, int kind, const void *const *src, const StatScratch &sc, double *mean_amp,
                   double *max_amp, int maxTp, cudaStream_t s);
void launch_refresh_ghosts(const LwsbView &v, cudaStream_t s);
void launch_reextend(const LwsbView &v, const StatScratch &sc, double *mean_amp, double *max_amp, int maxTp,
                     cudaStream_t s);
void launch_crop(const LwsbView &v, void *const *dst, int maxT, cudaStream_t s);
void launch_sweeps_generic(const LwsbView &v, const LwsbW &w, int fold, int rframe, int cframe, const double *thr,
                           int iters, cudaStream_t s);
void launch_online_generic(const LwsbView &v, const LwsbW *w3, int fold, const double *thr, int iters, int LA,
                           cudaStream_t s, long long j0 = 0, long long j1 = -1);
void launch_stream_extend(const LwsbView &v, int kind, const void *src, int m0, int n, cudaStream_t s);
// kernels_online.cu
bool launch_online_ring(const LwsbView &v, const double *const *wr_host, const double *const *wi_host, int fold,
                        const double *thr, int iters, int LA, const int *T_host, size_t smem_limit, unsigned *status,
                        cudaStream_t s, cudaError_t *err, int *which_kernel);
void launch_nofuture_q4(const LwsbView &v, const LwsbW &w, const double *thr, int iters, cudaStream_t s);

// kernels_batch.cu
struct StripPlan {
    int smem_limit; // shared-memory budget the plan was made for
    int SBK, LAGB;  // bins per block (8 or 4) and blocks between consecutive frames (2 or 3)
    int C, NBr, NBV, NS, G, R, pitch, nthreads, smem_bytes, QS, GFAST, TM; // TM: 0 one thread per task, LWSB_VARIANT_TM, LWSB_VARIANT_PAIR + window mode
    int GX;   // sweep slots g >= GX run one more frame behind (sweep offset QS g + 1); GX == G: none
    int LEAD; // frames of TMA look-ahead (2; 1 when the extra frame of GX takes its ring row)
};
bool plan_strips(int Nreal, int Q, int L, int iters, int maxT, int B, size_t smem_limit, int sm_count, StripPlan *out,
                 int force_cluster = 0, int max_sweeps = 0, int force_lag = 0, int variant = 0, int fold = 0, int force_block = 0,
                 double avg_iters = 0.0);
int strips_min_pitch(int Nreal, int c0);
cudaError_t launch_batch_strips(const LwsbView &v, const double *wr_host, const double *wi_host, int fold,
                                const double *thr, const double *max_amp, int iters, const StripPlan &pl,
                                unsigned *status, const int *items, int n_items, int max_pass, unsigned *done, unsigned long long *trace,
                                cudaStream_t s);
constexpr int STRIP_MAX_CLUSTER = 8; // stride of the per-utterance progress counters

cudaError_t launch_debug_fast_math(long long n, unsigned long long seed, unsigned long long *out4, cudaStream_t s);

// kernels_fft.cu
cudaError_t launch_stft(const double *x, int B, int nsamples, const double *awin, int fsize, int hop, int N, int logN,
                        int pre, const double2 *tw, double2 *S, int M, cudaStream_t s);
cudaError_t launch_istft(const double2 *S, int B, int M, int N, int logN, const double *swin, int nswin, int hop,
                         const double2 *tw, double *frames, double *signal, cudaStream_t s);
cudaError_t launch_cabs(const double2 *S, double *A, long long n, cudaStream_t s);
cudaError_t launch_sq_norms(const double2 *S, const double2 *R, int B, long long n, double *partial, int nblk, double *out,
                            cudaStream_t s);
inline int online_generic_max_nreal(int L) { return (1024 - 1) * (L + 1); }
int strip_launch_mode(); // last strip-kernel launch of the process: 1 cooperative, 0 plain, -1 none

} // namespace lwsb

/* lws_b200.h -- C-ABI of the B200-native LWS phase-recovery library (liblws_b200.so).
 *
 * This is the drop-in boundary for the hot path of Jonathan-LeRoux/lws.  The reference's
 * own native boundary is the nine `void f(double *Sr, double *Si, double *wr, ...)` functions
 * that python/lwslib.pxd:1-13 binds out of lwslib/lwslib.h:6-26 (in-place, split re/im
 * planes, caller-owned host buffers, no error channel), driven by the three binding
 * functions python/lws.pyx:209-258 (batch_lws), 261-311 (nofuture_lws), 314-375 (online_lws).
 * On a GPU the unit of work has to be a whole call on a whole batch of utterances, so each
 * entry point below replaces one *binding function* (the pre/post-processing included:
 * extspec, |.|, mean, threshold scaling, variant dispatch, crop) rather than one sweep.
 *
 * Conventions
 *   - plain C, plain pointers and sizes; every function returns 0 on success or a negative
 *     lwsb_status; nothing throws across the boundary; lwsb_last_error() gives the text.
 *   - host pointers are never retained; device state lives in the context.
 *   - spectrograms are row-major (T, Nreal) like the reference (python/lws.pyx:221-222),
 *     element type selected by `kind`: LWSB_C128 = interleaved re,im doubles (numpy
 *     complex128), LWSB_F64 = real doubles (a magnitude spectrogram; imaginary part 0).
 *   - weights are (Qprime, Q, L+1) row-major fp64 re / im planes exactly as
 *     python/lws.pyx:227-228 hands them to the C core; the |W| > 1e-12 mask
 *     (lws.pyx:231-232) is rebuilt inside.
 *   - all arithmetic is IEEE fp64 (the reference's type; fp32 cannot hold the 1e-5 parity
 *     bound).
 *   - there is NO CPU fallback: without a CUDA device lwsb_create() fails.
 */
#ifndef LWS_B200_H_INCLUDED
#define LWS_B200_H_INCLUDED

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lwsb_ctx lwsb_ctx;

typedef enum {
    LWSB_OK = 0,
    LWSB_ERR_CUDA = -1,        /* CUDA runtime / driver failure (text in lwsb_last_error)      */
    LWSB_ERR_ARG = -2,         /* bad argument (NULL, negative size, unknown enum)              */
    LWSB_ERR_EVEN_NREAL = -3,  /* Nreal even: reference raises ValueError (lws.pyx:223-224)     */
    LWSB_ERR_UNSUPPORTED = -4, /* a shape this build has no kernel for (text in lwsb_last_error) */
    LWSB_ERR_STATE = -5,       /* call order: weights or spectrograms not loaded                */
    LWSB_ERR_NOMEM = -6
} lwsb_status;

enum { LWSB_C128 = 0, LWSB_F64 = 1 };           /* element kind of a spectrogram buffer        */
enum { LWSB_W = 0, LWSB_W_AI = 1, LWSB_W_AF = 2 }; /* weight sets (lws.pyx:426-429)             */
enum { LWSB_HOST = 0, LWSB_DEVICE = 1 };        /* where the caller's buffers live             */

/* flags for the compute calls */
enum {
    LWSB_FORCE_GENERIC = 1, /* use the generic wavefront kernels even where a tuned one exists  */
    LWSB_FORCE_ANYQ = 2,    /* use the anyQ formulas for Q = 2 / 4 (debug: variant equivalence) */
    LWSB_FRACTIONAL = 4     /* per-frequency weight rows (the reference's *fractionalQ variants, lws.pyx:246-247, 299-300;
                               lwslib.cpp:1441): implied whenever the weights were set with Qprime != Q */
};

/* ---- library / context ------------------------------------------------------------- */
int lwsb_version(void);                      /* 10000*major + 100*minor + patch               */
int lwsb_has_experiments(void);              /* 1 when built with -DLWSB_EXPERIMENTS (extra kernel variants) */
int lwsb_strip_launch_mode(void);            /* last launch of the cluster strip kernel in this process: 1 cooperative (co-residency
                                                guaranteed by the driver), 0 plain, -1 none yet */
const char *lwsb_last_error(const lwsb_ctx *ctx); /* ctx may be NULL: error of the last failed
                                                lwsb_create() on this thread                  */
/* `stream` is a cudaStream_t to launch on (e.g. the caller's current stream) or NULL to let
 * the context create its own non-blocking stream. */
int lwsb_create(int device, void *stream, lwsb_ctx **out);
int lwsb_destroy(lwsb_ctx *ctx);
int lwsb_sync(lwsb_ctx *ctx);                /* wait for everything queued on the stream       */
/* Page-locked host memory (cudaHostAlloc, portable across devices): buffers the DMA engines reach directly.  Pageable
 * host buffers are accepted everywhere too -- the library stages them through pinned chunks with a few host threads
 * (env LWSB_HOST_THREADS) -- but a pinned result buffer saves that pass and the page faults of fresh memory. */
int lwsb_host_alloc(unsigned long long bytes, void **out);
int lwsb_host_free(void *p);

/* ---- weights: replaces the Wr/Wi/Wflag marshalling of lws.pyx:227-232, 341-352 -------- */
int lwsb_set_weights(lwsb_ctx *ctx, int which, const double *wr, const double *wi, int Qprime, int Q, int L);

/* create_weights (lws.pyx:160-181) on the host side of the library: writes (Qprime,Q,L+1)
 * re / im planes; Qprime = Q when fshift divides T and use_summarized_weights, else T. */
int lwsb_create_weights(const double *awin, const double *swin, int T, int fshift, int L,
                        int use_summarized_weights, double *wr, double *wi, int *Qprime_out, int *Q_out);

/* ---- staged interface: keep a batch of utterances resident in HBM ----------------------
 * lwsb_load     : B spectrograms (T[b], Nreal) -> extended spectrogram (extspec, lws.pyx:146-157,
 *                 235-237), amplitude plane and mean amplitude (lws.pyx:239-240) on the device.
 *                 Needs LWSB_W to be set (it fixes Q and L).
 * lwsb_batch / lwsb_nofuture / lwsb_online : the iteration loops of lws.pyx:244-253,
 *                 297-306, 365-370 on the resident batch, in place.  `thresholds` are the
 *                 UNSCALED values (get_thresholds, lws.pyx:203-206); they are multiplied by
 *                 each utterance's mean amplitude inside, as lws.pyx:245,298,361 do.
 *                 iterations == 0 is a no-op (lws.pyx:219-220).
 *                 lwsb_nofuture uses the set selected by `which` (the class passes W_ai,
 *                 lws.pyx:475).  lwsb_online needs all three sets.
 * lwsb_store    : crop + recombine (lws.pyx:256) -> B complex128 (T[b], Nreal) arrays.
 */
int lwsb_load(lwsb_ctx *ctx, const void *const *S_in, const int *T, int B, int Nreal, int kind, int where);
int lwsb_batch(lwsb_ctx *ctx, const double *thresholds, int iterations, int flags);
int lwsb_nofuture(lwsb_ctx *ctx, int which, const double *thresholds, int iterations, int flags);
int lwsb_online(lwsb_ctx *ctx, const double *thresholds, int iterations, int look_ahead, int flags);
int lwsb_store(lwsb_ctx *ctx, void *const *S_out, int where);
/* what happens between two chained reference calls (lws.pyx:256 then 235-240 of the next one): ghost frames become copies of
 * the UPDATED edge frames, |.| and its mean are recomputed.  Call it between two stages run on the resident batch to get
 * run_lws (lws.pyx:495-499); lwsb_run_lws does. */
int lwsb_restage(lwsb_ctx *ctx);

/* ---- one-shot interface: one call == one reference binding call on B utterances --------
 * (host or device buffers in, complex128 out; synchronous on return) */
int lwsb_batch_lws(lwsb_ctx *ctx, const void *const *S_in, void *const *S_out, const int *T, int B, int Nreal,
                   int kind, int where, const double *thresholds, int iterations, int flags);
int lwsb_nofuture_lws(lwsb_ctx *ctx, int which, const void *const *S_in, void *const *S_out, const int *T, int B,
                      int Nreal, int kind, int where, const double *thresholds, int iterations, int flags);
int lwsb_online_lws(lwsb_ctx *ctx, const void *const *S_in, void *const *S_out, const int *T, int B, int Nreal,
                    int kind, int where, const double *thresholds, int iterations, int look_ahead, int flags);
/* run_lws (lws.pyx:495-499): nofuture(W_ai) -> online -> batch without leaving the device.
 * The ghost frames / mean amplitude are re-derived between the stages exactly as the three
 * chained reference calls do. */
int lwsb_run_lws(lwsb_ctx *ctx, const void *const *S_in, void *const *S_out, const int *T, int B, int Nreal,
                 int kind, int where, const double *nofuture_thr, int nofuture_it, const double *online_thr,
                 int online_it, int look_ahead, const double *batch_thr, int batch_it, int flags);

/* ---- streaming online_lws (frame in, frame out) --------------------------------------------------
 * TF_RTISI_LA (lwslib.cpp:1432-1491) is causal: the row updates of frame m read nothing beyond frame m.  A stream is one
 * growing utterance kept in HBM; lwsb_stream_push appends frames and runs their row updates, after which every frame
 * older than the last `look_ahead` ones is final (lwsb_stream_frames).  The reference scales the thresholds by the mean
 * amplitude of the WHOLE utterance (lws.pyx:360-361): a stream takes that number (or the caller's estimate) up front;
 * given the true mean the frames are bit-identical to lwsb_online_lws on the complete spectrogram.
 * The three weight sets must have been set; the stream replaces the context's resident batch. */
int lwsb_stream_begin(lwsb_ctx *ctx, int Nreal, int max_frames, int kind, double mean_amp, const double *thresholds,
                      int iterations, int look_ahead, int flags);
int lwsb_stream_push(lwsb_ctx *ctx, const void *frames, int nframes, int where);        /* (nframes, Nreal) of `kind`   */
int lwsb_stream_frames(const lwsb_ctx *ctx, int *pushed, int *final_frames);            /* frames 0 .. final-1 are final */
int lwsb_stream_read(lwsb_ctx *ctx, void *out, int first_frame, int nframes, int where); /* complex128 (nframes, Nreal)  */
int lwsb_stream_end(lwsb_ctx *ctx);                                                      /* returns the frame count; all final */

/* ---- stft / istft (lws.pyx:43-90, 93-137), batched over signals of equal length ---------
 * lwsb_stft : x (B, nsamples) real -> S (B, M, fftsize/2+1) complex128.  Frame m covers the
 *             padded samples [m*fshift, m*fshift + fsize) where padded sample p is x[p - pre_pad]
 *             (zero outside the signal): pre_pad and M are what lws.pyx:54-77 derive from
 *             `perfectrec` (lwsb_stft_frames / lwsb_stft_prepad compute them).  The frame is
 *             multiplied by awin[fsize] and transformed with an fftsize-point DFT (zero-padded
 *             or cropped like np.fft.fft(frame, n=fftsize), lws.pyx:85-88).
 * lwsb_istft: S (B, M, Nreal) -> overlap-added signal (B, fshift*(M-1) + 2*(Nreal-1)), every
 *             inverse frame multiplied by swin (zero beyond nswin), lws.pyx:116-126.  The
 *             perfect-reconstruction crop of lws.pyx:128-135 is a slice the caller takes.
 * awin / swin are host pointers; x / S live where `where` says. */
int lwsb_stft_frames(int nsamples, int fsize, int fshift, int perfectrec);
int lwsb_stft_prepad(int fsize, int fshift, int perfectrec);
int lwsb_stft(lwsb_ctx *ctx, const double *x, int B, int nsamples, const double *awin, int fsize, int fshift,
              int fftsize, int pre_pad, int M, void *S_out, int where);
int lwsb_istft(lwsb_ctx *ctx, const void *S_in, int B, int M, int Nreal, const double *swin, int nswin, int fshift,
               double *x_out, int where);

/* ---- fused calls -----------------------------------------------------
 * lwsb_reconstruct : waveform -> waveform without leaving the device: y = istft(run_lws(|stft(x)|)) for B signals of
 *                    equal length (x: (B, nsamples) doubles, y: (B, lwsb_reconstruct_length(...)) doubles, both where
 *                    `where` says; fftsize = fsize).  The stages are the ones above, chained on the device: the
 *                    magnitudes are numpy's |.| of the STFT, the windows are host pointers of fsize doubles.
 *                    consistency_db (host, [B]) is optional: the consistency of the result (below).
 * lwsb_consistency : get_consistency (lws.pyx:140-144) of B spectrograms (B, M, Nreal) complex128:
 *                    20 log10(|S| / |stft(istft(S)) - S|), Frobenius norms, one value per spectrogram (host, [B]). */
/* lwsb_resident_consistency : the same number for every utterance of the RESIDENT batch as it stands (host, [B]) -- between
 *                    lwsb_batch calls it gives the per-sweep trace of the metric without moving the batch. */
int lwsb_resident_consistency(lwsb_ctx *ctx, const double *awin, const double *swin, int nswin, int fshift, int perfectrec,
                              double *out_db);
long long lwsb_reconstruct_length(int nsamples, int fsize, int fshift, int perfectrec);
int lwsb_reconstruct(lwsb_ctx *ctx, const double *x, int B, int nsamples, const double *awin, const double *swin, int fsize,
                     int fshift, int perfectrec, const double *nofuture_thr, int nofuture_it, const double *online_thr,
                     int online_it, int look_ahead, const double *batch_thr, int batch_it, int flags, double *y_out, int where,
                     double *consistency_db);
int lwsb_consistency(lwsb_ctx *ctx, const void *S, int B, int M, int Nreal, const double *awin, const double *swin, int nswin,
                     int fshift, int perfectrec, int where, double *out_db);

/* ---- introspection used by bench.py / tests -------------------------------------------- */
/* device time (ms, CUDA events on the context's stream) of the compute kernels of the last
 * lwsb_batch / lwsb_nofuture / lwsb_online call, and how many kernels it launched. */
int lwsb_last_compute_ms(lwsb_ctx *ctx, float *ms);
long long lwsb_launch_count(const lwsb_ctx *ctx); /* kernels launched by this context so far */
/* device time (ms) of each stage run since the last lwsb_load -- ms3[0] no-future sweeps, ms3[1] online chain,
 * ms3[2] batch sweeps; -1 for a stage that did not run (lwsb_run_lws runs up to three) */
int lwsb_last_stage_ms(lwsb_ctx *ctx, float *ms3);
/* work of the last lwsb_batch call: out4 = {bin-iterations asked for (bins x iterations), bin-iterations of the sweeps
 * that can move a bin (threshold below max|S|; the others are dropped before launch), work items, passes} */
int lwsb_last_batch_work(const lwsb_ctx *ctx, long long *out4);
/* which kernel the last lwsb_online call ran: 0 generic (global memory), 1 shared-memory ring with one bin per step,
 * 2 ring with two bins per step and thread, 3 ring with two bins per step on two lanes (env LWSB_ONLINE_DUO=0 selects 2),
 * 4 ring with four warps per row update taking turns (the default for Q <= 4; LWSB_ONLINE_FLOW=0 selects 3),
 * 5 value warps + chain warps (experiments build) */
int lwsb_last_online_kernel(const lwsb_ctx *ctx);
/* 1 and the plan {cluster size, blocks per strip, virtual blocks, frame slots, sweeps per pass, ring rows,
 * ring pitch, threads, shared-memory bytes, frames between sweeps, thread order (0 frame slots fastest, 1 sweep slots
 * fastest, 2 rotating), kernel variant, bins per block, first sweep slot that runs one more frame behind (= sweeps per
 * pass: none), frames of TMA look-ahead} (15 ints; pass room for 16) when the last lwsb_batch ran the cluster strip
 * kernel, 0 when it ran the generic wavefront kernel */
int lwsb_last_batch_plan(const lwsb_ctx *ctx, int *out9);
/* tuning knobs of the strip kernel's planner (0 = automatic): shared-memory budget per CTA in bytes, cluster
 * size (1, 2, 4, 8) and sweeps in flight per pass.  Also settable through the environment variables
 * LWSB_STRIP_SMEM / LWSB_STRIP_CLUSTER / LWSB_STRIP_SWEEPS (and LWSB_STRIP_LAG, LWSB_STRIP_TM) read by lwsb_create().  Results do not depend on them. */
int lwsb_set_tuning(lwsb_ctx *ctx, long long smem_limit, int cluster, int sweeps_per_pass);
/* kernel variant knobs: frames between consecutive sweeps (0 = automatic, else >= Q) and the variant of the strip
 * kernel: 0 = automatic (two lanes per task -- "pair-split" -- where the folded Q = 2 / Q = 4 updates allow it),
 * 1 = experimental tensor-memory producer/consumer warps, 2 = one thread per task, 10..15 = pair-split with
 * register-window mode (0..2) + 3 * explicit software pipelining (modes other than the default one exist only in
 * builds with -DLWSB_PAIR_EXPERIMENTS and otherwise select the default).  Results do not depend on it. */
int lwsb_set_variant(lwsb_ctx *ctx, int sweep_lag, int tensor_memory);
/* bins per block of the strip kernel: 0 = automatic, 8 (frames 2 blocks apart) or 4 (frames 3 blocks apart; Q <= 4).  Also env LWSB_STRIP_BLOCK. */
int lwsb_set_block_bins(lwsb_ctx *ctx, int bins);
/* time line of the strip kernel's work items -- one per (utterance, pass); the passes of an utterance run on different
 * clusters at the same time, each a few frames behind the previous one.  enable != 0 switches recording on for the
 * following lwsb_batch calls (also: env LWSB_STRIP_TRACE=1); returns the number of items of the last call copied out:
 * utt_pass[2i], utt_pass[2i+1] and stamps[8i..8i+7] = GPU globaltimer (ns) when the item was taken, its ring primed,
 * its last macro-step done, its frames written back; then cycles of strip 0: control lane waiting for the previous
 * pass's rows / polling its neighbours, compute warp 0 at work / waiting for the control warp. */
int lwsb_last_batch_trace(lwsb_ctx *ctx, int enable, int max_items, int *utt_pass, unsigned long long *stamps);
/* self-check of the branch-free sqrt / division of the pair-split kernels against the CUDA library functions on
 * n pseudo-random inputs: out4 = {sqrt samples in the fast range, of which differing, division samples, differing} */
int lwsb_debug_fast_math(lwsb_ctx *ctx, long long n, unsigned long long seed, unsigned long long *out4);
/* cycle accounting of cluster 0 in the last strip-kernel launch, summed over its CTAs: control lane {publish,
 * poll neighbours, TMA housekeeping}, compute warps {work, wait for the strip, wait for the neighbours, warps} */
int lwsb_last_batch_cycles(lwsb_ctx *ctx, unsigned long long *out13); /* + 6 consumer-phase counters (TM kernels) */
int lwsb_device_info(lwsb_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor, long long *hbm_bytes);
/* per-utterance statistics of the resident batch (mean / max of |S|, lws.pyx:240) */
int lwsb_get_stats(lwsb_ctx *ctx, double *mean_amp, double *max_amp);

/* Host-side mirrors of the device schedule, exported so that the CPU test-suite can replay
 * the exact dependency order without a GPU (tests/test_schedule.py).  No device needed.
 *  - lwsb_debug_terms: the linear stencil  acc = sum_e coef[e] * E[m+dr[e]][n+dk[e]]  that the
 *    kernels use for (weights, Q, L, fold, rframe, cframe, residue p); returns the count.
 *  - lwsb_debug_online_task: decode position j of the TF_RTISI_LA chain (lwslib.cpp:1432-1491)
 *    into (row, weight set, rframe, cframe, threshold index or -1). */
int lwsb_debug_terms(const double *wr, const double *wi, int Q, int L, int fold, int rframe, int cframe, int p,
                     int max_terms, int *dr, int *dk, double *cr, double *ci);
/*  - lwsb_debug_plan_strips: the plan the cluster strip kernel would use (same 15 numbers as
 *    lwsb_last_batch_plan) for a shape and a shared-memory / SM budget; returns 0 when the generic kernel serves it. */
int lwsb_debug_plan_strips(int Nreal, int Q, int L, int iterations, int maxT, int B, long long smem_limit,
                           int sm_count, int force_cluster, int max_sweeps, int force_block, int *out9);
/*  - lwsb_debug_work_items: the strip kernel's work list for B utterances with active_sweeps[b] sweeps that can move a
 *    bin: (utterance, pass) pairs, pass-major (utt_pass[2i], utt_pass[2i+1]); returns the number of items. */
int lwsb_debug_work_items(const int *active_sweeps, int B, int sweeps_per_pass, int max_items, int *utt_pass);
long long lwsb_debug_online_chain_length(int T, int iterations, int look_ahead);
int lwsb_debug_online_task(int T, int iterations, int look_ahead, int Q, long long j, int *row, int *which,
                           int *rframe, int *cframe, int *thr_index);

#ifdef __cplusplus
}
#endif
#endif /* LWS_B200_H_INCLUDED */

// api.cu -- the extern "C" boundary (include/lws_b200.h) and the host runtime behind it:
// context, device buffers, stencil cache, stage sequencing.  No torch types, no exceptions
// across the ABI, no CPU compute path: every entry point that does arithmetic on
// spectrograms launches CUDA kernels or fails.
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <chrono>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/lws_b200.h"
#include "kernels.h"
#include "lwsb_common.h"
#include "stencil.h"

using namespace lwsb;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

// Work list of the strip kernel: one item (utterance, pass) per pass of G sweeps an utterance needs, PASS-MAJOR, so that
// an item's producer -- the previous pass of the same utterance -- always has a smaller index: clusters take items in
// increasing order, hence the producer is finished or running whenever an item waits for it.  Returns the largest
// number of passes of any utterance.
// `group` > 0 lists the utterances in groups of that many, pass-major inside a group: the groups then finish one after the
// other and their results can leave the device while the later groups are still being worked on (lwsb_batch_lws).
int build_work_items(const int *nact, int B, int G, std::vector<int> &items, int group = 0)
{
    int max_pass = 0;
    items.clear();
    if (group <= 0 || group > B) group = B;
    for (int g0 = 0; g0 < B; g0 += group)
        for (int pass = 0, more = 1; more; ++pass) {
            more = 0;
            for (int b = g0; b < std::min(B, g0 + group); ++b)
                if (pass * G < nact[b]) { items.push_back(b); items.push_back(pass); more = 1; max_pass = std::max(max_pass, pass + 1); }
        }
    return max_pass;
}

std::mutex &strip_mutex(int device)
{
    static std::mutex m[64];
    return m[(unsigned)device % 64u];
}

// ---------------------------------------------------------------- host staging
// A caller of the reference hands pageable numpy arrays.  cudaMemcpyAsync from pageable memory goes through the
// driver's own staging at 5-6 GB/s (measured: 85 ms for the 495 MB of BASELINE configs[1], against 10 ms from pinned
// memory).  Pageable sources / destinations are therefore staged through two pinned buffers of the context: a few host
// threads copy chunk k + 1 while the DMA engine moves chunk k.
int host_threads()
{
    static int n = [] {
        if (const char *e = getenv("LWSB_HOST_THREADS")) return std::max(1, atoi(e));
        const unsigned hw = std::thread::hardware_concurrency();
        return (int)std::min(8u, std::max(1u, hw / 4));
    }();
    return n;
}

bool is_pageable(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}

} // namespace

struct lwsb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    cudaDeviceProp prop{};

    WeightSet w[3];
    DevBuf raww[3]; // device copy of wr | wi (2 * Q*Q*(L+1) doubles) followed by the int mask, reference layout

    // resident batch
    int B = 0, Nreal = 0, Q = 0, L = 0, P = 0, c0 = 0, maxT = 0;
    long long total_rows = 0, total_bins = 0;
    std::vector<int> T;
    std::vector<long long> rowbase, binbase;
    DevBuf E, A, row_max, leaf_tab, tab_of, leaf_sum, mean_amp, max_amp, dT, drowbase, stage, dptr, dthr, flags;
    long long leaf_stride = 0;
    std::vector<int> stat_tabs;                         // host copy of the summation trees, concatenated
    std::map<long long, std::pair<int, int>> stat_index; // array length -> (first leaf, leaf count)
    DevBuf fx, fS, fwin, fframes;          // stft / istft staging
    DevBuf rx, rS, rR, ry, rn;             // fused waveform -> waveform call and consistency: signals, spectrograms, norms
    DevBuf status;                         // watchdog word of the strip kernel
    DevBuf items, done, trace;             // strip kernel: (utterance, pass) work list, per-strip progress counters, optional time stamps
    bool want_trace = false;               // env LWSB_STRIP_TRACE=1 / lwsb_last_batch_trace
    int trace_items = 0;
    std::vector<int> trace_list;
    int last_kernel = 0;                   // 0 generic, 1 strips (introspection)
    int last_online_kernel = 0;            // 0 generic, 1 ring (one bin per step), 2 ring (two bins per step), 3 ring (two bins per step on two lanes)
    long long tune_smem = 0;               // tuning knobs (lwsb_set_tuning): shared-memory budget, cluster size,
    int tune_cluster = 0, tune_sweeps = 0; // sweeps per pass; 0 = automatic
    int tune_block = 0;                    // bins per block of the strip kernel: 0 automatic, 4 or 8 (env LWSB_STRIP_BLOCK, lwsb_set_block_bins)
    int tune_lag = 0;                      // frames between sweeps (env LWSB_STRIP_LAG only)
    int tune_tm = 0;                       // tensor-memory producer/consumer kernel (env LWSB_STRIP_TM=1, lwsb_set_tuning2);
                                           // off by default: measured slower than the single-warp pipeline
    StripPlan last_plan{};
    // streaming online_lws (lwsb_stream_*): the resident "batch" is one growing utterance
    bool stream_on = false, stream_ended = false;
    int stream_cap = 0, stream_kind = 0, stream_iters = 0, stream_LA = 0, stream_flags = 0;
    DevBuf sthr, sframes;
    std::map<int, DevBuf> twiddles;        // exp(-2 pi i j / N) tables by N
    std::vector<void *> hptr;

    cudaStream_t copy_stream = nullptr;    // results leave on it while the strip kernel still runs (lwsb_batch_lws)
    void *hdone = nullptr; size_t hdone_cap = 0; // pinned copy of the strip kernel's progress counters
    bool early_stored = false;             // the last lwsb_batch already copied the results out
    void *pin[2] = {nullptr, nullptr};     // pinned staging for pageable host buffers (lwsb_load / lwsb_store)
    size_t pin_cap = 0;
    cudaEvent_t pin_ev[2] = {nullptr, nullptr};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timing_valid = false;
    // per-stage device times since the last lwsb_load: 0 nofuture, 1 online, 2 batch (lwsb_last_stage_ms)
    cudaEvent_t evs[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
    bool stage_valid[3] = {false, false, false};
    int cur_stage = -1;
    long long launches = 0;
    // work of the last lwsb_batch call: bin-iterations asked for, bin-iterations of the sweeps that can move a bin
    // (threshold below max|S|: the others are dropped before launch), work items, passes
    long long last_work[4] = {0, 0, 0, 0};

    LwsbView view() const
    {
        LwsbView v;
        v.E = E.as<double2>(); v.A = A.as<double>();
        v.rowbase = drowbase.as<const long long>(); v.T = dT.as<const int>();
        v.mean_amp = mean_amp.as<const double>();
        v.P = P; v.c0 = c0; v.Nreal = Nreal; v.L = L; v.Q = Q; v.B = B;
        return v;
    }
    StatScratch scratch() const
    {
        return StatScratch{row_max.as<double>(), leaf_tab.as<const int>(), tab_of.as<const int2>(), leaf_sum.as<double>(), leaf_stride};
    }
    bool fractional() const { return w[LWSB_W].Qp != w[LWSB_W].Q; } // per-frequency weight rows: the *fractionalQ variants
    LwsbW devw(int which) const
    {
        const bool frac = w[which].Qp != w[which].Q;
        const size_t n = (size_t)(w[which].Qp + (frac ? 1 : 0)) * w[which].Q * (w[which].L + 1);
        const double *d = raww[which].as<const double>();
        return LwsbW{d, d + n, reinterpret_cast<const int *>(d + 2 * n), frac ? w[which].Qp : 0};
    }
};

namespace {

int fail(lwsb_ctx *c, int code, const std::string &msg)
{
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}

#define CU(ctx, call)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return fail(ctx, LWSB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));       \
    } while (0)

#define CHECK_CTX(c) if (!(c)) return LWSB_ERR_ARG

int use_device(lwsb_ctx *c) { CU(c, cudaSetDevice(c->device)); return LWSB_OK; }

int fold_for(int Q, int flags)
{
    if (flags & (LWSB_FORCE_ANYQ | LWSB_FRACTIONAL)) return LWSB_FOLD_ANY; // *fractionalQ = the anyQ formulas, other weight rows
    return Q == 2 ? LWSB_FOLD_Q2 : (Q == 4 ? LWSB_FOLD_Q4 : LWSB_FOLD_ANY); // lws.pyx:246-253
}

int upload_thresholds(lwsb_ctx *c, const double *thr, int n)
{
    CU(c, c->dthr.reserve(std::max(n, 1) * sizeof(double)));
    if (n > 0) CU(c, cudaMemcpyAsync(c->dthr.p, thr, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    return LWSB_OK;
}

int begin_compute(lwsb_ctx *c, int stage)
{
    CU(c, cudaEventRecord(c->ev0, c->stream));
    c->cur_stage = stage;
    if (stage >= 0) CU(c, cudaEventRecord(c->evs[stage][0], c->stream));
    return LWSB_OK;
}
int end_compute(lwsb_ctx *c)
{
    CU(c, cudaGetLastError());
    CU(c, cudaEventRecord(c->ev1, c->stream));
    if (c->cur_stage >= 0) {
        CU(c, cudaEventRecord(c->evs[c->cur_stage][1], c->stream));
        c->stage_valid[c->cur_stage] = true;
    }
    c->timing_valid = true;
    return LWSB_OK;
}

int check_resident(lwsb_ctx *c)
{
    if (c->B <= 0) return fail(c, LWSB_ERR_STATE, "no spectrograms loaded (call lwsb_load first)");
    return LWSB_OK;
}

constexpr size_t PIN_CHUNK = 24u << 20; // bytes per staging buffer

int ensure_pinned(lwsb_ctx *c)
{
    if (c->pin[0]) return LWSB_OK;
    for (int i = 0; i < 2; ++i) {
        CU(c, cudaHostAlloc(&c->pin[i], PIN_CHUNK, cudaHostAllocDefault));
        CU(c, cudaEventCreateWithFlags(&c->pin_ev[i], cudaEventDisableTiming));
    }
    c->pin_cap = PIN_CHUNK;
    return LWSB_OK;
}

// B host arrays (sizes[b] bytes each) <-> consecutive device ranges dev + offs[b].  Pinned host arrays are copied
// directly; pageable ones go through the two pinned buffers in chunks, host threads copying one chunk while the DMA
// engine moves the other.  to_device = false: the device data must already be complete on the stream.
int staged_copy(lwsb_ctx *c, char *dev, const std::vector<size_t> &offs, const std::vector<size_t> &sizes, void *const *host, bool to_device)
{
    const int B = (int)sizes.size();
    bool any_pageable = false;
    for (int b = 0; b < B && !any_pageable; ++b) any_pageable = is_pageable(host[b]);
    if (!any_pageable) {
        for (int b = 0; b < B; ++b)
            CU(c, to_device ? cudaMemcpyAsync(dev + offs[b], host[b], sizes[b], cudaMemcpyHostToDevice, c->stream)
                            : cudaMemcpyAsync(host[b], dev + offs[b], sizes[b], cudaMemcpyDeviceToHost, c->stream));
        return LWSB_OK;
    }
    if (int r = ensure_pinned(c)) return r;
    // pieces of at most PIN_CHUNK bytes: (array, offset inside it, bytes)
    struct Piece { int b; size_t off, n; };
    std::vector<std::vector<Piece>> chunks(1);
    size_t fill = 0;
    for (int b = 0; b < B; ++b)
        for (size_t o = 0; o < sizes[b];) {
            if (fill == PIN_CHUNK) { chunks.emplace_back(); fill = 0; }
            const size_t n = std::min(sizes[b] - o, PIN_CHUNK - fill);
            chunks.back().push_back(Piece{b, o, n});
            o += n; fill += n;
        }
    const int nc = (int)chunks.size();
    auto host_side = [&](int k) { // copy chunk k between the caller's arrays and pinned buffer k % 2, the chunk's bytes split over host threads
        char *pb = (char *)c->pin[k % 2];
        size_t total = 0;
        for (const Piece &pc : chunks[k]) total += pc.n;
        const int nt = total < (4u << 20) ? 1 : host_threads();
        auto work = [&, pb](size_t lo, size_t hi) { // bytes [lo, hi) of the chunk
            size_t at = 0;
            for (const Piece &pc : chunks[k]) {
                const size_t a = std::max(lo, at), b = std::min(hi, at + pc.n);
                if (a < b) {
                    char *hp = (char *)host[pc.b] + pc.off + (a - at);
                    if (to_device) memcpy(pb + a, hp, b - a);
                    else memcpy(hp, pb + a, b - a);
                }
                at += pc.n;
            }
        };
        if (nt == 1) { work(0, total); return; }
        std::vector<std::thread> th;
        const size_t per = ((total + nt - 1) / nt + 4095) & ~(size_t)4095;
        for (int i = 0; i < nt; ++i) {
            const size_t lo = std::min(total, per * i), hi = std::min(total, per * (i + 1));
            if (hi > lo) th.emplace_back(work, lo, hi);
        }
        for (auto &t : th) t.join();
    };
    auto dma = [&](int k) -> cudaError_t {
        char *pb = (char *)c->pin[k % 2];
        size_t at = 0;
        for (const Piece &pc : chunks[k]) {
            cudaError_t e = to_device ? cudaMemcpyAsync(dev + offs[pc.b] + pc.off, pb + at, pc.n, cudaMemcpyHostToDevice, c->stream)
                                      : cudaMemcpyAsync(pb + at, dev + offs[pc.b] + pc.off, pc.n, cudaMemcpyDeviceToHost, c->stream);
            if (e != cudaSuccess) return e;
            at += pc.n;
        }
        return cudaEventRecord(c->pin_ev[k % 2], c->stream);
    };
    if (to_device) {
        for (int k = 0; k < nc; ++k) {
            if (k >= 2) CU(c, cudaEventSynchronize(c->pin_ev[k % 2])); // the buffer's previous chunk has left
            host_side(k);
            CU(c, dma(k));
        }
        CU(c, cudaEventSynchronize(c->pin_ev[(nc - 1) % 2]));
        if (nc >= 2) CU(c, cudaEventSynchronize(c->pin_ev[(nc - 2) % 2]));
    } else {
        CU(c, dma(0));
        for (int k = 0; k < nc; ++k) {
            if (k + 1 < nc) {
                // buffer (k + 1) % 2 was emptied by host_side(k - 1): safe to refill while chunk k is copied out
                CU(c, dma(k + 1));
            }
            CU(c, cudaEventSynchronize(c->pin_ev[k % 2]));
            host_side(k);
        }
    }
    return LWSB_OK;
}

} // namespace

// ============================================================================ library / context
extern "C" int lwsb_version(void) { return 200; }

extern "C" int lwsb_strip_launch_mode(void) { return lwsb::strip_launch_mode(); }

extern "C" int lwsb_has_experiments(void)
{
#ifdef LWSB_EXPERIMENTS
    return 1;
#else
    return 0;
#endif
}

extern "C" const char *lwsb_last_error(const lwsb_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int lwsb_create(int device, void *stream, lwsb_ctx **out)
{
    if (!out) return fail(nullptr, LWSB_ERR_ARG, "out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, LWSB_ERR_CUDA,
                    std::string("no CUDA device available (this library has no CPU path): ") + cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(nullptr, LWSB_ERR_ARG, "device index out of range");
    lwsb_ctx *c = new lwsb_ctx();
    c->device = device;
    auto bail = [&](const char *what, cudaError_t err) {
        int code = fail(nullptr, LWSB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(err));
        delete c;
        return code;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
    if ((e = cudaGetDeviceProperties(&c->prop, device)) != cudaSuccess) return bail("cudaGetDeviceProperties", e);
    if (stream) { c->stream = (cudaStream_t)stream; c->own_stream = false; }
    else {
        if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
        c->own_stream = true;
    }
    if ((e = cudaEventCreate(&c->ev0)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreate(&c->ev1)) != cudaSuccess) return bail("cudaEventCreate", e);
    for (int i = 0; i < 3; ++i)
        for (int k = 0; k < 2; ++k)
            if ((e = cudaEventCreate(&c->evs[i][k])) != cudaSuccess) return bail("cudaEventCreate", e);
    if (const char *e1 = getenv("LWSB_STRIP_SMEM")) c->tune_smem = atoll(e1);
    if (const char *e2 = getenv("LWSB_STRIP_CLUSTER")) c->tune_cluster = atoi(e2);
    if (const char *e3 = getenv("LWSB_STRIP_SWEEPS")) c->tune_sweeps = atoi(e3);
    if (const char *e4 = getenv("LWSB_STRIP_LAG")) c->tune_lag = atoi(e4);
    if (const char *e5 = getenv("LWSB_STRIP_TM")) c->tune_tm = atoi(e5);
    if (const char *e6 = getenv("LWSB_STRIP_TRACE")) c->want_trace = atoi(e6) != 0;
#ifdef LWSB_EXPERIMENTS
    if (const char *e7 = getenv("LWSB_STRIP_BLOCK")) c->tune_block = atoi(e7);
#endif
    *out = c;
    return LWSB_OK;
}

extern "C" int lwsb_destroy(lwsb_ctx *c)
{
    CHECK_CTX(c);
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (DevBuf *b : {&c->sthr, &c->sframes, &c->rx, &c->rS, &c->rR, &c->ry, &c->rn, &c->items, &c->done, &c->trace, &c->E, &c->A, &c->row_max, &c->leaf_tab, &c->tab_of, &c->leaf_sum, &c->mean_amp, &c->max_amp, &c->dT,
                      &c->drowbase, &c->stage, &c->dptr, &c->dthr, &c->flags, &c->fx, &c->fS, &c->fwin, &c->fframes, &c->status})
        b->release();
    for (auto &kv : c->twiddles) kv.second.release();
    for (int i = 0; i < 3; ++i) c->raww[i].release();
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->hdone) cudaFreeHost(c->hdone);
    for (int i = 0; i < 2; ++i) {
        if (c->pin[i]) cudaFreeHost(c->pin[i]);
        if (c->pin_ev[i]) cudaEventDestroy(c->pin_ev[i]);
    }
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    for (int i = 0; i < 3; ++i)
        for (int k = 0; k < 2; ++k)
            if (c->evs[i][k]) cudaEventDestroy(c->evs[i][k]);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return LWSB_OK;
}

// pinned (page-locked) host memory for callers that want the DMA engine to reach their buffers directly
extern "C" int lwsb_host_alloc(unsigned long long bytes, void **out)
{
    if (!out) return LWSB_ERR_ARG;
    *out = nullptr;
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(nullptr, e == cudaErrorMemoryAllocation ? LWSB_ERR_NOMEM : LWSB_ERR_CUDA, cudaGetErrorString(e)); }
    return LWSB_OK;
}

extern "C" int lwsb_host_free(void *p)
{
    if (p && cudaFreeHost(p) != cudaSuccess) { cudaGetLastError(); return LWSB_ERR_CUDA; }
    return LWSB_OK;
}

extern "C" int lwsb_sync(lwsb_ctx *c)
{
    CHECK_CTX(c);
    if (int r = use_device(c)) return r;
    CU(c, cudaStreamSynchronize(c->stream));
    return LWSB_OK;
}

// ============================================================================ weights
extern "C" int lwsb_set_weights(lwsb_ctx *c, int which, const double *wr, const double *wi, int Qprime, int Q, int L)
{
    CHECK_CTX(c);
    if (which < 0 || which > 2 || !wr || !wi || Q < 1 || L < 0) return fail(c, LWSB_ERR_ARG, "bad weight arguments");
    if (Qprime < 1) return fail(c, LWSB_ERR_ARG, "bad weight arguments");
    if (int r = use_device(c)) return r;
    WeightSet &w = c->w[which];
    w.Q = Q; w.L = L; w.Qp = Qprime;
    // Qprime != Q: one weight row per FFT bin (lws.pyx:166-169), the reference's *fractionalQ variants.  They index row
    // Qprime at the DC bin, one past the table (lwslib.cpp:408); the device table gets that row, zero with the mask clear.
    const bool frac = Qprime != Q;
    const size_t n_in = (size_t)Qprime * Q * (L + 1), n = (size_t)(Qprime + (frac ? 1 : 0)) * Q * (L + 1);
    w.wr.assign(wr, wr + n_in);
    w.wi.assign(wi, wi + n_in);
    w.wr.resize(n, 0.0); w.wi.resize(n, 0.0);
    std::vector<int> wf(n);
    for (size_t i = 0; i < n; ++i) wf[i] = std::hypot(w.wr[i], w.wi[i]) > 1.0e-12 ? 1 : 0; // lws.pyx:231-232
    CU(c, c->raww[which].reserve(2 * n * sizeof(double) + n * sizeof(int)));
    char *d = c->raww[which].as<char>();
    CU(c, cudaMemcpyAsync(d, w.wr.data(), n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(d + n * sizeof(double), w.wi.data(), n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(d + 2 * n * sizeof(double), wf.data(), n * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return LWSB_OK;
}

extern "C" int lwsb_create_weights(const double *awin, const double *swin, int T, int fshift, int L,
                                   int use_summarized_weights, double *wr, double *wi, int *Qprime_out, int *Q_out)
{
    // lws.pyx:160-181
    if (!awin || !swin || T < 1 || fshift < 1 || L < 0) return LWSB_ERR_ARG;
    const int Q = (T + fshift - 1) / fshift;
    const double Qf = (double)T / (double)fshift;
    const int Qp = (T % fshift == 0 && use_summarized_weights) ? Q : T;
    if (Qprime_out) *Qprime_out = Qp;
    if (Q_out) *Q_out = Q;
    if (!wr || !wi) return LWSB_OK; // shape query
    const double twopi = 2.0 * M_PI;
    std::vector<double> w0r((size_t)(L + 1) * Q), w0i((size_t)(L + 1) * Q);
    for (int k = 0; k <= L; ++k)
        for (int q = 0; q < Q; ++q) {
            double sr = 0.0, si = 0.0;
            for (int t = 0; t < T - q * fshift; ++t) {
                const double wp = awin[t] * swin[t + q * fshift] / T;
                const double ang = -twopi * k * t / T;
                sr += std::cos(ang) * wp;
                si += std::sin(ang) * wp;
            }
            const double ang = -twopi * k * q / Qf;
            const double cr = std::cos(ang), ci = std::sin(ang);
            w0r[(size_t)k * Q + q] = sr * cr - si * ci;
            w0i[(size_t)k * Q + q] = sr * ci + si * cr;
        }
    w0r[0] -= 1.0;
    for (int n = 0; n < Qp; ++n)
        for (int q = 0; q < Q; ++q) {
            const double ang = twopi * n * q / Qf;
            const double cr = std::cos(ang), ci = std::sin(ang);
            for (int k = 0; k <= L; ++k) {
                const double a = w0r[(size_t)k * Q + q], b = w0i[(size_t)k * Q + q];
                wr[((size_t)n * Q + q) * (L + 1) + k] = a * cr - b * ci;
                wi[((size_t)n * Q + q) * (L + 1) + k] = a * ci + b * cr;
            }
        }
    return LWSB_OK;
}

// ============================================================================ staged interface
extern "C" int lwsb_load(lwsb_ctx *c, const void *const *S_in, const int *T, int B, int Nreal, int kind, int where)
{
    CHECK_CTX(c);
    if (!S_in || !T || B < 1 || Nreal < 1 || (kind != LWSB_C128 && kind != LWSB_F64) ||
        (where != LWSB_HOST && where != LWSB_DEVICE))
        return fail(c, LWSB_ERR_ARG, "bad lwsb_load arguments");
    if (Nreal % 2 == 0)
        return fail(c, LWSB_ERR_EVEN_NREAL, "Please only include non-negative frequencies in the input spectrogram.");
    if (!c->w[LWSB_W].valid()) return fail(c, LWSB_ERR_STATE, "set LWSB_W before loading spectrograms");
    if (Nreal <= c->w[LWSB_W].L)
        // extspec's mirror columns (lws.pyx:153-154) would reach outside the real bins: the reference picks up
        // zeros / already mirrored cells there, an fsize <= 2L corner nobody uses -- refused rather than restated
        return fail(c, LWSB_ERR_UNSUPPORTED, "spectrum narrower than the stencil reach (Nreal <= L) is not supported");
    if (c->fractional() && c->w[LWSB_W].Qp != 2 * (Nreal - 1))
        return fail(c, LWSB_ERR_ARG, "per-frequency weights need one row per FFT bin: Qprime == 2 * (Nreal - 1)");
    if (int r = use_device(c)) return r;
    c->B = 0; // invalid until everything below succeeded
    c->stream_on = false;
    c->stage_valid[0] = c->stage_valid[1] = c->stage_valid[2] = false;
    const int Q = c->w[LWSB_W].Q, L = c->w[LWSB_W].L;
    const int Np = Nreal + 2 * L;
    const int coff = (4 - L % 4) % 4; // bin 0 lands on a 64-byte boundary
    const int P = (std::max(coff + Np, strips_min_pitch(Nreal, coff + L)) + 3) / 4 * 4; // strip kernel reads whole blocks
    std::vector<long long> rowbase(B), binbase(B);
    long long rows = 0, bins = 0;
    int maxT = 0;
    for (int b = 0; b < B; ++b) {
        if (T[b] < 1 || !S_in[b]) return fail(c, LWSB_ERR_ARG, "empty utterance in batch");
        rowbase[b] = rows; binbase[b] = bins;
        rows += T[b] + 2 * (Q - 1);
        bins += (long long)T[b] * Nreal;
        maxT = std::max(maxT, T[b]);
    }
    CU(c, c->E.reserve((size_t)rows * P * sizeof(double2)));
    CU(c, c->A.reserve((size_t)rows * P * sizeof(double)));
    CU(c, c->row_max.reserve((size_t)rows * sizeof(double)));
    // numpy's summation tree per distinct array length (cached for the life of the context)
    long long stride = 0;
    std::vector<int2> tab_of(B);
    for (int b = 0; b < B; ++b) {
        const long long n = (long long)T[b] * Nreal;
        auto it = c->stat_index.find(n);
        if (it == c->stat_index.end()) {
            const int first = (int)(c->stat_tabs.size() / 3);
            stat_tree(n, c->stat_tabs);
            it = c->stat_index.emplace(n, std::make_pair(first, (int)(c->stat_tabs.size() / 3) - first)).first;
        }
        tab_of[b] = make_int2(it->second.first, it->second.second);
        stride = std::max(stride, (long long)it->second.second);
    }
    CU(c, c->leaf_tab.reserve(std::max<size_t>(c->stat_tabs.size(), 3) * sizeof(int)));
    // (re)sent on every load: a few tens of KB, and reserve() may have moved the buffer
    CU(c, cudaMemcpyAsync(c->leaf_tab.p, c->stat_tabs.data(), c->stat_tabs.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CU(c, c->tab_of.reserve(B * sizeof(int2)));
    CU(c, cudaMemcpyAsync(c->tab_of.p, tab_of.data(), B * sizeof(int2), cudaMemcpyHostToDevice, c->stream));
    CU(c, c->leaf_sum.reserve((size_t)B * stride * sizeof(double)));
    c->leaf_stride = stride;
    CU(c, c->mean_amp.reserve(B * sizeof(double)));
    CU(c, c->max_amp.reserve(B * sizeof(double)));
    CU(c, c->dT.reserve(B * sizeof(int)));
    CU(c, c->drowbase.reserve(B * sizeof(long long)));
    CU(c, c->dptr.reserve(B * sizeof(void *)));
    CU(c, cudaMemcpyAsync(c->dT.p, T, B * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(c->drowbase.p, rowbase.data(), B * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
    const size_t esz = kind == LWSB_C128 ? sizeof(double2) : sizeof(double);
    c->hptr.resize(B);
    if (where == LWSB_HOST) {
        CU(c, c->stage.reserve((size_t)bins * sizeof(double2))); // sized for the complex128 output as well
        std::vector<size_t> offs(B), sizes(B);
        for (int b = 0; b < B; ++b) {
            offs[b] = (size_t)binbase[b] * esz; sizes[b] = (size_t)T[b] * Nreal * esz;
            c->hptr[b] = c->stage.as<char>() + offs[b];
        }
        if (int r = staged_copy(c, c->stage.as<char>(), offs, sizes, const_cast<void *const *>(S_in), true)) return r;
    } else {
        for (int b = 0; b < B; ++b) c->hptr[b] = const_cast<void *>(S_in[b]);
    }
    CU(c, cudaMemcpyAsync(c->dptr.p, c->hptr.data(), B * sizeof(void *), cudaMemcpyHostToDevice, c->stream));
    c->T.assign(T, T + B);
    c->rowbase = rowbase; c->binbase = binbase;
    c->Nreal = Nreal; c->Q = Q; c->L = L; c->P = P; c->c0 = coff + L; c->maxT = maxT;
    c->total_rows = rows; c->total_bins = bins;
    c->B = B;
    LwsbView v = c->view();
    launch_extend(v, kind, c->dptr.as<const void *const>(), c->scratch(), c->mean_amp.as<double>(),
                  c->max_amp.as<double>(), maxT + 2 * (Q - 1), c->stream);
    c->launches += 2;
    CU(c, cudaGetLastError());
    // rowbase / hptr host vectors are read by the async copies above: make them safe to reuse
    CU(c, cudaStreamSynchronize(c->stream));
    return LWSB_OK;
}

extern "C" int lwsb_store(lwsb_ctx *c, void *const *S_out, int where)
{
    CHECK_CTX(c);
    if (!S_out || (where != LWSB_HOST && where != LWSB_DEVICE)) return fail(c, LWSB_ERR_ARG, "bad lwsb_store arguments");
    if (int r = check_resident(c)) return r;
    if (int r = use_device(c)) return r;
    const int B = c->B;
    if (where == LWSB_HOST) {
        CU(c, c->stage.reserve((size_t)c->total_bins * sizeof(double2)));
        for (int b = 0; b < B; ++b) c->hptr[b] = c->stage.as<double2>() + c->binbase[b];
    } else {
        for (int b = 0; b < B; ++b) {
            if (!S_out[b]) return fail(c, LWSB_ERR_ARG, "NULL output pointer");
            c->hptr[b] = S_out[b];
        }
    }
    CU(c, cudaMemcpyAsync(c->dptr.p, c->hptr.data(), B * sizeof(void *), cudaMemcpyHostToDevice, c->stream));
    launch_crop(c->view(), c->dptr.as<void *const>(), c->maxT, c->stream);
    c->launches += 1;
    CU(c, cudaGetLastError());
    if (where == LWSB_HOST) {
        std::vector<size_t> offs(B), sizes(B);
        for (int b = 0; b < B; ++b) {
            if (!S_out[b]) return fail(c, LWSB_ERR_ARG, "NULL output pointer");
            offs[b] = (size_t)c->binbase[b] * sizeof(double2); sizes[b] = (size_t)c->T[b] * c->Nreal * sizeof(double2);
        }
        if (int r = staged_copy(c, c->stage.as<char>(), offs, sizes, S_out, false)) return r;
    }
    CU(c, cudaStreamSynchronize(c->stream));
    return LWSB_OK;
}

// early_out != NULL (one-shot call, page-locked host result buffers): the utterances are listed in groups, and as soon as
// the last pass of an utterance has written all its frames back (the strips' progress counters, read through a side
// stream) its rows are copied to early_out[u] by the DMA engine, while the clusters work on the later groups.
static int batch_impl(lwsb_ctx *c, const double *thresholds, int iterations, int flags, void *const *early_out)
{
    CHECK_CTX(c);
    c->early_stored = false;
    if (iterations < 0 || (iterations > 0 && !thresholds)) return fail(c, LWSB_ERR_ARG, "bad thresholds");
    if (int r = check_resident(c)) return r;
    if (iterations == 0) return LWSB_OK; // lws.pyx:219-220
    if (int r = use_device(c)) return r;
    if (int r = upload_thresholds(c, thresholds, iterations)) return r;
    if (c->fractional()) flags |= LWSB_FRACTIONAL | LWSB_FORCE_GENERIC; // lws.pyx:246-247: Q != Qprime
    if (flags & LWSB_FRACTIONAL) {
        if (!c->fractional()) return fail(c, LWSB_ERR_STATE, "the per-frequency update (use_simplifications=False) needs per-frequency weights");
        flags |= LWSB_FORCE_GENERIC;
    }
    const int fold = fold_for(c->Q, flags);
    // sweeps whose threshold is not below max|S| cannot move a bin (lwslib.cpp:295-296): the strip kernel drops
    // them, and the plan is sized for the number that remain (largest over the batch)
    int active = iterations;
    double avg_active = 0.0;
    std::vector<int> nact; // sweeps per utterance that can move a bin
    if (!(flags & LWSB_FORCE_GENERIC)) {
        std::vector<double> mean(c->B), mx(c->B);
        CU(c, cudaMemcpyAsync(mean.data(), c->mean_amp.p, c->B * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaMemcpyAsync(mx.data(), c->max_amp.p, c->B * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
        active = 0;
        nact.resize(c->B);
        for (int b = 0; b < c->B; ++b) {
            int n = 0;
            for (int i = 0; i < iterations; ++i) n += (thresholds[i] * mean[b] < mx[b]) ? 1 : 0;
            nact[b] = n;
            active = std::max(active, n);
            avg_active += (double)n / c->B;
        }
        active = std::max(active, 1);
    }
    StripPlan pl;
    const bool strips = !(flags & LWSB_FORCE_GENERIC) &&
                        plan_strips(c->Nreal, c->Q, c->L, active, c->maxT, c->B,
                                    c->tune_smem > 0 ? std::min((size_t)c->tune_smem, c->prop.sharedMemPerBlockOptin)
                                                     : c->prop.sharedMemPerBlockOptin,
                                    c->prop.multiProcessorCount, &pl, c->tune_cluster, c->tune_sweeps, c->tune_lag, c->tune_tm, fold,
                                    c->tune_block, avg_active) &&
                        c->P >= strips_min_pitch(c->Nreal, c->c0);
    if (strips) {
        CU(c, c->status.reserve(256));
        CU(c, cudaMemsetAsync(c->status.p, 0, 256, c->stream));
        // work list: one item per (utterance, pass of pl.G sweeps), pass-major, so that the passes of one utterance
        // run on different clusters at the same time, each a few frames behind the previous one
        std::vector<int> items;
        if (early_out && getenv("LWSB_EARLY_STORE") && atoi(getenv("LWSB_EARLY_STORE")) == 0) early_out = nullptr;
        static const int n_groups = [] { const char *e = getenv("LWSB_EARLY_GROUPS"); return e ? std::max(1, atoi(e)) : 4; }();
        const int group = (early_out && c->B >= 32) ? (c->B + n_groups - 1) / n_groups : 0;
        const int max_pass = build_work_items(nact.data(), c->B, pl.G, items, group);
        const int n_items = (int)(items.size() / 2);
        c->last_work[0] = c->total_bins * iterations; c->last_work[1] = 0; c->last_work[2] = n_items; c->last_work[3] = max_pass;
        for (int b = 0; b < c->B; ++b) c->last_work[1] += (long long)nact[b] * c->T[b] * c->Nreal;
        const size_t done_bytes = (size_t)c->B * std::max(max_pass, 1) * STRIP_MAX_CLUSTER * sizeof(unsigned);
        if (n_items > 0) {
            CU(c, c->items.reserve(items.size() * sizeof(int)));
            CU(c, c->done.reserve(done_bytes));
            CU(c, cudaMemcpyAsync(c->items.p, items.data(), items.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
            CU(c, cudaMemsetAsync(c->done.p, 0, done_bytes, c->stream));
        }
        c->trace_items = 0;
        if (c->want_trace && n_items > 0) {
            CU(c, c->trace.reserve((size_t)n_items * 8 * sizeof(unsigned long long)));
            CU(c, cudaMemsetAsync(c->trace.p, 0, (size_t)n_items * 8 * sizeof(unsigned long long), c->stream));
            c->trace_items = n_items;
            c->trace_list = items;
        }
        // One strip kernel at a time per device: its clusters wait for one another (a pass reads what the previous
        // pass of the utterance wrote), which is dead-lock free only while every cluster of the launch is resident.
        std::lock_guard<std::mutex> strip_guard(strip_mutex(c->device));
        if (int r = begin_compute(c, 2)) return r;
        if (n_items > 0) {
            CU(c, launch_batch_strips(c->view(), c->w[LWSB_W].wr.data(), c->w[LWSB_W].wi.data(), fold,
                                      c->dthr.as<const double>(), c->max_amp.as<const double>(), iterations, pl,
                                      c->status.as<unsigned>(), c->items.as<const int>(), n_items, max_pass, c->done.as<unsigned>(),
                                      c->want_trace ? c->trace.as<unsigned long long>() : nullptr, c->stream));
            c->launches += 1;
        }
        c->last_kernel = 1; c->last_plan = pl;
        if (int r = end_compute(c)) return r;
        if (early_out && n_items > 0) {
            if (!c->copy_stream) CU(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
            if (c->hdone_cap < done_bytes) {
                if (c->hdone) cudaFreeHost(c->hdone);
                c->hdone = nullptr; c->hdone_cap = 0;
                CU(c, cudaHostAlloc(&c->hdone, done_bytes, cudaHostAllocDefault));
                c->hdone_cap = done_bytes;
            }
            CU(c, cudaStreamWaitEvent(c->copy_stream, c->evs[2][0], 0)); // everything before the launch (the load) is complete
            const unsigned *hd = reinterpret_cast<const unsigned *>(c->hdone);
            std::vector<char> copied(c->B, 0);
            int ncopied = 0;
            const size_t rowb = (size_t)c->Nreal * sizeof(double2);
            while (ncopied < c->B) {
                const bool finished = cudaEventQuery(c->ev1) == cudaSuccess;
                if (!finished) {
                    CU(c, cudaMemcpyAsync(c->hdone, c->done.p, done_bytes, cudaMemcpyDeviceToHost, c->copy_stream));
                    CU(c, cudaStreamSynchronize(c->copy_stream));
                }
                for (int u = 0; u < c->B; ++u) {
                    if (copied[u]) continue;
                    bool ready = finished || nact[u] == 0;
                    if (!ready) {
                        const int lastp = (nact[u] + pl.G - 1) / pl.G - 1;
                        ready = true;
                        for (int k = 0; k < pl.C && ready; ++k)
                            ready = hd[((size_t)u * max_pass + lastp) * STRIP_MAX_CLUSTER + k] >= (unsigned)c->T[u];
                    }
                    if (!ready) continue;
                    const double2 *src = c->E.as<double2>() + (c->rowbase[u] + c->Q - 1) * (long long)c->P + c->c0;
                    CU(c, cudaMemcpy2DAsync(early_out[u], rowb, src, (size_t)c->P * sizeof(double2), rowb, c->T[u], cudaMemcpyDeviceToHost,
                                            c->copy_stream));
                    copied[u] = 1; ++ncopied;
                }
                if (!finished && ncopied < c->B) std::this_thread::sleep_for(std::chrono::microseconds(300));
            }
            CU(c, cudaStreamSynchronize(c->copy_stream));
            c->early_stored = true;
        }
        unsigned st = 0;
        CU(c, cudaMemcpyAsync(&st, c->status.p, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
        if (st != 0) {
            char msg[96];
            snprintf(msg, sizeof msg, "strip kernel watchdog fired (code 0x%x): results are invalid", st);
            return fail(c, LWSB_ERR_CUDA, msg);
        }
        return LWSB_OK;
    }
    c->last_work[0] = c->last_work[1] = c->total_bins * iterations; c->last_work[2] = c->B; c->last_work[3] = 1;
    if (int r = begin_compute(c, 2)) return r;
    launch_sweeps_generic(c->view(), c->devw(LWSB_W), fold, c->Q, 1, c->dthr.as<const double>(), iterations, c->stream);
    c->launches += 1;
    c->last_kernel = 0;
    return end_compute(c);
}

extern "C" int lwsb_batch(lwsb_ctx *c, const double *thresholds, int iterations, int flags)
{
    return batch_impl(c, thresholds, iterations, flags, nullptr);
}

extern "C" int lwsb_nofuture(lwsb_ctx *c, int which, const double *thresholds, int iterations, int flags)
{
    CHECK_CTX(c);
    if (which < 0 || which > 2 || iterations < 0 || (iterations > 0 && !thresholds))
        return fail(c, LWSB_ERR_ARG, "bad lwsb_nofuture arguments");
    if (int r = check_resident(c)) return r;
    if (iterations == 0) return LWSB_OK; // lws.pyx:272-273
    if (!c->w[which].valid() || c->w[which].Q != c->Q || c->w[which].L != c->L || c->w[which].Qp != c->w[LWSB_W].Qp)
        return fail(c, LWSB_ERR_STATE, "no-future weight set missing or of a different shape than LWSB_W");
    if (int r = use_device(c)) return r;
    if (int r = upload_thresholds(c, thresholds, iterations)) return r;
    if (c->fractional()) flags |= LWSB_FRACTIONAL; // lws.pyx:299-300
    const int fold = fold_for(c->Q, flags);
    if (int r = begin_compute(c, 0)) return r;
    if (fold == LWSB_FOLD_Q4) // NoFuture_LWSQ4, reproduced as written (lwslib.cpp:538-617)
        launch_nofuture_q4(c->view(), c->devw(which), c->dthr.as<const double>(), iterations, c->stream);
    else
        launch_sweeps_generic(c->view(), c->devw(which), fold, 1, 0, c->dthr.as<const double>(), iterations, c->stream);
    c->launches += 1;
    return end_compute(c);
}

extern "C" int lwsb_online(lwsb_ctx *c, const double *thresholds, int iterations, int look_ahead, int flags)
{
    CHECK_CTX(c);
    if (iterations < 0 || (iterations > 0 && !thresholds) || look_ahead < 0)
        return fail(c, LWSB_ERR_ARG, "bad lwsb_online arguments");
    if (int r = check_resident(c)) return r;
    if (iterations == 0) return LWSB_OK; // lws.pyx:332-333
    for (int i = 1; i < 3; ++i)
        if (!c->w[i].valid() || c->w[i].Q != c->Q || c->w[i].L != c->L || c->w[i].Qp != c->w[LWSB_W].Qp)
            return fail(c, LWSB_ERR_STATE, "online mode needs W, W_ai and W_af of the same shape");
    if (c->fractional()) flags |= LWSB_FRACTIONAL | LWSB_FORCE_GENERIC; // lwslib.cpp:1441: !use_summarized_weights
    if (c->Nreal > online_generic_max_nreal(c->L))
        return fail(c, LWSB_ERR_UNSUPPORTED, "spectrum too wide for the online kernel");
    if (int r = use_device(c)) return r;
    if (int r = upload_thresholds(c, thresholds, iterations)) return r;
    const LwsbW w3[3] = {c->devw(LWSB_W), c->devw(LWSB_W_AI), c->devw(LWSB_W_AF)};
    if (int r = begin_compute(c, 1)) return r;
    bool ring = false;
    c->last_online_kernel = 0;
    if (!(flags & LWSB_FORCE_GENERIC)) {
        cudaError_t e = cudaSuccess;
        CU(c, c->status.reserve(256));
        CU(c, cudaMemsetAsync(c->status.p, 0, 256, c->stream));
        const double *wrh[3] = {c->w[0].wr.data(), c->w[1].wr.data(), c->w[2].wr.data()};
        const double *wih[3] = {c->w[0].wi.data(), c->w[1].wi.data(), c->w[2].wi.data()};
        ring = launch_online_ring(c->view(), wrh, wih, fold_for(c->Q, flags), c->dthr.as<const double>(), iterations,
                                  look_ahead, c->T.data(), c->prop.sharedMemPerBlockOptin, c->status.as<unsigned>(),
                                  c->stream, &e, &c->last_online_kernel);
        CU(c, e);
    }
    if (!ring)
        launch_online_generic(c->view(), w3, fold_for(c->Q, flags), c->dthr.as<const double>(), iterations, look_ahead,
                              c->stream);
    c->launches += 1;
    if (int r = end_compute(c)) return r;
    if (ring) {
        unsigned st = 0;
        CU(c, cudaMemcpyAsync(&st, c->status.p, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
        if (st != 0) {
            char msg[112];
            snprintf(msg, sizeof msg, "online kernel: internal schedule error (code 0x%x): results are invalid", st);
            return fail(c, LWSB_ERR_CUDA, msg);
        }
    }
    return LWSB_OK;
}


// ============================================================================ streaming online_lws
// TF_RTISI_LA is causal (lwslib.cpp:1432-1491): the row updates of frame m -- its initial estimate, then `iterations` times
// [the look-ahead frames m-LA .. m-1, frame m] -- read nothing beyond frame m, so they can run as soon as frame m has
// arrived, and frame m - LA is final once they have.  A stream is one growing utterance resident in HBM; every push runs
// the chain positions of the new frames (the generic chain kernel on a range of positions).  Given the same mean amplitude
// (the reference scales the thresholds by the mean of |S| over the WHOLE utterance, lws.pyx:360-361: a stream needs it --
// or an estimate -- up front) the frames are bit-identical to online_lws on the complete spectrogram.
extern "C" int lwsb_stream_begin(lwsb_ctx *c, int Nreal, int max_frames, int kind, double mean_amp, const double *thresholds,
                                 int iterations, int look_ahead, int flags)
{
    CHECK_CTX(c);
    if (Nreal < 1 || max_frames < 1 || (kind != LWSB_C128 && kind != LWSB_F64) || iterations < 1 || !thresholds || look_ahead < 0)
        return fail(c, LWSB_ERR_ARG, "bad lwsb_stream_begin arguments");
    if (Nreal % 2 == 0)
        return fail(c, LWSB_ERR_EVEN_NREAL, "Please only include non-negative frequencies in the input spectrogram.");
    for (int i = 0; i < 3; ++i)
        if (!c->w[i].valid() || c->w[i].Q != c->w[0].Q || c->w[i].L != c->w[0].L || c->w[i].Qp != c->w[0].Qp)
            return fail(c, LWSB_ERR_STATE, "online mode needs W, W_ai and W_af of the same shape");
    if (Nreal <= c->w[0].L) return fail(c, LWSB_ERR_UNSUPPORTED, "spectrum narrower than the stencil reach (Nreal <= L) is not supported");
    if (c->fractional() && c->w[0].Qp != 2 * (Nreal - 1))
        return fail(c, LWSB_ERR_ARG, "per-frequency weights need one row per FFT bin: Qprime == 2 * (Nreal - 1)");
    if (Nreal > online_generic_max_nreal(c->w[0].L)) return fail(c, LWSB_ERR_UNSUPPORTED, "spectrum too wide for the online kernel");
    if (int r = use_device(c)) return r;
    c->B = 0; c->stream_on = false;
    const int Q = c->w[0].Q, L = c->w[0].L;
    const int Np = Nreal + 2 * L, coff = (4 - L % 4) % 4;
    const int P = (std::max(coff + Np, strips_min_pitch(Nreal, coff + L)) + 3) / 4 * 4;
    const long long rows = (long long)max_frames + 2 * (Q - 1);
    CU(c, c->E.reserve((size_t)rows * P * sizeof(double2)));
    CU(c, c->A.reserve((size_t)rows * P * sizeof(double)));
    CU(c, c->mean_amp.reserve(sizeof(double)));
    CU(c, c->max_amp.reserve(sizeof(double)));
    CU(c, c->dT.reserve(sizeof(int)));
    CU(c, c->drowbase.reserve(sizeof(long long)));
    CU(c, c->sthr.reserve((size_t)iterations * sizeof(double)));
    const long long zero = 0;
    const int t0 = 0;
    CU(c, cudaMemcpyAsync(c->drowbase.p, &zero, sizeof zero, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(c->dT.p, &t0, sizeof t0, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(c->mean_amp.p, &mean_amp, sizeof mean_amp, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(c->sthr.p, thresholds, (size_t)iterations * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    c->T.assign(1, 0); c->rowbase.assign(1, 0); c->binbase.assign(1, 0);
    c->Nreal = Nreal; c->Q = Q; c->L = L; c->P = P; c->c0 = coff + L; c->maxT = 0;
    c->total_rows = rows; c->total_bins = 0;
    c->B = 1;
    c->stream_on = true; c->stream_ended = false;
    c->stream_cap = max_frames; c->stream_kind = kind; c->stream_iters = iterations; c->stream_LA = look_ahead;
    c->stream_flags = flags | (c->fractional() ? LWSB_FRACTIONAL : 0);
    return LWSB_OK;
}

extern "C" int lwsb_stream_push(lwsb_ctx *c, const void *frames, int nframes, int where)
{
    CHECK_CTX(c);
    if (!c->stream_on || c->stream_ended) return fail(c, LWSB_ERR_STATE, "no open stream (call lwsb_stream_begin first)");
    if (!frames || nframes < 1 || (where != LWSB_HOST && where != LWSB_DEVICE)) return fail(c, LWSB_ERR_ARG, "bad lwsb_stream_push arguments");
    const int T0 = c->T[0], T1 = T0 + nframes;
    if (T1 > c->stream_cap) return fail(c, LWSB_ERR_ARG, "stream is longer than the max_frames it was opened with");
    if (int r = use_device(c)) return r;
    const size_t esz = c->stream_kind == LWSB_C128 ? sizeof(double2) : sizeof(double);
    const size_t bytes = (size_t)nframes * c->Nreal * esz;
    const void *src = frames;
    if (where == LWSB_HOST) {
        CU(c, c->sframes.reserve(bytes));
        std::vector<size_t> offs(1, 0), sizes(1, bytes);
        void *hp[1] = {const_cast<void *>(frames)};
        if (int r = staged_copy(c, c->sframes.as<char>(), offs, sizes, hp, true)) return r;
        src = c->sframes.p;
    }
    launch_stream_extend(c->view(), c->stream_kind, src, T0, nframes, c->stream);
    CU(c, cudaMemcpyAsync(c->dT.p, &T1, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    const LwsbW w3[3] = {c->devw(LWSB_W), c->devw(LWSB_W_AI), c->devw(LWSB_W_AF)};
    if (int r = begin_compute(c, 1)) return r;
    launch_online_generic(c->view(), w3, fold_for(c->Q, c->stream_flags), c->sthr.as<const double>(), c->stream_iters, c->stream_LA, c->stream,
                          lwsb_online_chain_len(T0, c->stream_iters, c->stream_LA), lwsb_online_chain_len(T1, c->stream_iters, c->stream_LA));
    c->launches += 2;
    if (int r = end_compute(c)) return r;
    CU(c, cudaStreamSynchronize(c->stream)); // T1 (a stack variable) has been read; the caller may reuse `frames`
    c->T[0] = T1; c->maxT = T1; c->total_bins = (long long)T1 * c->Nreal;
    return LWSB_OK;
}

extern "C" int lwsb_stream_frames(const lwsb_ctx *c, int *pushed, int *final_frames)
{
    if (!c || !c->stream_on) return LWSB_ERR_STATE;
    const int T = c->T[0];
    if (pushed) *pushed = T;
    if (final_frames) *final_frames = c->stream_ended ? T : std::max(0, T - c->stream_LA);
    return LWSB_OK;
}

extern "C" int lwsb_stream_read(lwsb_ctx *c, void *out, int first_frame, int nframes, int where)
{
    CHECK_CTX(c);
    if (!c->stream_on) return fail(c, LWSB_ERR_STATE, "no stream");
    if (!out || first_frame < 0 || nframes < 0 || first_frame + nframes > c->T[0] || (where != LWSB_HOST && where != LWSB_DEVICE))
        return fail(c, LWSB_ERR_ARG, "bad lwsb_stream_read arguments");
    if (nframes == 0) return LWSB_OK;
    if (int r = use_device(c)) return r;
    const double2 *src = c->E.as<double2>() + (long long)(first_frame + c->Q - 1) * c->P + c->c0;
    CU(c, cudaMemcpy2DAsync(out, (size_t)c->Nreal * sizeof(double2), src, (size_t)c->P * sizeof(double2), (size_t)c->Nreal * sizeof(double2),
                            nframes, where == LWSB_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return LWSB_OK;
}

extern "C" int lwsb_stream_end(lwsb_ctx *c)
{
    CHECK_CTX(c);
    if (!c->stream_on) return fail(c, LWSB_ERR_STATE, "no stream");
    c->stream_ended = true; // nothing left to compute: every frame has had its row updates; all frames are final
    return c->T[0];
}

// Between two chained stages the reference crops and re-extends (lws.pyx:256 then 235-240 of the
// next call): ghost frames become copies of the *updated* edge frames, |.| and its mean are
// recomputed from the updated values.
static int restage(lwsb_ctx *c)
{
    launch_reextend(c->view(), c->scratch(), c->mean_amp.as<double>(), c->max_amp.as<double>(),
                    c->maxT + 2 * (c->Q - 1), c->stream);
    c->launches += 3;
    CU(c, cudaGetLastError());
    return LWSB_OK;
}

extern "C" int lwsb_restage(lwsb_ctx *c)
{
    CHECK_CTX(c);
    if (int r = check_resident(c)) return r;
    if (int r = use_device(c)) return r;
    return restage(c);
}

// ============================================================================ one-shot interface
extern "C" int lwsb_batch_lws(lwsb_ctx *c, const void *const *S_in, void *const *S_out, const int *T, int B, int Nreal,
                              int kind, int where, const double *thresholds, int iterations, int flags)
{
    if (int r = lwsb_load(c, S_in, T, B, Nreal, kind, where)) return r;
    bool pinned_out = where == LWSB_HOST && S_out != nullptr;
    for (int b = 0; b < B && pinned_out; ++b) pinned_out = S_out[b] && !is_pageable(S_out[b]);
    if (int r = batch_impl(c, thresholds, iterations, flags, pinned_out ? S_out : nullptr)) return r;
    if (c->early_stored) return LWSB_OK; // the results left while the kernel was running
    return lwsb_store(c, S_out, where);
}

extern "C" int lwsb_nofuture_lws(lwsb_ctx *c, int which, const void *const *S_in, void *const *S_out, const int *T,
                                 int B, int Nreal, int kind, int where, const double *thresholds, int iterations,
                                 int flags)
{
    if (int r = lwsb_load(c, S_in, T, B, Nreal, kind, where)) return r;
    if (int r = lwsb_nofuture(c, which, thresholds, iterations, flags)) return r;
    return lwsb_store(c, S_out, where);
}

extern "C" int lwsb_online_lws(lwsb_ctx *c, const void *const *S_in, void *const *S_out, const int *T, int B, int Nreal,
                               int kind, int where, const double *thresholds, int iterations, int look_ahead, int flags)
{
    if (int r = lwsb_load(c, S_in, T, B, Nreal, kind, where)) return r;
    if (int r = lwsb_online(c, thresholds, iterations, look_ahead, flags)) return r;
    return lwsb_store(c, S_out, where);
}

extern "C" int lwsb_run_lws(lwsb_ctx *c, const void *const *S_in, void *const *S_out, const int *T, int B, int Nreal,
                            int kind, int where, const double *nofuture_thr, int nofuture_it, const double *online_thr,
                            int online_it, int look_ahead, const double *batch_thr, int batch_it, int flags)
{
    if (int r = lwsb_load(c, S_in, T, B, Nreal, kind, where)) return r;
    bool dirty = false;
    if (nofuture_it > 0) {
        if (int r = lwsb_nofuture(c, LWSB_W_AI, nofuture_thr, nofuture_it, flags)) return r;
        dirty = true;
    }
    if (online_it > 0) {
        if (dirty) if (int r = restage(c)) return r;
        if (int r = lwsb_online(c, online_thr, online_it, look_ahead, flags)) return r;
        dirty = true;
    }
    c->early_stored = false;
    if (batch_it > 0) {
        if (dirty) if (int r = restage(c)) return r;
        bool pinned_out = where == LWSB_HOST && S_out != nullptr;
        for (int b = 0; b < B && pinned_out; ++b) pinned_out = S_out[b] && !is_pageable(S_out[b]);
        if (int r = batch_impl(c, batch_thr, batch_it, flags, pinned_out ? S_out : nullptr)) return r;
        if (c->early_stored) return LWSB_OK;
    }
    return lwsb_store(c, S_out, where);
}

// ============================================================================ stft / istft
extern "C" int lwsb_stft_frames(int nsamples, int fsize, int fshift, int perfectrec)
{
    // lws.pyx:54-77
    if (nsamples < 0 || fsize < 1 || fshift < 1) return LWSB_ERR_ARG;
    if (perfectrec) {
        const int res = fsize % fshift;
        const int pre = res == 0 ? fsize - fshift : fsize - res;
        const int post = (fshift - nsamples % fshift) % fshift;
        return (pre + nsamples + post) / fshift;
    }
    const int d = nsamples - fsize;
    const int post = ((fshift - d % fshift) % fshift + fshift) % fshift;
    return (nsamples + post - fsize) / fshift + 1;
}

extern "C" int lwsb_stft_prepad(int fsize, int fshift, int perfectrec)
{
    // lws.pyx:54-60
    if (fsize < 1 || fshift < 1) return LWSB_ERR_ARG;
    if (!perfectrec) return 0;
    const int res = fsize % fshift;
    return res == 0 ? fsize - fshift : fsize - res;
}

namespace {

// exp(-2*pi*i*j/N), j in [0, N): host-computed once per N in fp64, the axis points set exactly
int get_twiddles(lwsb_ctx *c, int N, const double2 **out)
{
    auto it = c->twiddles.find(N);
    if (it == c->twiddles.end()) {
        std::vector<double2> h(N);
        for (int j = 0; j < N; ++j) {
            const double a = -2.0 * M_PI * (double)j / (double)N;
            h[j] = make_double2(std::cos(a), std::sin(a));
        }
        if (N % 4 == 0) { h[N / 4] = make_double2(0.0, -1.0); h[3 * N / 4] = make_double2(0.0, 1.0); }
        if (N % 2 == 0) h[N / 2] = make_double2(-1.0, 0.0);
        DevBuf b;
        CU(c, b.reserve((size_t)N * sizeof(double2)));
        CU(c, cudaMemcpyAsync(b.p, h.data(), (size_t)N * sizeof(double2), cudaMemcpyHostToDevice, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
        it = c->twiddles.emplace(N, b).first;
    }
    *out = it->second.as<const double2>();
    return LWSB_OK;
}

int ilog2_exact(int N)
{
    int l = 0;
    while ((1 << l) < N) ++l;
    return (1 << l) == N ? l : -1;
}

} // namespace

extern "C" int lwsb_stft(lwsb_ctx *c, const double *x, int B, int nsamples, const double *awin, int fsize, int fshift,
                         int fftsize, int pre_pad, int M, void *S_out, int where)
{
    CHECK_CTX(c);
    if (!x || !awin || !S_out || B < 1 || nsamples < 0 || fsize < 1 || fshift < 1 || fftsize < 2 || pre_pad < 0 || M < 1 ||
        (where != LWSB_HOST && where != LWSB_DEVICE))
        return fail(c, LWSB_ERR_ARG, "bad lwsb_stft arguments");
    if (fftsize % 2 == 1) return fail(c, LWSB_ERR_ARG, "Odd ffts not supported."); // lws.pyx:51-52
    if (int r = use_device(c)) return r;
    const int N = fftsize, Nb = N / 2 + 1;
    int logN = ilog2_exact(N);
    if (logN >= 0 && (size_t)N * sizeof(double2) > c->prop.sharedMemPerBlockOptin) logN = -1;
    if (logN < 0 && (size_t)N * sizeof(double) > c->prop.sharedMemPerBlockOptin)
        return fail(c, LWSB_ERR_UNSUPPORTED, "fftsize too large for the on-chip transform");
    const double2 *tw;
    if (int r = get_twiddles(c, N, &tw)) return r;
    const size_t nx = (size_t)B * std::max(nsamples, 1), nS = (size_t)B * M * Nb;
    CU(c, c->fwin.reserve((size_t)fsize * sizeof(double)));
    CU(c, cudaMemcpyAsync(c->fwin.p, awin, (size_t)fsize * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    const double *dx = x;
    double2 *dS = reinterpret_cast<double2 *>(S_out);
    if (where == LWSB_HOST) {
        CU(c, c->fx.reserve(nx * sizeof(double)));
        CU(c, c->fS.reserve(nS * sizeof(double2)));
        if (nsamples > 0)
            CU(c, cudaMemcpyAsync(c->fx.p, x, (size_t)B * nsamples * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        dx = c->fx.as<double>();
        dS = c->fS.as<double2>();
    }
    CU(c, launch_stft(dx, B, nsamples, c->fwin.as<double>(), fsize, fshift, N, logN, pre_pad, tw, dS, M, c->stream));
    c->launches += 1;
    if (where == LWSB_HOST) CU(c, cudaMemcpyAsync(S_out, dS, nS * sizeof(double2), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return LWSB_OK;
}

extern "C" int lwsb_istft(lwsb_ctx *c, const void *S_in, int B, int M, int Nreal, const double *swin, int nswin,
                          int fshift, double *x_out, int where)
{
    CHECK_CTX(c);
    if (!S_in || !swin || !x_out || B < 1 || M < 1 || Nreal < 2 || nswin < 1 || fshift < 1 ||
        (where != LWSB_HOST && where != LWSB_DEVICE))
        return fail(c, LWSB_ERR_ARG, "bad lwsb_istft arguments");
    if (Nreal % 2 != 1) // lws.pyx:100-101
        return fail(c, LWSB_ERR_EVEN_NREAL, "We expect the spectrogram to only have non-negative frequencies");
    if (int r = use_device(c)) return r;
    const int N = 2 * (Nreal - 1);
    int logN = ilog2_exact(N);
    if (logN >= 0 && (size_t)N * sizeof(double2) > c->prop.sharedMemPerBlockOptin) logN = -1;
    if (logN < 0 && (size_t)(N / 2 + 2) * sizeof(double2) > c->prop.sharedMemPerBlockOptin)
        return fail(c, LWSB_ERR_UNSUPPORTED, "spectrum too wide for the on-chip transform");
    const double2 *tw;
    if (int r = get_twiddles(c, N, &tw)) return r;
    const size_t len = (size_t)fshift * (M - 1) + N, nS = (size_t)B * M * Nreal;
    CU(c, c->fwin.reserve((size_t)nswin * sizeof(double)));
    CU(c, cudaMemcpyAsync(c->fwin.p, swin, (size_t)nswin * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(c, c->fframes.reserve((size_t)B * M * N * sizeof(double)));
    const double2 *dS = reinterpret_cast<const double2 *>(S_in);
    double *dx = x_out;
    if (where == LWSB_HOST) {
        CU(c, c->fS.reserve(nS * sizeof(double2)));
        CU(c, c->fx.reserve((size_t)B * len * sizeof(double)));
        CU(c, cudaMemcpyAsync(c->fS.p, S_in, nS * sizeof(double2), cudaMemcpyHostToDevice, c->stream));
        dS = c->fS.as<const double2>();
        dx = c->fx.as<double>();
    }
    CU(c, launch_istft(dS, B, M, N, logN, c->fwin.as<double>(), nswin, fshift, tw, c->fframes.as<double>(), dx, c->stream));
    c->launches += 2;
    if (where == LWSB_HOST) CU(c, cudaMemcpyAsync(x_out, dx, (size_t)B * len * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return LWSB_OK;
}

// ============================================================================ fused calls
namespace {

// length of the signal istft returns for M frames (transforms.istft: the perfect-reconstruction slice of lws.pyx:135)
long long istft_length(int M, int fsize, int fshift, int perfectrec, int *crop_lo)
{
    const long long full = (long long)fshift * (M - 1) + fsize;
    *crop_lo = 0;
    if (!perfectrec) return full;
    *crop_lo = lwsb_stft_prepad(fsize, fshift, 1);
    if (fsize == fshift) return 0; // sig[:, pre:0] is empty
    return std::max(0LL, full - *crop_lo - (fsize - fshift));
}

// consistency of the B spectrograms at dS (device, (B, M, Nreal) complex128): 20 log10(|S| / |stft(istft(S)) - S|)
int consistency_device(lwsb_ctx *c, const double2 *dS, int B, int M, int Nreal, const double *awin, const double *swin, int nswin,
                       int fsize, int fshift, int perfectrec, double *out_db)
{
    int lo = 0;
    const long long ylen = istft_length(M, fsize, fshift, perfectrec, &lo);
    const long long full = (long long)fshift * (M - 1) + fsize;
    if (ylen <= 0 || lwsb_stft_frames((int)ylen, fsize, fshift, perfectrec) != M)
        return fail(c, LWSB_ERR_UNSUPPORTED, "consistency: stft(istft(S)) does not have the shape of S for these parameters");
    CU(c, c->ry.reserve((size_t)B * full * sizeof(double)));
    if (int r = lwsb_istft(c, dS, B, M, Nreal, swin, nswin, fshift, c->ry.as<double>(), LWSB_DEVICE)) return r;
    CU(c, c->rx.reserve((size_t)B * ylen * sizeof(double)));
    CU(c, cudaMemcpy2DAsync(c->rx.p, (size_t)ylen * sizeof(double), c->ry.as<double>() + lo, (size_t)full * sizeof(double),
                            (size_t)ylen * sizeof(double), B, cudaMemcpyDeviceToDevice, c->stream));
    const size_t nS = (size_t)B * M * Nreal;
    CU(c, c->rR.reserve(nS * sizeof(double2)));
    if (int r = lwsb_stft(c, c->rx.as<double>(), B, (int)ylen, awin, fsize, fshift, fsize, lwsb_stft_prepad(fsize, fshift, perfectrec), M,
                          c->rR.p, LWSB_DEVICE))
        return r;
    const int nblk = 64;
    CU(c, c->rn.reserve(((size_t)B * nblk * 2 + (size_t)B * 2) * sizeof(double)));
    double *partial = c->rn.as<double>(), *sums = partial + (size_t)B * nblk * 2;
    CU(c, launch_sq_norms(dS, c->rR.as<double2>(), B, (long long)M * Nreal, partial, nblk, sums, c->stream));
    c->launches += 2;
    std::vector<double> h(2 * (size_t)B);
    CU(c, cudaMemcpyAsync(h.data(), sums, h.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    for (int b = 0; b < B; ++b) out_db[b] = 20.0 * std::log10(std::sqrt(h[2 * b]) / std::sqrt(h[2 * b + 1]));
    return LWSB_OK;
}

} // namespace

extern "C" long long lwsb_reconstruct_length(int nsamples, int fsize, int fshift, int perfectrec)
{
    const int M = lwsb_stft_frames(nsamples, fsize, fshift, perfectrec);
    if (M < 1) return LWSB_ERR_ARG;
    int lo;
    return istft_length(M, fsize, fshift, perfectrec, &lo);
}

extern "C" int lwsb_consistency(lwsb_ctx *c, const void *S, int B, int M, int Nreal, const double *awin, const double *swin,
                                int nswin, int fshift, int perfectrec, int where, double *out_db)
{
    CHECK_CTX(c);
    if (!S || !awin || !swin || !out_db || B < 1 || M < 1 || Nreal < 2 || nswin < 1 || fshift < 1 ||
        (where != LWSB_HOST && where != LWSB_DEVICE))
        return fail(c, LWSB_ERR_ARG, "bad lwsb_consistency arguments");
    if (Nreal % 2 != 1) return fail(c, LWSB_ERR_EVEN_NREAL, "We expect the spectrogram to only have non-negative frequencies");
    if (int r = use_device(c)) return r;
    const int fsize = 2 * (Nreal - 1);
    const double2 *dS = reinterpret_cast<const double2 *>(S);
    if (where == LWSB_HOST) {
        const size_t nS = (size_t)B * M * Nreal;
        CU(c, c->rS.reserve(nS * sizeof(double2)));
        CU(c, cudaMemcpyAsync(c->rS.p, S, nS * sizeof(double2), cudaMemcpyHostToDevice, c->stream));
        dS = c->rS.as<double2>();
    }
    return consistency_device(c, dS, B, M, Nreal, awin, swin, nswin, fsize, fshift, perfectrec, out_db);
}

// consistency (dB) of every utterance of the RESIDENT batch as it stands (between lwsb_batch calls: a per-sweep trace of
// the reference's only quality metric, lws.pyx:140-144); fftsize = fsize = 2 (Nreal - 1)
extern "C" int lwsb_resident_consistency(lwsb_ctx *c, const double *awin, const double *swin, int nswin, int fshift, int perfectrec,
                                         double *out_db)
{
    CHECK_CTX(c);
    if (!awin || !swin || !out_db || nswin < 1 || fshift < 1) return fail(c, LWSB_ERR_ARG, "bad lwsb_resident_consistency arguments");
    if (int r = check_resident(c)) return r;
    if (int r = use_device(c)) return r;
    const int B = c->B, fsize = 2 * (c->Nreal - 1);
    CU(c, c->stage.reserve((size_t)c->total_bins * sizeof(double2)));
    for (int b = 0; b < B; ++b) c->hptr[b] = c->stage.as<double2>() + c->binbase[b];
    CU(c, cudaMemcpyAsync(c->dptr.p, c->hptr.data(), B * sizeof(void *), cudaMemcpyHostToDevice, c->stream));
    launch_crop(c->view(), c->dptr.as<void *const>(), c->maxT, c->stream);
    c->launches += 1;
    CU(c, cudaGetLastError());
    CU(c, cudaStreamSynchronize(c->stream));
    for (int b = 0; b < B; ++b)
        if (int r = consistency_device(c, c->stage.as<double2>() + c->binbase[b], 1, c->T[b], c->Nreal, awin, swin, nswin, fsize, fshift,
                                       perfectrec, out_db + b))
            return r;
    return LWSB_OK;
}

// y = istft(run_lws(|stft(x)|)) for B signals of equal length without leaving the device (lws.pyx:43-137, 495-499 chained)
extern "C" int lwsb_reconstruct(lwsb_ctx *c, const double *x, int B, int nsamples, const double *awin, const double *swin,
                                int fsize, int fshift, int perfectrec, const double *nofuture_thr, int nofuture_it,
                                const double *online_thr, int online_it, int look_ahead, const double *batch_thr, int batch_it,
                                int flags, double *y_out, int where, double *consistency_db)
{
    CHECK_CTX(c);
    if (!x || !awin || !swin || !y_out || B < 1 || nsamples < 1 || fsize < 2 || fsize % 2 || fshift < 1 ||
        (where != LWSB_HOST && where != LWSB_DEVICE))
        return fail(c, LWSB_ERR_ARG, "bad lwsb_reconstruct arguments");
    if (int r = use_device(c)) return r;
    const int M = lwsb_stft_frames(nsamples, fsize, fshift, perfectrec), Nreal = fsize / 2 + 1;
    if (M < 1) return fail(c, LWSB_ERR_ARG, "signal too short for one frame");
    int lo = 0;
    const long long ylen = istft_length(M, fsize, fshift, perfectrec, &lo);
    const long long full = (long long)fshift * (M - 1) + fsize;
    const double *dx = x;
    if (where == LWSB_HOST) {
        CU(c, c->rx.reserve((size_t)B * nsamples * sizeof(double)));
        CU(c, cudaMemcpyAsync(c->rx.p, x, (size_t)B * nsamples * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        dx = c->rx.as<double>();
    }
    const size_t per = (size_t)M * Nreal;
    CU(c, c->rS.reserve((size_t)B * per * sizeof(double2)));
    if (int r = lwsb_stft(c, dx, B, nsamples, awin, fsize, fshift, fsize, lwsb_stft_prepad(fsize, fshift, perfectrec), M, c->rS.p, LWSB_DEVICE))
        return r;
    // the magnitudes np.abs(stft(x)) a caller of the reference hands to run_lws (the phases are discarded)
    CU(c, c->rR.reserve((size_t)B * per * sizeof(double)));
    CU(c, launch_cabs(c->rS.as<double2>(), c->rR.as<double>(), (long long)B * per, c->stream));
    c->launches += 1;
    std::vector<const void *> in(B);
    std::vector<void *> out(B);
    std::vector<int> T(B, M);
    for (int b = 0; b < B; ++b) { in[b] = c->rR.as<double>() + b * per; out[b] = c->rS.as<double2>() + b * per; }
    if (int r = lwsb_run_lws(c, in.data(), out.data(), T.data(), B, Nreal, LWSB_F64, LWSB_DEVICE, nofuture_thr, nofuture_it, online_thr,
                             online_it, look_ahead, batch_thr, batch_it, flags))
        return r;
    if (consistency_db)
        if (int r = consistency_device(c, c->rS.as<double2>(), B, M, Nreal, awin, swin, fsize, fsize, fshift, perfectrec, consistency_db))
            return r;
    CU(c, c->ry.reserve((size_t)B * full * sizeof(double)));
    if (int r = lwsb_istft(c, c->rS.p, B, M, Nreal, swin, fsize, fshift, c->ry.as<double>(), LWSB_DEVICE)) return r;
    if (ylen > 0)
        CU(c, cudaMemcpy2DAsync(y_out, (size_t)ylen * sizeof(double), c->ry.as<double>() + lo, (size_t)full * sizeof(double),
                                (size_t)ylen * sizeof(double), B, where == LWSB_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice,
                                c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return LWSB_OK;
}

// ============================================================================ introspection
extern "C" int lwsb_last_compute_ms(lwsb_ctx *c, float *ms)
{
    CHECK_CTX(c);
    if (!ms) return fail(c, LWSB_ERR_ARG, "ms is NULL");
    if (!c->timing_valid) return fail(c, LWSB_ERR_STATE, "no compute call timed yet");
    if (int r = use_device(c)) return r;
    CU(c, cudaEventSynchronize(c->ev1));
    CU(c, cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return LWSB_OK;
}

extern "C" long long lwsb_launch_count(const lwsb_ctx *c) { return c ? c->launches : 0; }

extern "C" int lwsb_last_stage_ms(lwsb_ctx *c, float *ms3)
{
    CHECK_CTX(c);
    if (!ms3) return fail(c, LWSB_ERR_ARG, "ms3 is NULL");
    if (int r = use_device(c)) return r;
    for (int i = 0; i < 3; ++i) {
        ms3[i] = -1.0f;
        if (!c->stage_valid[i]) continue;
        CU(c, cudaEventSynchronize(c->evs[i][1]));
        CU(c, cudaEventElapsedTime(&ms3[i], c->evs[i][0], c->evs[i][1]));
    }
    return LWSB_OK;
}

extern "C" int lwsb_last_online_kernel(const lwsb_ctx *c) { return c ? c->last_online_kernel : LWSB_ERR_ARG; }

extern "C" int lwsb_last_batch_work(const lwsb_ctx *c, long long *out4)
{
    if (!c || !out4) return LWSB_ERR_ARG;
    for (int i = 0; i < 4; ++i) out4[i] = c->last_work[i];
    return LWSB_OK;
}

extern "C" int lwsb_last_batch_cycles(lwsb_ctx *c, unsigned long long *out7)
{
    CHECK_CTX(c);
    if (!out7) return fail(c, LWSB_ERR_ARG, "out7 is NULL");
    if (c->last_kernel != 1 || !c->status.p) return 0;
    if (int r = use_device(c)) return r;
    CU(c, cudaMemcpyAsync(out7, c->status.as<char>() + 8, 13 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return 1;
}

extern "C" int lwsb_last_batch_trace(lwsb_ctx *c, int enable, int max_items, int *utt_pass, unsigned long long *stamps)
{
    CHECK_CTX(c);
    c->want_trace = enable != 0;
    if (!utt_pass || !stamps || max_items <= 0 || c->trace_items == 0 || c->last_kernel != 1) return 0;
    if (int r = use_device(c)) return r;
    const int n = std::min(max_items, c->trace_items);
    CU(c, cudaMemcpyAsync(stamps, c->trace.p, (size_t)n * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    for (int i = 0; i < 2 * n; ++i) utt_pass[i] = c->trace_list[i];
    return n;
}

extern "C" int lwsb_debug_fast_math(lwsb_ctx *c, long long n, unsigned long long seed, unsigned long long *out4)
{
    CHECK_CTX(c);
    if (!out4 || n < 0) return fail(c, LWSB_ERR_ARG, "bad lwsb_debug_fast_math arguments");
    if (int r = use_device(c)) return r;
    CU(c, c->status.reserve(256));
    CU(c, cudaMemsetAsync(c->status.p, 0, 256, c->stream));
    CU(c, launch_debug_fast_math(n, seed, c->status.as<unsigned long long>(), c->stream));
    c->launches += 1;
    CU(c, cudaMemcpyAsync(out4, c->status.p, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    c->last_kernel = 0;
    return LWSB_OK;
}

extern "C" int lwsb_set_tuning(lwsb_ctx *c, long long smem_limit, int cluster, int sweeps_per_pass)
{
    CHECK_CTX(c);
    if (smem_limit < 0 || cluster < 0 || cluster > 8 || (cluster & (cluster - 1)))
        return fail(c, LWSB_ERR_ARG, "bad tuning values");
    c->tune_smem = smem_limit; c->tune_cluster = cluster; c->tune_sweeps = sweeps_per_pass;
    return LWSB_OK;
}

extern "C" int lwsb_set_block_bins(lwsb_ctx *c, int bins)
{
    CHECK_CTX(c);
    if (bins != 0 && bins != 4 && bins != 8) return fail(c, LWSB_ERR_ARG, "bins per block: 0 (automatic), 4 or 8");
#ifndef LWSB_EXPERIMENTS
    if (bins == 4) return fail(c, LWSB_ERR_UNSUPPORTED, "4-bin blocks are built only with -DLWSB_EXPERIMENTS");
#endif
    c->tune_block = bins;
    return LWSB_OK;
}

extern "C" int lwsb_set_variant(lwsb_ctx *c, int sweep_lag, int tensor_memory)
{
    CHECK_CTX(c);
    if (sweep_lag < 0 || tensor_memory < 0 || (tensor_memory > LWSB_VARIANT_DUO && (tensor_memory < LWSB_VARIANT_PAIR || tensor_memory > LWSB_VARIANT_PAIR + 5)))
        return fail(c, LWSB_ERR_ARG, "bad variant values");
    c->tune_lag = sweep_lag; c->tune_tm = tensor_memory;
    return LWSB_OK;
}

extern "C" int lwsb_last_batch_plan(const lwsb_ctx *c, int *out9)
{
    if (!c || !out9) return LWSB_ERR_ARG;
    if (c->last_kernel != 1) return 0;
    const StripPlan &p = c->last_plan;
    const int v[15] = {p.C, p.NBr, p.NBV, p.NS, p.G, p.R, p.pitch, p.nthreads, p.smem_bytes, p.QS, p.GFAST, p.TM, p.SBK, p.GX, p.LEAD};
    for (int i = 0; i < 15; ++i) out9[i] = v[i];
    return 1;
}

extern "C" int lwsb_device_info(lwsb_ctx *c, int *sm_count, int *cc_major, int *cc_minor, long long *hbm_bytes)
{
    CHECK_CTX(c);
    if (sm_count) *sm_count = c->prop.multiProcessorCount;
    if (cc_major) *cc_major = c->prop.major;
    if (cc_minor) *cc_minor = c->prop.minor;
    if (hbm_bytes) *hbm_bytes = (long long)c->prop.totalGlobalMem;
    return LWSB_OK;
}

extern "C" int lwsb_get_stats(lwsb_ctx *c, double *mean_amp, double *max_amp)
{
    CHECK_CTX(c);
    if (int r = check_resident(c)) return r;
    if (int r = use_device(c)) return r;
    if (mean_amp) CU(c, cudaMemcpyAsync(mean_amp, c->mean_amp.p, c->B * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (max_amp) CU(c, cudaMemcpyAsync(max_amp, c->max_amp.p, c->B * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return LWSB_OK;
}

// ============================================================================ host-only debug mirrors
extern "C" int lwsb_debug_terms(const double *wr, const double *wi, int Q, int L, int fold, int rframe, int cframe,
                                int p, int max_terms, int *dr, int *dk, double *cr, double *ci)
{
    if (!wr || !wi || Q < 1 || L < 0 || p < 0 || p >= Q) return LWSB_ERR_ARG;
    WeightSet w;
    w.Q = Q; w.L = L;
    const size_t n = (size_t)Q * Q * (L + 1);
    w.wr.assign(wr, wr + n);
    w.wi.assign(wi, wi + n);
    std::vector<LwsbTerm> t;
    build_terms(w, fold, rframe, cframe, p, t);
    const int cnt = (int)t.size();
    for (int i = 0; i < cnt && i < max_terms; ++i) {
        if (dr) dr[i] = t[i].dr;
        if (dk) dk[i] = t[i].dk;
        if (cr) cr[i] = t[i].cr;
        if (ci) ci[i] = t[i].ci;
    }
    return cnt;
}

extern "C" int lwsb_debug_plan_strips(int Nreal, int Q, int L, int iterations, int maxT, int B, long long smem_limit,
                                      int sm_count, int force_cluster, int max_sweeps, int force_block, int *out9)
{
    if (!out9) return LWSB_ERR_ARG;
    StripPlan p;
    if (!plan_strips(Nreal, Q, L, iterations, maxT, B, (size_t)smem_limit, sm_count, &p, force_cluster, max_sweeps, 0, 0, 0, force_block)) return 0;
    const int v[15] = {p.C, p.NBr, p.NBV, p.NS, p.G, p.R, p.pitch, p.nthreads, p.smem_bytes, p.QS, p.GFAST, p.TM, p.SBK, p.GX, p.LEAD};
    for (int i = 0; i < 15; ++i) out9[i] = v[i];
    return 1;
}

extern "C" int lwsb_debug_work_items(const int *active_sweeps, int B, int sweeps_per_pass, int max_items, int *utt_pass)
{
    if (!active_sweeps || B < 1 || sweeps_per_pass < 1 || max_items < 0 || (max_items > 0 && !utt_pass)) return LWSB_ERR_ARG;
    std::vector<int> items;
    build_work_items(active_sweeps, B, sweeps_per_pass, items);
    const int n = (int)(items.size() / 2);
    for (int i = 0; i < 2 * std::min(n, max_items); ++i) utt_pass[i] = items[i];
    return n;
}

extern "C" long long lwsb_debug_online_chain_length(int T, int iterations, int look_ahead)
{
    return lwsb_online_chain_len(T, iterations, look_ahead);
}

extern "C" int lwsb_debug_online_task(int T, int iterations, int look_ahead, int Q, long long j, int *row, int *which,
                                      int *rframe, int *cframe, int *thr_index)
{
    if (j < 0 || j >= lwsb_online_chain_len(T, iterations, look_ahead)) return LWSB_ERR_ARG;
    const LwsbOnlineTask t = lwsb_online_decode(T, iterations, look_ahead, Q, j);
    if (row) *row = t.row;
    if (which) *which = t.which;
    if (rframe) *rframe = t.rframe;
    if (cframe) *cframe = t.cframe;
    if (thr_index) *thr_index = t.thr;
    return LWSB_OK;
}

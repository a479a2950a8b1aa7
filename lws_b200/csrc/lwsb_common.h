// lwsb_common.h -- definitions shared by the host runtime and the CUDA kernels.
//
// Vocabulary (follows the reference):
//   frame m / bin c      : a row / column of one utterance's spectrogram, c in [0, Nreal)
//   extended spectrogram : (T + 2(Q-1)) x (Nreal + 2L) array with frozen ghost frames and
//                          mirrored bins (python/lws.pyx:146-157); extended column e = c + L
//   row update           : one in-place left-to-right pass over the bins of one frame
//   sweep                : row updates of all T frames in order (one "iteration")
//   chain                : the sequence of row updates an entry point performs
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define LWSB_HD __host__ __device__ __forceinline__
#else
#define LWSB_HD inline
#endif

// Formula family of the reference's C variants (how the +k / -k neighbours share a weight).
enum { LWSB_FOLD_ANY = 0, LWSB_FOLD_Q2 = 2, LWSB_FOLD_Q4 = 4, LWSB_FOLD_NF4 = 5 };
// variants of the cluster strip kernel (StripPlan::TM, lwsb_set_variant): one thread per task (SCALAR), two lanes per
// task alternating bins (DUO); builds with -DLWSB_EXPERIMENTS also have the tensor-memory producer / consumer warps (TM)
// and two lanes per task split by component (LWSB_VARIANT_PAIR + register-window mode 0..2)
enum { LWSB_VARIANT_AUTO = 0, LWSB_VARIANT_TM = 1, LWSB_VARIANT_SCALAR = 2, LWSB_VARIANT_DUO = 3, LWSB_VARIANT_PAIR = 10 };

// One term of the linear stencil   acc += (cr + i*ci) * E[m + dr][e + dk].
// For LWSB_FOLD_NF4 (the reference's NoFuture_LWSQ4 with its doubled bin offset) dk is
// relative to the flat offset (m+dr)*Np + 2*e of the reference layout.
struct LwsbTerm {
    int dr, dk;
    double cr, ci;
};

// ------------------------------------------------------------------ online (TF-RTISI-LA) chain
// lwslib.cpp:1432-1491: for every frame m: one initial row update of frame m (W_ai, threshold
// 0), then `it` times [row updates of the look-ahead frames max(m-LA,0)..m-1 with W, then
// frame m with W_af].  Position j of that chain is decoded arithmetically.
struct LwsbOnlineTask {
    int row;     // extended row index (frame + Q - 1)
    int which;   // LWSB_W / LWSB_W_AI / LWSB_W_AF
    int rframe;  // frames r < rframe are used on both sides, r >= rframe on the left only
    int cframe;  // centre-frame +-k terms used
    int thr;     // index into the threshold array, -1 = threshold 0
};

LWSB_HD long long lwsb_online_frame_base(int m, int it, int LA)
{
    // number of row updates before frame m
    if (LA <= 0) return (long long)m * (1 + it);
    if (m <= LA) return (long long)m + (long long)it * m * (m + 1) / 2;
    long long bLA = (long long)LA + (long long)it * LA * (LA + 1) / 2;
    return bLA + (long long)(m - LA) * (1 + (long long)it * (LA + 1));
}

LWSB_HD long long lwsb_online_chain_len(int T, int it, int LA) { return lwsb_online_frame_base(T, it, LA); }

// frame whose block of row updates contains chain position j
LWSB_HD int lwsb_online_frame(int it, int LA, long long j)
{
    if (LA <= 0) return (int)(j / (1 + it));
    const long long bLA = lwsb_online_frame_base(LA, it, LA);
    if (j >= bLA) return LA + (int)((j - bLA) / (1 + (long long)it * (LA + 1)));
    int m = 0;
    while (lwsb_online_frame_base(m + 1, it, LA) <= j) ++m;
    return m;
}

LWSB_HD LwsbOnlineTask lwsb_online_decode(int T, int it, int LA, int Q, long long j)
{
    (void)T;
    const int m = lwsb_online_frame(it, LA, j);
    const int nf = (LA > 0) ? (m < LA ? m : LA) : 0;
    long long rem = j - lwsb_online_frame_base(m, it, LA);
    LwsbOnlineTask t;
    if (rem == 0) { // Asym_UpdatePhase*(.., M=1, M0=0, .., threshold 0): lwslib.cpp:1467
        t.row = m + Q - 1; t.which = 1; t.rframe = 1; t.cframe = 0; t.thr = -1;
        return t;
    }
    rem -= 1;
    const int h = (int)(rem / (nf + 1));
    const int a = (int)(rem % (nf + 1));
    t.thr = h;
    t.cframe = 1;
    if (a < nf) { // look-ahead block, M = nf, M0 = nf + 1: lwslib.cpp:1472, 917-925
        t.row = (m - nf) + a + Q - 1;
        t.which = 0;
        t.rframe = nf + 1 - a;
        if (t.rframe > Q) t.rframe = Q;
    } else {      // newest frame with the full asymmetric window, M = 1, M0 = 1: lwslib.cpp:1475
        t.row = m + Q - 1;
        t.which = 2;
        t.rframe = 1;
    }
    return t;
}

// ------------------------------------------------------------------ device-side views
#ifdef __CUDACC__
struct LwsbStencil {       // one linear stencil per residue p = c mod Q
    const LwsbTerm *terms; // [Q][maxt]
    const int *count;      // [Q]
    int maxt;
};

struct LwsbView {              // a resident batch of utterances
    double2 *E;                // extended spectrograms, all utterances stacked, row pitch P (complex)
    double *A;                 // amplitudes, same indexing
    const long long *rowbase;  // [B] first row of utterance u in E / A
    const int *T;              // [B] frames
    const double *mean_amp;    // [B] mean |S| (lws.pyx:240)
    int P;                     // row pitch in elements
    int c0;                    // physical column of bin 0 (= coff + L)
    int Nreal, L, Q, B;
};
#endif

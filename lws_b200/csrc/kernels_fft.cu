// kernels_fft.cu -- stft / istft on the device (python/lws.pyx:43-90, 93-137).
//
// One CTA per frame.  fftsize a power of two: in-place radix-2 FFT in shared memory (fp64,
// twiddles from a host-computed table so that accuracy matches a library FFT); any other
// even fftsize: direct DFT with the same table (exact index arithmetic k*t mod N) -- slow,
// O(N^2) per frame, but the reference accepts such sizes so the drop-in must too.
// The inverse transform uses the real-signal shortcut the reference takes implicitly
// (Hermitian extension, lws.pyx:121-122) and a gather-form overlap-add that sums the frames
// in increasing frame order like the reference's `signal[...] += ...` loop (lws.pyx:126).
#include <cuda_runtime.h>
#include <algorithm>
#include "exact.cuh"
#include "kernels.h"

namespace lwsb {

extern __shared__ double2 fft_smem[];

// in-place radix-2 decimation-in-time butterflies on bit-reversed input; sign = -1 forward, +1 inverse
__device__ __forceinline__ void fft_pow2_inplace(double2 *s, int N, int logN, const double2 *tw, double sign)
{
    for (int st = 0; st < logN; ++st) {
        const int half = 1 << st;
        __syncthreads();
        for (int j = threadIdx.x; j < (N >> 1); j += blockDim.x) {
            const int grp = j >> st, pos = j & (half - 1);
            const int i0 = (grp << (st + 1)) + pos, i1 = i0 + half;
            double2 w = tw[pos << (logN - 1 - st)]; // exp(-2*pi*i*pos/(2*half))
            w.y *= -sign;                            // table holds the forward sign
            const double2 a = s[i0], b = s[i1];
            const double2 t = make_double2(b.x * w.x - b.y * w.y, b.x * w.y + b.y * w.x);
            s[i0] = make_double2(a.x + t.x, a.y + t.y);
            s[i1] = make_double2(a.x - t.x, a.y - t.y);
        }
    }
    __syncthreads();
}

// grid (M, B).  x: (B, nsamples); frame m covers padded samples [m*hop, m*hop + fsize), padded
// sample p is x[p - pre] (0 outside).  S: (B, M, N/2+1) complex128.
__global__ void k_stft(const double *x, int nsamples, const double *awin, int fsize, int hop, int N, int logN,
                       int pre, const double2 *tw, double2 *S, int M)
{
    const int m = blockIdx.x, b = blockIdx.y;
    const double *xb = x + (long long)b * nsamples;
    const int nuse = fsize < N ? fsize : N; // np.fft.fft(frame, n=N) crops or zero-pads
    double2 *out = S + ((long long)b * M + m) * (N / 2 + 1);
    if (logN >= 0) {
        for (int t = threadIdx.x; t < N; t += blockDim.x) {
            double v = 0.0;
            if (t < nuse) {
                const long long p = (long long)m * hop + t - pre;
                if (p >= 0 && p < nsamples) v = xb[p] * awin[t];
            }
            fft_smem[__brev((unsigned)t) >> (32 - logN)] = make_double2(v, 0.0);
        }
        fft_pow2_inplace(fft_smem, N, logN, tw, -1.0);
        for (int k = threadIdx.x; k <= N / 2; k += blockDim.x) out[k] = fft_smem[k];
    } else {
        double *fr = reinterpret_cast<double *>(fft_smem);
        for (int t = threadIdx.x; t < nuse; t += blockDim.x) {
            const long long p = (long long)m * hop + t - pre;
            fr[t] = (p >= 0 && p < nsamples) ? xb[p] * awin[t] : 0.0;
        }
        __syncthreads();
        for (int k = threadIdx.x; k <= N / 2; k += blockDim.x) {
            double re = 0.0, im = 0.0;
            int idx = 0; // k*t mod N
            for (int t = 0; t < nuse; ++t) {
                const double2 w = tw[idx];
                re = fma(fr[t], w.x, re);
                im = fma(fr[t], w.y, im);
                idx += k; if (idx >= N) idx -= N;
            }
            out[k] = make_double2(re, im);
        }
    }
}

// grid (M, B).  S: (B, M, N/2+1) -> windowed frames fr: (B, M, N) real
__global__ void k_istft_frames(const double2 *S, const double *swin, int nswin, int N, int logN, const double2 *tw,
                               double *frames, int M)
{
    const int m = blockIdx.x, b = blockIdx.y;
    const double2 *in = S + ((long long)b * M + m) * (N / 2 + 1);
    double *out = frames + ((long long)b * M + m) * N;
    const double inv = 1.0 / (double)N;
    if (logN >= 0) {
        for (int k = threadIdx.x; k < N; k += blockDim.x) {
            double2 v;
            if (k <= N / 2) v = in[k];
            else { v = in[N - k]; v.y = -v.y; }
            fft_smem[__brev((unsigned)k) >> (32 - logN)] = v;
        }
        fft_pow2_inplace(fft_smem, N, logN, tw, +1.0);
        for (int t = threadIdx.x; t < N; t += blockDim.x) out[t] = fft_smem[t].x * inv * (t < nswin ? swin[t] : 0.0);
    } else {
        for (int k = threadIdx.x; k <= N / 2; k += blockDim.x) fft_smem[k] = in[k];
        __syncthreads();
        for (int t = threadIdx.x; t < N; t += blockDim.x) {
            // real part of sum_k X[k] e^{+2 pi i k t / N} over the Hermitian-extended spectrum
            double acc = fft_smem[0].x;
            int idx = 0;
            for (int k = 1; k < N / 2; ++k) {
                idx += t; if (idx >= N) idx -= N;
                const double2 w = tw[idx]; // e^{-i a}: cos a = w.x, sin a = -w.y
                const double2 X = fft_smem[k];
                acc += 2.0 * (X.x * w.x + X.y * w.y);
            }
            acc += fft_smem[N / 2].x * ((t & 1) ? -1.0 : 1.0);
            out[t] = acc * inv * (t < nswin ? swin[t] : 0.0);
        }
    }
}

// overlap-add, gather form: signal[b][p] = sum_s frames[b][s][p - hop*s], s increasing
__global__ void k_overlap_add(const double *frames, int N, int hop, int M, double *signal, long long len)
{
    const int b = blockIdx.y;
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= len) return;
    const long long s_lo = p < N ? 0 : (p - N + hop) / hop; // smallest s with p - hop*s < N
    long long s_hi = p / hop;
    if (s_hi > M - 1) s_hi = M - 1;
    double acc = 0.0;
    const double *fb = frames + (long long)b * M * N;
    for (long long s = s_lo; s <= s_hi; ++s) acc += fb[s * N + (p - hop * s)];
    signal[(long long)b * len + p] = acc;
}

size_t fft_smem_bytes(int N, int logN) { return logN >= 0 ? (size_t)N * sizeof(double2) : (size_t)(N / 2 + 1) * sizeof(double2) + 16; }

cudaError_t launch_stft(const double *x, int B, int nsamples, const double *awin, int fsize, int hop, int N, int logN,
                        int pre, const double2 *tw, double2 *S, int M, cudaStream_t s)
{
    size_t sm = fft_smem_bytes(N, logN);
    if (logN < 0) sm = (size_t)N * sizeof(double);
    if (sm > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_stft, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return e;
    }
    k_stft<<<dim3(M, B), 256, sm, s>>>(x, nsamples, awin, fsize, hop, N, logN, pre, tw, S, M);
    return cudaGetLastError();
}

cudaError_t launch_istft(const double2 *S, int B, int M, int N, int logN, const double *swin, int nswin, int hop,
                         const double2 *tw, double *frames, double *signal, cudaStream_t s)
{
    const size_t sm = fft_smem_bytes(N, logN);
    if (sm > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_istft_frames, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return e;
    }
    k_istft_frames<<<dim3(M, B), 256, sm, s>>>(S, swin, nswin, N, logN, tw, frames, M);
    const long long len = (long long)hop * (M - 1) + N;
    k_overlap_add<<<dim3((unsigned)((len + 255) / 256), B), 256, 0, s>>>(frames, N, hop, M, signal, len);
    return cudaGetLastError();
}

// Per-utterance squared norms for get_consistency (lws.pyx:140-144): out[2b] = sum |S|^2, out[2b+1] = sum |R - S|^2.
// One partial sum per block (fixed tree), added in block order by a second tiny kernel: the result does not depend
// on the order in which blocks finish.
__global__ void k_sq_norms(const double2 *S, const double2 *R, long long n, double *partial, int nblk)
{
    __shared__ double sa[256], sb[256];
    const int b = blockIdx.y;
    const double2 *s = S + (long long)b * n, *r = R + (long long)b * n;
    double a = 0.0, d = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double2 x = s[i], y = r[i];
        a += x.x * x.x + x.y * x.y;
        const double ex = y.x - x.x, ey = y.y - x.y;
        d += ex * ex + ey * ey;
    }
    sa[threadIdx.x] = a; sb[threadIdx.x] = d;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) { sa[threadIdx.x] += sa[threadIdx.x + w]; sb[threadIdx.x] += sb[threadIdx.x + w]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        partial[((long long)b * nblk + blockIdx.x) * 2] = sa[0];
        partial[((long long)b * nblk + blockIdx.x) * 2 + 1] = sb[0];
    }
}
__global__ void k_sq_norms_finish(const double *partial, int nblk, double *out)
{
    const int b = blockIdx.x;
    if (threadIdx.x == 0) {
        double a = 0.0, d = 0.0;
        for (int i = 0; i < nblk; ++i) { a += partial[((long long)b * nblk + i) * 2]; d += partial[((long long)b * nblk + i) * 2 + 1]; }
        out[2 * b] = a; out[2 * b + 1] = d;
    }
}

// |S| as numpy computes it (np.abs of a complex128 array: x_cabs) -- the magnitudes a caller of the reference would
// hand to run_lws
__global__ void k_cabs(const double2 *S, double *A, long long n)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double2 z = S[i];
        A[i] = x_cabs(z.x, z.y);
    }
}
cudaError_t launch_cabs(const double2 *S, double *A, long long n, cudaStream_t s)
{
    k_cabs<<<(unsigned)std::min<long long>((n + 255) / 256, 148 * 8), 256, 0, s>>>(S, A, n);
    return cudaGetLastError();
}

cudaError_t launch_sq_norms(const double2 *S, const double2 *R, int B, long long n, double *partial, int nblk, double *out,
                            cudaStream_t s)
{
    k_sq_norms<<<dim3(nblk, B), 256, 0, s>>>(S, R, n, partial, nblk);
    k_sq_norms_finish<<<B, 32, 0, s>>>(partial, nblk, out);
    return cudaGetLastError();
}

} // namespace lwsb

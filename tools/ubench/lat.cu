// micro-benchmark: dependent-issue latency of the fp64 operations the LWS chain is made of (one warp)
#include <cstdio>
#include <cuda_runtime.h>
template <int OP> __global__ void k(double *out, double a, double b, int n, long long *cyc)
{
    double x = a + threadIdx.x, y = b;
    long long t0 = clock64();
    #pragma unroll 1
    for (int i = 0; i < n; ++i) {
        #pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (OP == 0) x = __dadd_rn(x, y);
            if (OP == 1) x = __dmul_rn(x, y);
            if (OP == 2) x = __fma_rn(x, y, y);
            if (OP == 3) x = __dsqrt_rn(x) + 1.5;
            if (OP == 4) x = __ddiv_rn(y, x) + 1.5;
            if (OP == 5) { x = __dadd_rn(x, y); y = __dadd_rn(y, 1e-9); }   // 2 independent chains
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x + y;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_lds(double *out, int n, long long *cyc)
{
    __shared__ int idx[1024];
    for (int i = threadIdx.x; i < 1024; i += 32) idx[i] = (i * 17 + 1) & 1023;
    __syncwarp();
    int p = threadIdx.x;
    long long t0 = clock64();
    #pragma unroll 1
    for (int i = 0; i < n * 16; ++i) p = idx[p];
    long long t1 = clock64();
    out[threadIdx.x] = p;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
// throughput: W warps, each with 8 independent DADD chains
__global__ void k_tp(double *out, double b, int n, long long *cyc)
{
    double x[8];
    for (int u = 0; u < 8; ++u) x[u] = threadIdx.x + u;
    __syncthreads();
    long long t0 = clock64();
    #pragma unroll 1
    for (int i = 0; i < n; ++i)
        #pragma unroll
        for (int u = 0; u < 8; ++u) { x[u] = __dadd_rn(x[u], b); x[u] = __dmul_rn(x[u], b); }
    __syncthreads();
    long long t1 = clock64();
    double s = 0; for (int u = 0; u < 8; ++u) s += x[u];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, 8192 * 8); cudaMalloc(&cyc, 8);
    const int n = 2000;
    const char *names[] = {"DADD", "DMUL", "DFMA", "DSQRT(+DADD)", "DDIV(+DADD)", "2xDADD chains"};
    #define RUN(OP) k<OP><<<1, 32>>>(out, 1.0, 1.0000001, n, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-16s %.1f cycles per dependent op\n", names[OP], (double)h / (n * 16.0));
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5)
    k_lds<<<1, 32>>>(out, n, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-16s %.1f cycles per dependent LDS.32\n", "LDS", (double)h / (n * 16.0));
    for (int w = 1; w <= 32; w *= 2) {
        k_tp<<<1, 32 * w>>>(out, 1.0000001, n, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("throughput %2d warps x 8 chains: %.2f fp64 warp-instr per cycle per SM\n", w, (double)w * n * 16.0 / h);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

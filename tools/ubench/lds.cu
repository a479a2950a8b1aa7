// micro-benchmark: LDS.128 throughput per SM for the access patterns of the strip kernel's ring
//   pattern 0: 32 lanes, 32 distinct consecutive 16-byte cells (conflict free: 4 wavefronts)
//   pattern 1: lanes 2p and 2p+1 read the SAME cell (16 distinct cells per warp)
//   pattern 2: lanes 4p..4p+3 read the same cell (8 distinct cells)
//   pattern 3: all lanes the same cell
//   pattern 4: two-way bank conflict inside every quarter-warp (lane l reads cell (l % 4) * 1 + (l / 4 % 2) * 8 + ...)
//   pattern 5: lanes p and p+16 read the same cell (duplicates in different quarter-warps)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, int pattern, int n, long long *cyc)
{
    extern __shared__ double2 sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_double2(i, -i);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int cell;
    switch (pattern) {
    case 0: cell = lane; break;
    case 1: cell = lane >> 1; break;
    case 2: cell = lane >> 2; break;
    case 3: cell = 0; break;
    case 4: cell = (lane & 3) + ((lane >> 2) & 1) * 8 + (lane >> 3) * 16; break; // within a quarter: banks 0-3 twice
    default: cell = lane & 15; break;
    }
    const double2 *p = sm + (warp & 3) * 64 + cell;
    unsigned ax = 0, ay = 0;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            unsigned a, b, c, d;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"((unsigned)__cvta_generic_to_shared(p + u * 64)) : "memory");
            ax ^= a ^ c; ay ^= b ^ d;
        }
    }
    __syncthreads();
    long long t1 = clock64();
    out[threadIdx.x] = (double)(ax + ay);
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    const int n = 2000;
    for (int w = 4; w <= 16; w *= 2)
        for (int pat = 0; pat < 6; ++pat) {
            k<<<1, 32 * w, 65536>>>(out, pat, n, cyc);
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("warps %2d pattern %d: %.2f cycles per warp-level LDS.128 (SM-wide)\n", w, pat, (double)h / (n * 16.0 * w));
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

// The hand-over pattern of k_online_flow in isolation, for compute-sanitizer --tool synccheck: K warps take turns; warp r waits on
// barrier 1 + (r + K - 1) % K for the warp that did the step before and then meets the next warp on barrier 1 + r (64 threads each).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/barpat tools/ubench/barpat.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(int steps, int *out)
{
    constexpr int K = 4;
    __shared__ int cell;
    const int role = threadIdx.x >> 5;
    const int mine = 1 + role, prev = 1 + (role + K - 1) % K;
    if (threadIdx.x == 0) cell = 0;
    __syncthreads();
    for (int b = role; b < steps; b += K) {
        if (b > 0) {
            if (MODE == 0) asm volatile("bar.sync %0, %1;" ::"r"(prev), "r"(64) : "memory");
            else asm volatile("bar.sync %0, 64;" ::"r"(prev) : "memory");
        }
        if ((threadIdx.x & 31) == 0) cell = cell + 1; // the step: must see the previous warp's update
        if (b + 1 < steps) {
            if (MODE == 0) asm volatile("bar.sync %0, %1;" ::"r"(mine), "r"(64) : "memory");
            else if (MODE == 1) asm volatile("bar.sync %0, 64;" ::"r"(mine) : "memory");
            else asm volatile("bar.arrive %0, 64;" ::"r"(mine) : "memory");
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) *out = cell;
}
int main()
{
    int *d, h[3];
    cudaMalloc(&d, 12);
    k<0><<<1, 128>>>(1000, d);
    k<1><<<1, 128>>>(1000, d + 1);
    k<2><<<1, 128>>>(1000, d + 2);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, 12, cudaMemcpyDeviceToHost);
    printf("%s: %d %d %d (expect 1000 each)\n", cudaGetErrorString(e), h[0], h[1], h[2]);
    return 0;
}

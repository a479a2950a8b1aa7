// micro-benchmark / semantics check: tensor memory (TMEM) as a per-lane scratch pad.
// Producer warps 4..7 write with tcgen05.st, consumer warps 0..3 (same lane quarters) read with tcgen05.ld.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define ST16(taddr, v)                                                                                          \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" \
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),   \
                   "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory")
#define LD16(taddr, v)                                                                                          \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),       \
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])  \
                 : "r"(taddr) : "memory")

__global__ void __launch_bounds__(256, 1) k(unsigned *bad, long long *cyc, int reps)
{
    __shared__ uint32_t tbase;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tbase)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tbase;
    const uint32_t quarter = (uint32_t)(warp & 3) * 32u;
    const uint32_t taddr0 = base + (quarter << 16);
    uint32_t v[16];
    unsigned nbad = 0;
    if (warp >= 4) { // producer: lane l of quarter q writes word (q*32+l)*1000 + column
        for (int c = 0; c < 512; c += 16) {
            for (int i = 0; i < 16; ++i) v[i] = (quarter + lane) * 1000u + c + i;
            ST16(taddr0 + c, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp < 4) { // consumer
        for (int c = 0; c < 512; c += 16) {
            LD16(taddr0 + c, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int i = 0; i < 16; ++i) nbad += v[i] != (quarter + lane) * 1000u + c + i;
        }
        atomicAdd(bad, nbad);
    }
    __syncthreads();
    // timing: dependent LD latency (one warp), LD throughput (4 warps), ST throughput (4 warps)
    long long t0 = clock64();
    if (warp == 0) {
        uint32_t a = 0;
        for (int r = 0; r < reps; ++r) {
            LD16(taddr0 + (a & 255u), v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            a += v[0] & 16u;
        }
        if (a == 0xffffffffu) bad[1] = a;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    __syncthreads();
    t0 = clock64();
    if (warp < 4) {
        uint32_t acc = 0;
        for (int r = 0; r < reps; ++r) {
            for (int c = 0; c < 512; c += 16) { LD16(taddr0 + c, v); acc += v[3]; }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
        if (acc == 0x12345u) bad[1] = acc;
    }
    __syncthreads();
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = t1 - t0;
    t0 = clock64();
    if (warp >= 4) {
        for (int r = 0; r < reps; ++r) {
            for (int c = 0; c < 512; c += 16) ST16(taddr0 + c, v);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
    }
    __syncthreads();
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}

int main()
{
    unsigned *bad, hb[2] = {0, 0};
    long long *cyc, hc[3];
    cudaMalloc(&bad, 8); cudaMalloc(&cyc, 24); cudaMemset(bad, 0, 8);
    const int reps = 200;
    k<<<1, 256>>>(bad, cyc, reps);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(hb, bad, 8, cudaMemcpyDeviceToHost); cudaMemcpy(hc, cyc, 24, cudaMemcpyDeviceToHost);
    printf("status %s, mismatches %u\n", cudaGetErrorString(e), hb[0]);
    printf("dependent LD.x16+wait latency: %.1f cycles\n", (double)hc[0] / reps);
    printf("LD throughput, 4 warps: %.1f bytes/cycle/SM\n", 4.0 * 32 * 512 * 4 * reps / (double)hc[1]);
    printf("ST throughput, 4 warps: %.1f bytes/cycle/SM\n", 4.0 * 32 * 512 * 4 * reps / (double)hc[2]);
    return 0;
}

// tcgen05.ld throughput probe (B200): W warps of one CTA per SM read TMEM back to back; bytes per clock per SM
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(512, 1) k(int iters, unsigned long long *cyc, unsigned *sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t r[32];
    unsigned acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint32_t col = (uint32_t)((it * 32 + (warp >> 2) * 64) & 511) & ~31u;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
              "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
              "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(tmem + lane_base + col));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc += r[0] ^ r[31];
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = (unsigned long long)(t1 - t0);
    if (acc == 0xdeadbeefu) *sink = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}
int main() {
    unsigned long long *cyc; unsigned *sink;
    cudaMalloc(&cyc, 8); cudaMalloc(&sink, 4);
    const int iters = 20000;
    for (int warps : {4, 8, 16}) {
        k<<<148, warps * 32>>>(iters, cyc, sink);
        cudaError_t e = cudaDeviceSynchronize();
        unsigned long long c = 0; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        const double bytes = (double)iters * warps * 32 * 32 * 4;
        printf("%2d warps: %s, %llu cycles, %.1f bytes/clk/SM (one load in flight per warp)\n", warps, cudaGetErrorString(e), c, bytes / (double)c);
    }
    return 0;
}

// micro-benchmark: int8 tensor peak of tcgen05.mma kind::i8 on sm_100a (the prefill GEMM's roofline denominator).
// One CTA per SM, one thread issues back-to-back MMAs (M=128, N=256, K=32 per instruction, operands in shared memory in the canonical
// no-swizzle K-major layout, int32 accumulators in TMEM; two accumulators alternate); no loads, no epilogue: the pipe's ceiling.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o i8_peak.bin i8_peak.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
template <int N>
__global__ void __launch_bounds__(128) k(int iters, long long *out) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint32_t tmem_base;
    __shared__ uint64_t bar;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (128 * 128 + N * 128) / 4; i += blockDim.x) ((uint32_t *)sm)[i] = 0x01010101u * (i & 3);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t t0addr = tmem_base;
    long long t0 = 0, t1 = 0;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a = smem_u32(sm), b = smem_u32(sm + 128 * 128);
        t0 = clock64();
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int ks = 0; ks < 4; ks++) {            // K = 128 per tile: 4 x K32
                const uint64_t ad = make_desc(a + ks * 256, 128, 1024), bd = make_desc(b + ks * 256, 128, 1024);
                const uint32_t d = t0addr + (uint32_t)((it & 1) * (N <= 256 ? 256 : 0));
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(1) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(t0addr), "r"(512) : "memory");
}
template <int N> void run(int sms) {
    long long *d; cudaMalloc(&d, 8);
    const int iters = 4000;
    const size_t smem = 128 * 128 + N * 128 + 1024;
    cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<N><<<sms, 128, smem>>>(iters, d); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<N><<<sms, 128, smem>>>(iters, d);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long clk; cudaMemcpy(&clk, d, 8, cudaMemcpyDeviceToHost);
    const double ops = 2.0 * 128 * N * 128 * (double)iters * sms;
    printf("kind::i8 M=128 N=%d K=32 x4 per tile, %d CTAs: %.3f ms, %.1f TOPS (%.1f clk per K32 MMA; err=%s)\n", N, sms, ms, ops / (ms * 1e-3) / 1e12,
           (double)clk / iters / 4, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    run<128>(sms); run<256>(sms);
    return 0;
}

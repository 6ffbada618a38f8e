// micro-benchmark: cost of grid-wide barrier variants on B200 (148 CTAs x 1024 threads), with and without a producer warp
// keeping ~170 KB of cp.async.bulk loads in flight per SM (as the decode-step kernel does).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/barrier_bench tools/micro/barrier_bench.cu && /tmp/barrier_bench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned ld_acquire(const unsigned *p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_relaxed(const unsigned *p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_volatile(const unsigned *p) { unsigned v; asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void red_release(unsigned *p, unsigned v) { asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void red_relaxed(unsigned *p, unsigned v) { asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void st_release(unsigned *p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void st_relaxed(unsigned *p, unsigned v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

template <int VAR>
__device__ __forceinline__ void barrier(unsigned *sync, unsigned *flags, int &nbar, float *data) {
    asm volatile("bar.sync 1, 992;" ::: "memory");
    nbar++;
    const unsigned G = gridDim.x;
    if (VAR == 0) {            // cooperative-groups style: fence, atomic, acquire spin, fence
        if (threadIdx.x == 0) { __threadfence(); atomicAdd(sync, 1u); while (ld_acquire(sync) < nbar * G) {} __threadfence(); }
    } else if (VAR == 1) {     // release reduction + acquire spin (no explicit fences)
        if (threadIdx.x == 0) { red_release(sync, 1u); while (ld_acquire(sync) < nbar * G) {} }
    } else if (VAR == 2) {     // relaxed everything (NOT a correct barrier for data: latency floor)
        if (threadIdx.x == 0) { red_relaxed(sync, 1u); while (ld_relaxed(sync) < nbar * G) {} }
    } else if (VAR == 3) {     // per-CTA flags, all-to-all: release store, warp 0 polls 148 flags with acquire loads
        if (threadIdx.x == 0) st_release(flags + blockIdx.x, (unsigned)nbar);
        if (threadIdx.x < 32) {
            bool ok;
            do { ok = true; for (unsigned i = threadIdx.x; i < G; i += 32) ok &= ld_acquire(flags + i) >= (unsigned)nbar; } while (!__all_sync(0xffffffffu, ok));
        }
    } else if (VAR == 4) {     // per-CTA flags, relaxed (latency floor of the flag scheme)
        if (threadIdx.x == 0) st_relaxed(flags + blockIdx.x, (unsigned)nbar);
        if (threadIdx.x < 32) {
            bool ok;
            do { ok = true; for (unsigned i = threadIdx.x; i < G; i += 32) ok &= ld_relaxed(flags + i) >= (unsigned)nbar; } while (!__all_sync(0xffffffffu, ok));
        }
    } else if (VAR == 5) {     // fence once + relaxed atomic + relaxed spin + no trailing fence (consumers read data with ld.cg)
        if (threadIdx.x == 0) { __threadfence(); red_relaxed(sync, 1u); while (ld_relaxed(sync) < nbar * G) {} }
    }
    asm volatile("bar.sync 1, 992;" ::: "memory");
}

template <int VAR>
__global__ void __launch_bounds__(1024, 1) k(unsigned *sync, unsigned *flags, float *data, const uint8_t *big, size_t big_bytes, int iters, int stream, long long *out) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *full = (uint64_t *)smem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 31) {          // producer: keep 31 x 5.5 KB bulk loads in flight, re-arming as each lands (nobody consumes)
        if (!stream) return;
        if (lane < 31) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[lane])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncwarp();
        if (lane >= 31) return;
        const size_t per = big_bytes / gridDim.x;
        const uint8_t *base = big + per * blockIdx.x;
        size_t off = (size_t)lane * 5632;
        unsigned parity = 0;
        const volatile unsigned *stop = (const volatile unsigned *)(sync + 8);
        while (!*stop) {
            asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(&full[lane])), "r"(5632) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem + 1024 + lane * 5632)), "l"(base + off), "r"(5632), "r"(smem_u32(&full[lane])) : "memory");
            unsigned ok = 0;
            while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&full[lane])), "r"(parity) : "memory");
            parity ^= 1;
            off += 31 * 5632; if (off + 5632 > per) off = (size_t)lane * 5632;
        }
        return;
    }
    int nbar = 0;
    long long t0 = 0;
    for (int it = 0; it < iters + 10; it++) {
        if (it == 10) t0 = clock64();
        // every thread writes a little data before the barrier (like row results) and reads another CTA's after it
        data[(size_t)blockIdx.x * 1024 + threadIdx.x] = (float)it;
        barrier<VAR>(sync, flags, nbar, data);
        const float v = __ldcg(&data[(size_t)((blockIdx.x + 1) % gridDim.x) * 1024 + threadIdx.x]);
        if (v != (float)it && out) atomicAdd((unsigned long long *)&out[1], 1ull);       // stale read counter
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = clock64() - t0; sync[8] = 1; }
    asm volatile("bar.sync 1, 992;" ::: "memory");
    if (threadIdx.x == 0) { __threadfence(); atomicAdd(&sync[9], 1u); while (ld_acquire(&sync[9]) < gridDim.x) {} sync[8] = 1; }
}

template <int VAR> void run(const char *name, int stream, unsigned *sync, unsigned *flags, float *data, uint8_t *big, size_t bb, long long *out, int clock_khz) {
    cudaMemset(sync, 0, 256); cudaMemset(flags, 0, 4096); cudaMemset(out, 0, 16);
    cudaFuncSetAttribute(k<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 2000;
    void *args[] = {&sync, &flags, &data, &big, &bb, (void *)&iters, &stream, &out};
    cudaError_t e = cudaLaunchCooperativeKernel((void *)k<VAR>, dim3(148), dim3(1024), args, 200 * 1024, 0);
    cudaError_t e2 = cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("%-58s streaming=%d : %7.0f cycles = %.2f us per barrier, stale reads %lld  (%s %s)\n", name, stream, (double)h[0] / iters, (double)h[0] / iters / (clock_khz / 1e3), h[1], cudaGetErrorString(e), cudaGetErrorString(e2));
}
int main() {
    unsigned *sync, *flags; float *data; uint8_t *big; long long *out;
    const size_t bb = (size_t)2 << 30;
    cudaMalloc(&sync, 256); cudaMalloc(&flags, 4096); cudaMalloc(&data, 148 * 1024 * 4); cudaMalloc(&big, bb); cudaMalloc(&out, 16);
    cudaMemset(big, 1, bb);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    for (int stream = 0; stream < 2; stream++) {
        run<0>("0 fence + atomicAdd + acquire spin + fence", stream, sync, flags, data, big, bb, out, khz);
        run<1>("1 red.release + ld.acquire spin", stream, sync, flags, data, big, bb, out, khz);
        run<2>("2 red.relaxed + ld.relaxed spin (no ordering)", stream, sync, flags, data, big, bb, out, khz);
        run<3>("3 per-CTA flags: st.release, all-to-all ld.acquire", stream, sync, flags, data, big, bb, out, khz);
        run<4>("4 per-CTA flags relaxed (no ordering)", stream, sync, flags, data, big, bb, out, khz);
        run<5>("5 fence + red.relaxed + ld.relaxed spin", stream, sync, flags, data, big, bb, out, khz);
    }
    return 0;
}

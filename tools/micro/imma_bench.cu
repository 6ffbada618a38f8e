// micro-benchmark: legacy mma.sync int8 (IMMA) latency / throughput on sm_100a, vs dp4a.  nvcc -arch=sm_100a -o imma_bench.bin imma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma_k32(int (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void hmma(float (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
template <int ILP, int KIND>
__global__ void k(int iters, long long *out, int *sink) {
    int c[ILP][4]; float f[ILP][4];
    for (int i = 0; i < ILP; i++) for (int j = 0; j < 4; j++) { c[i][j] = 0; f[i][j] = 0; }
    unsigned a = threadIdx.x * 0x01010101u, b = threadIdx.x + 3;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (KIND == 0) mma_k32(c[i], a, a + 1, a + 2, a + 3, b, b + 1);
            else if (KIND == 1) hmma(f[i], a, a + 1, a + 2, a + 3, b, b + 1);
            else { c[i][0] = __dp4a((int)a, (int)b, c[i][0]); c[i][1] = __dp4a((int)a + 1, (int)b, c[i][1]); c[i][2] = __dp4a((int)a + 2, (int)b, c[i][2]); c[i][3] = __dp4a((int)a + 3, (int)b, c[i][3]); }
        }
    }
    long long t1 = clock64();
    int s = 0; for (int i = 0; i < ILP; i++) for (int j = 0; j < 4; j++) s += c[i][j] + (int)f[i][j];
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    if (s == 0x7fffffff) *sink = s;
}
template <int ILP, int KIND> void run(const char *name, int warps) {
    long long *d; int *s; cudaMalloc(&d, 8); cudaMalloc(&s, 4);
    const int iters = 2000;
    k<ILP, KIND><<<148, warps * 32>>>(iters, d, s); cudaDeviceSynchronize();
    k<ILP, KIND><<<148, warps * 32>>>(iters, d, s); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-6s ILP=%d warps/SM=%2d: %7.2f clk per instr per warp; SM throughput %.3f instr/clk\n", name, ILP, warps, (double)h / iters / ILP, (double)iters * ILP * warps / h);
}
int main() {
    run<1, 0>("imma", 1); run<4, 0>("imma", 1); run<4, 0>("imma", 4); run<4, 0>("imma", 8); run<4, 0>("imma", 16); run<8, 0>("imma", 16);
    run<1, 1>("hmma", 1); run<4, 1>("hmma", 4); run<4, 1>("hmma", 16);
    run<1, 2>("4xdp4a", 1); run<4, 2>("4xdp4a", 4); run<4, 2>("4xdp4a", 16);
    return 0;
}

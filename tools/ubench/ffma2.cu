// Micro-benchmark: FP32 FMA throughput per SM per clock with scalar FFMA, packed FFMA2 (fma.rn.f32x2),
// and each of them interleaved with integer ALU work (issue-slot pressure).  Build: nvcc -arch=sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define ITERS 4096
template <int MODE>
__global__ void __launch_bounds__(256) bench(float* out, float s, int n, long long* clk) {
    float a[8], b = s, c = s * 0.5f;
    u64 p[8];
    unsigned m[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.001f + i; float2 t = make_float2(a[i], a[i] + 1.f); p[i] = *reinterpret_cast<u64*>(&t); }
#pragma unroll
    for (int i = 0; i < 4; ++i) m[i] = threadIdx.x + i;
    float2 bb = make_float2(b, b), cc = make_float2(c, c);
    u64 b2 = *reinterpret_cast<u64*>(&bb), c2 = *reinterpret_cast<u64*>(&cc);
    long long t0 = clock64();
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0 || MODE == 2) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
            if (MODE == 1 || MODE == 3) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(b2), "l"(c2));
            if ((MODE == 2 || MODE == 3) && (i & 1)) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(m[i >> 1]) : "r"(m[(i >> 1) ^ 1]), "r"(it));
        }
    }
    long long t1 = clock64();
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float2 t = *reinterpret_cast<float2*>(&p[i]); r += a[i] + t.x + t.y; }
#pragma unroll
    for (int i = 0; i < 4; ++i) r += (float)m[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
int main() {
    float* out; long long* clk; cudaMalloc(&out, 148 * 8 * 256 * 4); cudaMalloc(&clk, 8);
    const char* names[4] = {"FFMA", "FFMA2", "FFMA + 0.5 LOP3", "FFMA2 + 0.5 LOP3"};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int ctas = 1; ctas <= 8; ctas *= 2)
    for (int mode = 0; mode < 4; ++mode) {
        float ms = 0;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) bench<0><<<148 * ctas, 256>>>(out, 1.0001f, ITERS * 8, clk);
            if (mode == 1) bench<1><<<148 * ctas, 256>>>(out, 1.0001f, ITERS * 8, clk);
            if (mode == 2) bench<2><<<148 * ctas, 256>>>(out, 1.0001f, ITERS * 8, clk);
            if (mode == 3) bench<3><<<148 * ctas, 256>>>(out, 1.0001f, ITERS * 8, clk);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
        }
        double fma_per_thread = (double)ITERS * 8 * 8 * ((mode & 1) ? 2 : 1);
        double cycles = ms * 1e-3 * 1.965e9;
        printf("%-18s warps/SM=%2d  ms=%.3f  FMA/clk/SM=%.1f (at 1965 MHz)  warp-instr/clk/SMSP=%.2f\n", names[mode], ctas * 8, ms,
               fma_per_thread * 256 * ctas / cycles, (double)ITERS * 8 * (8 + ((mode >= 2) ? 4 : 0)) * 8 * ctas / 4 / cycles);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

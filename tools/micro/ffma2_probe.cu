// Development microbenchmark: issue rate of the packed fp32 instructions of sm_100a (FFMA2 / FMUL2 / FADD2)
// against scalar FFMA, alone and mixed with ALU-pipe work.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe ffma2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 pack(float lo, float hi) { u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ float lo_of(u64 v) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); return lo + hi; }

template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, int iters, float a, float b) {
    float t = threadIdx.x * 1e-3f;
    if (MODE == 0) {            // scalar FFMA, 8 chains
        float x[8];
#pragma unroll
        for (int k = 0; k < 8; k++) x[k] = t + k;
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int u = 0; u < 16; u++)
#pragma unroll
                for (int k = 0; k < 8; k++) x[k] = fmaf(x[k], a, b);
        }
        float s = 0; for (int k = 0; k < 8; k++) s += x[k];
        if (s == 123.456f) out[0] = s;
    } else if (MODE == 1) {     // FFMA2, 8 chains of pairs
        u64 x[8]; u64 A = pack(a, a), B = pack(b, b);
#pragma unroll
        for (int k = 0; k < 8; k++) x[k] = pack(t + k, t - k);
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int u = 0; u < 16; u++)
#pragma unroll
                for (int k = 0; k < 8; k++) x[k] = fma2(x[k], A, B);
        }
        float s = 0; for (int k = 0; k < 8; k++) s += lo_of(x[k]);
        if (s == 123.456f) out[0] = s;
    } else if (MODE == 3) {     // scalar FFMA x8 + 8 FMNMX (reference for mode 2)
        float x[8]; int mi[8];
#pragma unroll
        for (int k = 0; k < 8; k++) { x[k] = t + k; mi[k] = threadIdx.x * k; }
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int u = 0; u < 16; u++)
#pragma unroll
                for (int k = 0; k < 8; k++) { x[k] = fmaf(x[k], a, b); asm volatile("min.s32 %0, %0, %1;" : "+r"(mi[k]) : "r"(mi[(k + 3) & 7])); }
        }
        float s = 0; for (int k = 0; k < 8; k++) s += x[k] + mi[k];
        if (s == 123.456f) out[0] = s;
    } else if (MODE == 4) {     // FFMA2 x8 + 8 FMNMX on independent data
        u64 x[8]; u64 A = pack(a, a), B = pack(b, b); int mi[8];
#pragma unroll
        for (int k = 0; k < 8; k++) { x[k] = pack(t + k, t - k); mi[k] = threadIdx.x * k; }
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int u = 0; u < 16; u++)
#pragma unroll
                for (int k = 0; k < 8; k++) { x[k] = fma2(x[k], A, B); asm volatile("min.s32 %0, %0, %1;" : "+r"(mi[k]) : "r"(mi[(k + 3) & 7])); }
        }
        float s = 0; for (int k = 0; k < 8; k++) s += lo_of(x[k]) + mi[k];
        if (s == 123.456f) out[0] = s;
    }
}
template <int MODE> double run(const char* name, double lane_flops_per_inner, double instrs_per_inner) {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    float* out; cudaMalloc(&out, 256);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 2048;
    double best_ms = 1e30;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0); probe<MODE><<<blocks, threads>>>(out, iters, 0.999f, 0.001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best_ms) best_ms = ms;
    }
    double inner = (double)blocks * threads * iters * 16.0 * 8.0;
    double clk = prop.clockRate * 1e3;   // Hz (max)
    double warp_instr_per_clk_per_sm = inner * instrs_per_inner / 32.0 / (best_ms * 1e-3) / clk / prop.multiProcessorCount;
    printf("%-44s %8.3f ms  %7.2f TFLOP/s  %.2f warp-instr/clk/SM (at %.0f MHz nominal)\n", name, best_ms, inner * lane_flops_per_inner / (best_ms * 1e-3) / 1e12,
           warp_instr_per_clk_per_sm, clk / 1e6);
    cudaFree(out);
    return best_ms;
}
int main() {
    run<0>("scalar FFMA", 2, 1);
    run<1>("FFMA2 (packed f32x2)", 4, 1);
    run<3>("scalar FFMA + IMNMX (1:1)", 2, 2);
    run<4>("FFMA2 + IMNMX (1:1)", 4, 2);
    return 0;
}

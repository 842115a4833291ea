// Development microbenchmark: does a mixed stream of scalar FP32 instructions (FFMA / FMUL / FADD) and packed FFMA2
// keep the FMA pipe of an sm_100a SM sub-partition full?  Every pattern below is a sequence of INDEPENDENT dependency chains
// (8 scalar chains s0..s7 and 8 packed chains p0..p7), so the only limits are issue and pipe occupancy.
// Reported: FMA-pipe "lane-cycles" per clock per sub-partition, counting a scalar instruction as 1 and an FFMA2 as 2
// (1.00 = the pipe never idles).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_mix_probe fma_mix_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define S_FFMA(k) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[k]) : "f"(a), "f"(b))
#define S_FMUL(k) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(s[k]) : "f"(a))
#define S_FADD(k) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(s[k]) : "f"(b))
#define P_(k)     asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[k]) : "l"(A), "l"(B))
#define I_(k)     asm volatile("min.s32 %0, %0, %1;" : "+r"(m[k]) : "r"(m[(k + 3) & 7]))

template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, int iters, float a, float b) {
    float s[8]; u64 p[8]; int m[8];
    u64 A, B;
    asm("mov.b64 %0, {%1, %1};" : "=l"(A) : "f"(a)); asm("mov.b64 %0, {%1, %1};" : "=l"(B) : "f"(b));
#pragma unroll
    for (int k = 0; k < 8; k++) { s[k] = threadIdx.x * 1e-3f + k; asm("mov.b64 %0, {%1, %2};" : "=l"(p[k]) : "f"(s[k]), "f"(-s[k])); m[k] = threadIdx.x * (k + 1); }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (MODE == 0) { S_FFMA(0); S_FFMA(1); S_FFMA(2); S_FFMA(3); S_FFMA(4); S_FFMA(5); S_FFMA(6); S_FFMA(7); }                       // 8 S
            if (MODE == 1) { P_(0); P_(1); P_(2); P_(3); P_(4); P_(5); P_(6); P_(7); }                                                       // 8 P
            if (MODE == 2) { S_FFMA(0); P_(0); S_FFMA(1); P_(1); S_FFMA(2); P_(2); S_FFMA(3); P_(3); S_FFMA(4); P_(4); S_FFMA(5); P_(5); S_FFMA(6); P_(6); S_FFMA(7); P_(7); }   // SPSP
            if (MODE == 3) { S_FFMA(0); S_FFMA(1); P_(0); P_(1); S_FFMA(2); S_FFMA(3); P_(2); P_(3); S_FFMA(4); S_FFMA(5); P_(4); P_(5); S_FFMA(6); S_FFMA(7); P_(6); P_(7); }   // SSPP
            if (MODE == 4) { S_FFMA(0); S_FFMA(1); S_FFMA(2); S_FFMA(3); S_FFMA(4); S_FFMA(5); S_FFMA(6); S_FFMA(7); P_(0); P_(1); P_(2); P_(3); P_(4); P_(5); P_(6); P_(7); }   // S8 P8
            if (MODE == 5) { S_FMUL(0); S_FADD(1); P_(0); S_FMUL(2); P_(1); S_FADD(3); S_FMUL(4); P_(2); S_FADD(5); S_FMUL(6); P_(3); S_FADD(7); P_(4); S_FMUL(0); S_FADD(1); P_(5); }  // DK-like 10 S : 6 P, FMUL/FADD
            if (MODE == 6) { S_FFMA(0); P_(0); I_(0); S_FFMA(1); P_(1); I_(1); S_FFMA(2); P_(2); I_(2); S_FFMA(3); P_(3); I_(3); S_FFMA(4); P_(4); I_(4); S_FFMA(5); P_(5); I_(5); S_FFMA(6); P_(6); I_(6); S_FFMA(7); P_(7); I_(7); }  // S P I
            if (MODE == 7) { S_FFMA(0); S_FFMA(1); S_FFMA(2); P_(0); S_FFMA(3); S_FFMA(4); S_FFMA(5); P_(1); S_FFMA(6); S_FFMA(7); S_FFMA(0); P_(2); S_FFMA(1); S_FFMA(2); S_FFMA(3); P_(3); }  // SSSP
            if (MODE == 8) { S_FFMA(0); P_(0); P_(1); P_(2); S_FFMA(1); P_(3); P_(4); P_(5); S_FFMA(2); P_(6); P_(7); P_(0); S_FFMA(3); P_(1); P_(2); P_(3); }                  // SPPP
        }
    }
    float r = 0; for (int k = 0; k < 8; k++) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[k])); r += s[k] + lo + hi + m[k]; }
    if (r == 123.456f) out[0] = r;
}
struct Mix { const char* name; int n_s, n_p, n_i; };
template <int MODE> void run(const Mix& mx, int blocks_per_sm, int threads) {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    float* out; cudaMalloc(&out, 256);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = prop.multiProcessorCount * blocks_per_sm, iters = 4096;
    double best_ms = 1e30;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0); probe<MODE><<<blocks, threads>>>(out, iters, 0.999f, 0.001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best_ms) best_ms = ms;
    }
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double clk = clk_khz * 1e3;
    const double warps_per_smsp = blocks_per_sm * threads / 32.0 / 4.0;
    const double units = (double)iters * 8.0 * warps_per_smsp;                  // pattern units executed per sub-partition
    const double cycles_per_unit = best_ms * 1e-3 * clk / units;
    const double pipe = mx.n_s + 2.0 * mx.n_p, issue = mx.n_s + mx.n_p + mx.n_i;
    printf("%-34s %2.0f warps/SMSP  %8.3f ms  %6.2f clk per unit (issue slots %2.0f, FMA-pipe cycles %2.0f)  pipe busy %.3f  issue busy %.3f\n",
           mx.name, warps_per_smsp, best_ms, cycles_per_unit, issue, pipe, pipe / cycles_per_unit, issue / cycles_per_unit);
    cudaFree(out);
}
template <int MODE> void both(const Mix& mx) { run<MODE>(mx, 4, 256); run<MODE>(mx, 1, 128); }
int main() {
    both<0>({"8 FFMA", 8, 0, 0});
    both<1>({"8 FFMA2", 0, 8, 0});
    both<2>({"(FFMA FFMA2) x8", 8, 8, 0});
    both<3>({"(FFMA FFMA FFMA2 FFMA2) x4", 8, 8, 0});
    both<4>({"FFMA x8, FFMA2 x8", 8, 8, 0});
    both<5>({"DK-like 10 FMUL/FADD : 6 FFMA2", 10, 6, 0});
    both<6>({"(FFMA FFMA2 IMNMX) x8", 8, 8, 8});
    both<7>({"(FFMA FFMA FFMA FFMA2) x4", 12, 4, 0});
    both<8>({"(FFMA FFMA2 FFMA2 FFMA2) x4", 4, 12, 0});
    return 0;
}

// Development microbenchmark: FFMA2 throughput and latency by OPERAND FORM on sm_100a (register pair, scalar register broadcast
// R.F32, swapped halves LO_HI, uniform-register broadcast UR.F32).  8 independent chains per thread for throughput, 1 for latency.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_forms_probe ffma2_forms_probe.cu ; check the forms with cuobjdump -sass
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float lo, float hi) { u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ float lo_(u64 v) { float l, h; asm("mov.b64 {%0, %1}, %2;" : "=f"(l), "=f"(h) : "l"(v)); return l; }
__device__ __forceinline__ float hi_(u64 v) { float l, h; asm("mov.b64 {%0, %1}, %2;" : "=f"(l), "=f"(h) : "l"(v)); return h; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

template <int MODE, int CH>
__global__ void __launch_bounds__(256) probe(float* out, int iters, float ka, float kb, const float* tab) {
    u64 x[CH];
    const float s0 = tab[threadIdx.x & 7], s1 = tab[(threadIdx.x + 1) & 7];       // per-thread scalars (registers, not uniform)
    const u64 A = pk(s0, s1), B = pk(s1, s0), UA = pk(ka, ka), UB = pk(kb, kb), SA = pk(s0, s0), SB = pk(s1, s1);
#pragma unroll
    for (int k = 0; k < CH; k++) x[k] = pk(threadIdx.x * 1e-3f + k, 1.f - k);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++)
#pragma unroll
            for (int k = 0; k < CH; k++) {
                if (MODE == 0) x[k] = fma2(x[k], A, B);                         // pair, pair, pair
                if (MODE == 1) x[k] = fma2(x[k], UA, UB);                       // pair, UR.F32, UR.F32
                if (MODE == 2) x[k] = fma2(x[k], SA, B);                        // pair, R.F32, pair
                if (MODE == 3) x[k] = fma2(x[k], SA, SB);                       // pair, R.F32, R.F32
                if (MODE == 4) x[k] = fma2(pk(hi_(x[k]), lo_(x[k])), A, B);     // pair.LO_HI, pair, pair
                if (MODE == 5) x[k] = fma2(pk(lo_(x[k]), lo_(x[k])), A, UB);    // R.F32 (own lo half), pair, UR.F32
                if (MODE == 6) x[k] = fma2(x[k], SA, UB);                       // pair, R.F32, UR.F32   (the commonest form in the solver)
                if (MODE == 7) x[k] = fma2(x[k], UA, x[k]);                     // pair, UR.F32, pair (add2 form)
            }
    }
    float r = 0;
#pragma unroll
    for (int k = 0; k < CH; k++) r += lo_(x[k]) + hi_(x[k]);
    if (r == 123.456f) out[0] = r;
}
template <int MODE, int CH> void run(const char* name, int blocks_per_sm, int threads) {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    float* out; cudaMalloc(&out, 256);
    float h[8] = {0.999f, 0.998f, 0.997f, 0.996f, 0.995f, 0.994f, 0.993f, 0.992f}; float* tab; cudaMalloc(&tab, 32); cudaMemcpy(tab, h, 32, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = prop.multiProcessorCount * blocks_per_sm, iters = 2048;
    double best = 1e30;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0); probe<MODE, CH><<<blocks, threads>>>(out, iters, 0.999f, 0.001f, tab); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
    }
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double warps = blocks_per_sm * threads / 32.0 / 4.0, n = (double)iters * 16 * CH * warps;
    printf("%-40s chains %d, %2.0f warps/SMSP: %6.2f clk per FFMA2 per sub-partition\n", name, CH, warps, best * 1e-3 * khz * 1e3 / n);
    cudaFree(out); cudaFree(tab);
}
#define ALL(M, NAME) run<M, 8>(NAME, 4, 256); run<M, 1>(NAME, 1, 128);
int main() {
    ALL(0, "pair, pair, pair") ALL(1, "pair, UR.F32, UR.F32") ALL(2, "pair, R.F32, pair") ALL(3, "pair, R.F32, R.F32")
    ALL(4, "pair.LO_HI, pair, pair") ALL(5, "R.F32 (lo of pair), pair, UR.F32") ALL(6, "pair, R.F32, UR.F32") ALL(7, "pair, UR.F32, pair")
    return 0;
}

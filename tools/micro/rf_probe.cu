// Development microbenchmark: how many REGISTER-FILE operands per cycle can an sm_100a sub-partition feed to the FMA pipe?
// The FFMA peak probe (fma_mix_probe.cu, rtb_measure_fp32_peak) uses acc = fma(acc, a, b) with a, b shared by all chains: ptxas marks them
// .reuse, so each FFMA reads ONE register from the file.  Real code (the Durand-Kerner trip) reads up to three fresh registers per FFMA.
// Every mode below runs 8 independent accumulator chains per thread; the SASS (cuobjdump) tells the register numbers and .reuse flags.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/rf_probe tools/micro/rf_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#define FMA3(d, x, y, z) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(x), "f"(y), "f"(z))
#define MUL2(d, x, y)    asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(x), "f"(y))
#define ADD2(d, x, y)    asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(x), "f"(y))

template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, const float* in, int iters) {
    float acc[8], p[8], q[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { acc[k] = in[threadIdx.x + k]; p[k] = in[64 + threadIdx.x + 3 * k]; q[k] = in[128 + threadIdx.x + 5 * k]; }
    const float a = in[300], b = in[301];
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (MODE == 0) FMA3(acc[k], acc[k], a, b);                  // 1 fresh register per FFMA (a, b reused)
                if (MODE == 1) FMA3(acc[k], acc[k], p[k], b);               // 2 fresh
                if (MODE == 2) FMA3(acc[k], p[k], q[k], acc[k]);            // 3 fresh, all distinct
                if (MODE == 3) FMA3(acc[k], p[k], q[(k + 3) & 7], acc[k]);  // 3 fresh, another pairing (other parities)
                if (MODE == 4) MUL2(acc[k], acc[k], p[k]);                  // FMUL, 2 fresh
                if (MODE == 5) ADD2(acc[k], acc[k], p[k]);                  // FADD, 2 fresh
                if (MODE == 6) FMA3(acc[k], p[k], p[k], acc[k]);            // 3 operands, 2 distinct
                if (MODE == 7) { if (k & 1) FMA3(acc[k], p[k], q[k], acc[k]); else MUL2(acc[k], acc[k], p[k]); }   // alternating 3-fresh FFMA / 2-fresh FMUL
                if (MODE == 8) { if (k & 1) FMA3(acc[k], p[k], q[k], acc[k]); else MUL2(acc[k], acc[k], a); }      // alternating 3-fresh FFMA / 1-fresh FMUL
                if (MODE == 9) FMA3(acc[k], p[k], a, acc[k]);               // 2 fresh + reused a
            }
        }
    }
    float r = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) r += acc[k] + p[k] + q[k];
    if (r == 123.456f) out[0] = r;
}
template <int MODE> void run(const char* name, float* out, float* in) {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int iters = 4096;
    for (int bps : {4, 2, 1}) {
        double best = 1e30;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0); probe<MODE><<<prop.multiProcessorCount * bps, 256>>>(out, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
        }
        const double warps = bps * 256 / 32.0 / 4.0, instr = (double)iters * 32.0 * warps;
        printf("mode %d  %-52s %2.0f warps/SMSP  %.3f ms  %.3f clk per instruction\n", MODE, name, warps, best, best * 1e-3 * clk_khz * 1e3 / instr);
    }
}
int main() {
    float *out, *in; cudaMalloc(&out, 256); cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096);
    run<0>("FFMA acc = acc*a + b        (1 fresh operand)", out, in);
    run<1>("FFMA acc = acc*p_k + b      (2 fresh)", out, in);
    run<9>("FFMA acc = p_k*a + acc      (2 fresh)", out, in);
    run<2>("FFMA acc = p_k*q_k + acc    (3 fresh)", out, in);
    run<3>("FFMA acc = p_k*q_k+3 + acc  (3 fresh, other pairing)", out, in);
    run<6>("FFMA acc = p_k*p_k + acc    (2 distinct)", out, in);
    run<4>("FMUL acc = acc*p_k          (2 fresh)", out, in);
    run<5>("FADD acc = acc+p_k          (2 fresh)", out, in);
    run<7>("FFMA 3 fresh / FMUL 2 fresh alternating", out, in);
    run<8>("FFMA 3 fresh / FMUL 1 fresh alternating", out, in);
    return 0;
}

/* fp32_peak.cu — FFMA-only microbenchmark: the measured fp32 CUDA-core peak that the
 * ray-trace kernels' roofline fraction is quoted against (MEASURED_PEAKS.json carries
 * HBM and bf16-tensor figures only; this path has no dense contraction). */
#include <cuda_runtime.h>
#include "../../include/rtb200.h"

namespace {
__global__ void __launch_bounds__(256) ffma_kernel(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    float s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456f) out[0] = s;       /* keeps the chain live without a store in the common case */
}
/* The same chains with THREE distinct register operands per FFMA (acc = p_k * q_k + acc; p_k, q_k loop-invariant registers):
 * the operand-delivery limit of a sub-partition.  The kernel above reads one register per FFMA (a is a uniform register, b sits
 * in the operand-reuse cache); code whose FMAs combine three live values cannot go faster than this one
 * (tools/micro/rf_probe.cu, profiles/r2_rf_probe.txt: 1.07 clk per FFMA there, 1.53-1.77 here). */
__global__ void __launch_bounds__(256) ffma3_kernel(float* out, const float* in, int iters) {
    float acc[8], p[8], q[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { acc[k] = threadIdx.x * 1e-3f + k; p[k] = in[threadIdx.x + 3 * k]; q[k] = in[64 + threadIdx.x + 5 * k]; }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
#pragma unroll
            for (int k = 0; k < 8; k++) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[k]) : "f"(p[k]), "f"(q[k]));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; k++) s += acc[k];
    if (s == 123.456f) out[0] = s;
}
}  // namespace

static int measure(int device, double* tflops, bool three_registers) {
    if (!tflops) return RTB_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return RTB_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return RTB_ERR_CUDA;
    float* out = nullptr;
    if (cudaMalloc(&out, 4096) != cudaSuccess) return RTB_ERR_CUDA;
    cudaMemset(out, 0, 4096);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    double best = 0.0;
    for (int rep = 0; rep < 6; rep++) {
        cudaEventRecord(e0);
        if (three_registers) ffma3_kernel<<<blocks, threads>>>(out, out + 64, iters);
        else ffma_kernel<<<blocks, threads>>>(out, iters, 0.999f, 0.001f);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(out); return RTB_ERR_CUDA; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        double fl = (double)blocks * threads * (double)iters * 16.0 * 8.0 * 2.0;
        double tf = fl / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;           /* first launch is warm-up */
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = best;
    return RTB_OK;
}

extern "C" int rtb_measure_fp32_peak(int device, double* tflops) { return measure(device, tflops, false); }
extern "C" int rtb_measure_fp32_peak3(int device, double* tflops) { return measure(device, tflops, true); }

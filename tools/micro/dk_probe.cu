/* dk_probe.cu — the fused build's Durand-Kerner trip in isolation: issue-slot efficiency against warps per sub-partition.
 * Every thread runs TRIPS trips (4 DKstep_f each, no convergence test) on its own torus invariants; the kernel reports
 * SM cycles per warp-trip.  176 instructions per trip means 176 cycles per warp-trip and sub-partition at full issue rate.
 * build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I raytracing-opengl_b200/csrc -o tools/micro/dk_probe tools/micro/dk_probe.cu
 * run:   tools/micro/dk_probe            (prints one line per warps-per-SM setting) */
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define RTB_STRICT 0
#define RTB_NS probe
#include "rt_device.cuh"

using namespace probe;

#ifndef VARIANT
#define VARIANT 0
#endif

/* a copy of DKstep_f (rt_fused.cuh) with switches that remove one suspect each: VARIANT bit 0 = no step-size accumulation
 * (FMNMX3, ALU pipe), bit 1 = the reciprocal replaced by a multiply (no MUFU), bit 2 = no 2^-20 prescale */
DEV void DKstep_v(float& x, float& y, float x1, float y1, float x2, float y2, float x3, float y3, const TorusState& T, float& E) {
    const float u = fmaf(-y, y, x * x), w = x * y;
    const float Ax = fmaf(u, T.al, fmaf(x, T.be, T.k0)), Ay = fmaf(w, T.al2, y * T.be);
    const float fx = fmaf(Ax, Ax, -fmaf(Ay, Ay, fmaf(u, T.ga, fmaf(x, T.de, T.rho))));
    const float fy = fmaf(Ax + Ax, Ay, -fmaf(w, T.ga2, y * T.de));
    const float ax = x - x1, ay = y - y1, bx = x - x2, by = y - y2, cx = x - x3, cy = y - y3;
    const float qx = fmaf(bx, cx, -(by * cy)), qy = fmaf(bx, cy, by * cx);
    const float px = fmaf(ax, qx, -(ay * qy)), py = fmaf(ax, qy, ay * qx);
#if VARIANT & 4
    const float sx = px, sy = py;
#else
    const float sx = px * 9.5367431640625e-07f, sy = py * 9.5367431640625e-07f;
#endif
#if VARIANT & 2
    const float r = fmaf(sx, px, sy * py) * 0.999f;
#else
    const float r = rcp_mufu(fmaf(sx, px, sy * py));
#endif
    const float ix = sx * r, iy = -sy * r;
    const float gx = fmaf(fx, ix, -(fy * iy)), gy = fmaf(fx, iy, fy * ix);
    x -= gx; y -= gy;
#if VARIANT & 1
    E += gx;
#else
    E = max3_nan_abs(E, gx, gy);
#endif
}
/* VARIANT 8: the same step with EVERY multiply and add issued as an FFMA (a*b = fma(a, b, -0), a+b = fma(a, 1, b): exact identities;
 * the constants come from kernel-argument registers so that ptxas cannot fold them back) — does the FMUL/FADD/FFMA mix cost issue slots? */
__device__ float g_one, g_nzero;
DEV float FM(float a, float b, float nz) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(nz)); return r; }
DEV float FA(float a, float b, float one) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(one), "f"(b)); return r; }
DEV void DKstep_a(float& x, float& y, float x1, float y1, float x2, float y2, float x3, float y3, const TorusState& T, float& E, float one, float nz) {
    const float u = fmaf(-y, y, FM(x, x, nz)), w = FM(x, y, nz);
    const float Ax = fmaf(u, T.al, fmaf(x, T.be, T.k0)), Ay = fmaf(w, T.al2, FM(y, T.be, nz));
    const float fx = fmaf(Ax, Ax, -fmaf(Ay, Ay, fmaf(u, T.ga, fmaf(x, T.de, T.rho))));
    const float fy = fmaf(FA(Ax, Ax, one), Ay, -fmaf(w, T.ga2, FM(y, T.de, nz)));
    const float ax = FA(x, -x1, one), ay = FA(y, -y1, one), bx = FA(x, -x2, one), by = FA(y, -y2, one), cx = FA(x, -x3, one), cy = FA(y, -y3, one);
    const float qx = fmaf(bx, cx, -FM(by, cy, nz)), qy = fmaf(bx, cy, FM(by, cx, nz));
    const float px = fmaf(ax, qx, -FM(ay, qy, nz)), py = fmaf(ax, qy, FM(ay, qx, nz));
    const float sx = FM(px, 9.5367431640625e-07f, nz), sy = FM(py, 9.5367431640625e-07f, nz);
    const float r = rcp_mufu(fmaf(sx, px, FM(sy, py, nz)));
    const float ix = FM(sx, r, nz), iy = FM(-sy, r, nz);
    const float gx = fmaf(fx, ix, -FM(fy, iy, nz)), gy = fmaf(fx, iy, FM(fy, ix, nz));
    x = FA(x, -gx, one); y = FA(y, -gy, one);
    E = max3_nan_abs(E, gx, gy);
}
/* VARIANT 16: the FFMA-heavy, naturally paired parts as packed f32x2 instructions (an FFMA2 reads three 64-bit register pairs for two
 * FMAs: half the register-file traffic per FMA): roots live in (re, im) pairs, A and B of cTorus are evaluated as pairs, the three
 * differences and the update are FADD2; the complex products stay scalar (their packed form needs a swap and a sign). */
DEV f2 ffma2(f2 a, f2 b, f2 c) { return fma2(a, b, c); }
DEV f2 fadd2(f2 a, f2 b) { f2 r; asm("add.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
DEV f2 fsub2(f2 a, f2 b) { f2 r; asm("sub.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
DEV f2 fmul2(f2 a, f2 b) { f2 r; asm("mul.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
struct TorusP { f2 be, k0, al, de, rho, ga; };
DEV void DKstep_p2(f2& c, f2 c1, f2 c2, f2 c3, const TorusP& T, float& E) {
    const float x = lo(c), y = hi(c);
    const f2 uw = pk(fmaf(-y, y, x * x), x * y);
    const f2 A = ffma2(uw, T.al, ffma2(c, T.be, T.k0));                 /* (al u + be x + k0, al2 w + be y) */
    const f2 B = ffma2(uw, T.ga, ffma2(c, T.de, T.rho));                /* (ga u + de x + rho, ga2 w + de y) */
    const float Ax = lo(A), Ay = hi(A);
    const float fx = fmaf(Ax, Ax, -fmaf(Ay, Ay, lo(B)));
    const float fy = fmaf(Ax + Ax, Ay, -hi(B));
    const f2 a = fsub2(c, c1), b = fsub2(c, c2), d = fsub2(c, c3);
    const float ax = lo(a), ay = hi(a), bx = lo(b), by = hi(b), cx = lo(d), cy = hi(d);
    const float qx = fmaf(bx, cx, -(by * cy)), qy = fmaf(bx, cy, by * cx);
    const float px = fmaf(ax, qx, -(ay * qy)), py = fmaf(ax, qy, ay * qx);
    const float sx = px * 9.5367431640625e-07f, sy = py * 9.5367431640625e-07f;
    const float r = rcp_mufu(fmaf(sx, px, sy * py));
    const float ix = sx * r, iy = -sy * r;
    const float gx = fmaf(fx, ix, -(fy * iy)), gy = fmaf(fx, iy, fy * ix);
    c = fsub2(c, pk(gx, gy));
    E = max3_nan_abs(E, gx, gy);
}
/* VARIANT 17: only cTorus packed — (x, x) * (x, y) = (x^2, x y) as one FMUL2, u = x^2 - y^2 written over the low half, then A and B as two
 * FFMA2 each on the pairs (u, w) and (x, y); differences, products, inverse and update scalar.  Roots live in (re, im) pairs. */
DEV void DKstep_p3(f2& c, f2 c1, f2 c2, f2 c3, const TorusP& T, float& E) {
    const float x = lo(c), y = hi(c);
    const f2 sq = fmul2(pk(x, x), c);                                  /* (x^2, x y) */
    const f2 uw = pk(fmaf(-y, y, lo(sq)), hi(sq));
    const f2 A = ffma2(uw, T.al, ffma2(c, T.be, T.k0));
    const f2 B = ffma2(uw, T.ga, ffma2(c, T.de, T.rho));
    const float Ax = lo(A), Ay = hi(A);
    const float fx = fmaf(Ax, Ax, -fmaf(Ay, Ay, lo(B)));
    const float fy = fmaf(Ax + Ax, Ay, -hi(B));
    const float ax = x - lo(c1), ay = y - hi(c1), bx = x - lo(c2), by = y - hi(c2), cx = x - lo(c3), cy = y - hi(c3);
    const float qx = fmaf(bx, cx, -(by * cy)), qy = fmaf(bx, cy, by * cx);
    const float px = fmaf(ax, qx, -(ay * qy)), py = fmaf(ax, qy, ay * qx);
    const float sx = px * 9.5367431640625e-07f, sy = py * 9.5367431640625e-07f;
    const float r = rcp_mufu(fmaf(sx, px, sy * py));
    const float ix = sx * r, iy = -sy * r;
    const float gx = fmaf(fx, ix, -(fy * iy)), gy = fmaf(fx, iy, fy * ix);
    c = pk(x - gx, y - gy);
    E = max3_nan_abs(E, gx, gy);
}
#if VARIANT == 8
#define DKstep_f(a, b, c, d, e, f, g, h, T, E) DKstep_a(a, b, c, d, e, f, g, h, T, E, one, nz)
#elif VARIANT && VARIANT != 17
#define DKstep_f DKstep_v
#endif

template <int TRIPS>
__global__ void __launch_bounds__(1024) dk_kernel(const float* __restrict__ in, float* __restrict__ out, unsigned long long* cycles) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    TorusState T;
    const float s = in[tid & 1023];
    T.al = 1.0f + 1e-7f * s; T.al2 = 2.f * T.al; T.be = -60.f + s; T.k0 = 900.f + s; T.ga = 3.9f + 0.01f * s; T.ga2 = 2.f * T.ga; T.de = -200.f + s; T.rho = 3000.f + s;
    float x0 = 1.f, y0 = 0.f, x1 = 0.4f, y1 = 0.9f, x2 = -0.65f, y2 = 0.72f, x3 = -0.908f, y3 = -0.297f;
    float E = 0.f;
    const float one = in[1023] + 1.0f - in[1023], nz = -0.0f * in[1022];       /* 1 and -0 the compiler cannot see */
    (void)one; (void)nz;
    __syncthreads();
    const long long t0 = clock64();
#if VARIANT == 16 || VARIANT == 17
#if VARIANT == 17
#define DKstep_p2 DKstep_p3
#endif
    TorusP TP;
    TP.be = pk(T.be, T.be); TP.k0 = pk(T.k0, 0.f); TP.al = pk(T.al, T.al2); TP.de = pk(T.de, T.de); TP.rho = pk(T.rho, 0.f); TP.ga = pk(T.ga, T.ga2);
    f2 c0 = pk(x0, y0), c1 = pk(x1, y1), c2 = pk(x2, y2), c3 = pk(x3, y3);
#pragma unroll 1
    for (int k = 0; k < TRIPS; k++) {
        E = 0.f;
        DKstep_p2(c0, c1, c2, c3, TP, E);
        DKstep_p2(c1, c2, c3, c0, TP, E);
        DKstep_p2(c2, c3, c0, c1, TP, E);
        DKstep_p2(c3, c0, c1, c2, TP, E);
        if (E == 12345.f) break;
    }
    x0 = lo(c0) + lo(c1) + lo(c2) + lo(c3); y0 = hi(c0) + hi(c1) + hi(c2) + hi(c3);
#else
#pragma unroll 1
    for (int k = 0; k < TRIPS; k++) {
        E = 0.f;
        DKstep_f(x0, y0, x1, y1, x2, y2, x3, y3, T, E);
        DKstep_f(x1, y1, x2, y2, x3, y3, x0, y0, T, E);
        DKstep_f(x2, y2, x3, y3, x0, y0, x1, y1, T, E);
        DKstep_f(x3, y3, x0, y0, x1, y1, x2, y2, T, E);
        if (E == 12345.f) break;                     /* keeps the loop shape of the real solve: one compare + branch per trip */
    }
#endif
    const long long t1 = clock64();
    out[tid] = x0 + y0 + x1 + y1 + x2 + y2 + x3 + y3 + E;
    if (threadIdx.x == 0) atomicMax(cycles + 0, (unsigned long long)(t1 - t0));
}

int main() {
    float *in, *out; unsigned long long* cyc;
    cudaMalloc(&in, 1024 * 4); cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    float h[1024]; for (int i = 0; i < 1024; i++) h[i] = (float)(i % 97) * 0.01f;
    cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    constexpr int TRIPS = 4000;
    for (int threads : { 128, 256, 384, 512, 640, 768, 1024 }) {
        for (int rep = 0; rep < 2; rep++) {
            cudaMemset(cyc, 0, 8);
            dk_kernel<TRIPS><<<148, threads>>>(in, out, cyc);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        }
        unsigned long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        const double warps_per_smsp = threads / 32.0 / 4.0;
        const double cyc_per_warp_trip = (double)c / TRIPS / warps_per_smsp;
        printf("variant %d  threads/SM %4d  warps/SMSP %.0f  cycles per warp-trip %.1f  (176 = every issue slot used)  issue efficiency %.3f\n",
               VARIANT, threads, warps_per_smsp, cyc_per_warp_trip, 176.0 / cyc_per_warp_trip);
    }
    return 0;
}

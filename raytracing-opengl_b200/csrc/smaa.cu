/* smaa.cu — the SMAA 1x post-pass of the reference as three sm_100a kernels.
 *
 * The reference chains three full-screen draws behind the ray-trace pass (GLWrapper.cpp:173-204) with the pixel shaders of
 * assets/shaders/SMAA.h: luma edge detection (:689), blending-weight calculation (:1145, with the diagonal :919 and corner
 * :1108 patterns of the HIGH / ULTRA presets) and neighbourhood blending (:1252); render targets RGBA8 / RG8 / RGBA8
 * (GLWrapper.cpp:127-129), lookup tables AreaTex (RG8 160x560) and SearchTex (R8 64x16) (SMAA_Builder.h:45-79).
 * main.cpp:32 enables the ULTRA preset.
 *
 * Here: one thread per pixel in all three passes; the "textures" are plain unorm8 arrays in HBM read through the
 * read-only path, filtered in fp32 exactly as oracle/smaa_prelude.h states the sampler (LINEAR, CLAMP_TO_EDGE, centres at
 * (i + 0.5) / size) — a fetch whose bilinear weight is exactly 0 is skipped, which cannot change the value.  The passes are
 * HBM / L2-latency bound (no arithmetic to speak of): pass 1 reads 4 B and writes 2 B per pixel, pass 2 reads 2 B and
 * writes 4 B (plus the searches of the few edge pixels), pass 3 reads 8 B and writes 4 B — 24 B per pixel algorithmic.
 * Compiled with -fmad=false: the same separately rounded fp32 operations as the CPU checker, so the 8-bit results agree.
 */
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "rtb_ctx.h"

namespace {

struct SmaaParams {
    const uchar4* color;      /* RGBA8 [h][w] */
    uchar2* edges;            /* RG8   [h][w] */
    uchar4* blend;            /* RGBA8 [h][w] */
    uchar4* out;              /* RGBA8 [h][w] */
    const uchar2* area;       /* RG8 160 x 560 */
    const unsigned char* search;   /* R8 64 x 16 */
    unsigned* edge_list;      /* pass 2: indices of the pixels whose edges fetch is not zero */
    unsigned* edge_count;
    const float* uv;          /* texel-centre coordinates: u[x] = (x + 0.5) / w for x < w, then v[y] = (y + 0.5) / h (the vertex stage's
                                 interpolated texcoord; one IEEE division per column / row instead of two per pixel and pass) */
    int w, h;
    float rt_x, rt_y, rt_z, rt_w;  /* SMAA_RT_METRICS = (1/W, 1/H, W, H), SMAA_Builder.h:33-35 */
    float threshold;          /* presets, SMAA.h:304-324 */
    int max_steps, max_steps_diag, corner_rounding;   /* max_steps_diag = 0: no diagonal detection; corner_rounding < 0: no corner detection */
};

constexpr int AREA_W = 160, AREA_H = 560, SEARCH_W = 64, SEARCH_H = 16;
constexpr float AREATEX_MAX_DISTANCE = 16.f, AREATEX_MAX_DISTANCE_DIAG = 20.f, AREATEX_SUBTEX_SIZE = 1.0f / 7.0f;

/* unorm8 -> float: v / 255.0f, correctly rounded, without the division: one product and one Markstein correction (two FFMA —
 * explicit fmaf() is not a contraction, so -fmad=false leaves it alone).  Equal to the IEEE quotient for all 256 inputs
 * (tests/test_smaa.py::test_unorm8_decode_is_the_ieee_quotient runs the same sequence on the CPU). */
__device__ __forceinline__ float u2f(unsigned char v) {
    /* (float)v without the conversion unit (the XU pipe ran at 60-70 % in these passes): 2^23 + v is exact in fp32 */
    const float x = __uint_as_float(0x4B000000u | (unsigned)v) - 8388608.0f, rcp = 1.0f / 255.0f;
    const float q = x * rcp;
    return fmaf(fmaf(-q, 255.0f, x), rcp, q);
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ unsigned char unorm8(float v) {
    v = v < 0.f ? 0.f : (v > 1.f ? 1.f : v);
    if (!(v == v)) v = 0.f;
    /* (unsigned char)(v * 255 + 0.5) without the conversion unit: RD(t + 2^23) = floor(t) + 2^23 for t in [0.5, 255.5] */
    return (unsigned char)(__float_as_uint(__fadd_rd(v * 255.0f + 0.5f, 8388608.0f)) & 0xffu);
}

/* ---- the sampler: LINEAR + CLAMP_TO_EDGE on unorm8 texels, one template per texel type ---- */
struct F2 { float x, y; };
struct F4 { float x, y, z, w; };
__device__ __forceinline__ F4 texel4(const uchar4* t, int w, int h, int x, int y) {
    const uchar4 p = __ldg(t + (size_t)clampi(y, 0, h - 1) * w + clampi(x, 0, w - 1));
    return { u2f(p.x), u2f(p.y), u2f(p.z), u2f(p.w) };
}
__device__ __forceinline__ F2 texel2(const uchar2* t, int w, int h, int x, int y) {
    const uchar2 p = __ldg(t + (size_t)clampi(y, 0, h - 1) * w + clampi(x, 0, w - 1));
    return { u2f(p.x), u2f(p.y) };
}
__device__ __forceinline__ float texel1(const unsigned char* t, int w, int h, int x, int y) {
    return u2f(__ldg(t + (size_t)clampi(y, 0, h - 1) * w + clampi(x, 0, w - 1)));
}
struct Footprint { int x0, y0; float ax, ay; };
__device__ __forceinline__ Footprint footprint(float u, float v, int w, int h) {
    const float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
    /* floor() and the conversion to int through one round-down add (FMA pipe instead of two XU instructions): for
     * |x| < 2^22, RD(x + 1.5 * 2^23) = floor(x) + 1.5 * 2^23 exactly, with the integer in the low mantissa bits */
    const float tx = __fadd_rd(x, 12582912.0f), ty = __fadd_rd(y, 12582912.0f);
    const float fx = tx - 12582912.0f, fy = ty - 12582912.0f;
    return { __float_as_int(tx) - 0x4B400000, __float_as_int(ty) - 0x4B400000, x - fx, y - fy };
}
/* (a (1 - t) + b t): with t == 0 this is a exactly, whatever b is — the fetch of b is skipped then */
#define LERP1(a, b, t) ((a) * (1.0f - (t)) + (b) * (t))
__device__ __forceinline__ F2 sample2(const uchar2* t, int w, int h, float u, float v) {
    const Footprint f = footprint(u, v, w, h);
    F2 top = texel2(t, w, h, f.x0, f.y0);
    if (f.ax != 0.f) { const F2 b = texel2(t, w, h, f.x0 + 1, f.y0); top = { LERP1(top.x, b.x, f.ax), LERP1(top.y, b.y, f.ax) }; }
    if (f.ay == 0.f) return top;
    F2 bot = texel2(t, w, h, f.x0, f.y0 + 1);
    if (f.ax != 0.f) { const F2 b = texel2(t, w, h, f.x0 + 1, f.y0 + 1); bot = { LERP1(bot.x, b.x, f.ax), LERP1(bot.y, b.y, f.ax) }; }
    return { LERP1(top.x, bot.x, f.ay), LERP1(top.y, bot.y, f.ay) };
}
__device__ __forceinline__ F4 sample4(const uchar4* t, int w, int h, float u, float v) {
    const Footprint f = footprint(u, v, w, h);
    F4 top = texel4(t, w, h, f.x0, f.y0);
    if (f.ax != 0.f) { const F4 b = texel4(t, w, h, f.x0 + 1, f.y0); top = { LERP1(top.x, b.x, f.ax), LERP1(top.y, b.y, f.ax), LERP1(top.z, b.z, f.ax), LERP1(top.w, b.w, f.ax) }; }
    if (f.ay == 0.f) return top;
    F4 bot = texel4(t, w, h, f.x0, f.y0 + 1);
    if (f.ax != 0.f) { const F4 b = texel4(t, w, h, f.x0 + 1, f.y0 + 1); bot = { LERP1(bot.x, b.x, f.ax), LERP1(bot.y, b.y, f.ax), LERP1(bot.z, b.z, f.ax), LERP1(bot.w, b.w, f.ax) }; }
    return { LERP1(top.x, bot.x, f.ay), LERP1(top.y, bot.y, f.ay), LERP1(top.z, bot.z, f.ay), LERP1(top.w, bot.w, f.ay) };
}
__device__ __forceinline__ float sample1(const unsigned char* t, int w, int h, float u, float v) {
    const Footprint f = footprint(u, v, w, h);
    float top = texel1(t, w, h, f.x0, f.y0);
    if (f.ax != 0.f) top = LERP1(top, texel1(t, w, h, f.x0 + 1, f.y0), f.ax);
    if (f.ay == 0.f) return top;
    float bot = texel1(t, w, h, f.x0, f.y0 + 1);
    if (f.ax != 0.f) bot = LERP1(bot, texel1(t, w, h, f.x0 + 1, f.y0 + 1), f.ax);
    return LERP1(top, bot, f.ay);
}
__device__ __forceinline__ float roundh(float a) { return floorf(a + 0.5f); }
__device__ __forceinline__ float stepf(float edge, float x) { return x < edge ? 0.f : 1.f; }
__device__ __forceinline__ float sat(float a) { return a < 0.f ? 0.f : (a > 1.f ? 1.f : a); }

/* ------------------------------------------------------------------ pass 1: luma edge detection, SMAA.h:689-742 */
__device__ __forceinline__ float luma(const SmaaParams& P, float u, float v) {
    const F4 c = sample4(P.color, P.w, P.h, u, v);
    return c.x * 0.2126f + c.y * 0.7152f + c.z * 0.0722f;
}
__global__ void smaa_edge_kernel(const SmaaParams P) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= P.w || y >= P.h) return;
    const float u = __ldg(P.uv + x), v = __ldg(P.uv + P.w + y);
    /* neighbour coordinates as the vertex stage forms them (SMAA.h:645-650): metrics * offset + texcoord */
    const float ul = P.rt_x * -1.0f + u, vt = P.rt_y * -1.0f + v, ur = P.rt_x * 1.0f + u, vb = P.rt_y * 1.0f + v;
    const float ull = P.rt_x * -2.0f + u, vtt = P.rt_y * -2.0f + v;
    const float L = luma(P, u, v), Lleft = luma(P, ul, v), Ltop = luma(P, u, vt);
    const float dx = fabsf(L - Lleft), dy = fabsf(L - Ltop);
    float ex = stepf(P.threshold, dx), ey = stepf(P.threshold, dy);
    uchar2 o = make_uchar2(0, 0);                                    /* `discard` leaves the cleared target */
    if (ex * 1.0f + ey * 1.0f != 0.0f) {
        const float Lright = luma(P, ur, v), Lbottom = luma(P, u, vb);
        float mx = fmaxf(dx, fabsf(L - Lright)), my = fmaxf(dy, fabsf(L - Lbottom));
        const float Lleftleft = luma(P, ull, v), Ltoptop = luma(P, u, vtt);
        mx = fmaxf(mx, fabsf(Lleft - Lleftleft));
        my = fmaxf(my, fabsf(Ltop - Ltoptop));
        const float final_delta = fmaxf(mx, my);
        ex *= stepf(final_delta, 2.0f * dx);                         /* local contrast adaptation, factor 2 (SMAA.h:404) */
        ey *= stepf(final_delta, 2.0f * dy);
        o = make_uchar2(unorm8(ex), unorm8(ey));
    }
    P.edges[(size_t)y * P.w + x] = o;
}

/* ------------------------------------------------------------------ pass 2: blending weights, SMAA.h:836-1247 */
/* pass 2 fetches edges at ~40 places: one out-of-line copy of the sampler keeps the kernel inside the instruction cache
 * (inlined everywhere it was 4 000 instructions and stalled on instruction fetch 12 warps per issue) */
#ifndef SMAA_EDGES_INLINE
#define SMAA_EDGES_INLINE 0
#endif
#if SMAA_EDGES_INLINE
#define SMAA_P2_INLINE __forceinline__
#else
#define SMAA_P2_INLINE __noinline__
#endif
__device__ SMAA_P2_INLINE F2 edges_at(const SmaaParams& P, float u, float v) { return sample2(P.edges, P.w, P.h, u, v); }
__device__ __forceinline__ F2 edges_off(const SmaaParams& P, float u, float v, int ox, int oy) {      /* textureLodOffset */
    return edges_at(P, u + (float)ox * P.rt_x, v + (float)oy * P.rt_y);
}
__device__ SMAA_P2_INLINE F2 area_at(const SmaaParams& P, float u, float v) { return sample2(P.area, AREA_W, AREA_H, u, v); }
/* two binary values out of one bilinear fetch at a quarter-pixel offset, SMAA.h:836-858 */
__device__ __forceinline__ float decode_r(float r) { return roundh(r * fabsf(5.0f * r - 5.0f * 0.75f)); }

/* diagonal searches, SMAA.h:862-892: returns (steps, last weight); e = the last edges fetched.
 * (Fetching several steps ahead — the coordinates do not depend on what was fetched — was measured and is slower: 0.38 / 0.43 /
 * 0.49 / 0.57 ms per 4K frame for 1 / 2 / 4 / 8 steps at once.  The pass is bound by instruction issue and fetch, not by the
 * memory round trips.) */
template <bool DECODE>
__device__ F2 search_diag(const SmaaParams& P, float cx, float cy, float dirx, float diry, F2& e) {
    float cz = -1.0f, cw = 1.0f;
    const float last = (float)(P.max_steps_diag - 1);
    while (cz < last && cw > 0.9f) {
        cx = P.rt_x * dirx + cx; cy = P.rt_y * diry + cy; cz = 1.0f * 1.0f + cz;
        e = edges_at(P, cx, cy);
        if (DECODE) e = { decode_r(e.x), roundh(e.y) };
        cw = e.x * 0.5f + e.y * 0.5f;
    }
    return { cz, cw };
}
__device__ __forceinline__ F2 search_diag1(const SmaaParams& P, float u, float v, float dirx, float diry, F2& e) {
    return search_diag<false>(P, u, v, dirx, diry, e);
}
__device__ __forceinline__ F2 search_diag2(const SmaaParams& P, float u, float v, float dirx, float diry, F2& e) {
    return search_diag<true>(P, u + 0.25f * P.rt_x, v, dirx, diry, e);
}
__device__ __forceinline__ F2 area_diag(const SmaaParams& P, float distx, float disty, float ex, float ey, float offset) {   /* SMAA.h:900-914 */
    float tx = AREATEX_MAX_DISTANCE_DIAG * ex + distx, ty = AREATEX_MAX_DISTANCE_DIAG * ey + disty;
    const float psx = 1.0f / 160.0f, psy = 1.0f / 560.0f;
    tx = psx * tx + 0.5f * psx; ty = psy * ty + 0.5f * psy;
    tx += 0.5f;
    ty += AREATEX_SUBTEX_SIZE * offset;
    return area_at(P, tx, ty);
}
__device__ F2 diag_weights(const SmaaParams& P, float u, float v, F2 e) {                       /* SMAA.h:919-989, subsampleIndices = 0 */
    F2 weights = { 0.f, 0.f };
    float dx, dy, dz, dw;
    F2 end = { 0.f, 0.f };
    if (e.x > 0.0f) {
        const F2 r = search_diag1(P, u, v, -1.0f, 1.0f, end);
        dx = r.x; dz = r.y;
        dx += (float)(end.y > 0.9f);
    } else { dx = 0.f; dz = 0.f; }
    { const F2 r = search_diag1(P, u, v, 1.0f, -1.0f, end); dy = r.x; dw = r.y; }
    if (dx + dy > 2.0f) {
        const float c0x = (-dx + 0.25f) * P.rt_x + u, c0y = dx * P.rt_y + v, c1x = dy * P.rt_x + u, c1y = (-dy - 0.25f) * P.rt_y + v;
        const F2 a = edges_off(P, c0x, c0y, -1, 0), b = edges_off(P, c1x, c1y, 1, 0);
        /* c.yxwz = decode(c.xyzw): red / blue (x, z) go through the quarter-offset decoder, all four are rounded */
        const float cy_ = decode_r(a.x), cx_ = roundh(a.y), cw_ = decode_r(b.x), cz_ = roundh(b.y);
        float ccx = 2.0f * cx_ + cy_, ccy = 2.0f * cz_ + cw_;
        if (stepf(0.9f, dz) != 0.f) ccx = 0.f;
        if (stepf(0.9f, dw) != 0.f) ccy = 0.f;
        const F2 ar = area_diag(P, dx, dy, ccx, ccy, 0.f);
        weights.x += ar.x; weights.y += ar.y;
    }
    { const F2 r = search_diag2(P, u, v, -1.0f, -1.0f, end); dx = r.x; dz = r.y; }
    if (edges_off(P, u, v, 1, 0).x > 0.0f) {
        const F2 r = search_diag2(P, u, v, 1.0f, 1.0f, end);
        dy = r.x; dw = r.y;
        dy += (float)(end.y > 0.9f);
    } else { dy = 0.f; dw = 0.f; }
    if (dx + dy > 2.0f) {
        const float c0x = -dx * P.rt_x + u, c0y = -dx * P.rt_y + v, c1x = dy * P.rt_x + u, c1y = dy * P.rt_y + v;
        const float cx_ = edges_off(P, c0x, c0y, -1, 0).y, cy_ = edges_off(P, c0x, c0y, 0, -1).x;
        const F2 zw = edges_off(P, c1x, c1y, 1, 0);
        const float cz_ = zw.y, cw_ = zw.x;
        float ccx = 2.0f * cx_ + cy_, ccy = 2.0f * cz_ + cw_;
        if (stepf(0.9f, dz) != 0.f) ccx = 0.f;
        if (stepf(0.9f, dw) != 0.f) ccy = 0.f;
        const F2 ar = area_diag(P, dx, dy, ccx, ccy, 0.f);
        weights.x += ar.y; weights.y += ar.x;                        /* .gr */
    }
    return weights;
}
/* how far the last search step overshot, from SearchTex (SMAA.h:997-1013) */
__device__ __forceinline__ float search_length(const SmaaParams& P, float ex, float ey, float offset) {
    float sx = 66.0f * 0.5f, sy = 33.0f * -1.0f, bx = 66.0f * offset, by = 33.0f * 1.0f;
    sx += -1.0f; sy += 1.0f; bx += 0.5f; by += -0.5f;
    sx *= 1.0f / 64.0f; sy *= 1.0f / 16.0f; bx *= 1.0f / 64.0f; by *= 1.0f / 16.0f;
    return sample1(P.search, SEARCH_W, SEARCH_H, sx * ex + bx, sy * ey + by);
}
/* the four axis searches, SMAA.h:1019-1085: two pixels per step through one bilinear fetch between them.
 * HORZ: walks u (else v); NEG: towards smaller coordinates.  Returns the coordinate after the last accepted step and the
 * edges fetched there. */
template <bool HORZ, bool NEG>
__device__ float search_axis(const SmaaParams& P, float u, float v, float end, F2& e) {
    const float step = NEG ? -2.0f : 2.0f;
    e = HORZ ? F2{ 0.f, 1.f } : F2{ 1.f, 0.f };
    for (;;) {
        const float pos = HORZ ? u : v, along = HORZ ? e.y : e.x, across = HORZ ? e.x : e.y;
        if (!((NEG ? pos > end : pos < end) && along > 0.8281f && across == 0.0f)) return pos;
        e = edges_at(P, u, v);
        /* texcoord = mad((-2, -0) or (0, 2) ..., metrics, texcoord): the zero product is a signed zero and c + (+-0) = c */
        if (HORZ) u = step * P.rt_x + u;
        else v = step * P.rt_y + v;
    }
}
__device__ float search_x_left(const SmaaParams& P, float u, float v, float end) {
    F2 e;
    u = search_axis<true, true>(P, u, v, end, e);
    const float offset = -(255.0f / 127.0f) * search_length(P, e.x, e.y, 0.0f) + 3.25f;
    return P.rt_x * offset + u;
}
__device__ float search_x_right(const SmaaParams& P, float u, float v, float end) {
    F2 e;
    u = search_axis<true, false>(P, u, v, end, e);
    const float offset = -(255.0f / 127.0f) * search_length(P, e.x, e.y, 0.5f) + 3.25f;
    return -P.rt_x * offset + u;
}
__device__ float search_y_up(const SmaaParams& P, float u, float v, float end) {
    F2 e;
    v = search_axis<false, true>(P, u, v, end, e);
    const float offset = -(255.0f / 127.0f) * search_length(P, e.y, e.x, 0.0f) + 3.25f;
    return P.rt_y * offset + v;
}
__device__ float search_y_down(const SmaaParams& P, float u, float v, float end) {
    F2 e;
    v = search_axis<false, false>(P, u, v, end, e);
    const float offset = -(255.0f / 127.0f) * search_length(P, e.y, e.x, 0.5f) + 3.25f;
    return -P.rt_y * offset + v;
}
__device__ __forceinline__ F2 area(const SmaaParams& P, float distx, float disty, float e1, float e2, float offset) {     /* SMAA.h:1091-1103 */
    float tx = AREATEX_MAX_DISTANCE * roundh(4.0f * e1) + distx, ty = AREATEX_MAX_DISTANCE * roundh(4.0f * e2) + disty;
    const float psx = 1.0f / 160.0f, psy = 1.0f / 560.0f;
    tx = psx * tx + 0.5f * psx; ty = psy * ty + 0.5f * psy;
    ty = AREATEX_SUBTEX_SIZE * offset + ty;
    return area_at(P, tx, ty);
}
/* corner patterns, SMAA.h:1108-1140: which == 0 horizontal (red edges above / below), 1 vertical (green edges left / right) */
__device__ __forceinline__ void corner_pattern(const SmaaParams& P, int which, F2& weights, float c0x, float c0y, float c1x, float c1y, float dx, float dy) {
    if (P.corner_rounding < 0) return;
    const float lx = stepf(dx, dy), ly = stepf(dy, dx);
    const float norm = (float)P.corner_rounding / 100.0f;
    float rx = (1.0f - norm) * lx, ry = (1.0f - norm) * ly;
    const float s = lx + ly;
    rx /= s; ry /= s;
    float fx = 1.0f, fy = 1.0f;
    if (which == 0) {
        fx -= rx * edges_off(P, c0x, c0y, 0, 1).x;
        fx -= ry * edges_off(P, c1x, c1y, 1, 1).x;
        fy -= rx * edges_off(P, c0x, c0y, 0, -2).x;
        fy -= ry * edges_off(P, c1x, c1y, 1, -2).x;
    } else {
        fx -= rx * edges_off(P, c0x, c0y, 1, 0).y;
        fx -= ry * edges_off(P, c1x, c1y, 1, 1).y;
        fy -= rx * edges_off(P, c0x, c0y, -2, 0).y;
        fy -= ry * edges_off(P, c1x, c1y, -2, 1).y;
    }
    weights.x *= sat(fx); weights.y *= sat(fy);
}
/* the pixel shader body of pass 2 for one pixel whose edges fetch `e` is not zero (SMAA.h:1145-1247) */
__device__ uchar4 blend_weights(const SmaaParams& P, int x, int y, F2 e) {
    const float u = __ldg(P.uv + x), v = __ldg(P.uv + P.w + y);
    F4 wgt = { 0.f, 0.f, 0.f, 0.f };
    /* the vertex stage's offsets, SMAA.h:655-668 */
    const float pixx = u * P.rt_z, pixy = v * P.rt_w;
    const float o0x = P.rt_x * -0.25f + u, o0y = P.rt_y * -0.125f + v, o0z = P.rt_x * 1.25f + u, o0w = P.rt_y * -0.125f + v;
    const float o1x = P.rt_x * -0.125f + u, o1y = P.rt_y * -0.25f + v, o1z = P.rt_x * -0.125f + u, o1w = P.rt_y * 1.25f + v;
    const float ms = (float)P.max_steps;
    const float o2x = P.rt_x * (-2.0f * ms) + o0x, o2y = P.rt_x * (2.0f * ms) + o0z, o2z = P.rt_y * (-2.0f * ms) + o1y, o2w = P.rt_y * (2.0f * ms) + o1w;
    if (e.y > 0.0f) {                                                /* edge at north */
        bool axis = true;
        if (P.max_steps_diag > 0) {
            const F2 dw = diag_weights(P, u, v, e);
            wgt.x = dw.x; wgt.y = dw.y;
            axis = wgt.x == -wgt.y;                                  /* no diagonal found: horizontal / vertical processing */
        }
        if (axis) {
            const float cx = search_x_left(P, o0x, o0y, o2x), cy = o1y;
            float dx = cx;
            const float e1 = edges_at(P, cx, cy).x;
            const float cz = search_x_right(P, o0z, o0w, o2y);
            float dy = cz;
            dx = fabsf(roundh(P.rt_z * dx + -pixx)); dy = fabsf(roundh(P.rt_z * dy + -pixx));
            const float sdx = sqrtf(dx), sdy = sqrtf(dy);
            const float e2 = edges_off(P, cz, cy, 1, 0).x;
            F2 w2 = area(P, sdx, sdy, e1, e2, 0.f);
            corner_pattern(P, 0, w2, cx, v, cz, v, dx, dy);
            wgt.x = w2.x; wgt.y = w2.y;
        } else {
            e.x = 0.0f;                                              /* a diagonal was found: skip vertical processing */
        }
    }
    if (e.x > 0.0f) {                                                /* edge at west */
        const float cy = search_y_up(P, o1x, o1y, o2z), cx = o0x;
        float dx = cy;
        const float e1 = edges_at(P, cx, cy).y;
        const float cz = search_y_down(P, o1z, o1w, o2w);
        float dy = cz;
        dx = fabsf(roundh(P.rt_w * dx + -pixy)); dy = fabsf(roundh(P.rt_w * dy + -pixy));
        const float sdx = sqrtf(dx), sdy = sqrtf(dy);
        const float e2 = edges_off(P, cx, cz, 0, 1).y;
        F2 w2 = area(P, sdx, sdy, e1, e2, 0.f);
        corner_pattern(P, 1, w2, u, cy, u, cz, dx, dy);
        wgt.z = w2.x; wgt.w = w2.y;
    }
    return make_uchar4(unorm8(wgt.x), unorm8(wgt.y), unorm8(wgt.z), unorm8(wgt.w));
}
/* pass 2 as the reference draws it: every pixel runs the shader, the few edge pixels with their divergent searches
 * (option "smaa_compact" = 0; kept as the A/B partner and cross-check of the two kernels below) */
__global__ void smaa_blend_kernel(const SmaaParams P) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= P.w || y >= P.h) return;
    const float u = __ldg(P.uv + x), v = __ldg(P.uv + P.w + y);
    const F2 e = edges_at(P, u, v);
    uchar4 o = make_uchar4(0, 0, 0, 0);
    if (e.x > 0.f || e.y > 0.f) o = blend_weights(P, x, y, e);
    P.blend[(size_t)y * P.w + x] = o;
}
/* pass 2, compacted (the default).  2a streams: every pixel evaluates the shader's own predicate (its edges fetch — which, at
 * the quarter-ulp the coordinate arithmetic is off a texel centre, may pick up a neighbour's edge), stores the zero weights
 * of a non-edge pixel and appends an edge pixel to a list (one atomic per warp).  2b runs the searches with every lane on an
 * edge pixel: the latency chains of up to ~200 dependent fetches no longer hold 31 idle lanes each. */
constexpr int CLASSIFY_ROWS = 4;                                     /* rows per thread: four independent fetches in flight */
__global__ void smaa_classify_kernel(const SmaaParams P) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, lane = (threadIdx.y * blockDim.x + threadIdx.x) & 31;
    bool edge[CLASSIFY_ROWS];
    int ys[CLASSIFY_ROWS];
#pragma unroll
    for (int k = 0; k < CLASSIFY_ROWS; k++) {
        const int y = ys[k] = (blockIdx.y * CLASSIFY_ROWS + k) * blockDim.y + threadIdx.y;
        edge[k] = false;
        if (x < P.w && y < P.h) {
            const float u = __ldg(P.uv + x), v = __ldg(P.uv + P.w + y);
            const F2 e = sample2(P.edges, P.w, P.h, u, v);
            edge[k] = e.x > 0.f || e.y > 0.f;
            if (!edge[k]) P.blend[(size_t)y * P.w + x] = make_uchar4(0, 0, 0, 0);
        }
    }
#pragma unroll
    for (int k = 0; k < CLASSIFY_ROWS; k++) {
        const unsigned m = __ballot_sync(0xffffffffu, edge[k]);
        if (m) {
            const int leader = __ffs(m) - 1;
            unsigned base = 0;
            if (lane == leader) base = atomicAdd(P.edge_count, (unsigned)__popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (edge[k]) P.edge_list[base + __popc(m & ((1u << lane) - 1u))] = (unsigned)ys[k] * (unsigned)P.w + (unsigned)x;
        }
    }
}
__global__ void smaa_uv_kernel(float* uv, int w, int h) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < w) uv[i] = ((float)i + 0.5f) / (float)w;
    else if (i < w + h) uv[i] = ((float)(i - w) + 0.5f) / (float)h;
}
__global__ void __launch_bounds__(128) smaa_blend_list_kernel(const SmaaParams P) {
    const unsigned n = *P.edge_count;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned idx = P.edge_list[i];
        const int y = (int)(idx / (unsigned)P.w), x = (int)(idx - (unsigned)y * (unsigned)P.w);
        const float u = __ldg(P.uv + x), v = __ldg(P.uv + P.w + y);
        P.blend[idx] = blend_weights(P, x, y, edges_at(P, u, v));
    }
}

/* ------------------------------------------------------------------ pass 3: neighbourhood blending, SMAA.h:1252-1308 */
__global__ void smaa_neighborhood_kernel(const SmaaParams P) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= P.w || y >= P.h) return;
    const float u = __ldg(P.uv + x), v = __ldg(P.uv + P.w + y);
    /* offset = mad(metrics.xyxy, (1, 0, 0, 1), texcoord.xyxy): the two zero products are +0 (the metrics are positive and
     * finite) and +0 + c = c for the positive coordinates, so .y and .z are v and u themselves — which lets the three fetches
     * at this pixel's own column / row share their footprints */
    const float ox = P.rt_x * 1.0f + u, oy = v, oz = u, ow = P.rt_y * 1.0f + v;
    const float ax = sample4(P.blend, P.w, P.h, ox, oy).w;           /* right  */
    const float ay = sample4(P.blend, P.w, P.h, oz, ow).y;           /* top    */
    const F4 self = sample4(P.blend, P.w, P.h, u, v);
    const float aw = self.x, az = self.z;                            /* bottom / left */
    F4 c;
    if (ax * 1.0f + ay * 1.0f + az * 1.0f + aw * 1.0f < 1e-5f) {
        c = sample4(P.color, P.w, P.h, u, v);
    } else {
        const bool hz = fmaxf(ax, az) > fmaxf(ay, aw);               /* max(horizontal) > max(vertical) */
        float bo_x = 0.f, bo_y = ay, bo_z = 0.f, bo_w = aw, bw_x = ay, bw_y = aw;
        if (hz) { bo_x = ax; bo_y = 0.f; bo_z = az; bo_w = 0.f; bw_x = ax; bw_y = az; }
        const float s = bw_x * 1.0f + bw_y * 1.0f;
        bw_x /= s; bw_y /= s;
        const float c0x = bo_x * P.rt_x + u, c0y = bo_y * P.rt_y + v, c1x = bo_z * -P.rt_x + u, c1y = bo_w * -P.rt_y + v;
        const F4 a = sample4(P.color, P.w, P.h, c0x, c0y), b = sample4(P.color, P.w, P.h, c1x, c1y);
        c = { bw_x * a.x, bw_x * a.y, bw_x * a.z, bw_x * a.w };
        c.x += bw_y * b.x; c.y += bw_y * b.y; c.z += bw_y * b.z; c.w += bw_y * b.w;
    }
    P.out[(size_t)y * P.w + x] = make_uchar4(unorm8(c.x), unorm8(c.y), unorm8(c.z), unorm8(c.w));
}

/* float frame -> the RGBA8 colour target the reference renders into (GLWrapper.cpp:127,161): clamp, scale, round */
__global__ void smaa_quantize_kernel(const float4* __restrict__ src, uchar4* __restrict__ dst, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 v = src[i];
    dst[i] = make_uchar4(unorm8(v.x), unorm8(v.y), unorm8(v.z), unorm8(v.w));
}

}  // namespace

/* presets, SMAA.h:304-324 */
static bool smaa_preset(int preset, SmaaParams& P) {
    switch (preset) {
        case 0: P.threshold = 0.15f; P.max_steps = 4; P.max_steps_diag = 0; P.corner_rounding = -1; return true;
        case 1: P.threshold = 0.1f; P.max_steps = 8; P.max_steps_diag = 0; P.corner_rounding = -1; return true;
        case 2: P.threshold = 0.1f; P.max_steps = 16; P.max_steps_diag = 8; P.corner_rounding = 25; return true;
        case 3: P.threshold = 0.05f; P.max_steps = 32; P.max_steps_diag = 16; P.corner_rounding = 25; return true;
    }
    return false;
}

/* the three passes on ctx->smaa_color (already RGBA8) -> ctx->smaa_out, on `st` */
int rtb_smaa_run(rtb_ctx* ctx, cudaStream_t st) {
    SmaaParams P;
    if (!smaa_preset(ctx->smaa_preset, P)) return rtb_fail(ctx, RTB_ERR_STATE, "SMAA preset %d", ctx->smaa_preset);
    if (!ctx->smaa_area || !ctx->smaa_search) return rtb_fail(ctx, RTB_ERR_STATE, "SMAA is enabled but the area / search tables were never set (rtb_smaa_set_tables)");
    P.color = (const uchar4*)ctx->smaa_color; P.edges = (uchar2*)ctx->smaa_edges; P.blend = (uchar4*)ctx->smaa_blend; P.out = (uchar4*)ctx->smaa_out;
    P.area = (const uchar2*)ctx->smaa_area; P.search = ctx->smaa_search;
    P.edge_list = nullptr; P.edge_count = nullptr; P.uv = ctx->smaa_uv;
    P.w = ctx->width; P.h = ctx->height;
    P.rt_x = 1.0f / (float)P.w; P.rt_y = 1.0f / (float)P.h; P.rt_z = (float)P.w; P.rt_w = (float)P.h;
    dim3 block(32, 8);
    if (const char* e = getenv("RTB_SMAA_BLOCK")) {                      /* development: "WxH", W a multiple of 32 (one warp = one row segment) */
        int bw = 0, bh = 0;
        if (sscanf(e, "%dx%d", &bw, &bh) == 2 && bw >= 32 && bw % 32 == 0 && bh >= 1 && bw * bh <= 1024) block = dim3(bw, bh);
    }
    const dim3 grid((P.w + block.x - 1) / block.x, (P.h + block.y - 1) / block.y);
    if (ctx->smaa_timed) CU(cudaEventRecord(ctx->ev_s0, st));
    smaa_edge_kernel<<<grid, block, 0, st>>>(P);
    if (ctx->opt_smaa_compact) {
        P.edge_list = ctx->smaa_list; P.edge_count = ctx->smaa_count;
        CU(cudaMemsetAsync(ctx->smaa_count, 0, sizeof(unsigned), st));
        smaa_classify_kernel<<<dim3(grid.x, (P.h + block.y * CLASSIFY_ROWS - 1) / (block.y * CLASSIFY_ROWS)), block, 0, st>>>(P);
        smaa_blend_list_kernel<<<ctx->n_sm * 8, 128, 0, st>>>(P);
    } else {
        smaa_blend_kernel<<<grid, block, 0, st>>>(P);
    }
    smaa_neighborhood_kernel<<<grid, block, 0, st>>>(P);
    CU(cudaGetLastError());
    if (ctx->smaa_timed) CU(cudaEventRecord(ctx->ev_s1, st));
    return RTB_OK;
}

int rtb_smaa_alloc(rtb_ctx* ctx) {
    const size_t px = (size_t)ctx->width * ctx->height;
    if (!ctx->smaa_color) CU(cudaMalloc(&ctx->smaa_color, px * 4));
    if (!ctx->smaa_edges) CU(cudaMalloc(&ctx->smaa_edges, px * 2));
    if (!ctx->smaa_list) CU(cudaMalloc(&ctx->smaa_list, px * sizeof(unsigned)));
    if (!ctx->smaa_count) CU(cudaMalloc(&ctx->smaa_count, sizeof(unsigned)));
    if (!ctx->smaa_uv) {
        CU(cudaMalloc(&ctx->smaa_uv, (size_t)(ctx->width + ctx->height) * sizeof(float)));
        smaa_uv_kernel<<<(ctx->width + ctx->height + 255) / 256, 256, 0, ctx->stream>>>(ctx->smaa_uv, ctx->width, ctx->height);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(ctx->stream));      /* the passes may run on a caller's stream */
    }
    if (!ctx->smaa_blend) CU(cudaMalloc(&ctx->smaa_blend, px * 4));
    if (!ctx->smaa_out) CU(cudaMalloc(&ctx->smaa_out, px * 4));
    if (!ctx->ev_s0) { CU(cudaEventCreate(&ctx->ev_s0)); CU(cudaEventCreate(&ctx->ev_s1)); }
    return RTB_OK;
}

/* after a frame: quantise the float frame into the RGBA8 colour target, then the three passes */
int rtb_smaa_after_frame(rtb_ctx* ctx, const float* frame, cudaStream_t st) {
    int rc = rtb_smaa_alloc(ctx);
    if (rc) return rc;
    const size_t n = (size_t)ctx->width * ctx->height;
    smaa_quantize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const float4*)frame, (uchar4*)ctx->smaa_color, n);
    CU(cudaGetLastError());
    ctx->smaa_timed = true;
    return rtb_smaa_run(ctx, st);
}

void rtb_smaa_release(rtb_ctx* ctx) {
    for (uint8_t** p : { &ctx->smaa_color, &ctx->smaa_edges, &ctx->smaa_blend, &ctx->smaa_out, &ctx->smaa_area, &ctx->smaa_search }) { if (*p) cudaFree(*p); *p = nullptr; }
    for (unsigned** p : { &ctx->smaa_list, &ctx->smaa_count }) { if (*p) cudaFree(*p); *p = nullptr; }
    if (ctx->smaa_uv) { cudaFree(ctx->smaa_uv); ctx->smaa_uv = nullptr; }
    if (ctx->ev_s0) { cudaEventDestroy(ctx->ev_s0); cudaEventDestroy(ctx->ev_s1); ctx->ev_s0 = ctx->ev_s1 = nullptr; }
}

extern "C" {

int rtb_smaa_set_tables(rtb_ctx* ctx, const uint8_t* area_rg8, const uint8_t* search_r8) {
    if (!ctx || !area_rg8 || !search_r8) return rtb_fail(ctx, RTB_ERR_INVALID, "null argument");
    CU(cudaSetDevice(ctx->device));
    if (!ctx->smaa_area) CU(cudaMalloc(&ctx->smaa_area, (size_t)AREA_W * AREA_H * 2));
    if (!ctx->smaa_search) CU(cudaMalloc(&ctx->smaa_search, (size_t)SEARCH_W * SEARCH_H));
    CU(cudaMemcpy(ctx->smaa_area, area_rg8, (size_t)AREA_W * AREA_H * 2, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->smaa_search, search_r8, (size_t)SEARCH_W * SEARCH_H, cudaMemcpyHostToDevice));
    return RTB_OK;
}

int rtb_enable_smaa(rtb_ctx* ctx, int preset) {
    if (!ctx) return rtb_fail(nullptr, RTB_ERR_INVALID, "null context");
    if (preset < -1 || preset > 3) return rtb_fail(ctx, RTB_ERR_INVALID, "SMAA preset must be -1 (off) or 0..3 (LOW, MEDIUM, HIGH, ULTRA)");
    ctx->smaa_preset = preset;
    return RTB_OK;
}

int rtb_smaa_apply(rtb_ctx* ctx, const uint8_t* rgba8_host, uint8_t* out_host, uint8_t* edges_host, uint8_t* blend_host, float* ms) {
    if (!ctx || !rgba8_host) return rtb_fail(ctx, RTB_ERR_INVALID, "null argument");
    if (ctx->smaa_preset < 0) return rtb_fail(ctx, RTB_ERR_STATE, "SMAA is off (rtb_enable_smaa)");
    CU(cudaSetDevice(ctx->device));
    int rc = rtb_smaa_alloc(ctx);
    if (rc) return rc;
    const size_t px = (size_t)ctx->width * ctx->height;
    CU(cudaMemcpyAsync(ctx->smaa_color, rgba8_host, px * 4, cudaMemcpyHostToDevice, ctx->stream));
    ctx->smaa_timed = true;
    rc = rtb_smaa_run(ctx, ctx->stream);
    if (rc) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    if (ms) CU(cudaEventElapsedTime(ms, ctx->ev_s0, ctx->ev_s1));
    if (out_host) CU(cudaMemcpy(out_host, ctx->smaa_out, px * 4, cudaMemcpyDeviceToHost));
    if (edges_host) CU(cudaMemcpy(edges_host, ctx->smaa_edges, px * 2, cudaMemcpyDeviceToHost));
    if (blend_host) CU(cudaMemcpy(blend_host, ctx->smaa_blend, px * 4, cudaMemcpyDeviceToHost));
    return RTB_OK;
}

int rtb_smaa_last_ms(rtb_ctx* ctx, float* ms) {
    if (!ctx || !ms) return rtb_fail(ctx, RTB_ERR_INVALID, "null argument");
    if (!ctx->smaa_timed || !ctx->ev_s0) return rtb_fail(ctx, RTB_ERR_STATE, "no SMAA pass has run");
    int rc = rtb_sync(ctx);
    if (rc) return rc;
    CU(cudaEventElapsedTime(ms, ctx->ev_s0, ctx->ev_s1));
    return RTB_OK;
}

}  // extern "C"

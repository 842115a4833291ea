/* rt_fused.cuh — the intersectors of the FUSED build (namespace rtb_fast, RTB_STRICT == 0).
 *
 * Same algorithm as the shader, function by function (rt.frag lines cited), evaluated the way a GLSL compiler that
 * contracts a*b+c into FMAs would: every multiply-add is ONE fused instruction, reciprocals are MUFU.RCP (1 ulp), and the
 * quaternion sandwich of every rotated primitive (rt.frag:305-311, two Hamilton products = 94 issue cycles for the ray's
 * direction and origin) is the 3x3 matrix of that sandwich, computed once per upload by pack_kernel (21 instructions).
 * Results are NOT bit-identical to the strict build: parity of this build is the envelope criterion of DESIGN.md section 2
 * (tests/test_envelope.py): every pixel within 1e-4 of the fp32 or the fp64 evaluation of the shader, except pixels that
 * the shader's own arithmetic does not determine to 1e-4 (stochastic-rounding ensemble of the oracle).
 *
 * The kernels are ISSUE bound (profiles/README.md): an FFMA2 holds the issue port two cycles, so packing buys nothing once
 * a multiply and its add fuse — everything here is scalar FFMA/FMUL/FADD, and the cost of a test is its instruction count.
 */
#pragma once

namespace RTB_NS {

/* rd' = M rd, ro' = M (ro - p): the record starts with m[9], px, py, pz (three LDS.128) */
struct LocalRay { vec3 rd, ro; };
template <class T>
DEV LocalRay to_local(SPtr<T> rec, vec3 ro, vec3 rd) {
    const float4 a = lds4(rec, 0), b = lds4(rec, 1), c = lds4(rec, 2);      /* m0 m1 m2 m3 | m4 m5 m6 m7 | m8 px py pz */
    const vec3 o = ro - mk3(c.y, c.z, c.w);
    LocalRay L;
    L.rd = mk3(fmaf(a.x, rd.x, fmaf(a.y, rd.y, a.z * rd.z)), fmaf(a.w, rd.x, fmaf(b.x, rd.y, b.y * rd.z)), fmaf(b.z, rd.x, fmaf(b.w, rd.y, c.x * rd.z)));
    L.ro = mk3(fmaf(a.x, o.x, fmaf(a.y, o.y, a.z * o.z)), fmaf(a.w, o.x, fmaf(b.x, o.y, b.y * o.z)), fmaf(b.z, o.x, fmaf(b.w, o.y, c.x * o.z)));
    return L;
}

/* rt.frag:372-390; uv = opt_uv */
DEV bool intersectRing(const PackK&, vec3 ro, vec3 rd, SPtr<PRingM> R, float tmin, float& t, vec2& uv) {
    const LocalRay L = to_local(R, ro, rd);
    const float r1 = ldsf(R, offsetof(PRingM, r1)), r2 = ldsf(R, offsetof(PRingM, r2));
    t = -L.ro.z * rcp_mufu(L.rd.z);
    const float x = fmaf(L.rd.x, t, L.ro.x);
    const float y = fmaf(L.rd.y, t, L.ro.y);
    const float p = fmaf(x, x, y * y);
    if (t > 0 && t < tmin && p < r2 && p > r1) {
        uv = mk2((p - r1) / (r2 - r1), x * rsqrtf(p));           /* dot(normalize(vec2(x, y)), vec2(1, 0)) */
        return true;
    }
    return false;
}

/* rt.frag:399-421, the slab test (the accept rule and its NaN behaviour stay in box_accept, rt_device.cuh).
 * The max / min chains keep GLSL's operand order: they decide what a NaN slab distance (a ray parallel to a face that
 * starts in the face's plane) does, and the scan order depends on it. */
DEV bool box_candidate(const PackK&, vec3 ro, vec3 rd, SPtr<PBoxM> B, float& tN) {
    const LocalRay L = to_local(B, ro, rd);
    const float4 f4 = lds4(B, 3);                                 /* fx fy fz tex */
    const vec3 m = mk3(rcp_mufu(L.rd.x), rcp_mufu(L.rd.y), rcp_mufu(L.rd.z));
    const vec3 n = m * L.ro;
    const vec3 k = mk3(fabsf(m.x) * f4.x, fabsf(m.y) * f4.y, fabsf(m.z) * f4.z);
    const vec3 t1 = -n - k;
    const vec3 t2 = -n + k;
    tN = gmax(gmax(t1.x, t1.y), t1.z);
    const float tF = gmin(gmin(t2.x, t2.y), t2.z);
    return !(tN > tF || tF < 0.0f);
}

/* intersectSurface, rt.frag:513-572: see surface_candidate of the strict build for the three outcomes */
DEV int surface_candidate(const PackK&, vec3 ro, vec3 rd, SPtr<PSurfM> S, float& t) {
    const LocalRay L = to_local(S, ro, rd);
    const float4 c4 = lds4(S, 3);                                 /* a b c d */
    const float e = ldsf(S, offsetof(PSurfM, e)), f = ldsf(S, offsetof(PSurfM, f));
    const float a = c4.x, b = c4.y, c = c4.z, d = c4.w;
    const float d1 = L.rd.x, d2 = L.rd.y, d3 = L.rd.z, o1 = L.ro.x, o2 = L.ro.y, o3 = L.ro.z;
    const float ad1 = a * d1, bd2 = b * d2, cd3 = c * d3;
    const float p2 = fmaf(ad1, d1, fmaf(bd2, d2, cd3 * d3));
    const float p1 = fmaf(2.f, fmaf(ad1, o1, fmaf(bd2, o2, cd3 * o3)), fmaf(d, d3, e * d2));
    const float p3 = fmaf(a * o1, o1, fmaf(b * o2, o2, fmaf(c * o3, o3, fmaf(d, o3, fmaf(e, o2, f)))));
    if (fabsf(p2) < 1e-6f) {                                      /* quirk Q2 */
        t = -p3 / p1;
        return 2;
    }
    const float disc = fmaf(p1, p1, -4.f * p2 * p3);
    if (!(disc >= 0.f)) return 0;                                 /* the shader's NaN path ends in `FLT_MAX < tmin` (strict build, same place) */
    const float p4s = sqrt_mufu(disc);
    const float inv = rcp_mufu(p2 + p2);
    const float t1 = (-p1 - p4s) * inv, t2 = (-p1 + p4s) * inv;
    float mn = 3.402823466e+38f, mx = 3.402823466e+38f;
    const float epsilon = 1e-4f;
    if (t1 > epsilon && t1 < mn) { mn = t1; mx = t2; }
    if (t2 > epsilon && t2 < mn) { mn = t2; mx = t1; }
    const float4 lo4 = lds4(S, 4), hi4 = lds4(S, 5);              /* e f minx miny | minz maxx maxy maxz */
    if (!checkSurfaceEdges(ro, rd, mn, mx, mk3(lo4.z, lo4.w, hi4.x), mk3(hi4.y, hi4.z, hi4.w), epsilon)) return 0;
    t = mn;
    return 1;
}

/* ---- torus, rt.frag:439-487 ----
 * cTorus(t) = A^2 - 4R^2 B with A = rdrd t^2 + 2 rord t + k0, B = rdxy t^2 + 2 roxy2 t + roxy0 (rt.frag:445-455): the
 * nested form is kept (the expanded quartic loses five digits for a torus 30 units away); the invariants carry the
 * factors 2 and 4R^2:  with u = Re t^2, w = x y = Im t^2 / 2:
 *     A  = (al u + be x + k0,  al2 w + be y)                      al = rdrd, al2 = 2 al, be = 2 rord
 *     fx = A.x^2 - (A.y^2 + ga u + de x + rho)                    ga = 4R^2 rdxy, de = 8R^2 roxy2, rho = 4R^2 roxy0
 *     fy = 2 A.x A.y - (ga2 w + de y)                             ga2 = 2 ga
 * 15 instructions (the strict build: 32 separately rounded operations). */
struct TorusState { float al, al2, be, k0, ga, ga2, de, rho; };
DEV bool torus_setup(const PackK&, vec3 ro, vec3 rd, SPtr<PTorusM> P, int cull, TorusState& T) {
    const LocalRay L = to_local(P, ro, rd);
    const float4 f4 = lds4(P, 3);                                 /* R2 r2 k fourR2 */
    const float rdxy = fmaf(L.rd.x, L.rd.x, L.rd.y * L.rd.y);
    const float roxy2 = fmaf(L.ro.x, L.rd.x, L.ro.y * L.rd.y);
    const float roxy0 = fmaf(L.ro.x, L.ro.x, L.ro.y * L.ro.y);
    const float rdrd = fmaf(L.rd.z, L.rd.z, rdxy);
    const float rord = fmaf(L.ro.z, L.rd.z, roxy2);
    const float roro = fmaf(L.ro.z, L.ro.z, roxy0);
    if (cull) {
        /* conservative reject (option "cull"), as in the strict build */
        const float bs = sqrtf(f4.x) + sqrtf(f4.y);
        const float rr = bs * bs * 1.1025f;
        const float tc = -rord / rdrd;
        const float d2 = roro - rord * rord / rdrd;
        if (d2 > rr || (tc < 0.f && roro > rr)) return false;
    }
    T.al = rdrd; T.al2 = rdrd + rdrd; T.be = rord + rord; T.k0 = roro + f4.z;
    T.ga = f4.w * rdxy; T.ga2 = T.ga + T.ga; T.de = (f4.w + f4.w) * roxy2; T.rho = f4.w * roxy0;
    return true;
}
/* DKstep, rt.frag:456-461: c0 -= cTorus(c0) / ((c0-c1)(c0-c2)(c0-c3)); E = max(E, |fc.x|, |fc.y|) NaN-propagating.
 * The inverse of the product p is conj(p) / |p|^2 through ONE MUFU.RCP.  |p|^2 leaves the range in which MUFU.RCP returns a
 * normal number for tori farther than ~37 units (Durand-Kerner's first step throws a root to ~k0^2): p is scaled by 2^-20 for
 * the norm, which keeps conj(p s) / (p s . p) = conj(p) / |p|^2 accurate for |p|^2 in [2^-106, 2^128) — and +inf gives 0 like the
 * IEEE division of the shader.  43 + 2 instructions. */
DEV void DKstep_f(float& x, float& y, float x1, float y1, float x2, float y2, float x3, float y3, const TorusState& T, float& E) {
    const float u = fmaf(-y, y, x * x), w = x * y;
    const float Ax = fmaf(u, T.al, fmaf(x, T.be, T.k0)), Ay = fmaf(w, T.al2, y * T.be);
    const float fx = fmaf(Ax, Ax, -fmaf(Ay, Ay, fmaf(u, T.ga, fmaf(x, T.de, T.rho))));
    const float fy = fmaf(Ax + Ax, Ay, -fmaf(w, T.ga2, y * T.de));
    const float ax = x - x1, ay = y - y1, bx = x - x2, by = y - y2, cx = x - x3, cy = y - y3;
    const float qx = fmaf(bx, cx, -(by * cy)), qy = fmaf(bx, cy, by * cx);
    const float px = fmaf(ax, qx, -(ay * qy)), py = fmaf(ax, qy, ay * qx);
    const float sx = px * 9.5367431640625e-07f, sy = py * 9.5367431640625e-07f;     /* 2^-20 */
    const float r = rcp_mufu(fmaf(sx, px, sy * py));
    const float ix = sx * r, iy = -sy * r;
    const float gx = fmaf(fx, ix, -(fy * iy)), gy = fmaf(fx, iy, fy * ix);
    x -= gx; y -= gy;
    E = max3_nan_abs(E, gx, gy);
}
/* intersectTorus rt.frag:462-485: the root t and the trip count.  (The template parameter selects the deferred-witness trips
 * in the strict build; there is one form here.) */
template <bool DEFERRED>
DEV float torus_solve(const PackK&, const TorusState& T, int& iters) {
    float x0 = 1.f, y0 = 0.f, x1 = 0.4f, y1 = 0.9f;               /* rt.frag:467-470: c2 = c1 c1, c3 = c2 c1 */
    float x2 = 0.4f * 0.4f - 0.9f * 0.9f, y2 = 0.4f * 0.9f + 0.9f * 0.4f;
    float x3 = x2 * 0.4f - y2 * 0.9f, y3 = x2 * 0.9f + y2 * 0.4f;
    iters = 0;
    for (;;) {                                                    /* rt.frag:471-477 */
        float E = 0.f;
        DKstep_f(x0, y0, x1, y1, x2, y2, x3, y3, T, E);
        DKstep_f(x1, y1, x2, y2, x3, y3, x0, y0, T, E);
        DKstep_f(x2, y2, x3, y3, x0, y0, x1, y1, T, E);
        DKstep_f(x3, y3, x0, y0, x1, y1, x2, y2, T, E);
        iters++;
        if (E < 0.001f || iters >= 60) break;
    }
    const float eps = 0.001f;                                     /* rt.frag:478-485 */
    if (fabsf(y0) > eps || x0 < 0.f) x0 = 10000.f;
    if (fabsf(y1) > eps || x1 < 0.f) x1 = 10000.f;
    if (fabsf(y2) > eps || x2 < 0.f) x2 = 10000.f;
    if (fabsf(y3) > eps || x3 < 0.f) x3 = 10000.f;
    return gmin(gmin(x0, x1), gmin(x2, x3));
}
DEV float torus_solve_scalar(const TorusState& T, int& iters) { const PackK K = {}; return torus_solve<true>(K, T, iters); }

}  // namespace RTB_NS

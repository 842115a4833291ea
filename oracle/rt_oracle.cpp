/* rt_oracle.cpp — CPU restatement of the reference's per-pixel ray tracer.
 *
 * TEST INFRASTRUCTURE (see rt_oracle.h): the checker for the CUDA path, never
 * the product.  Every function below restates one function of
 * /root/reference/assets/shaders/rt.frag (cited as rt.frag:LINE) with the same
 * operation order, constants and control flow, in IEEE fp32 (build with
 * -ffp-contract=off, no fast-math).  GLSL built-ins follow the GLSL 3.30
 * specification's defining formulas (same ones glm 0.9.9.7, vendored by the
 * reference, implements).
 *
 * PINNING: the reference has no tests, golden vectors or fixtures for this path
 *  and its GL render cannot run in this image.  The restatement
 * is pinned instead against oracle/_ref — the reference's OWN rt.frag text
 * compiled as C++ where it lies (oracle/build_ref.py) — by tests/test_oracle_ref.py
 * and by the golden fixtures under tests/golden/ that oracle/_ref generated.
 * Texel filtering (gl_sampler.h) is driver-defined in the reference: parity
 * unpinned there.
 *
 * PRECISION VARIANTS: the arithmetic is written on the type `real` (real_types.h).  ORC_VARIANT 0 (float) is the
 * restatement proper and the only one that is pinned; variants 1 (double, orc64_*) and 2 (stochastic rounding,
 * orcsr_*) run the same control flow and classify pixels whose colour the fp32 arithmetic does not determine.
 * Every pixel also reports its discrete path (hash of the hit-id / shadow-outcome sequence) and its Durand-Kerner
 * trip total through the *_ex entry points.
 *
 * Pins for behaviour the GLSL leaves undefined (SURVEY.md 8a "quirks"):
 *   Q1  `int num, type;` in getReflectedColor (rt.frag:791) are uninitialised and
 *       read after a miss (rt.frag:793): pinned to 0 (=> never TYPE_POINT_LIGHT).
 *   Q4  `i--` for refractive hits (rt.frag:870-872) can loop without bound:
 *       capped at MAX_GLASS_EVENTS refractive events per pixel.
 *   Q9  screen-space derivatives (fwidth rt.frag:326, implicit LOD rt.frag:396,
 *       433-435) are taken across the 2x2 quad (x&~1, y&~1); a neighbour
 *       contributes only if it executes the SAME texture site in the SAME loop
 *       trip (SIMT lock-step); otherwise that derivative is 0.
 *   vec4 color in getSphereTexture (rt.frag:329) is uninitialised for
 *       texNum outside 1..3: pinned to (0,0,0,0).
 */
#include "rt_oracle.h"
#include "gl_sampler.h"
#include "real_types.h"      /* `real` = float (this file's pinned form), double or stochastically rounded fp32: see there */

#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace {

constexpr int MAX_GLASS_EVENTS = 64;
constexpr real PI_F = 3.14159265358979f;      /* rt.frag:5 */
constexpr real maxDist = 1000000.0f;          /* rt.frag:145 */

/* ---------------- GLSL vector types and built-ins ---------------- */
struct vec2 { real x, y; };
struct vec3 { real x, y, z; };
struct vec4 { real x, y, z, w; };

inline vec2 operator+(vec2 a, vec2 b) { return { a.x + b.x, a.y + b.y }; }
inline vec2 operator-(vec2 a, vec2 b) { return { a.x - b.x, a.y - b.y }; }
inline vec2 operator*(vec2 a, real s) { return { a.x * s, a.y * s }; }
inline vec2 operator*(real s, vec2 a) { return { s * a.x, s * a.y }; }
inline vec2 operator/(vec2 a, real s) { return { a.x / s, a.y / s }; }

inline vec3 operator+(vec3 a, vec3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
inline vec3 operator-(vec3 a, vec3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline vec3 operator-(vec3 a) { return { -a.x, -a.y, -a.z }; }
inline vec3 operator*(vec3 a, vec3 b) { return { a.x * b.x, a.y * b.y, a.z * b.z }; }
inline vec3 operator*(vec3 a, real s) { return { a.x * s, a.y * s, a.z * s }; }
inline vec3 operator*(real s, vec3 a) { return { s * a.x, s * a.y, s * a.z }; }
inline vec3 operator/(vec3 a, real s) { return { a.x / s, a.y / s, a.z / s }; }
inline vec3 operator/(real s, vec3 a) { return { s / a.x, s / a.y, s / a.z }; }
inline vec3& operator+=(vec3& a, vec3 b) { a = a + b; return a; }
inline vec3& operator*=(vec3& a, vec3 b) { a = a * b; return a; }
inline vec3& operator*=(vec3& a, real s) { a = a * s; return a; }

inline vec4 operator*(vec4 a, real s) { return { a.x * s, a.y * s, a.z * s, a.w * s }; }
inline bool operator!=(vec4 a, vec4 b) { return a.x != b.x || a.y != b.y || a.z != b.z || a.w != b.w; }

inline real gmin(real x, real y) { return (y < x) ? y : x; }           /* GLSL min */
inline real gmax(real x, real y) { return (x < y) ? y : x; }           /* GLSL max */
inline real clampf(real x, real lo, real hi) { return gmin(gmax(x, lo), hi); }
inline real dot(vec2 a, vec2 b) { return a.x * b.x + a.y * b.y; }
inline real dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline real dot(vec4 a, vec4 b) { return (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w); }
inline real inversesqrt(real x) { return 1.0f / r_sqrt(x); }
inline real length(vec3 v) { return r_sqrt(dot(v, v)); }
inline vec3 normalize(vec3 v) { return v * inversesqrt(dot(v, v)); }
inline vec2 normalize(vec2 v) { return v * inversesqrt(dot(v, v)); }
inline vec3 reflect(vec3 I, vec3 N) { return I - N * dot(N, I) * 2.0f; }
inline vec3 refract(vec3 I, vec3 N, real eta) {
    real d = dot(N, I);
    real k = 1.0f - eta * eta * (1.0f - d * d);
    if (k >= 0.0f) return eta * I - (eta * d + r_sqrt(k)) * N;
    return { 0.0f, 0.0f, 0.0f };
}
inline real signf(real x) { return (real)((0.0f < x) - (x < 0.0f)); }
inline real stepf(real edge, real x) { return x < edge ? 0.0f : 1.0f; }
inline vec3 vabs(vec3 a) { return { r_abs(a.x), r_abs(a.y), r_abs(a.z) }; }
inline vec3 vmax(vec3 a, vec3 b) { return { gmax(a.x, b.x), gmax(a.y, b.y), gmax(a.z, b.z) }; }
inline vec3 vexp(vec3 a) { return { r_exp(a.x), r_exp(a.y), r_exp(a.z) }; }
inline vec3 v3(const float* p) { return { p[0], p[1], p[2] }; }
inline vec4 v4(const float* p) { return { p[0], p[1], p[2], p[3] }; }
inline vec3 xyz(vec4 a) { return { a.x, a.y, a.z }; }

/* hit_record, rt.frag:115-120 (material fields used by the shader only) */
struct Material { vec3 color, absorb; real diffuse, reflection, refraction; int specular; real kd, ks; };
inline Material mat_of(const rtb_material& m) {
    return { v3(m.color), v3(m.absorb), m.diffuse, m.reflect, m.refract, m.specular, m.kd, m.ks };
}
struct HitRecord { Material mat; vec3 normal; real bias_mult; real alpha; };

/* ---------------- scene ---------------- */
struct Scene {
    rtb_defines def;
    rtb_scene scene;
    std::vector<rtb_sphere> spheres;
    std::vector<rtb_plane> planes;
    std::vector<rtb_surface> surfaces;
    std::vector<rtb_box> boxes;
    std::vector<rtb_torus> toruses;
    std::vector<rtb_ring> rings;
    std::vector<rtb_light_point> lights_point;
    std::vector<rtb_light_direct> lights_direct;
    vec3 AMBIENT_COLOR, SHADOW_AMBIENT;
    int ITERATIONS;
    glsim::CubeMap skybox;
    glsim::Texture2D tex[6];
    int pairing = ORC_PAIR_PROGRAM_ORDER;
};

/* GLWrapper::to_string, GLWrapper.cpp:279-282: std::to_string(real) == "%f", then
 * parsed by the GLSL compiler as a real literal. */
float round_through_percent_f(float v) {
    char buf[64];
    snprintf(buf, sizeof buf, "%f", v);
    return strtof(buf, nullptr);
}

/* ---------------- derivative bookkeeping (pin Q9) ---------------- */
enum SiteKind { SITE_SPHERE = 0, SITE_RING = 1, SITE_BOX = 2 };
struct SiteRec { uint64_t key; float u, v; };
struct Derivs { float dudx, dvdx, dudy, dvdy; };

struct QuadCtx {
    std::vector<SiteRec> prev[4], cur[4];
};

struct Stats {
    uint64_t rays_nearest = 0, rays_shadow = 0, tests[7] = { 0 }, dk = 0, shaded[7] = { 0 }, light_evals = 0;
    uint64_t dk_hist[61] = { 0 };
};

/* ---------------- one fragment-shader invocation ---------------- */
struct Frag {
    const Scene& S;
    Stats* st;
    QuadCtx* quad;          /* may be null (KAT entry points): derivatives are 0 */
    int lane = 0;
    real fragx = 0.5f, fragy = 0.5f;   /* gl_FragCoord.xy */

    /* globals of the shader, rt.frag:148-149 */
    vec3 opt_normal = { 0, 0, 0 };
    vec2 opt_uv = { 0, 0 };

    /* program position for derivative pairing */
    int k_trip = 0, k_ctx = 0, k_stage = 0, k_light = 0, k_ring = 0;
    int ord[3] = { 0, 0, 0 };
    int last_dk = 0;
    /* discrete path of this invocation: FNV-1a over (type, num) of every calcInter result and the occlusion outcome of every inShadow */
    uint64_t path = 1469598103934665603ull;
    uint32_t dk_pix = 0;
    void path_mix(uint32_t v) { path = (path ^ v) * 1099511628211ull; }
    std::vector<float>* trace = nullptr;      /* debugging aid (orc*_trace): kind (0 calcInter / 1 inShadow), type, num, t or shadow */

    Frag(const Scene& s, Stats* stats, QuadCtx* q) : S(s), st(stats), quad(q) {}

    uint64_t site_key(int kind, int fetch) {
        if (S.pairing == ORC_PAIR_ORDINAL) return ((uint64_t)kind << 40) | (uint64_t)(ord[kind]++);
        return ((uint64_t)k_trip << 44) | ((uint64_t)k_ctx << 43) | ((uint64_t)k_stage << 42) | ((uint64_t)k_light << 30) |
               ((uint64_t)k_ring << 8) | ((uint64_t)kind << 4) | (uint64_t)fetch;
    }
    /* dFdx / dFdy of uv at this texture site across the 2x2 quad */
    Derivs site_derivs(int kind, int fetch, vec2 uv) {
        Derivs d = { 0, 0, 0, 0 };
        if (!quad) return d;
        uint64_t key = site_key(kind, fetch);
        const float fu = to_f(uv.x), fv = to_f(uv.y);
        quad->cur[lane].push_back({ key, fu, fv });
        auto find = [&](int other, float& u, float& v) {
            for (const SiteRec& r : quad->prev[other]) if (r.key == key) { u = r.u; v = r.v; return true; }
            return false;
        };
        float u, v;
        if (find(lane ^ 1, u, v)) {     /* dFdx = right - left */
            if (lane & 1) { d.dudx = fu - u; d.dvdx = fv - v; } else { d.dudx = u - fu; d.dvdx = v - fv; }
        }
        if (find(lane ^ 2, u, v)) {     /* dFdy = top - bottom */
            if (lane & 2) { d.dudy = fu - u; d.dvdy = fv - v; } else { d.dudy = u - fu; d.dvdy = v - fv; }
        }
        return d;
    }

    /* ---- quaternions, rt.frag:285-311 ---- */
    static vec4 quat_conj(vec4 q) { return { -q.x, -q.y, -q.z, q.w }; }
    static vec4 quat_inv(vec4 q) { return quat_conj(q) * (1 / dot(q, q)); }
    static vec4 quat_mult(vec4 q1, vec4 q2) {
        vec4 qr;
        qr.x = (q1.w * q2.x) + (q1.x * q2.w) + (q1.y * q2.z) - (q1.z * q2.y);
        qr.y = (q1.w * q2.y) - (q1.x * q2.z) + (q1.y * q2.w) + (q1.z * q2.x);
        qr.z = (q1.w * q2.z) + (q1.x * q2.y) - (q1.y * q2.x) + (q1.z * q2.w);
        qr.w = (q1.w * q2.w) - (q1.x * q2.x) - (q1.y * q2.y) - (q1.z * q2.z);
        return qr;
    }
    static vec3 rotate(vec4 qr, vec3 v) {
        vec4 qr_conj = quat_conj(qr);
        vec4 q_pos = { v.x, v.y, v.z, 0 };
        vec4 q_tmp = quat_mult(qr, q_pos);
        return xyz(quat_mult(q_tmp, qr_conj));
    }

    /* rt.frag:313-317 */
    vec3 getRayDir() {
        vec2 half = vec2{ (real)S.scene.canvas_width, (real)S.scene.canvas_height } / 2.0f;
        vec2 p = (vec2{ fragx, fragy } - half) / (real)S.scene.canvas_height;
        vec3 result = { p.x, p.y, 1.0f };
        return normalize(rotate(v4(S.scene.quat_camera_rotation), result));
    }

    /* rt.frag:319-340 */
    vec4 getSphereTexture(vec3 sphereNormal, vec4 quat, int texNum) {
        if (quat != vec4{ 0, 0, 0, 1 }) sphereNormal = rotate(quat, sphereNormal);
        real u = 0.5f + r_atan2(sphereNormal.z, sphereNormal.x) / (2.f * PI_F);
        real v = 0.5f - r_asin(sphereNormal.y) / PI_F;
        vec2 uv = { u, v };
        Derivs d = site_derivs(SITE_SPHERE, 0, uv);
        vec2 df = { r_abs(d.dudx) + r_abs(d.dudy), r_abs(d.dvdx) + r_abs(d.dvdy) };   /* fwidth */
        if (df.x > 0.5f) df.x = 0.f;
        vec4 color = { 0, 0, 0, 0 };
        if (texNum >= 1 && texNum <= 3) {
            glsim::rgba c = glsim::texture_lod(S.tex[texNum], to_f(uv.x), to_f(uv.y), to_f(r_log2(gmax(df.x, df.y) * 1024.f)));
            color = { c.r, c.g, c.b, c.a };
        }
        return color;
    }

    /* rt.frag:342-354 */
    bool intersectSphere(vec3 ro, vec3 rd, vec4 object, bool hollow, real tmin, real& t) {
        vec3 oc = ro - xyz(object);
        real b = dot(oc, rd);
        real c = dot(oc, oc) - object.w * object.w;
        real h = b * b - c;
        if (h < 0.0f) return false;
        real h_sqrt = r_sqrt(h);
        t = -b - h_sqrt;
        if (hollow && t < 0.0f) t = -b + h_sqrt;
        return t > 0 && t < tmin;
    }

    /* rt.frag:356-370, PLANE_ONESIDE defined (rt.frag:21) */
    bool intersectPlane(vec3 ro, vec3 rd, vec3 n, vec3 p, real tmin, real& t) {
        real denom = clampf(dot(n, rd), -1, 1);
        if (denom < -1e-6f) {
            vec3 p_ro = p - ro;
            t = dot(p_ro, n) / denom;
            return (t > 0) && (t < tmin);
        }
        return false;
    }

    /* rt.frag:372-390 */
    bool intersectRing(vec3 ro, vec3 rd, int num, real tmin, real& t) {
        const rtb_ring& ring = S.rings[num];
        vec4 q = v4(ring.quat_rotation);
        rd = rotate(q, rd);
        ro = rotate(q, ro - v3(ring.pos));
        t = -ro.z / rd.z;
        real x = ro.x + rd.x * t;
        real y = ro.y + rd.y * t;
        real p = x * x + y * y;
        if (t > 0 && t < tmin && p < ring.r2 && p > ring.r1) {
            real cosv = dot(normalize(vec2{ x, y }), vec2{ 1, 0 });
            opt_uv = { (p - ring.r1) / (ring.r2 - ring.r1), cosv };
            return true;
        }
        return false;
    }
    /* rt.frag:391-394 */
    vec3 getRingNormal(int num) { return rotate(quat_inv(v4(S.rings[num].quat_rotation)), vec3{ 0, 0, -1 }); }
    /* rt.frag:395-397: texture(texture_ring, uv), implicit LOD; `num` is ignored by the shader */
    vec4 getRingTexture(int /*num*/, vec2 uv) {
        Derivs d = site_derivs(SITE_RING, 0, uv);
        float lod = glsim::implicit_lod(S.tex[4], d.dudx, d.dvdx, d.dudy, d.dvdy);
        glsim::rgba c = glsim::texture_lod(S.tex[4], to_f(uv.x), to_f(uv.y), lod);
        return { c.r, c.g, c.b, c.a };
    }

    /* rt.frag:399-427 */
    bool intersectBox(vec3 ro, vec3 rd, int num, real tmin, real& t) {
        const rtb_box& box = S.boxes[num];
        vec4 q = v4(box.quat_rotation);
        vec3 rdd = rotate(q, rd);
        vec3 roo = rotate(q, ro - v3(box.pos));
        vec3 m = 1.0f / rdd;
        vec3 n = m * roo;
        vec3 k = vabs(m) * v3(box.form);
        vec3 t1 = -n - k;
        vec3 t2 = -n + k;
        real tN = gmax(gmax(t1.x, t1.y), t1.z);
        real tF = gmin(gmin(t2.x, t2.y), t2.z);
        if (tN > tF || tF < 0.0f) return false;
        if (tN >= tmin) return false;
        /* nor = -sign(rdd)*step(t1.yzx,t1.xyz)*step(t1.zxy,t1.xyz) */
        vec3 sg = { signf(rdd.x), signf(rdd.y), signf(rdd.z) };
        vec3 s1 = { stepf(t1.y, t1.x), stepf(t1.z, t1.y), stepf(t1.x, t1.z) };
        vec3 s2 = { stepf(t1.z, t1.x), stepf(t1.x, t1.y), stepf(t1.y, t1.z) };
        vec3 nor = -sg * s1 * s2;
        t = tN;
        opt_normal = rotate(quat_inv(q), nor);
        return true;
    }
    /* rt.frag:428-436 */
    vec4 getBoxTexture(vec3 pt, vec3 normal, int num) {
        const rtb_box& box = S.boxes[num];
        vec4 q = v4(box.quat_rotation);
        vec3 pos = rotate(q, v3(box.pos));
        pt = rotate(q, pt);
        normal = rotate(q, normal);
        vec2 uv0 = { 0.5f * (pt.z - pos.z) - 0.5f, 0.5f * (pt.y - pos.y) - 0.5f };   /* pt.zy */
        vec2 uv1 = { 0.5f * (pt.z - pos.z) - 0.5f, 0.5f * (pt.x - pos.x) - 0.5f };   /* pt.zx */
        vec2 uv2 = { 0.5f * (pt.x - pos.x) - 0.5f, 0.5f * (pt.y - pos.y) - 0.5f };   /* pt.xy */
        vec2 uvs[3] = { uv0, uv1, uv2 };
        real wgt[3] = { r_abs(normal.x), r_abs(normal.y), r_abs(normal.z) };
        vec4 acc = { 0, 0, 0, 0 };
        for (int f = 0; f < 3; f++) {
            Derivs d = site_derivs(SITE_BOX, f, uvs[f]);
            float lod = glsim::implicit_lod(S.tex[5], d.dudx, d.dvdx, d.dudy, d.dvdy);
            glsim::rgba c = glsim::texture_lod(S.tex[5], to_f(uvs[f].x), to_f(uvs[f].y), lod);
            vec4 term = { wgt[f] * c.r, wgt[f] * c.g, wgt[f] * c.b, wgt[f] * c.a };
            if (f == 0) acc = term;
            else acc = { acc.x + term.x, acc.y + term.y, acc.z + term.z, acc.w + term.w };
        }
        return acc;
    }

    /* ---- torus, rt.frag:439-496 ---- */
    static vec2 cmul(vec2 c1, vec2 c2) { return { c1.x * c2.x - c1.y * c2.y, c1.x * c2.y + c1.y * c2.x }; }
    static vec2 cinv(vec2 c) { return vec2{ c.x, -c.y } / dot(c, c); }
    static vec2 cTorus(vec2 t, vec3 ro, vec3 rd, vec2 torus) {
        real R2 = torus.x * torus.x;
        real r2 = torus.y * torus.y;
        vec2 t2 = { t.x * t.x - t.y * t.y, 2.f * t.x * t.y };
        vec2 res = t2 * dot(rd, rd) + 2.f * t * dot(ro, rd) + vec2{ dot(ro, ro) + R2 - r2, 0.f };
        res = cmul(res, res);
        vec2 roxy = { ro.x, ro.y }, rdxy = { rd.x, rd.y };
        vec2 res2 = 4.f * R2 * (t2 * dot(rdxy, rdxy) + 2.f * t * dot(roxy, rdxy) + vec2{ dot(roxy, roxy), 0.f });
        return res - res2;
    }
    static real DKstep(vec2& c0, vec2 c1, vec2 c2, vec2 c3, vec3 ro, vec3 rd, vec2 torus) {
        vec2 fc = cTorus(c0, ro, rd, torus);
        fc = cmul(fc, cinv(cmul(c0 - c1, cmul(c0 - c2, c0 - c3))));
        c0 = c0 - fc;
        return gmax(r_abs(fc.x), r_abs(fc.y));
    }
    bool intersectTorus(vec3 ro, vec3 rd, int num, real tmin, real& t) {
        real eps = 0.001f;
        const rtb_torus& torus = S.toruses[num];
        vec4 q = v4(torus.quat_rotation);
        vec2 form = { torus.form[0], torus.form[1] };
        ro = rotate(q, ro - v3(torus.pos));
        rd = rotate(q, rd);
        vec2 c0 = { 1.f, 0.f };
        vec2 c1 = { 0.4f, 0.9f };
        vec2 c2 = cmul(c1, vec2{ 0.4f, 0.9f });
        vec2 c3 = cmul(c2, vec2{ 0.4f, 0.9f });
        int iters = 0;
        for (int i = 0; i < 60; i++) {
            iters++;
            real e = DKstep(c0, c1, c2, c3, ro, rd, form);
            e = gmax(e, DKstep(c1, c2, c3, c0, ro, rd, form));
            e = gmax(e, DKstep(c2, c3, c0, c1, ro, rd, form));
            e = gmax(e, DKstep(c3, c0, c1, c2, ro, rd, form));
            if (e < eps) break;
        }
        last_dk = iters;
        dk_pix += (uint32_t)iters;
        if (st) { st->dk += iters; st->dk_hist[iters]++; }
        vec4 rs = { c0.x, c1.x, c2.x, c3.x };
        vec4 ri = { r_abs(c0.y), r_abs(c1.y), r_abs(c2.y), r_abs(c3.y) };
        if (ri.x > eps || rs.x < 0.f) rs.x = 10000.f;
        if (ri.y > eps || rs.y < 0.f) rs.y = 10000.f;
        if (ri.z > eps || rs.z < 0.f) rs.z = 10000.f;
        if (ri.w > eps || rs.w < 0.f) rs.w = 10000.f;
        t = gmin(gmin(rs.x, rs.y), gmin(rs.z, rs.w));
        return t > 0 && t < 100 && t < tmin;
    }
    vec3 getTorusNormal(vec3 ro, vec3 rd, real t, int num) {
        const rtb_torus& torus = S.toruses[num];
        vec4 q = v4(torus.quat_rotation);
        ro = rotate(q, ro - v3(torus.pos));
        rd = rotate(q, rd);
        vec3 pos = ro + rd * t;
        real fy = torus.form[1], fx = torus.form[0];
        vec3 normal = pos * (vec3{ 1, 1, 1 } * (dot(pos, pos) - fy * fy) - fx * fx * vec3{ 1.0f, 1.0f, -1.0f });
        return normalize(rotate(quat_inv(q), normal));
    }

    /* ---- quadric surfaces, rt.frag:500-584 ---- */
    static bool isBetween(vec3 value, vec3 mn, vec3 mx) {        /* rt.frag:280-283 */
        return (value.x > mn.x && value.y > mn.y && value.z > mn.z) && (value.x < mx.x && value.y < mx.y && value.z < mx.z);
    }
    static bool checkSurfaceEdges(vec3 o, vec3 d, real& tMin, real& tMax, vec3 v_min, vec3 v_max, real epsilon) {
        vec3 pt = d * tMin + o;
        if (!isBetween(pt, v_min, v_max)) {
            if (tMax < epsilon) return false;
            pt = d * tMax + o;
            if (!isBetween(pt, v_min, v_max)) return false;
            real tmp = tMin; tMin = tMax; tMax = tmp;
        }
        return true;
    }
    bool intersectSurface(vec3 ro, vec3 rd, int num, real tmin, real& t) {
        vec3 orig_ro = ro;
        vec3 orig_rd = rd;
        const rtb_surface& surface = S.surfaces[num];
        vec4 q = v4(surface.quat_rotation);
        ro = rotate(q, ro - v3(surface.pos));
        rd = rotate(q, rd);
        real a = surface.a, b = surface.b, c = surface.c, d = surface.d, e = surface.e, f = surface.f;
        real d1 = rd.x, d2 = rd.y, d3 = rd.z;
        real o1 = ro.x, o2 = ro.y, o3 = ro.z;
        real p1 = 2 * a * d1 * o1 + 2 * b * d2 * o2 + 2 * c * d3 * o3 + d * d3 + d2 * e;
        real p2 = a * d1 * d1 + b * d2 * d2 + c * d3 * d3;
        real p3 = a * o1 * o1 + b * o2 * o2 + c * o3 * o3 + d * o3 + e * o2 + f;
        real p4 = r_sqrt(p1 * p1 - 4 * p2 * p3);
        if (r_abs(p2) < 1e-6f) {                 /* quirk Q2: accepts t GREATER than tmin, rt.frag:541-545 */
            t = -p3 / p1;
            return t > tmin;
        }
        real mn = FLT_MAX;
        real mx = FLT_MAX;
        real t1 = (-p1 - p4) / (2 * p2);
        real t2 = (-p1 + p4) / (2 * p2);
        real epsilon = 1e-4f;
        if (t1 > epsilon && t1 < mn) { mn = t1; mx = t2; }
        if (t2 > epsilon && t2 < mn) { mn = t2; mx = t1; }
        if (!checkSurfaceEdges(orig_ro, orig_rd, mn, mx, v3(surface.v_min), v3(surface.v_max), epsilon)) return false;
        t = mn;
        return t < tmin;
    }
    vec3 getSurfaceNormal(vec3 ro, vec3 rd, real t, int num) {
        const rtb_surface& surface = S.surfaces[num];
        vec4 q = v4(surface.quat_rotation);
        ro = ro - v3(surface.pos);
        ro = rotate(q, ro);
        rd = rotate(q, rd);
        vec3 tm = rd * t + ro;
        vec3 normal = { 2 * surface.a * tm.x, 2 * surface.b * tm.y + surface.e, 2 * surface.c * tm.z + surface.d };
        normal = rotate(quat_inv(q), normal);
        return normalize(normal);
    }

    /* rt.frag:587-628 */
    real calcInter(vec3 ro, vec3 rd, int& num, int& type) {
        if (st) st->rays_nearest++;
        real tmin = maxDist;
        real t;
        const rtb_defines& D = S.def;
        if (st) { st->tests[RTB_TYPE_PLANE] += D.plane_size; st->tests[RTB_TYPE_SPHERE] += D.sphere_size;
                  st->tests[RTB_TYPE_SURFACE] += D.surface_size; st->tests[RTB_TYPE_BOX] += D.box_size;
                  st->tests[RTB_TYPE_TORUS] += D.torus_size; st->tests[RTB_TYPE_RING] += D.ring_size;
                  st->tests[RTB_TYPE_POINT_LIGHT] += D.light_point_size; }
        for (int i = 0; i < D.plane_size; i++)
            if (intersectPlane(ro, rd, v3(S.planes[i].normal), v3(S.planes[i].pos), tmin, t)) { num = i; tmin = t; type = RTB_TYPE_PLANE; }
        for (int i = 0; i < D.sphere_size; i++)
            if (intersectSphere(ro, rd, v4(S.spheres[i].obj), S.spheres[i].hollow != 0, tmin, t)) { num = i; tmin = t; type = RTB_TYPE_SPHERE; }
        for (int i = 0; i < D.surface_size; i++)
            if (intersectSurface(ro, rd, i, tmin, t)) { num = i; tmin = t; type = RTB_TYPE_SURFACE; }
        for (int i = 0; i < D.box_size; i++)
            if (intersectBox(ro, rd, i, tmin, t)) { num = i; tmin = t; type = RTB_TYPE_BOX; }
        for (int i = 0; i < D.torus_size; i++)
            if (intersectTorus(ro, rd, i, tmin, t)) { num = i; tmin = t; type = RTB_TYPE_TORUS; }
        for (int i = 0; i < D.ring_size; i++)
            if (intersectRing(ro, rd, i, tmin, t)) { num = i; tmin = t; type = RTB_TYPE_RING; }
        for (int i = 0; i < D.light_point_size; i++)
            if (intersectSphere(ro, rd, v4(S.lights_point[i].pos), false, tmin, t)) { num = i; tmin = t; type = RTB_TYPE_POINT_LIGHT; }
        path_mix(tmin < maxDist ? (uint32_t)((type << 24) | (num & 0xffffff)) : 0xffffffffu);
        if (trace) { trace->push_back(0.f); trace->push_back((float)type); trace->push_back((float)num); trace->push_back(to_f(tmin)); }
        return tmin;
    }

    /* rt.frag:630-658 (PLANE_ONESIDE == 1: planes never shadow; no early exit) */
    real inShadow(vec3 ro, vec3 rd, real dist) {
        if (st) st->rays_shadow++;
        const rtb_defines& D = S.def;
        if (st) { st->tests[RTB_TYPE_SPHERE] += D.sphere_size; st->tests[RTB_TYPE_SURFACE] += D.surface_size;
                  st->tests[RTB_TYPE_BOX] += D.box_size; st->tests[RTB_TYPE_TORUS] += D.torus_size;
                  st->tests[RTB_TYPE_RING] += D.ring_size; }
        real t;
        real shadow = 0;
        for (int i = 0; i < D.sphere_size; i++)
            if (intersectSphere(ro, rd, v4(S.spheres[i].obj), false, dist, t)) shadow = 1;
        for (int i = 0; i < D.surface_size; i++)
            if (intersectSurface(ro, rd, i, dist, t)) shadow = 1;
        for (int i = 0; i < D.box_size; i++)
            if (intersectBox(ro, rd, i, dist, t)) shadow = 1;
        for (int i = 0; i < D.torus_size; i++)
            if (intersectTorus(ro, rd, i, dist, t)) shadow = 1;
        for (int i = 0; i < D.ring_size; i++)
            if (intersectRing(ro, rd, i, dist, t)) {
                const rtb_ring& ring = S.rings[i];
                if (ring.textureNum > 0) {
                    k_stage = 1; k_ring = i;
                    shadow += getRingTexture(ring.textureNum, opt_uv).w;
                    k_stage = 0; k_ring = 0;
                } else {
                    shadow = 1;
                }
            }
        path_mix(shadow > 0 ? 0x5ad0u : 0x11e7u);
        if (trace) { trace->push_back(1.f); trace->push_back(0.f); trace->push_back(0.f); trace->push_back(to_f(shadow)); }
        return gmin(shadow, 1);
    }

    /* rt.frag:660-679 */
    void calcShade2(vec3 light_dir, vec3 light_color, real intensity, vec3 pt, vec3 rd, const Material& material, vec3 normal,
                    bool doShadow, real dist, real distDiv, vec3& diffuse, vec3& specular) {
        if (st) st->light_evals++;
        light_dir = normalize(light_dir);
        real dp = clampf(dot(normal, light_dir), 0.0f, 1.0f);
        light_color *= dp;
        if (doShadow) {                                         /* SHADOW_ENABLED 1, rt.frag:15 */
            real sh = 1 - inShadow(pt, light_dir, dist);
            vec3 shadow = { sh, sh, sh };
            light_color *= vmax(shadow, S.SHADOW_AMBIENT);
        }
        diffuse += light_color * material.color * material.diffuse * intensity / distDiv;
        if (material.specular > 0) {
            vec3 reflection = reflect(light_dir, normal);
            real specDp = clampf(dot(rd, reflection), 0.0f, 1.0f);
            specular += light_color * r_pow(specDp, (real)material.specular) * intensity / distDiv;
        }
    }

    /* rt.frag:681-709 */
    vec3 calcShade(vec3 pt, vec3 rd, const Material& material, vec3 normal, bool doShadow) {
        real dist, distDiv;
        vec3 light_color, light_dir;
        vec3 diffuse = { 0, 0, 0 };
        vec3 specular = { 0, 0, 0 };
        vec3 pixelColor = S.AMBIENT_COLOR * material.color;
        for (int i = 0; i < S.def.light_point_size; i++) {
            const rtb_light_point& light = S.lights_point[i];
            light_color = v3(light.color);
            light_dir = xyz(v4(light.pos)) - pt;
            dist = length(light_dir);
            distDiv = 1 + light.linear_k * dist + light.quadratic_k * dist * dist;
            k_light = i;
            calcShade2(light_dir, light_color, light.intensity, pt, rd, material, normal, doShadow, dist, distDiv, diffuse, specular);
        }
        for (int i = 0; i < S.def.light_direct_size; i++) {
            light_color = v3(S.lights_direct[i].color);
            light_dir = -v3(S.lights_direct[i].direction);
            dist = maxDist;
            distDiv = 1;
            k_light = S.def.light_point_size + i;
            calcShade2(light_dir, light_color, S.lights_direct[i].intensity, pt, rd, material, normal, doShadow, dist, distDiv, diffuse, specular);
        }
        k_light = 0;
        pixelColor += diffuse * material.kd + specular * material.ks;
        return pixelColor;
    }

    /* rt.frag:711-715 */
    static real getFresnel(vec3 normal, vec3 rd, real reflection) {
        real ndotv = clampf(dot(normal, -rd), 0.0f, 1.0f);
        return reflection + (1.0f - reflection) * r_pow(1.0f - ndotv, (real)5.0f);
    }
    /* rt.frag:717-742, DO_FRESNEL 1 */
    static real FresnelReflectAmount(real n1, real n2, vec3 normal, vec3 incident, real refl) {
        real r0 = (n1 - n2) / (n1 + n2);
        r0 *= r0;
        real cosX = -dot(normal, incident);
        if (n1 > n2) {
            real n = n1 / n2;
            real sinT2 = n * n * (1.0f - cosX * cosX);
            if (sinT2 > 1.0f) return 1.0f;
            cosX = r_sqrt(1.0f - sinT2);
        }
        real x = 1.0f - cosX;
        real ret = r0 + (1.0f - r0) * x * x * x * x * x;
        ret = (refl + (1.0f - refl) * ret);
        return ret;
    }

    /* rt.frag:744-784 */
    HitRecord get_hit_info(vec3 ro, vec3 rd, vec3 pt, real t, int num, int type) {
        HitRecord hr = {};
        if (st && type >= 0 && type < 7) st->shaded[type]++;
        if (type == RTB_TYPE_SPHERE) {
            const rtb_sphere& sphere = S.spheres[num];
            hr = { mat_of(sphere.material), normalize(pt - xyz(v4(sphere.obj))), 0, 1 };
            if (sphere.textureNum != 0) {
                vec4 texColor = getSphereTexture(hr.normal, v4(sphere.quat_rotation), sphere.textureNum);
                hr.mat.color = xyz(texColor);
                hr.alpha = texColor.w;
            }
        }
        if (type == RTB_TYPE_PLANE) hr = { mat_of(S.planes[num].material), normalize(v3(S.planes[num].normal)), 0, 1 };
        if (type == RTB_TYPE_SURFACE) hr = { mat_of(S.surfaces[num].mat), getSurfaceNormal(ro, rd, t, num), 0, 1 };
        if (type == RTB_TYPE_BOX) {
            const rtb_box& box = S.boxes[num];
            hr = { mat_of(box.mat), opt_normal, 0, 1 };
            if (box.textureNum != 0) hr.mat.color = xyz(getBoxTexture(pt, opt_normal, num));
        }
        if (type == RTB_TYPE_TORUS) hr = { mat_of(S.toruses[num].mat), getTorusNormal(ro, rd, t, num), 0, 1 };
        if (type == RTB_TYPE_RING) {
            const rtb_ring& ring = S.rings[num];
            hr = { mat_of(ring.mat), getRingNormal(num), 0, 1 };
            if (ring.textureNum != 0) {
                vec4 texColor = getRingTexture(ring.textureNum, opt_uv);
                hr.mat.color = xyz(texColor);
                hr.alpha = texColor.w;
            }
        }
        real distance = length(pt - ro);
        hr.bias_mult = (9e-3f * distance + 35) / 35e3f;
        return hr;
    }

    /* rt.frag:787-802 */
    vec3 getReflectedColor(vec3 ro, vec3 rd) {
        vec3 color = { 0, 0, 0 };
        vec3 pt;
        int num = 0, type = 0;                                  /* pin Q1 */
        real t = calcInter(ro, rd, num, type);
        if (type == RTB_TYPE_POINT_LIGHT) return v3(S.lights_point[num].color);
        HitRecord hr;
        if (t < maxDist) {
            pt = ro + rd * t;
            hr = get_hit_info(ro, rd, pt, t, num, type);
            ro = dot(rd, hr.normal) < 0 ? pt + hr.normal * hr.bias_mult : pt - hr.normal * hr.bias_mult;
            color = calcShade(ro, rd, hr.mat, hr.normal, true);
        }
        return color;
    }

    /* rt.frag:804-902 */
    vec4 main_() {
        real reflectMultiplier, refractMultiplier, tm;
        Material mat;
        vec3 pt, n;
        vec3 mask = { 1.0f, 1.0f, 1.0f };
        vec3 color = { 0.0f, 0.0f, 0.0f };
        vec3 ro = v3(S.scene.camera_pos);
        vec3 rd = getRayDir();
        real absorbDistance = 0.0f;
        int type = 0;
        int num = 0;
        HitRecord hr;
        int glass_events = 0;

        for (int i = 0; i < S.ITERATIONS; i++) {
            k_ctx = 0; k_stage = 0; k_light = 0; k_ring = 0;
            tm = calcInter(ro, rd, num, type);
            if (tm < maxDist) {
                pt = ro + rd * tm;
                hr = get_hit_info(ro, rd, pt, tm, num, type);
                if (type == RTB_TYPE_POINT_LIGHT) {
                    color += v3(S.lights_point[num].color) * mask;
                    break;
                }
                mat = hr.mat;
                n = hr.normal;
                bool outside = dot(rd, n) < 0;
                n = outside ? n : -n;
                /* TOTAL_INTERNAL_REFLECTION 1 */
                if (mat.refraction > 0)
                    reflectMultiplier = FresnelReflectAmount(outside ? 1 : mat.refraction, outside ? mat.refraction : 1, rd, n, mat.reflection);
                else
                    reflectMultiplier = getFresnel(n, rd, mat.reflection);
                refractMultiplier = 1 - reflectMultiplier;

                if (mat.refraction > 0.0f) {                    /* refractive */
                    if (outside && mat.reflection > 0) {
                        k_ctx = 1;
                        color += getReflectedColor(pt + n * hr.bias_mult, reflect(rd, n)) * reflectMultiplier * mask;
                        k_ctx = 0;
                        mask *= refractMultiplier;
                    } else if (!outside) {
                        absorbDistance += tm;
                        vec3 absorb = vexp(-mat.absorb * absorbDistance);
                        mask *= absorb;
                    }
                    if (reflectMultiplier >= 1) break;
                    ro = pt - n * hr.bias_mult;
                    rd = refract(rd, n, outside ? 1 / mat.refraction : mat.refraction);
                    i--;                                        /* REFLECT_REDUCE_ITERATION, rt.frag:870-872 */
                    if (++glass_events >= MAX_GLASS_EVENTS) break;   /* pin Q4 */
                } else if (mat.reflection > 0.0f) {             /* reflective */
                    ro = pt + n * hr.bias_mult;
                    color += calcShade(ro, rd, mat, n, true) * refractMultiplier * mask;
                    rd = reflect(rd, n);
                    mask *= reflectMultiplier;
                } else {                                        /* diffuse */
                    color += calcShade(pt + n * hr.bias_mult, rd, mat, n, true) * mask * hr.alpha;
                    if (hr.alpha < 1) {
                        ro = pt - n * hr.bias_mult;
                        mask *= 1 - hr.alpha;
                    } else {
                        break;
                    }
                }
            } else {
                glsim::rgba c = glsim::texture_cube(S.skybox, to_f(rd.x), to_f(rd.y), to_f(rd.z));
                color += vec3{ c.r, c.g, c.b } * mask;
                break;
            }
            k_trip++;
        }
        return { color.x, color.y, color.z, 1.0f };
    }
};

void add_stats(orc_stats* dst, const Stats& s, uint64_t pixels) {
    dst->pixels += pixels;
    dst->rays_nearest += s.rays_nearest;
    dst->rays_shadow += s.rays_shadow;
    for (int i = 0; i < 7; i++) { dst->tests[i] += s.tests[i]; dst->shaded_hits[i] += s.shaded[i]; }
    dst->dk_iterations += s.dk;
    dst->light_evals += s.light_evals;
    for (int i = 0; i <= 60; i++) dst->dk_hist[i] += s.dk_hist[i];
}

/* what the *_ex entry points report per pixel besides the colour */
struct PixelInfo { uint64_t path; uint32_t dk; };

/* Run the four invocations of one 2x2 quad to the derivative fixed point (pin Q9). */
void render_quad(const Scene& S, int qx, int qy, float* out4x4, Stats* stats, PixelInfo* info4 = nullptr) {
    QuadCtx quad;
    vec4 col[4];
    PixelInfo info[4] = {};
    Stats last;
    for (int pass = 0; pass < 8; pass++) {
        Stats local;
        for (int l = 0; l < 4; l++) quad.cur[l].clear();
        for (int l = 0; l < 4; l++) {
            Frag f(S, &local, &quad);
            f.lane = l;
            f.fragx = (float)(qx + (l & 1)) + 0.5f;
            f.fragy = (float)(qy + (l >> 1)) + 0.5f;
            col[l] = f.main_();
            info[l] = { f.path, f.dk_pix };
        }
        last = local;
        bool any = false, same = true;
        for (int l = 0; l < 4; l++) {
            if (!quad.cur[l].empty()) any = true;
            if (quad.cur[l].size() != quad.prev[l].size()) same = false;
            else
                for (size_t i = 0; i < quad.cur[l].size(); i++) {
                    const SiteRec &a = quad.cur[l][i], &b = quad.prev[l][i];
                    if (a.key != b.key || memcmp(&a.u, &b.u, 4) || memcmp(&a.v, &b.v, 4)) { same = false; break; }
                }
        }
        if (!any || (pass > 0 && same)) break;
        for (int l = 0; l < 4; l++) quad.prev[l] = quad.cur[l];
    }
    if (stats) {
        stats->rays_nearest += last.rays_nearest; stats->rays_shadow += last.rays_shadow; stats->dk += last.dk;
        stats->light_evals += last.light_evals;
        for (int i = 0; i < 7; i++) { stats->tests[i] += last.tests[i]; stats->shaded[i] += last.shaded[i]; }
        for (int i = 0; i <= 60; i++) stats->dk_hist[i] += last.dk_hist[i];
    }
    for (int l = 0; l < 4; l++) {
        out4x4[l * 4 + 0] = to_f(col[l].x); out4x4[l * 4 + 1] = to_f(col[l].y); out4x4[l * 4 + 2] = to_f(col[l].z); out4x4[l * 4 + 3] = to_f(col[l].w);
        if (info4) info4[l] = info[l];
    }
}

int resolve_threads(int n) {
    if (n > 0) return n;
    unsigned hc = std::thread::hardware_concurrency();
    return hc ? (int)hc : 1;
}

struct Handle { Scene S; };

}  // namespace

extern "C" {

orc_handle* ORC_API(create)(const orc_scene_desc* d) {
    if (!d || !d->scene) return nullptr;
    Handle* h = new Handle();
    Scene& S = h->S;
    S.def = d->defines;
    S.scene = *d->scene;
    auto cp = [](auto& vec, const auto* src, int n) { if (n > 0 && src) vec.assign(src, src + n); };
    cp(S.spheres, d->spheres, d->defines.sphere_size);
    cp(S.planes, d->planes, d->defines.plane_size);
    cp(S.surfaces, d->surfaces, d->defines.surface_size);
    cp(S.boxes, d->boxes, d->defines.box_size);
    cp(S.toruses, d->toruses, d->defines.torus_size);
    cp(S.rings, d->rings, d->defines.ring_size);
    cp(S.lights_point, d->lights_point, d->defines.light_point_size);
    cp(S.lights_direct, d->lights_direct, d->defines.light_direct_size);
    S.ITERATIONS = d->defines.iterations;
    S.AMBIENT_COLOR = { round_through_percent_f(d->defines.ambient_color[0]), round_through_percent_f(d->defines.ambient_color[1]),
                        round_through_percent_f(d->defines.ambient_color[2]) };
    S.SHADOW_AMBIENT = { round_through_percent_f(d->defines.shadow_ambient[0]), round_through_percent_f(d->defines.shadow_ambient[1]),
                         round_through_percent_f(d->defines.shadow_ambient[2]) };
    if (d->cube[0].px) {
        S.skybox.w = d->cube[0].w; S.skybox.h = d->cube[0].h;
        for (int f = 0; f < 6; f++) glsim::expand_rgba8(d->cube[f].px, d->cube[f].w, d->cube[f].h, d->cube[f].ch, S.skybox.face[f]);
    }
    for (int u = 1; u <= 5; u++)
        if (d->tex2d[u].px) glsim::build_mips(S.tex[u], d->tex2d[u].px, d->tex2d[u].w, d->tex2d[u].h, d->tex2d[u].ch);
    return (orc_handle*)h;
}

void ORC_API(destroy)(orc_handle* h) { delete (Handle*)h; }
void ORC_API(set_pairing)(orc_handle* h, int rule) { ((Handle*)h)->S.pairing = rule; }

/* n 2x2 quads; besides the colours, per pixel: path[n][4] (hash of the discrete path) and dk[n][4] (Durand-Kerner trips); either may be NULL */
int ORC_API(render_quads_ex)(orc_handle* hh, int n, const int32_t* qx, const int32_t* qy, float* out, uint64_t* path, uint32_t* dk,
                             uint32_t sample, orc_stats* stats, int n_threads) {
    Handle* h = (Handle*)hh;
    if (!h || n < 0) return -1;
    orc_real::sr_set_sample(sample);                            /* selects the random-rounding stream of the stochastic variant; unused by the others */
    int nt = resolve_threads(n_threads);
    if (nt > n) nt = n > 0 ? n : 1;
    std::atomic<int> next(0);
    std::vector<Stats> tstats(nt);
    auto work = [&](int tid) {
        const int chunk = 16;
        for (;;) {
            int b = next.fetch_add(chunk);
            if (b >= n) break;
            int e = b + chunk < n ? b + chunk : n;
            for (int i = b; i < e; i++) {
                PixelInfo info[4];
                render_quad(h->S, qx[i], qy[i], out + (size_t)i * 16, &tstats[tid], info);
                for (int l = 0; l < 4; l++) {
                    if (path) path[(size_t)i * 4 + l] = info[l].path;
                    if (dk) dk[(size_t)i * 4 + l] = info[l].dk;
                }
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) th.emplace_back(work, t);
    work(0);
    for (auto& t : th) t.join();
    if (stats) for (int t = 0; t < nt; t++) add_stats(stats, tstats[t], 0);
    if (stats) stats->pixels += (uint64_t)n * 4;
    return 0;
}

int ORC_API(render_quads)(orc_handle* h, int n, const int32_t* qx, const int32_t* qy, float* out, orc_stats* stats, int n_threads) {
    return ORC_API(render_quads_ex)(h, n, qx, qy, out, nullptr, nullptr, 0, stats, n_threads);
}

/* a window; path[hgt][w] and dk[hgt][w] as above */
int ORC_API(render_ex)(orc_handle* hh, int x0, int y0, int w, int hgt, float* out, uint64_t* path, uint32_t* dk, uint32_t sample,
                       orc_stats* stats, int n_threads) {
    Handle* h = (Handle*)hh;
    if (!h || (x0 | y0 | w | hgt) & 1 || w <= 0 || hgt <= 0) return -1;
    orc_real::sr_set_sample(sample);
    int nt = resolve_threads(n_threads);
    int qrows = hgt / 2, qcols = w / 2;
    std::atomic<int> next(0);
    std::vector<Stats> tstats(nt);
    auto work = [&](int tid) {
        for (;;) {
            int r = next.fetch_add(1);
            if (r >= qrows) break;
            for (int c = 0; c < qcols; c++) {
                float px[16];
                PixelInfo info[4];
                render_quad(h->S, x0 + 2 * c, y0 + 2 * r, px, &tstats[tid], info);
                for (int l = 0; l < 4; l++) {
                    size_t p = (size_t)(2 * r + (l >> 1)) * w + (size_t)(2 * c + (l & 1));
                    memcpy(out + p * 4, px + l * 4, 16);
                    if (path) path[p] = info[l].path;
                    if (dk) dk[p] = info[l].dk;
                }
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) th.emplace_back(work, t);
    work(0);
    for (auto& t : th) t.join();
    if (stats) for (int t = 0; t < nt; t++) add_stats(stats, tstats[t], 0);
    if (stats) stats->pixels += (uint64_t)w * hgt;
    return 0;
}

int ORC_API(render)(orc_handle* h, int x0, int y0, int w, int hgt, float* out, orc_stats* stats, int n_threads) {
    return ORC_API(render_ex)(h, x0, y0, w, hgt, out, nullptr, nullptr, 0, stats, n_threads);
}

float ORC_API(calc_inter)(orc_handle* h, const float ro[3], const float rd[3], int32_t* num, int32_t* type) {
    Frag f(((Handle*)h)->S, nullptr, nullptr);
    int n = *num, t = *type;
    real tm = f.calcInter(v3(ro), v3(rd), n, t);
    *num = n; *type = t;
    return to_f(tm);
}

float ORC_API(in_shadow)(orc_handle* h, const float ro[3], const float rd[3], float dist) {
    Frag f(((Handle*)h)->S, nullptr, nullptr);
    return to_f(f.inShadow(v3(ro), v3(rd), dist));
}

int ORC_API(intersect)(orc_handle* h, int type, int index, const float ro_[3], const float rd_[3], float tmin, float* t, int32_t* dk_iters) {
    const Scene& S = ((Handle*)h)->S;
    Frag f(S, nullptr, nullptr);
    vec3 ro = v3(ro_), rd = v3(rd_);
    real tt = 0.0f;
    bool hit = false;
    switch (type) {
        case RTB_TYPE_SPHERE: hit = f.intersectSphere(ro, rd, v4(S.spheres[index].obj), S.spheres[index].hollow != 0, tmin, tt); break;
        case RTB_TYPE_PLANE: hit = f.intersectPlane(ro, rd, v3(S.planes[index].normal), v3(S.planes[index].pos), tmin, tt); break;
        case RTB_TYPE_SURFACE: hit = f.intersectSurface(ro, rd, index, tmin, tt); break;
        case RTB_TYPE_BOX: hit = f.intersectBox(ro, rd, index, tmin, tt); break;
        case RTB_TYPE_TORUS: hit = f.intersectTorus(ro, rd, index, tmin, tt); break;
        case RTB_TYPE_RING: hit = f.intersectRing(ro, rd, index, tmin, tt); break;
        case RTB_TYPE_POINT_LIGHT: hit = f.intersectSphere(ro, rd, v4(S.lights_point[index].pos), false, tmin, tt); break;
        default: return -1;
    }
    if (t) *t = to_f(tt);
    if (dk_iters) *dk_iters = f.last_dk;
    return hit ? 1 : 0;
}

/* debugging aid: the sequence of scene queries of pixel (x, y) — 4 floats per query (kind, type, num, t or shadow); returns the number of queries.
 * (No derivative context: texture LODs are those of an isolated invocation.) */
int ORC_API(trace)(orc_handle* h, int x, int y, uint32_t sample, float* out, int cap) {
    orc_real::sr_set_sample(sample);
    Frag f(((Handle*)h)->S, nullptr, nullptr);
    std::vector<float> tr;
    f.trace = &tr;
    f.fragx = (float)x + 0.5f; f.fragy = (float)y + 0.5f;
    f.main_();
    int n = (int)tr.size() / 4;
    for (int i = 0; i < n * 4 && i < cap * 4; i++) out[i] = tr[i];
    return n;
}

void ORC_API(ray_dir)(orc_handle* h, int x, int y, float out[3]) {
    Frag f(((Handle*)h)->S, nullptr, nullptr);
    f.fragx = (float)x + 0.5f; f.fragy = (float)y + 0.5f;
    vec3 d = f.getRayDir();
    out[0] = to_f(d.x); out[1] = to_f(d.y); out[2] = to_f(d.z);
}

#if ORC_VARIANT == 0    /* the sampler model is fp32 in every variant: probed once */
void orc_sample_cube(orc_handle* h, const float dir[3], float out[4]) {
    glsim::rgba c = glsim::texture_cube(((Handle*)h)->S.skybox, dir[0], dir[1], dir[2]);
    out[0] = c.r; out[1] = c.g; out[2] = c.b; out[3] = c.a;
}

void orc_sample_2d(orc_handle* h, int unit, float u, float v, float lod, float out[4]) {
    glsim::rgba c = glsim::texture_lod(((Handle*)h)->S.tex[unit], u, v, lod);
    out[0] = c.r; out[1] = c.g; out[2] = c.b; out[3] = c.a;
}

int orc_mip_levels(orc_handle* h, int unit) { return (int)((Handle*)h)->S.tex[unit].levels.size(); }
const uint8_t* orc_mip_level(orc_handle* h, int unit, int level, int32_t* w, int32_t* hgt) {
    const glsim::Level& L = ((Handle*)h)->S.tex[unit].levels[level];
    *w = L.w; *hgt = L.h;
    return L.px.data();
}
#endif

}  // extern "C"

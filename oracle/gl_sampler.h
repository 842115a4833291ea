/* gl_sampler.h — CPU model of the OpenGL fixed-function texture units the
 * reference shader calls.  TEST INFRASTRUCTURE (part of oracle/): only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may use it.
 *
 * The reference samples through the GL driver, whose filtering arithmetic is
 * NOT in /root/reference (driver-defined): PARITY UNPINNED for texel filtering.
 * This file restates the OpenGL 3.3 core specification with full fp32 weights:
 *   - sampler state: src/GLWrapper.cpp:308-314 (cubemap: LINEAR/LINEAR,
 *     CLAMP_TO_EDGE, no mips — load_cubemap(faces,false), main.cpp:147; seamless
 *     filtering never enabled) and src/GLWrapper.cpp:336-343 (2-D: wrap REPEAT,
 *     min LINEAR_MIPMAP_LINEAR, mag LINEAR, glGenerateMipmap);
 *   - cube face selection: GL 3.3 spec table 3.19 (major axis, sc, tc, ma);
 *   - level of detail: GL 3.3 spec 3.8.11, rho = max(|dP/dx|, |dP/dy|) in texels.
 * Mip levels: 2x2 box filter, dimension max(1, floor(d/2)), 8-bit storage with
 * round-half-up ((a+b+c+d+2)>>2) — glGenerateMipmap's filter is
 * implementation-defined; the CUDA library builds the identical chain.
 */
#ifndef ORACLE_GL_SAMPLER_H
#define ORACLE_GL_SAMPLER_H

#include <cmath>
#include <cstdint>
#include <vector>

namespace glsim {

struct rgba { float r, g, b, a; };

struct Level { int w = 0, h = 0; std::vector<uint8_t> px; /* RGBA8 */ };

struct Texture2D {
    std::vector<Level> levels;          /* empty => unit not bound: samples return 0,0,0,1 */
    bool bound() const { return !levels.empty(); }
};

struct CubeMap {
    int w = 0, h = 0;
    std::vector<uint8_t> face[6];       /* RGBA8, +X,-X,+Y,-Y,+Z,-Z */
    bool bound() const { return w > 0; }
};

/* any channel count -> RGBA8 the way GL expands GL_RED / GL_RGB / GL_RGBA uploads */
inline void expand_rgba8(const uint8_t* src, int w, int h, int ch, std::vector<uint8_t>& dst) {
    dst.resize((size_t)w * h * 4);
    for (size_t i = 0; i < (size_t)w * h; i++) {
        uint8_t r = src[i * ch], g = 0, b = 0, a = 255;
        if (ch >= 3) { g = src[i * ch + 1]; b = src[i * ch + 2]; }
        if (ch == 2) { g = src[i * ch + 1]; }
        if (ch == 4) a = src[i * ch + 3];
        dst[i * 4 + 0] = r; dst[i * 4 + 1] = g; dst[i * 4 + 2] = b; dst[i * 4 + 3] = a;
    }
}

inline void build_mips(Texture2D& t, const uint8_t* src, int w, int h, int ch) {
    t.levels.clear();
    t.levels.emplace_back();
    t.levels[0].w = w; t.levels[0].h = h;
    expand_rgba8(src, w, h, ch, t.levels[0].px);
    while (t.levels.back().w > 1 || t.levels.back().h > 1) {
        const Level& s = t.levels.back();
        Level d;
        d.w = s.w > 1 ? s.w / 2 : 1;
        d.h = s.h > 1 ? s.h / 2 : 1;
        d.px.resize((size_t)d.w * d.h * 4);
        for (int y = 0; y < d.h; y++)
            for (int x = 0; x < d.w; x++) {
                int x0 = s.w > 1 ? 2 * x : 0, x1 = s.w > 1 ? 2 * x + 1 : 0;
                int y0 = s.h > 1 ? 2 * y : 0, y1 = s.h > 1 ? 2 * y + 1 : 0;
                for (int c = 0; c < 4; c++) {
                    int sum = s.px[((size_t)y0 * s.w + x0) * 4 + c] + s.px[((size_t)y0 * s.w + x1) * 4 + c] +
                              s.px[((size_t)y1 * s.w + x0) * 4 + c] + s.px[((size_t)y1 * s.w + x1) * 4 + c];
                    d.px[((size_t)y * d.w + x) * 4 + c] = (uint8_t)((sum + 2) >> 2);
                }
            }
        t.levels.push_back(std::move(d));
    }
}

inline rgba texel(const uint8_t* px, int w, int x, int y) {
    const uint8_t* p = px + ((size_t)y * w + x) * 4;
    /* unorm8 -> float: c / 255 (GL 3.3 spec eq. 2.1) */
    return { p[0] / 255.0f, p[1] / 255.0f, p[2] / 255.0f, p[3] / 255.0f };
}

inline rgba lerp_bilinear(rgba c00, rgba c10, rgba c01, rgba c11, float fx, float fy) {
    float gx = 1.0f - fx, gy = 1.0f - fy;
    rgba top = { c00.r * gx + c10.r * fx, c00.g * gx + c10.g * fx, c00.b * gx + c10.b * fx, c00.a * gx + c10.a * fx };
    rgba bot = { c01.r * gx + c11.r * fx, c01.g * gx + c11.g * fx, c01.b * gx + c11.b * fx, c01.a * gx + c11.a * fx };
    return { top.r * gy + bot.r * fy, top.g * gy + bot.g * fy, top.b * gy + bot.b * fy, top.a * gy + bot.a * fy };
}

inline int wrap_repeat(int i, int n) { int m = i % n; return m < 0 ? m + n : m; }
inline int wrap_clamp(int i, int n) { return i < 0 ? 0 : (i >= n ? n - 1 : i); }

/* LINEAR filter on one level, wrap REPEAT (GL 3.3 spec 3.8.11 eq. 3.25-3.27) */
inline rgba bilinear_repeat(const Level& L, float s, float t) {
    float u = s * (float)L.w - 0.5f, v = t * (float)L.h - 0.5f;
    float fu = floorf(u), fv = floorf(v);
    float fx = u - fu, fy = v - fv;
    /* |u| can be huge for far-away repeats; reduce in float before the int cast */
    int i0 = wrap_repeat((int)fmodf(fu, (float)L.w), L.w), j0 = wrap_repeat((int)fmodf(fv, (float)L.h), L.h);
    int i1 = wrap_repeat(i0 + 1, L.w), j1 = wrap_repeat(j0 + 1, L.h);
    const uint8_t* p = L.px.data();
    return lerp_bilinear(texel(p, L.w, i0, j0), texel(p, L.w, i1, j0), texel(p, L.w, i0, j1), texel(p, L.w, i1, j1), fx, fy);
}

/* textureLod(sampler2D, uv, lod): explicit lambda, LINEAR_MIPMAP_LINEAR / LINEAR */
inline rgba texture_lod(const Texture2D& t, float s, float tt, float lod) {
    if (!t.bound()) return { 0, 0, 0, 1 };
    int q = (int)t.levels.size() - 1;
    float lam = lod;
    if (!(lam > 0.0f)) lam = 0.0f;          /* also catches NaN and -inf (log2(0)) */
    if (lam > (float)q) lam = (float)q;
    int d1 = (int)floorf(lam);
    float f = lam - (float)d1;
    rgba a = bilinear_repeat(t.levels[d1], s, tt);
    if (f == 0.0f || d1 >= q) return a;
    rgba b = bilinear_repeat(t.levels[d1 + 1], s, tt);
    float g = 1.0f - f;
    return { a.r * g + b.r * f, a.g * g + b.g * f, a.b * g + b.b * f, a.a * g + b.a * f };
}

/* texture(sampler2D, uv) with implicit LOD from screen-space derivatives of uv
 * (dudx.. are the differences across the 2x2 quad, in uv units). */
inline float implicit_lod(const Texture2D& t, float dudx, float dvdx, float dudy, float dvdy) {
    if (!t.bound()) return 0.0f;
    float w = (float)t.levels[0].w, h = (float)t.levels[0].h;
    float ax = dudx * w, bx = dvdx * h, ay = dudy * w, by = dvdy * h;
    float rx = sqrtf(ax * ax + bx * bx), ry = sqrtf(ay * ay + by * by);
    float rho = rx < ry ? ry : rx;
    return log2f(rho);
}

/* texture(samplerCube, dir): face select + LINEAR + CLAMP_TO_EDGE, per-face (non-seamless) */
inline rgba texture_cube(const CubeMap& c, float rx, float ry, float rz) {
    if (!c.bound()) return { 0, 0, 0, 1 };
    float ax = fabsf(rx), ay = fabsf(ry), az = fabsf(rz);
    int face; float sc, tc, ma;
    if (ax >= ay && ax >= az) { ma = ax; if (rx >= 0) { face = 0; sc = -rz; tc = -ry; } else { face = 1; sc = rz; tc = -ry; } }
    else if (ay >= az)        { ma = ay; if (ry >= 0) { face = 2; sc = rx; tc = rz; } else { face = 3; sc = rx; tc = -rz; } }
    else                      { ma = az; if (rz >= 0) { face = 4; sc = rx; tc = -ry; } else { face = 5; sc = -rx; tc = -ry; } }
    float s = 0.5f * (sc / ma + 1.0f), t = 0.5f * (tc / ma + 1.0f);
    float u = s * (float)c.w - 0.5f, v = t * (float)c.h - 0.5f;
    float fu = floorf(u), fv = floorf(v);
    float fx = u - fu, fy = v - fv;
    int i0 = wrap_clamp((int)fu, c.w), i1 = wrap_clamp((int)fu + 1, c.w);
    int j0 = wrap_clamp((int)fv, c.h), j1 = wrap_clamp((int)fv + 1, c.h);
    const uint8_t* p = c.face[face].data();
    return lerp_bilinear(texel(p, c.w, i0, j0), texel(p, c.w, i1, j0), texel(p, c.w, i0, j1), texel(p, c.w, i1, j1), fx, fy);
}

}  // namespace glsim
#endif

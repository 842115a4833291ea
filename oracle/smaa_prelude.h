/* smaa_prelude.h — just enough shading language in C++ to compile the reference's assets/shaders/SMAA.h TEXT on the CPU.
 *
 * TEST INFRASTRUCTURE; used only by smaa_ref_harness.cpp -> oracle/_ref/libsmaa_ref.so (build_smaa_ref.py).
 * SMAA.h is written against a porting layer (SMAA_CUSTOM_SL, SMAA.h:170-172): vector types with swizzles, mad / saturate /
 * lerp, and the SMAASample* texture macros.  This header supplies that layer:
 *   - float2/3/4, int2, bool2/4 with exactly the swizzles SMAA.h uses (proxy members that read and WRITE through, like GLSL's),
 *   - Ref2 / Ref4: what an `inout floatN` parameter becomes (build_smaa_ref.py rewrites the qualifier), so that swizzles can be
 *     passed by reference as in GLSL (SMAAMovc(cond.xy, variable.xy, ...), SMAA.h:627-635, :1203),
 *   - fp32 arithmetic, component-wise, nothing fused (SMAA_GLSL_3: `#define mad(a, b, c) (a * b + c)`, SMAA.h:576),
 *   - the sampler model of gl_sampler.h's conventions: unorm8 texels, LINEAR, CLAMP_TO_EDGE, texel centres at (i + 0.5) / size
 *     (GLWrapper.cpp:215-221, SMAA_Builder.h:51-79).  Bilinear weights are exact fp32; real GPUs quantise them to 8 bits, which
 *     SMAA's thresholds are designed to tolerate (SMAA.h:1106 "Rounding prevents precision errors of bilinear filtering").
 */
#ifndef SMAA_PRELUDE_H
#define SMAA_PRELUDE_H

#include <cmath>
#include <cstdint>

namespace smaa_sl {

struct float2; struct float3; struct float4;

/* ---- swizzle proxies: N components of a parent's storage; aliases the parent through a union (same trick as glm's) ---- */
template <int A, int B> struct Sw2 {
    float v[4];
    operator float2() const;
    Sw2& operator=(const float2& o);
    Sw2& operator=(const Sw2& o) { float a = o.v[A], b = o.v[B]; v[A] = a; v[B] = b; return *this; }
    template <int C, int D> Sw2& operator=(const Sw2<C, D>& o) { float a = o.v[C], b = o.v[D]; v[A] = a; v[B] = b; return *this; }
    Sw2& operator*=(const float2& o);
};
template <int A, int B, int C> struct Sw3 {
    float v[4];
    operator float3() const;
    Sw3& operator=(const float3& o);
};
template <int A, int B, int C, int D> struct Sw4 {
    float v[4];
    operator float4() const;
    Sw4& operator=(const float4& o);
};

struct float2 {
    union { struct { float x, y; }; struct { float r, g; }; float v[4];
            Sw2<0, 1> xy, rg; Sw2<1, 0> yx, gr; Sw2<0, 0> xx; Sw2<1, 1> yy;
            Sw4<0, 1, 0, 1> xyxy; Sw4<0, 0, 1, 1> xxyy; };
    float2() : x(0), y(0) {}
    float2(float x_, float y_) : x(x_), y(y_) {}
    float2(const float2& o) : x(o.x), y(o.y) {}
    float2& operator=(const float2& o) { x = o.x; y = o.y; return *this; }
    float2& operator*=(const float2& o) { x *= o.x; y *= o.y; return *this; }
    float2& operator/=(float s) { x /= s; y /= s; return *this; }
    float2& operator+=(const float2& o) { x += o.x; y += o.y; return *this; }
};
struct float3 {
    union { struct { float x, y, z; }; struct { float r, g, b; }; float v[4];
            Sw2<0, 1> xy; Sw2<1, 2> yz; Sw2<0, 0> xx; Sw2<2, 1> zy; Sw2<0, 2> xz; Sw3<0, 1, 2> xyz, rgb; Sw3<1, 0, 2> grb;
            Sw4<0, 1, 2, 1> xyzy; Sw4<0, 1, 0, 2> xyxz; };
    float3() : x(0), y(0), z(0) {}
    float3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    float3(const float2& a, float z_) : x(a.x), y(a.y), z(z_) {}
    float3(const float3& o) : x(o.x), y(o.y), z(o.z) {}
    float3& operator=(const float3& o) { x = o.x; y = o.y; z = o.z; return *this; }
};
struct float4 {
    union { struct { float x, y, z, w; }; struct { float r, g, b, a; }; float v[4];
            Sw2<0, 1> xy, rg; Sw2<2, 3> zw, ba; Sw2<0, 2> xz, rb; Sw2<1, 3> yw; Sw2<3, 2> wz; Sw2<0, 3> ra; Sw2<2, 2> zz; Sw2<3, 3> ww; Sw2<0, 0> xx; Sw2<1, 0> gr;
            Sw3<0, 1, 2> xyz, rgb; Sw4<0, 1, 2, 3> xyzw; Sw4<1, 0, 3, 2> yxwz; Sw4<0, 1, 0, 1> xyxy; Sw4<0, 0, 1, 1> xxyy; };
    float4() : x(0), y(0), z(0), w(0) {}
    float4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
    float4(const float2& a, float z_, float w_) : x(a.x), y(a.y), z(z_), w(w_) {}
    float4(const float2& a, const float2& b_) : x(a.x), y(a.y), z(b_.x), w(b_.y) {}
    float4(const float4& o) : x(o.x), y(o.y), z(o.z), w(o.w) {}
    float4& operator=(const float4& o) { x = o.x; y = o.y; z = o.z; w = o.w; return *this; }
    float4& operator+=(const float4& o) { x += o.x; y += o.y; z += o.z; w += o.w; return *this; }
};
struct int2 { int x, y; int2(int x_, int y_) : x(x_), y(y_) {} };
struct bool2 {
    bool x, y;
    bool2(bool x_, bool y_) : x(x_), y(y_) {}
    explicit bool2(const float2& f) : x(f.x != 0.f), y(f.y != 0.f) {}          /* bool2(step(...)): GLSL's float -> bool conversion */
};
struct bool4 {
    bool x, y, z, w;
    bool2 xy, zw;
    bool4(bool x_, bool y_, bool z_, bool w_) : x(x_), y(y_), z(z_), w(w_), xy(x_, y_), zw(z_, w_) {}
};

template <int A, int B> Sw2<A, B>::operator float2() const { return float2(v[A], v[B]); }
template <int A, int B> Sw2<A, B>& Sw2<A, B>::operator=(const float2& o) { v[A] = o.x; v[B] = o.y; return *this; }
template <int A, int B> Sw2<A, B>& Sw2<A, B>::operator*=(const float2& o) { v[A] *= o.x; v[B] *= o.y; return *this; }
template <int A, int B, int C> Sw3<A, B, C>::operator float3() const { return float3(v[A], v[B], v[C]); }
template <int A, int B, int C> Sw3<A, B, C>& Sw3<A, B, C>::operator=(const float3& o) { v[A] = o.x; v[B] = o.y; v[C] = o.z; return *this; }
template <int A, int B, int C, int D> Sw4<A, B, C, D>::operator float4() const { return float4(v[A], v[B], v[C], v[D]); }
template <int A, int B, int C, int D> Sw4<A, B, C, D>& Sw4<A, B, C, D>::operator=(const float4& o) {
    float t0 = o.x, t1 = o.y, t2 = o.z, t3 = o.w; v[A] = t0; v[B] = t1; v[C] = t2; v[D] = t3; return *this; }

/* `inout float2` / `inout float4` parameters: references to the components of whatever lvalue (vector or swizzle) is passed */
struct Ref2 {
    float& x; float& y;
    Ref2(float2& f) : x(f.x), y(f.y) {}
    template <int A, int B> Ref2(Sw2<A, B>& s) : x(s.v[A]), y(s.v[B]) {}
    Ref2(float& x_, float& y_) : x(x_), y(y_) {}
    Ref2& operator*=(const float2& o) { x *= o.x; y *= o.y; return *this; }
};
struct Ref4 {
    float& x; float& y; float& z; float& w;
    Ref2 xy, zw;
    Ref4(float4& f) : x(f.x), y(f.y), z(f.z), w(f.w), xy(f.x, f.y), zw(f.z, f.w) {}
};

/* ---- component-wise fp32 arithmetic ---- */
#define SMAA_OP2(op) \
    inline float2 operator op(const float2& a, const float2& b) { return float2(a.x op b.x, a.y op b.y); } \
    inline float2 operator op(const float2& a, float s) { return float2(a.x op s, a.y op s); } \
    inline float2 operator op(float s, const float2& a) { return float2(s op a.x, s op a.y); } \
    inline float3 operator op(const float3& a, const float3& b) { return float3(a.x op b.x, a.y op b.y, a.z op b.z); } \
    inline float3 operator op(const float3& a, float s) { return float3(a.x op s, a.y op s, a.z op s); } \
    inline float3 operator op(float s, const float3& a) { return float3(s op a.x, s op a.y, s op a.z); } \
    inline float4 operator op(const float4& a, const float4& b) { return float4(a.x op b.x, a.y op b.y, a.z op b.z, a.w op b.w); } \
    inline float4 operator op(const float4& a, float s) { return float4(a.x op s, a.y op s, a.z op s, a.w op s); } \
    inline float4 operator op(float s, const float4& a) { return float4(s op a.x, s op a.y, s op a.z, s op a.w); }
SMAA_OP2(+) SMAA_OP2(-) SMAA_OP2(*) SMAA_OP2(/)
#undef SMAA_OP2
inline float2 operator-(const float2& a) { return float2(-a.x, -a.y); }
inline float4 operator-(const float4& a) { return float4(-a.x, -a.y, -a.z, -a.w); }

inline float abs(float a) { return std::fabs(a); }
inline float2 abs(const float2& a) { return float2(std::fabs(a.x), std::fabs(a.y)); }
inline float3 abs(const float3& a) { return float3(std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)); }
inline float4 abs(const float4& a) { return float4(std::fabs(a.x), std::fabs(a.y), std::fabs(a.z), std::fabs(a.w)); }
inline float max(float a, float b) { return a < b ? b : a; }
inline float2 max(const float2& a, const float2& b) { return float2(max(a.x, b.x), max(a.y, b.y)); }
inline float step(float edge, float x) { return x < edge ? 0.0f : 1.0f; }
inline float2 step(const float2& e, const float2& x) { return float2(step(e.x, x.x), step(e.y, x.y)); }
inline float2 step(float e, const float2& x) { return float2(step(e, x.x), step(e, x.y)); }
inline float dot(const float2& a, const float2& b) { return a.x * b.x + a.y * b.y; }
inline float dot(const float3& a, const float3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float dot(const float4& a, const float4& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline float round(float a) { return std::floor(a + 0.5f); }                   /* GLSL round(): halves are implementation-defined; SMAA never produces one */
inline float2 round(const float2& a) { return float2(round(a.x), round(a.y)); }
inline float4 round(const float4& a) { return float4(round(a.x), round(a.y), round(a.z), round(a.w)); }
inline float sqrt(float a) { return std::sqrt(a); }
inline float2 sqrt(const float2& a) { return float2(std::sqrt(a.x), std::sqrt(a.y)); }
inline float length(const float2& a) { return std::sqrt(dot(a, a)); }
inline float clampf(float a, float lo, float hi) { return a < lo ? lo : (a > hi ? hi : a); }
inline float saturate(float a) { return clampf(a, 0.0f, 1.0f); }
inline float2 saturate(const float2& a) { return float2(saturate(a.x), saturate(a.y)); }
inline float4 lerp(const float4& a, const float4& b, float t) { return a * (1.0f - t) + b * t; }
#define mad(a, b, c) ((a) * (b) + (c))

/* ---- textures: unorm8, LINEAR, CLAMP_TO_EDGE ---- */
struct Tex {
    const uint8_t* px; int w, h, ch;        /* ch interleaved channels per texel */
    float4 texel(int x, int y) const {
        x = x < 0 ? 0 : (x >= w ? w - 1 : x); y = y < 0 ? 0 : (y >= h ? h - 1 : y);
        const uint8_t* p = px + ((size_t)y * w + x) * ch;
        return float4(p[0] / 255.0f, ch > 1 ? p[1] / 255.0f : 0.0f, ch > 2 ? p[2] / 255.0f : 0.0f, ch > 3 ? p[3] / 255.0f : 1.0f);
    }
    float4 sample(const float2& uv) const {
        const float u = uv.x * (float)w - 0.5f, v = uv.y * (float)h - 0.5f;
        const float fu = std::floor(u), fv = std::floor(v);
        const float ax = u - fu, ay = v - fv;
        const int x0 = (int)fu, y0 = (int)fv;
        const float4 top = texel(x0, y0) * (1.0f - ax) + texel(x0 + 1, y0) * ax;
        const float4 bot = texel(x0, y0 + 1) * (1.0f - ax) + texel(x0 + 1, y0 + 1) * ax;
        return top * (1.0f - ay) + bot * ay;
    }
};

}  // namespace smaa_sl
#endif

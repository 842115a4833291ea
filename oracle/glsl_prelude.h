/* glsl_prelude.h — just enough GLSL 3.30 in C++ to compile the reference's
 * assets/shaders/rt.frag TEXT (transformed mechanically by build_ref.py) on the
 * CPU.  TEST INFRASTRUCTURE; used only by ref_harness.cpp -> oracle/_ref/libref.so.
 *
 * Storage types are tiny structs with GLSL's implicit int->float conversions
 * (which glm's templates reject); every built-in with a defining formula in the
 * GLSL specification is DELEGATED to glm 0.9.9.7 — the library the reference
 * itself vendors (external_sources/glm) — so the arithmetic of normalize /
 * reflect / refract / dot / clamp / step / sign here is independent of the
 * restatement in rt_oracle.cpp.
 */
#ifndef GLSL_PRELUDE_H
#define GLSL_PRELUDE_H

#include <cmath>
#include <glm/glm.hpp>

namespace glsl {

struct vec2 {
    float x, y;
    vec2() : x(0), y(0) {}
    explicit vec2(float s) : x(s), y(s) {}
    vec2(float x_, float y_) : x(x_), y(y_) {}
    vec2(const glm::vec2& g) : x(g.x), y(g.y) {}
    operator glm::vec2() const { return glm::vec2(x, y); }
};
struct vec3 {
    union { struct { float x, y, z; }; struct { float r, g, b; }; };
    vec3() : x(0), y(0), z(0) {}
    explicit vec3(float s) : x(s), y(s), z(s) {}
    vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    vec3(vec2 v, float z_) : x(v.x), y(v.y), z(z_) {}
    vec3(const glm::vec3& g_) : x(g_.x), y(g_.y), z(g_.z) {}
    operator glm::vec3() const { return glm::vec3(x, y, z); }
    vec2 xy() const { return vec2(x, y); }
    vec2 zy() const { return vec2(z, y); }
    vec2 zx() const { return vec2(z, x); }
    vec3 xyz() const { return *this; }
    vec3 yzx() const { return vec3(y, z, x); }
    vec3 zxy() const { return vec3(z, x, y); }
};
struct vec4 {
    union { struct { float x, y, z, w; }; struct { float r, g, b, a; }; };
    vec4() : x(0), y(0), z(0), w(0) {}
    explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
    vec4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
    vec4(vec3 v, float w_) : x(v.x), y(v.y), z(v.z), w(w_) {}
    vec4(const glm::vec4& g_) : x(g_.x), y(g_.y), z(g_.z), w(g_.w) {}
    operator glm::vec4() const { return glm::vec4(x, y, z, w); }
    vec2 xy() const { return vec2(x, y); }
    vec3 xyz() const { return vec3(x, y, z); }
    vec3 rgb() const { return vec3(x, y, z); }
};
struct bvec3 {
    bool x, y, z;
    explicit bvec3(bool s) : x(s), y(s), z(s) {}
    bvec3(bool x_, bool y_, bool z_) : x(x_), y(y_), z(z_) {}
};
inline bool operator==(bvec3 a, bvec3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

#define G2(v) glm::vec2(v)
#define G3(v) glm::vec3(v)
#define G4(v) glm::vec4(v)

/* arithmetic (component-wise, GLSL 5.9) */
inline vec2 operator+(vec2 a, vec2 b) { return G2(a) + G2(b); }
inline vec2 operator-(vec2 a, vec2 b) { return G2(a) - G2(b); }
inline vec2 operator*(vec2 a, vec2 b) { return G2(a) * G2(b); }
inline vec2 operator*(vec2 a, float s) { return G2(a) * s; }
inline vec2 operator*(float s, vec2 a) { return s * G2(a); }
inline vec2 operator/(vec2 a, float s) { return G2(a) / s; }
inline vec2& operator-=(vec2& a, vec2 b) { a = a - b; return a; }

inline vec3 operator+(vec3 a, vec3 b) { return G3(a) + G3(b); }
inline vec3 operator-(vec3 a, vec3 b) { return G3(a) - G3(b); }
inline vec3 operator-(vec3 a) { return -G3(a); }
inline vec3 operator*(vec3 a, vec3 b) { return G3(a) * G3(b); }
inline vec3 operator*(vec3 a, float s) { return G3(a) * s; }
inline vec3 operator*(float s, vec3 a) { return s * G3(a); }
inline vec3 operator/(vec3 a, float s) { return G3(a) / s; }
inline vec3 operator/(float s, vec3 a) { return s / G3(a); }
inline vec3 operator-(float s, vec3 a) { return s - G3(a); }
inline vec3& operator+=(vec3& a, vec3 b) { a = a + b; return a; }
inline vec3& operator*=(vec3& a, vec3 b) { a = a * b; return a; }
inline vec3& operator*=(vec3& a, float s) { a = a * s; return a; }

inline vec4 operator+(vec4 a, vec4 b) { return G4(a) + G4(b); }
inline vec4 operator*(vec4 a, float s) { return G4(a) * s; }
inline vec4 operator*(float s, vec4 a) { return s * G4(a); }
inline bool operator!=(vec4 a, vec4 b) { return G4(a) != G4(b); }
inline bool operator==(vec4 a, vec4 b) { return G4(a) == G4(b); }

/* built-ins, delegated to glm */
inline float dot(vec2 a, vec2 b) { return glm::dot(G2(a), G2(b)); }
inline float dot(vec3 a, vec3 b) { return glm::dot(G3(a), G3(b)); }
inline float dot(vec4 a, vec4 b) { return glm::dot(G4(a), G4(b)); }
inline float length(vec3 a) { return glm::length(G3(a)); }
inline vec2 normalize(vec2 a) { return glm::normalize(G2(a)); }
inline vec3 normalize(vec3 a) { return glm::normalize(G3(a)); }
inline vec3 reflect(vec3 I, vec3 N) { return glm::reflect(G3(I), G3(N)); }
inline vec3 refract(vec3 I, vec3 N, float eta) { return glm::refract(G3(I), G3(N), eta); }
inline float clamp(float x, float lo, float hi) { return glm::clamp(x, lo, hi); }
inline vec3 clamp(vec3 x, vec3 lo, vec3 hi) { return glm::clamp(G3(x), G3(lo), G3(hi)); }
inline float min(float a, float b) { return glm::min(a, b); }
inline float max(float a, float b) { return glm::max(a, b); }
inline vec3 max(vec3 a, vec3 b) { return glm::max(G3(a), G3(b)); }
inline float abs(float a) { return glm::abs(a); }
inline vec3 abs(vec3 a) { return glm::abs(G3(a)); }
inline vec4 abs(vec4 a) { return glm::abs(G4(a)); }
inline vec3 sign(vec3 a) { return glm::sign(G3(a)); }
inline vec3 step(vec3 edge, vec3 x) { return glm::step(G3(edge), G3(x)); }
inline float sqrt(float a) { return glm::sqrt(a); }
inline float pow(float a, float b) { return glm::pow(a, b); }
inline float exp(float a) { return glm::exp(a); }
inline vec3 exp(vec3 a) { return glm::exp(G3(a)); }
inline float log2(float a) { return glm::log2(a); }
inline float atan(float y, float x) { return glm::atan(y, x); }
inline float asin(float a) { return glm::asin(a); }
inline bvec3 greaterThan(vec3 a, vec3 b) { glm::bvec3 r = glm::greaterThan(G3(a), G3(b)); return bvec3(r.x, r.y, r.z); }
inline bvec3 lessThan(vec3 a, vec3 b) { glm::bvec3 r = glm::lessThan(G3(a), G3(b)); return bvec3(r.x, r.y, r.z); }

#undef G2
#undef G3
#undef G4

/* sampler handles: the harness binds them to glsim textures */
struct sampler2D { int unit = 0; };
struct samplerCube { int unit = 0; };

}  // namespace glsl
#endif

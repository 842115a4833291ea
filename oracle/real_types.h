/* real_types.h — the scalar types the oracle restatement is instantiated with.
 *
 * TEST INFRASTRUCTURE (see rt_oracle.h).  rt_oracle.cpp is compiled three times:
 *
 *   ORC_VARIANT 0   real = float     the restatement proper, IEEE fp32, pinned bit for bit against oracle/_ref
 *                                    (the reference's own rt.frag compiled as C++); entry points orc_*
 *   ORC_VARIANT 1   real = double    the same control flow evaluated in fp64 (SURVEY.md 8c "fp64 classification mode");
 *                                    inputs are the same fp32 values, constants the same fp32 literals; entry points orc64_*
 *   ORC_VARIANT 2   real = sr32      fp32 with STOCHASTIC rounding: every + - * / sqrt and library call is computed in fp64
 *                                    and rounded to one of the two neighbouring fp32 values, chosen by a hash of the exact
 *                                    result keyed by a sample number (CESTAC-style random rounding, but a deterministic
 *                                    function: reproducible and thread-independent); entry points orcsr_*
 *
 * Variants 1 and 2 answer one question about a pixel: does the shader's arithmetic DETERMINE its colour to the 1e-4
 * tolerance, or would two conformant GLSL implementations (GLSL 4.x 4.7.1: a*b+c may be contracted, division and
 * library functions are only accurate to a few ulp) already disagree?  tests/envelope.py builds the envelope criterion
 * for the fused CUDA build from them.
 */
#ifndef ORC_REAL_TYPES_H
#define ORC_REAL_TYPES_H

#include <cmath>
#include <cstdint>

#ifndef ORC_VARIANT
#define ORC_VARIANT 0
#endif

namespace orc_real {

/* The rounding decision is a pseudo-random FUNCTION of the exact result and of the sample number (no generator state: the
 * same pixel gives the same colour whatever thread renders it, and sample k of a frame is reproducible): the top bits of a
 * multiplicative hash of the fp64 bit pattern, keyed by the sample.
 * `sample` = index | amplitude << 16.  Amplitude 0 or 1: the result is one of the two fp32 neighbours of the exact value
 * (what a conformant implementation may return).  Amplitude A > 1: round to nearest, then move by a uniform integer in
 * [-A, A] ulps — DELIBERATELY more noise than any implementation has; tests/envelope.py uses it to find every pixel whose
 * discrete path (which primitive is nearest, shadowed or not) hangs on a margin of a few ulps. */
struct SrKey { uint64_t key; uint32_t amp; };
inline SrKey& sr_key() { static SrKey k = { 0x9e3779b97f4a7c15ull, 1 }; return k; }      /* set once per render call, read-only while threads run */
inline void sr_set_sample(uint32_t sample) {
    sr_key().key = ((uint64_t)(sample & 0xffffu) + 1) * 0x9e3779b97f4a7c15ull | 1;
    sr_key().amp = (sample >> 16) > 1 ? (sample >> 16) : 1;
}
/* Exact values, infinities, NaN and values next to the subnormal range keep round-to-nearest.  Branch-free in the common
 * (amplitude 1) case: the decision bit is unpredictable by construction. */
inline float sr_round(double d) {
    const float f = (float)d;                                   /* round to nearest */
    const double fd = (double)f;
    uint64_t h; __builtin_memcpy(&h, &d, 8);
    h = (h ^ sr_key().key) * 0xbf58476d1ce4e5b9ull;
    uint32_t u; __builtin_memcpy(&u, &f, 4);
    const uint32_t mag = u & 0x7fffffffu;
    const uint32_t eligible = (uint32_t)(fd != d) & (uint32_t)(mag >= 0x01000000u) & (uint32_t)(mag < 0x7f000000u);
    const uint32_t amp = sr_key().amp;
    if (amp > 1) {
        const int32_t off = (int32_t)((h >> 33) % (2 * amp + 1)) - (int32_t)amp;
        u += eligible * (uint32_t)off;
    } else {
        const uint32_t away = (uint32_t)((fd < d) == ((int32_t)u >= 0));  /* one step from f towards d grows the magnitude */
        u += (eligible & (uint32_t)(h >> 63)) * (away * 2u - 1u);
    }
    float g; __builtin_memcpy(&g, &u, 4);
    return g;
}

struct sr32 {
    float v;
    sr32() = default;
    constexpr sr32(float x) : v(x) {}
    explicit constexpr operator float() const { return v; }
};
inline sr32 operator+(sr32 a, sr32 b) { return sr32(sr_round((double)a.v + (double)b.v)); }
inline sr32 operator-(sr32 a, sr32 b) { return sr32(sr_round((double)a.v - (double)b.v)); }
inline sr32 operator*(sr32 a, sr32 b) { return sr32(sr_round((double)a.v * (double)b.v)); }
inline sr32 operator/(sr32 a, sr32 b) { return sr32(sr_round((double)a.v / (double)b.v)); }
inline sr32 operator-(sr32 a) { return sr32(-a.v); }
inline sr32& operator+=(sr32& a, sr32 b) { a = a + b; return a; }
inline sr32& operator*=(sr32& a, sr32 b) { a = a * b; return a; }
inline bool operator<(sr32 a, sr32 b) { return a.v < b.v; }
inline bool operator>(sr32 a, sr32 b) { return a.v > b.v; }
inline bool operator<=(sr32 a, sr32 b) { return a.v <= b.v; }
inline bool operator>=(sr32 a, sr32 b) { return a.v >= b.v; }
inline bool operator==(sr32 a, sr32 b) { return a.v == b.v; }
inline bool operator!=(sr32 a, sr32 b) { return a.v != b.v; }

}  // namespace orc_real

#if ORC_VARIANT == 0
typedef float real;
#define ORC_API(name) orc_##name
#elif ORC_VARIANT == 1
typedef double real;
#define ORC_API(name) orc64_##name
#else
typedef orc_real::sr32 real;
#define ORC_API(name) orcsr_##name
#endif

/* library functions of GLSL on `real` (fp32: the same libm calls the restatement always made) */
inline float r_sqrt(float x) { return sqrtf(x); }
inline float r_abs(float x) { return fabsf(x); }
inline float r_pow(float x, float y) { return powf(x, y); }
inline float r_exp(float x) { return expf(x); }
inline float r_atan2(float y, float x) { return atan2f(y, x); }
inline float r_asin(float x) { return asinf(x); }
inline float r_log2(float x) { return log2f(x); }
inline float to_f(float x) { return x; }

inline double r_sqrt(double x) { return std::sqrt(x); }
inline double r_abs(double x) { return std::fabs(x); }
inline double r_pow(double x, double y) { return std::pow(x, y); }
inline double r_exp(double x) { return std::exp(x); }
inline double r_atan2(double y, double x) { return std::atan2(y, x); }
inline double r_asin(double x) { return std::asin(x); }
inline double r_log2(double x) { return std::log2(x); }
inline float to_f(double x) { return (float)x; }

inline orc_real::sr32 r_sqrt(orc_real::sr32 x) { return orc_real::sr32(orc_real::sr_round(std::sqrt((double)x.v))); }
inline orc_real::sr32 r_abs(orc_real::sr32 x) { return orc_real::sr32(fabsf(x.v)); }
inline orc_real::sr32 r_pow(orc_real::sr32 x, orc_real::sr32 y) { return orc_real::sr32(orc_real::sr_round(std::pow((double)x.v, (double)y.v))); }
inline orc_real::sr32 r_exp(orc_real::sr32 x) { return orc_real::sr32(orc_real::sr_round(std::exp((double)x.v))); }
inline orc_real::sr32 r_atan2(orc_real::sr32 y, orc_real::sr32 x) { return orc_real::sr32(orc_real::sr_round(std::atan2((double)y.v, (double)x.v))); }
inline orc_real::sr32 r_asin(orc_real::sr32 x) { return orc_real::sr32(orc_real::sr_round(std::asin((double)x.v))); }
inline orc_real::sr32 r_log2(orc_real::sr32 x) { return orc_real::sr32(orc_real::sr_round(std::log2((double)x.v))); }
inline float to_f(orc_real::sr32 x) { return x.v; }

#endif

/* smaa_ref_harness.cpp — runs the reference's OWN SMAA.h (assets/shaders/SMAA.h, transformed mechanically by
 * build_smaa_ref.py into _ref/smaa_gen.inc) on the CPU, the three passes exactly as GLWrapper::draw() chains them
 * (GLWrapper.cpp:173-204) with the shaders of SMAA_Builder.h:122-199 and the lookup tables of src/AreaTex.h / SearchTex.h.
 *
 * TEST INFRASTRUCTURE: the checker of the CUDA SMAA passes (csrc/smaa.cu); built only where /root/reference exists, into
 * oracle/_ref/libsmaa_ref.so (git-ignored; travels to the GPU box).  The texture sampler is ours (smaa_prelude.h).
 */
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

#include "smaa_prelude.h"
#include "AreaTex.h"          /* -I /root/reference/src: the reference's tables, compiled into the .so, never into the repository */
#include "SearchTex.h"

namespace smaa_sl {
static thread_local float4 SMAA_RT_METRICS;      /* SMAA_Builder.h:33-35: float4(1/W, 1/H, W, H) */
}

#define SMAA_CUSTOM_SL
#define SMAATexture2D(tex) const Tex& tex
#define SMAATexturePass2D(tex) tex
#define SMAASampleLevelZero(tex, coord) tex.sample(coord)
#define SMAASampleLevelZeroPoint(tex, coord) tex.sample(coord)
#define SMAASampleLevelZeroOffset(tex, coord, offset) tex.sample(float2(coord) + float2((float)(offset).x, (float)(offset).y) * float2(SMAA_RT_METRICS.xy))
#define SMAASample(tex, coord) tex.sample(coord)
#define SMAASamplePoint(tex, coord) tex.sample(coord)
#define SMAASampleOffset(tex, coord, offset) SMAASampleLevelZeroOffset(tex, coord, offset)
#define SMAA_FLATTEN
#define SMAA_BRANCH
#define discard return float2(0.0f, 0.0f)        /* the edge target is cleared to 0 before the pass (GLWrapper.cpp:177-178) */

/* the body once per preset (SMAA_Builder.h:22-28), each in its own namespace */
#define SMAA_BODY "smaa_gen.inc"
namespace p_low { using namespace smaa_sl;
#define SMAA_PRESET_LOW
#include SMAA_BODY
#undef SMAA_PRESET_LOW
}
#include "smaa_undef.inc"
namespace p_medium { using namespace smaa_sl;
#define SMAA_PRESET_MEDIUM
#include SMAA_BODY
#undef SMAA_PRESET_MEDIUM
}
#include "smaa_undef.inc"
namespace p_high { using namespace smaa_sl;
#define SMAA_PRESET_HIGH
#include SMAA_BODY
#undef SMAA_PRESET_HIGH
}
#include "smaa_undef.inc"
namespace p_ultra { using namespace smaa_sl;
#define SMAA_PRESET_ULTRA
#include SMAA_BODY
#undef SMAA_PRESET_ULTRA
}

namespace {
using namespace smaa_sl;

inline uint8_t unorm8(float v) { v = v < 0.f ? 0.f : (v > 1.f ? 1.f : v); if (!(v == v)) v = 0.f; return (uint8_t)(v * 255.0f + 0.5f); }

template <class F> void parallel_rows(int h, int threads, F f) {
    if (threads <= 0) { unsigned hc = std::thread::hardware_concurrency(); threads = hc ? (int)hc : 1; }
    std::atomic<int> next(0);
    auto work = [&] { for (;;) { int y = next.fetch_add(1); if (y >= h) break; f(y); } };
    std::vector<std::thread> th;
    for (int t = 1; t < threads; t++) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
}

#define RUN_PRESET(NS)                                                                                                               \
    static void run_##NS(const uint8_t* rgba, int w, int h, uint8_t* edges, uint8_t* blend, uint8_t* out, int threads) {             \
        const Tex color = { rgba, w, h, 4 }, edgesTex = { edges, w, h, 2 }, blendTex = { blend, w, h, 4 };                            \
        const Tex area = { areaTexBytes, AREATEX_WIDTH, AREATEX_HEIGHT, 2 }, search = { searchTexBytes, SEARCHTEX_WIDTH, SEARCHTEX_HEIGHT, 1 }; \
        const float4 metrics(1.0f / (float)w, 1.0f / (float)h, (float)w, (float)h);                                                  \
        auto tc = [&](int x, int y) { return float2(((float)x + 0.5f) / (float)w, ((float)y + 0.5f) / (float)h); };                  \
        parallel_rows(h, threads, [&](int y) {                                                                                       \
            SMAA_RT_METRICS = metrics;                                                                                               \
            for (int x = 0; x < w; x++) {                                                                                            \
                float4 offset[3];                                                                                                    \
                const float2 t = tc(x, y);                                                                                           \
                NS::SMAAEdgeDetectionVS(t, offset);                                                                                  \
                const float2 e = NS::SMAALumaEdgeDetectionPS(t, offset, color);                                                      \
                edges[((size_t)y * w + x) * 2] = unorm8(e.x); edges[((size_t)y * w + x) * 2 + 1] = unorm8(e.y);                      \
            }                                                                                                                        \
        });                                                                                                                          \
        parallel_rows(h, threads, [&](int y) {                                                                                       \
            SMAA_RT_METRICS = metrics;                                                                                               \
            for (int x = 0; x < w; x++) {                                                                                            \
                float4 offset[3]; float2 pix;                                                                                        \
                const float2 t = tc(x, y);                                                                                           \
                NS::SMAABlendingWeightCalculationVS(t, pix, offset);                                                                 \
                const float4 wgt = NS::SMAABlendingWeightCalculationPS(t, pix, offset, edgesTex, area, search, float4(0.f, 0.f, 0.f, 0.f)); \
                uint8_t* o = blend + ((size_t)y * w + x) * 4;                                                                        \
                o[0] = unorm8(wgt.x); o[1] = unorm8(wgt.y); o[2] = unorm8(wgt.z); o[3] = unorm8(wgt.w);                              \
            }                                                                                                                        \
        });                                                                                                                          \
        parallel_rows(h, threads, [&](int y) {                                                                                       \
            SMAA_RT_METRICS = metrics;                                                                                               \
            for (int x = 0; x < w; x++) {                                                                                            \
                float4 offset;                                                                                                       \
                const float2 t = tc(x, y);                                                                                           \
                NS::SMAANeighborhoodBlendingVS(t, offset);                                                                           \
                const float4 c = NS::SMAANeighborhoodBlendingPS(t, offset, color, blendTex);                                         \
                uint8_t* o = out + ((size_t)y * w + x) * 4;                                                                          \
                o[0] = unorm8(c.x); o[1] = unorm8(c.y); o[2] = unorm8(c.z); o[3] = unorm8(c.w);                                      \
            }                                                                                                                        \
        });                                                                                                                          \
    }
RUN_PRESET(p_low) RUN_PRESET(p_medium) RUN_PRESET(p_high) RUN_PRESET(p_ultra)
}  // namespace

extern "C" {

/* rgba8 [h][w][4] (row 0 = first row of the GL texture = bottom scanline) -> edges [h][w][2], blend [h][w][4], out [h][w][4].
 * preset: 0 LOW, 1 MEDIUM, 2 HIGH, 3 ULTRA (SMAA_Builder.h:9-12; main.cpp:32 uses ULTRA) */
int smaa_ref_run(const uint8_t* rgba8, int w, int h, int preset, uint8_t* edges, uint8_t* blend, uint8_t* out, int threads) {
    if (!rgba8 || !edges || !blend || !out || w <= 0 || h <= 0) return -1;
    switch (preset) {
        case 0: run_p_low(rgba8, w, h, edges, blend, out, threads); break;
        case 1: run_p_medium(rgba8, w, h, edges, blend, out, threads); break;
        case 2: run_p_high(rgba8, w, h, edges, blend, out, threads); break;
        case 3: run_p_ultra(rgba8, w, h, edges, blend, out, threads); break;
        default: return -1;
    }
    return 0;
}

/* the reference's lookup tables (src/AreaTex.h: RG8 160x560, src/SearchTex.h: R8 64x16), for handing them to the CUDA library in tests */
const uint8_t* smaa_ref_area_tex(int* w, int* h) { *w = AREATEX_WIDTH; *h = AREATEX_HEIGHT; return areaTexBytes; }
const uint8_t* smaa_ref_search_tex(int* w, int* h) { *w = SEARCHTEX_WIDTH; *h = SEARCHTEX_HEIGHT; return searchTexBytes; }

}  // extern "C"

/* rtb200.h — C-ABI of librtb200.so, the B200 (sm_100a) replacement for the
 * reference's GLSL ray-trace pass.
 *
 * The reference reaches its hot path (assets/shaders/rt.frag, one fragment-shader
 * invocation per pixel) only through the methods of class GLWrapper
 * (src/GLWrapper.h:17-38).  Every entry point below replaces one of those
 * methods; a reference-side binding is a replacement GLWrapper.cpp whose
 * methods forward here (shipped in raytracing-opengl_b200/host/, see
 * INTEGRATION.md).  Plain pointers and sizes only; no C++/torch types.
 *
 * Conventions: functions return 0 on success, a negative rtb_status otherwise;
 * rtb_last_error() gives the message.  The library copies host inputs before
 * returning (the reference's glBufferData/glTexImage2D semantics).  One context
 * is used from one thread at a time (the reference's GL-context rule).  There
 * is NO CPU fallback: without a CUDA device rtb_create() fails.
 */
#ifndef RTB200_H
#define RTB200_H

#include <stddef.h>
#include <stdint.h>
#include "rtb200_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rtb_ctx rtb_ctx;

enum rtb_status {
    RTB_OK = 0,
    RTB_ERR_INVALID = -1,      /* bad argument (unknown binding, null pointer, size mismatch) */
    RTB_ERR_CUDA = -2,         /* a CUDA runtime call failed */
    RTB_ERR_STATE = -3,        /* call order violated (render before set_defines, ...) */
    RTB_ERR_NO_DEVICE = -4
};

/* Which kernel rtb_render() launches. */
enum rtb_kernel {
    RTB_KERNEL_AUTO = 0,       /* persistent unless a 2-D texture is referenced by the scene */
    RTB_KERNEL_QUAD = 1,       /* warp = 8x4 pixel tile, 2x2 quads in lock step (needed for fwidth/implicit LOD) */
    RTB_KERNEL_PERSISTENT = 2  /* persistent threads, per-lane ray refill (scenes without 2-D textures) */
};

/* Per-frame work counters, filled by rtb_render_counted() (an instrumented,
 * untimed launch).  "ray" = one full-scene query: one calcInter (rt.frag:587)
 * or one inShadow (rt.frag:630) evaluation. */
typedef struct rtb_stats {
    uint64_t pixels;
    uint64_t rays_nearest;         /* calcInter evaluations  */
    uint64_t rays_shadow;          /* inShadow evaluations   */
    uint64_t tests[7];             /* ray-primitive tests per rtb_prim_type (6 = light spheres) */
    uint64_t dk_iterations;        /* Durand-Kerner iterations executed (4 DKsteps each), rt.frag:471-477 */
    uint64_t shaded_hits[7];       /* get_hit_info evaluations per type */
    uint64_t light_evals;          /* calcShade2 evaluations */
    double   flops;                /* algorithmic flops, SURVEY.md 8d constants applied to the counters */
    float    kernel_ms;            /* CUDA-event time of the last TIMED rtb_render on this context */
    int32_t  kernel_used;          /* rtb_kernel actually launched */
    int32_t  grid, block, smem_bytes;
} rtb_stats;

/* GLWrapper::GLWrapper(w,h,fullScreen) + init_window()  (GLWrapper.h:17,25; GLWrapper.cpp:12,61).
 * Creates the CUDA context/stream on `device` and the RGBA32F framebuffer.
 * Returns NULL on failure (rtb_last_error(NULL) has the reason). */
rtb_ctx* rtb_create(int width, int height, int device);

/* GLWrapper::~GLWrapper / stop()  (GLWrapper.h:19,29). */
void rtb_destroy(rtb_ctx* ctx);

/* Multi-GPU screen partition (no reference counterpart: the reference is single-GPU).
 * The frame is cut into blocks of `block_rows` scanlines; block b belongs to rank
 * b % world.  This context renders only its own blocks, packed in block order. */
int rtb_set_partition(rtb_ctx* ctx, int rank, int world, int block_rows);
/* Number of scanlines this context owns under the current partition. */
int rtb_local_rows(const rtb_ctx* ctx);

/* GLWrapper::init_shaders(rt_defines&)  (GLWrapper.h:26; GLWrapper.cpp:232-247).
 * Counts, ITERATIONS and the two colours are the shader's specialisation
 * constants; the colours are rounded through "%f" exactly like
 * GLWrapper::to_string (GLWrapper.cpp:279-282). */
int rtb_set_defines(rtb_ctx* ctx, const rtb_defines* defines);

/* GLWrapper::init_buffer / update_buffer  (GLWrapper.h:37-38; GLWrapper.cpp:365-386).
 * `binding` is the uniform-block binding point of SceneManager.cpp:246-254
 * (enum rtb_binding).  bytes may be 0 and data NULL (SceneManager.cpp:246). */
int rtb_upload(rtb_ctx* ctx, int binding, const void* data, size_t bytes);

/* GLWrapper::load_cubemap + set_skybox  (GLWrapper.h:27,35; GLWrapper.cpp:284-317).
 * Six decoded faces in GL order +X,-X,+Y,-Y,+Z,-Z, `channels` = 3 or 4 bytes per
 * texel, rows in file order (row 0 = t 0).  Bilinear, clamp-to-edge, no mips. */
int rtb_set_cubemap(rtb_ctx* ctx, const uint8_t* const faces[6], int w, int h, int channels);

/* GLWrapper::load_texture(unit, name, uniform)  (GLWrapper.h:36; GLWrapper.cpp:319-363)
 * after decoding.  unit 1..3 = texture_sphere_1..3, 4 = texture_ring,
 * 5 = texture_box (main.cpp:149-153).  Builds the mip chain (glGenerateMipmap);
 * sampling is REPEAT + trilinear. */
int rtb_set_texture2d(rtb_ctx* ctx, int unit, const uint8_t* pixels, int w, int h, int channels);

/* Options: "kernel" (enum rtb_kernel), "strict" (1 = no FMA contraction, IEEE
 * div/sqrt: operation-for-operation the shader's arithmetic; 0 = fast build),
 * "cull" (1 = conservative bounding-sphere reject before the torus solve;
 * result-preserving, reported separately from the roofline), "ctas_per_sm" (quad kernel), "coop" (0 = switch the
 * persistent kernel's cooperative drain off: an A/B and test switch, results are identical). */
int rtb_set_option(rtb_ctx* ctx, const char* key, int value);

/* GLWrapper::draw()  (GLWrapper.h:34; GLWrapper.cpp:155-165): render one frame
 * from the current buffers into the context's device framebuffer.  Asynchronous. */
int rtb_render(rtb_ctx* ctx);

/* Same, into a caller-owned DEVICE buffer (local_rows*W*4 floats) on a caller
 * stream (cudaStream_t passed as void*; NULL = the context's stream). */
int rtb_render_to(rtb_ctx* ctx, void* device_rgba32f, void* cuda_stream);

/* Instrumented launch: same image, plus exact work counters (untimed). */
int rtb_render_counted(rtb_ctx* ctx, rtb_stats* out);

/* Block until the context's stream is idle. */
int rtb_sync(rtb_ctx* ctx);

/* glReadPixels equivalents.  Row 0 = BOTTOM scanline (GL window coordinates,
 * rt.frag:315).  dst holds local_rows*W*4 values. */
int rtb_read_rgba32f(rtb_ctx* ctx, float* dst);
int rtb_read_rgba8(rtb_ctx* ctx, uint8_t* dst);     /* clamp to [0,1], *255, round */

/* Device pointer of the context framebuffer (for zero-copy consumers). */
void* rtb_device_framebuffer(rtb_ctx* ctx);

int rtb_get_stats(rtb_ctx* ctx, rtb_stats* out);

/* FFMA-only microbenchmark: measured fp32 CUDA-core peak of `device` in
 * TFLOP/s (the roofline denominator; MEASURED_PEAKS.json has none). */
int rtb_measure_fp32_peak(int device, double* tflops);

const char* rtb_last_error(const rtb_ctx* ctx);
const char* rtb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RTB200_H */

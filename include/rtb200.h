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

/* How the scanline blocks of the ranks reach the root's frame (multi-GPU). */
enum rtb_gather_mode {
    RTB_GATHER_NCCL = 0,       /* one grouped ncclSend/ncclRecv fan-in per frame + one de-interleave pass on the root (default) */
    RTB_GATHER_P2P = 1         /* single-process only: every rank's kernel stores its pixels straight into the root's frame over NVLink */
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

/* ---- multi-GPU (no reference counterpart: the reference is single-GPU) -------------------------------------
 * The frame is cut into blocks of `block_rows` scanlines (a multiple of 4); block b belongs to rank b % world.
 * A rank renders only its own blocks, packed in block order; every rank holds the whole scene.  Two ways to run it:
 *
 *  (1) ONE process, N devices — what a C++ host like the reference's main.cpp uses (host/GLWrapper.cpp: RT_GPUS=N):
 *      rtb_create_multi() returns a context that behaves exactly like a single-device one; every entry point fans
 *      out to the N devices, rtb_render() launches the N kernels and gathers the frame on device 0
 *      (ncclCommInitAll; per frame one ncclGroupStart .. ncclSend/ncclRecv .. ncclGroupEnd and a de-interleave pass;
 *      or, with option "gather" = RTB_GATHER_P2P, no collective at all: the kernels store into device 0's frame).
 *  (2) one process PER device (torchrun, MPI): each process creates a single-device context, rank 0 calls
 *      rtb_comm_unique_id() and ships the 128 bytes to the others by whatever the launcher offers, everyone calls
 *      rtb_comm_init(); per frame rtb_render_to() + rtb_gather().
 * NCCL is loaded at run time (dlopen libnccl.so.2) by the first of these calls; single-GPU use never needs it. */
rtb_ctx* rtb_create_multi(int width, int height, int n_gpus, int block_rows);
int rtb_n_gpus(const rtb_ctx* ctx);                    /* 1 for a single-device context */
/* kernel time of every rank of the last rtb_render() and the device-side frame time on the root (launch of the first
 * kernel to the end of the gather), both by CUDA events.  kernel_ms may be NULL; it holds rtb_n_gpus() values. */
int rtb_rank_times(rtb_ctx* ctx, float* kernel_ms, float* frame_ms);

#define RTB_COMM_ID_BYTES 128
int rtb_comm_unique_id(uint8_t id[RTB_COMM_ID_BYTES]);
/* joins the communicator and sets the partition (rank, world, block_rows) of this context */
int rtb_comm_init(rtb_ctx* ctx, const uint8_t id[RTB_COMM_ID_BYTES], int rank, int world, int block_rows);
/* Frame-end gather of mode (2): every rank passes its packed rows (device pointer; NULL = the context framebuffer), the
 * root (rank 0) also the destination of the assembled frame (H*W*4 floats, device).  Ordered on `cuda_stream`
 * (NULL = the context stream). */
int rtb_gather(rtb_ctx* ctx, const void* local_rows_device, void* full_frame_device, void* cuda_stream);

/* Lower level: the partition alone (a rank rendered on its own, tests). */
int rtb_set_partition(rtb_ctx* ctx, int rank, int world, int block_rows);
/* Number of scanlines this context owns under the current partition. */
int rtb_local_rows(const rtb_ctx* ctx);

/* GLWrapper::init_shaders(rt_defines&)  (GLWrapper.h:26; GLWrapper.cpp:232-247).
 * Counts, ITERATIONS and the two colours are the shader's specialisation
 * constants; the colours are rounded through "%f" exactly like
 * GLWrapper::to_string (GLWrapper.cpp:279-282). */
int rtb_set_defines(rtb_ctx* ctx, const rtb_defines* defines);

/* GLWrapper::init_buffer  (GLWrapper.h:37; GLWrapper.cpp:365-379: glBufferData).
 * `binding` is the uniform-block binding point of SceneManager.cpp:246-254
 * (enum rtb_binding).  Replaces the block: it now holds exactly `bytes` bytes.
 * bytes may be 0 and data NULL (SceneManager.cpp:246); data NULL with bytes > 0
 * allocates only.  The bytes are copied before the call returns (pinned staging
 * inside the library; the device copy itself is asynchronous). */
int rtb_upload(rtb_ctx* ctx, int binding, const void* data, size_t bytes);

/* GLWrapper::update_buffer  (GLWrapper.h:38; GLWrapper.cpp:381-386: glBufferSubData(0, bytes)).
 * Overwrites the first `bytes` bytes of the block and leaves the rest as it was;
 * bytes beyond the block's size are an error (GL_INVALID_VALUE in the reference). */
int rtb_update(rtb_ctx* ctx, int binding, const void* data, size_t bytes);

/* GLWrapper::load_cubemap + set_skybox  (GLWrapper.h:27,35; GLWrapper.cpp:284-317).
 * Six decoded faces in GL order +X,-X,+Y,-Y,+Z,-Z, `channels` = 3 or 4 bytes per
 * texel, rows in file order (row 0 = t 0).  Bilinear, clamp-to-edge, no mips. */
int rtb_set_cubemap(rtb_ctx* ctx, const uint8_t* const faces[6], int w, int h, int channels);

/* GLWrapper::load_texture(unit, name, uniform)  (GLWrapper.h:36; GLWrapper.cpp:319-363)
 * after decoding.  unit 1..3 = texture_sphere_1..3, 4 = texture_ring,
 * 5 = texture_box (main.cpp:149-153).  Builds the mip chain (glGenerateMipmap);
 * sampling is REPEAT + trilinear. */
int rtb_set_texture2d(rtb_ctx* ctx, int unit, const uint8_t* pixels, int w, int h, int channels);

/* Options: "kernel" (enum rtb_kernel); "strict" (1, the default = no FMA contraction, IEEE div/sqrt: operation for
 * operation the shader's arithmetic, bit-comparable with the CPU oracle; 0 = the FUSED build: FMA contraction, MUFU
 * reciprocals, rotation matrices — 1.6x faster, parity by the envelope criterion);
 * "cull" (1 = conservative bounding-sphere reject before the torus solve; result-preserving, reported separately from
 * the roofline); "ctas_per_sm" (quad kernel); "coop" (0 = switch the persistent kernel's cooperative drain off: an A/B
 * and test switch, results are identical); "gather" (enum rtb_gather_mode, multi-GPU contexts); "smaa_compact" (0 = run
 * SMAA's blending-weight pass over every pixel as the reference draws it instead of over the compacted edge pixels: an
 * A/B and test switch, results are identical); "lpt" (-1 = automatic, the default: on for frames of at least 4096 tiles on the
 * persistent kernel; 0 / 1 = off / on): from the second frame on, the cheapest tiles of the previous frame are handed out last so
 * that the frame ends on its shortest paths; the order never changes a pixel; "wide" (-1 = automatic, the default: scenes without
 * tori run the fused build's persistent kernel with 24 warps of 80 registers instead of 20 warps of 96; 0 / 1 = off / on; same bits). */
int rtb_set_option(rtb_ctx* ctx, const char* key, int value);

/* GLWrapper::draw()  (GLWrapper.h:34; GLWrapper.cpp:155-165): render one frame
 * from the current buffers into the context's device framebuffer.  Asynchronous. */
int rtb_render(rtb_ctx* ctx);

/* ---- SMAA post-pass -------------------------------------------------------------------------
 * GLWrapper::enable_SMAA(preset)  (GLWrapper.h:30; GLWrapper.cpp:149-153): preset 0..3 = LOW, MEDIUM, HIGH, ULTRA
 * (SMAA_Builder.h:9-12; main.cpp:32 uses ULTRA), -1 = off (the default of this library).  When on, rtb_render() follows
 * the ray-trace pass with the reference's three passes (GLWrapper.cpp:173-204: luma edge detection, blending weights,
 * neighbourhood blending) on the RGBA8-quantised frame, and rtb_read_rgba8() returns the post-processed image
 * (rtb_read_rgba32f() keeps returning the ray-traced floats).  Not applied by rtb_render_to(). */
int rtb_enable_smaa(rtb_ctx* ctx, int preset);
/* SMAA_Builder::load_area_texture / load_search_texture  (SMAA_Builder.h:45-79): the precomputed lookup tables of
 * src/AreaTex.h (RG8, 160 x 560) and src/SearchTex.h (R8, 64 x 16), rows in file order.  The library ships no copy of them. */
int rtb_smaa_set_tables(rtb_ctx* ctx, const uint8_t* area_rg8, const uint8_t* search_r8);
/* The three passes alone on a host RGBA8 image of the context's size (tests, measurements).  Any output may be NULL:
 * out RGBA8, edges RG8, blend RGBA8; *ms = CUDA-event time of the three kernels. */
int rtb_smaa_apply(rtb_ctx* ctx, const uint8_t* rgba8, uint8_t* out_rgba8, uint8_t* edges_rg8, uint8_t* blend_rgba8, float* ms);
/* CUDA-event time of the SMAA passes of the last frame / rtb_smaa_apply. */
int rtb_smaa_last_ms(rtb_ctx* ctx, float* ms);

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

/* Inspection of the cost-ordered tile hand-out (option "lpt"): the per-tile costs of the last frame (summed path lengths of the
 * tile's 32 pixels) and the hand-out order derived from them for the next frame, `capacity` entries each at most; returns the
 * number of tiles (0: the option is off or no frame has been rendered), negative on error.  Either pointer may be NULL. */
int rtb_tile_order(rtb_ctx* ctx, uint32_t* cost, uint32_t* order, int capacity);

/* FFMA-only microbenchmark: measured fp32 CUDA-core peak of `device` in
 * TFLOP/s (the roofline denominator; MEASURED_PEAKS.json has none). */
int rtb_measure_fp32_peak(int device, double* tflops);
/* The same chains with three distinct register operands per FFMA: what a sub-partition's register file can feed when every
 * FMA combines three live values (reported beside the peak; never the roofline denominator). */
int rtb_measure_fp32_peak3(int device, double* tflops);

const char* rtb_last_error(const rtb_ctx* ctx);
const char* rtb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RTB200_H */

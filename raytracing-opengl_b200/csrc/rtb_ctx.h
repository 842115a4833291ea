/* rtb_ctx.h — the context behind the C-ABI of include/rtb200.h (shared by rtb_api.cu and rtb_multi.cu; not installed). */
#ifndef RTB_CTX_H
#define RTB_CTX_H

#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/rtb200.h"
#include "rt_launch.h"

struct RtbTex2D { uint8_t* dev = nullptr; int w = 0, h = 0, levels = 0; uint32_t off[16] = { 0 }; };

/* Host staging of the uniform-block uploads: a ring of pinned arenas, one per frame in flight.  rtb_upload copies the
 * caller's bytes into the current arena (so the caller may reuse its buffer at once — glBufferData semantics) and queues an
 * asynchronous H2D copy on the context stream; rendering a frame closes the arena with an event and moves on to the next one,
 * which is reused only after its own event has completed (two frames later: never waits in a steady frame loop). */
struct RtbStageSlot { uint8_t* host = nullptr; size_t cap = 0, used = 0; cudaEvent_t done = nullptr; bool pending = false; };
enum { RTB_STAGE_SLOTS = 3 };

struct rtb_ctx {
    int device = 0, n_sm = 0;
    int width = 0, height = 0;
    int rank = 0, world = 1, block_rows = 16, local_rows = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_order = nullptr;
    bool have_defines = false;
    rtb_defines defines = {};
    void* raw[RTB_NUM_BINDINGS] = { nullptr };
    size_t raw_cap[RTB_NUM_BINDINGS] = { 0 }, raw_bytes[RTB_NUM_BINDINGS] = { 0 };
    rtb_scene scene_host = {};                   /* host copy of the scene uniform (camera, canvas): kernel arguments */
    bool have_scene = false;
    bool uses_tex[RTB_NUM_BINDINGS] = { false }; /* does the uploaded array reference a 2-D texture?  decided once per upload */
    RtbStageSlot stage[RTB_STAGE_SLOTS];
    int stage_cur = 0;
    uint8_t* packed = nullptr; size_t packed_cap = 0;
    bool dirty = true;
    unsigned int* tile_counter = nullptr;
    unsigned int *tile_cost = nullptr, *tile_perm = nullptr, *tile_hist = nullptr;   /* persistent kernel: per-tile path lengths of the last frame and the order derived from them */
    int lpt_tiles = 0, opt_lpt = -1, opt_wide = -1;
    bool lpt_valid = false;
    unsigned long long* counters = nullptr;
    unsigned long long* cta_times = nullptr;     /* RTB_DEBUG_TIMES=1: per-CTA start / drain / end stamps, printed by rtb_sync */
    int cta_times_n = 0, cta_times_cap = 0;
    float* fb = nullptr; size_t fb_floats = 0;   /* this rank's scanlines, packed in block order */
    uint8_t* fb8 = nullptr;                      /* RGBA8 copy of the frame, made on demand by rtb_read_rgba8 */
    uint8_t* cube = nullptr; int cube_w = 0, cube_h = 0;
    RtbTex2D tex[6];
    int opt_kernel = RTB_KERNEL_AUTO, opt_strict = 1, opt_cull = 0, opt_ctas_per_sm = 0, opt_coop = 1, opt_gather = RTB_GATHER_NCCL;
    rtb_stats stats = {};
    bool timed_pending = false;
    std::string err;

    /* ---- multi-GPU (rtb_multi.cu) ---- */
    std::vector<rtb_ctx*> peers;                 /* single-process multi-device root: the contexts of ranks 1..n-1 */
    rtb_ctx* root = nullptr;                     /* set in those peers */
    void* comm = nullptr;                        /* ncclComm_t of this rank (either mode) */
    float* gather_scratch = nullptr; size_t gather_scratch_floats = 0;   /* root: the packed rows of ranks 1..n-1 */
    float* fb_full = nullptr;                    /* single-process root: the assembled frame */
    bool p2p_ok = true;                          /* every peer can store into device 0's memory (RTB_GATHER_P2P) */
    cudaEvent_t ev_f0 = nullptr, ev_f1 = nullptr;
    bool frame_timed = false;
    float frame_ms = 0.f;

    /* ---- SMAA post-pass (smaa.cu) ---- */
    int smaa_preset = -1;                        /* -1 = off (the default here; main.cpp:32 asks for ULTRA = 3) */
    uint8_t *smaa_color = nullptr, *smaa_edges = nullptr, *smaa_blend = nullptr, *smaa_out = nullptr;   /* RGBA8 / RG8 / RGBA8 / RGBA8 targets */
    uint8_t *smaa_area = nullptr, *smaa_search = nullptr;                                             /* lookup tables */
    float* smaa_uv = nullptr;                    /* texel-centre coordinates per column / row */
    unsigned *smaa_list = nullptr, *smaa_count = nullptr;                                             /* pass 2: compacted edge pixels */
    int opt_smaa_compact = 1;                    /* 0: pass 2 as one full-screen kernel (A/B partner) */
    cudaEvent_t ev_s0 = nullptr, ev_s1 = nullptr;
    bool smaa_timed = false, smaa_valid = false; /* smaa_valid: smaa_out holds the post-processed version of the last frame */
};

int rtb_fail(rtb_ctx* c, int code, const char* fmt, ...);
#define CU(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) return rtb_fail(ctx, RTB_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

/* rtb_api.cu */
int rtb_do_render(rtb_ctx* ctx, float* target, bool target_global_rows, cudaStream_t st, bool counted, bool timed);
int rtb_compute_local_rows(int height, int rank, int world, int block_rows);
/* smaa.cu */
int rtb_smaa_after_frame(rtb_ctx* ctx, const float* frame, cudaStream_t st);
void rtb_smaa_release(rtb_ctx* ctx);
/* rtb_multi.cu */
int rtb_multi_render(rtb_ctx* root);
void rtb_multi_release(rtb_ctx* ctx);

#endif

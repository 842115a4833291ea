/* rtb_api.cu — the C-ABI of include/rtb200.h: context, uploads, sampler setup, launches.
 *
 * Each entry point replaces one GLWrapper method of the reference (cited in
 * include/rtb200.h).  There is no CPU rendering path in this library: every
 * frame is produced by the sm_100a kernels in rt_kernels.cu.
 */
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>

#include "rtb_ctx.h"

namespace {
thread_local std::string g_last_error;
}

int rtb_fail(rtb_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    if (c) c->err = buf;
    return code;
}
#define fail rtb_fail

namespace {

/* GL unorm8 conversion of the colour buffer (the reference renders into RGBA8, GLWrapper.cpp:127,216): clamp to [0,1],
 * scale by 255, round to nearest; NaN -> 0.  One float4 in, one uchar4 out per thread: a pure HBM stream. */
__global__ void rgba8_kernel(const float4* __restrict__ src, uchar4* __restrict__ dst, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 v = src[i];
    auto q = [](float x) -> unsigned char {
        x = x < 0.f ? 0.f : (x > 1.f ? 1.f : x);
        if (!(x == x)) x = 0.f;
        return (unsigned char)__fadd_rn(__fmul_rn(x, 255.0f), 0.5f);      /* no contraction: same value as the host formula */
    };
    dst[i] = make_uchar4(q(v.x), q(v.y), q(v.z), q(v.w));
}

const size_t ELEM_SIZE[RTB_NUM_BINDINGS] = { sizeof(rtb_scene), sizeof(rtb_sphere), sizeof(rtb_plane), sizeof(rtb_surface), sizeof(rtb_box),
                                             sizeof(rtb_torus), sizeof(rtb_ring), sizeof(rtb_light_point), sizeof(rtb_light_direct) };

uint32_t align16(uint32_t v) { return (v + 15u) & ~15u; }

}  // namespace
int rtb_compute_local_rows(int height, int rank, int world, int block_rows) {
    int rows = 0;
    for (int b = rank; b * block_rows < height; b += world) {
        int r = height - b * block_rows;
        rows += r < block_rows ? r : block_rows;
    }
    return rows;
}
namespace {

/* GLWrapper::to_string (GLWrapper.cpp:279-282): the colour reaches the shader as the text "%f" */
float round_through_percent_f(float v) {
    char buf[64];
    snprintf(buf, sizeof buf, "%f", v);
    return strtof(buf, nullptr);
}

/* glGenerateMipmap stand-in: 2x2 box filter, floor(d/2) sizes, 8-bit round-half-up (DESIGN.md "samplers") */
void build_mip_chain(const uint8_t* src, int w, int h, int ch, std::vector<uint8_t>& out, RtbTex2D& t) {
    std::vector<uint8_t> cur((size_t)w * h * 4);
    for (size_t i = 0; i < (size_t)w * h; i++) {
        uint8_t r = src[i * ch], g = 0, b = 0, a = 255;
        if (ch >= 3) { g = src[i * ch + 1]; b = src[i * ch + 2]; }
        if (ch == 2) g = src[i * ch + 1];
        if (ch == 4) a = src[i * ch + 3];
        cur[i * 4] = r; cur[i * 4 + 1] = g; cur[i * 4 + 2] = b; cur[i * 4 + 3] = a;
    }
    out.clear();
    t.w = w; t.h = h; t.levels = 0;
    int cw = w, chh = h;
    for (;;) {
        t.off[t.levels++] = (uint32_t)out.size();
        out.insert(out.end(), cur.begin(), cur.end());
        if ((cw == 1 && chh == 1) || t.levels >= 16) break;
        int nw = cw > 1 ? cw / 2 : 1, nh = chh > 1 ? chh / 2 : 1;
        std::vector<uint8_t> nxt((size_t)nw * nh * 4);
        for (int y = 0; y < nh; y++)
            for (int x = 0; x < nw; x++) {
                int x0 = cw > 1 ? 2 * x : 0, x1 = cw > 1 ? 2 * x + 1 : 0;
                int y0 = chh > 1 ? 2 * y : 0, y1 = chh > 1 ? 2 * y + 1 : 0;
                for (int c = 0; c < 4; c++) {
                    int s = cur[((size_t)y0 * cw + x0) * 4 + c] + cur[((size_t)y0 * cw + x1) * 4 + c] +
                            cur[((size_t)y1 * cw + x0) * 4 + c] + cur[((size_t)y1 * cw + x1) * 4 + c];
                    nxt[((size_t)y * nw + x) * 4 + c] = (uint8_t)((s + 2) >> 2);
                }
            }
        cur.swap(nxt);
        cw = nw; chh = nh;
    }
}

/* does an uploaded array reference a 2-D texture (textureNum != 0)?  Decided once per upload, on the caller's bytes */
bool array_uses_textures(int binding, const void* data, size_t bytes) {
    const size_t n = bytes / ELEM_SIZE[binding];
    if (binding == RTB_BIND_SPHERES) { for (size_t i = 0; i < n; i++) if (((const rtb_sphere*)data)[i].textureNum != 0) return true; }
    else if (binding == RTB_BIND_BOXES) { for (size_t i = 0; i < n; i++) if (((const rtb_box*)data)[i].textureNum != 0) return true; }
    else if (binding == RTB_BIND_RINGS) { for (size_t i = 0; i < n; i++) if (((const rtb_ring*)data)[i].textureNum != 0) return true; }
    return false;
}
bool scene_uses_textures(const rtb_ctx* c) { return c->uses_tex[RTB_BIND_SPHERES] || c->uses_tex[RTB_BIND_BOXES] || c->uses_tex[RTB_BIND_RINGS]; }

/* the current staging arena with room for `bytes` more (see RtbStageSlot in rtb_ctx.h) */
int stage_reserve(rtb_ctx* ctx, size_t bytes, uint8_t** out) {
    RtbStageSlot& s = ctx->stage[ctx->stage_cur];
    if (s.pending) {                                  /* first use since this arena's frame was queued: its copies must be done (normally long ago) */
        CU(cudaEventSynchronize(s.done));
        s.pending = false;
        s.used = 0;
    }
    const size_t need = s.used + ((bytes + 255) & ~(size_t)255);
    if (need > s.cap) {
        /* grow: copies already queued from the old arena must finish first (only while the scene is first being built) */
        if (s.used) CU(cudaStreamSynchronize(ctx->stream));
        if (s.host) cudaFreeHost(s.host);
        s.host = nullptr; s.used = 0;
        s.cap = std::max<size_t>(need, 256 * 1024) * 2;
        CU(cudaMallocHost(&s.host, s.cap));
    }
    *out = s.host + s.used;
    s.used += (bytes + 255) & ~(size_t)255;
    return RTB_OK;
}
/* a frame has been queued behind the uploads: close the arena, continue in the next one */
int stage_rotate(rtb_ctx* ctx) {
    RtbStageSlot& s = ctx->stage[ctx->stage_cur];
    if (s.used == 0) return RTB_OK;
    CU(cudaEventRecord(s.done, ctx->stream));
    s.pending = true;
    ctx->stage_cur = (ctx->stage_cur + 1) % RTB_STAGE_SLOTS;
    return RTB_OK;
}

/* the strict build stages quaternion records, the fused build matrix records (rt_params.h) */
PackedLayout make_layout(const rtb_defines& d, bool strict) {
    PackedLayout L;
    uint32_t o = 0;
    L.off_plane = o;  o += align16(d.plane_size * sizeof(PPlane));
    L.off_sphere = o; o += align16(d.sphere_size * sizeof(PSphere));
    L.off_surf = o;   o += align16(d.surface_size * (strict ? sizeof(PSurf) : sizeof(PSurfM)));
    L.off_box = o;    o += align16(d.box_size * (strict ? sizeof(PBox) : sizeof(PBoxM)));
    L.off_torus = o;  o += align16(d.torus_size * (strict ? sizeof(PTorus) : sizeof(PTorusM)));
    L.off_ring = o;   o += align16(d.ring_size * (strict ? sizeof(PRing) : sizeof(PRingM)));
    L.off_light = o;  o += align16(d.light_point_size * sizeof(PLight));
    L.total_bytes = o < 16 ? 16 : o;
    return L;
}

/* algorithmic flop constants of the brute-force algorithm */
double algorithmic_flops(const rtb_stats& s) {
    double f = 0.0;
    f += 19.0 * (double)s.tests[RTB_TYPE_SPHERE] + 16.0 * (double)s.tests[RTB_TYPE_PLANE] + 168.0 * (double)s.tests[RTB_TYPE_SURFACE] +
         134.0 * (double)s.tests[RTB_TYPE_BOX] + 123.0 * (double)s.tests[RTB_TYPE_RING] + 19.0 * (double)s.tests[RTB_TYPE_POINT_LIGHT];
    f += 130.0 * (double)s.tests[RTB_TYPE_TORUS] + 271.0 * (double)s.dk_iterations;
    f += 60.0 * (double)s.light_evals;
    f += 15.0 * (double)s.shaded_hits[RTB_TYPE_SPHERE] + 74.0 * (double)s.shaded_hits[RTB_TYPE_BOX] + 206.0 * (double)s.shaded_hits[RTB_TYPE_SURFACE] +
         200.0 * (double)s.shaded_hits[RTB_TYPE_TORUS] + 68.0 * (double)s.shaded_hits[RTB_TYPE_RING];
    f += 71.0 * (double)s.pixels;
    return f;
}

/* Hand-out order of the next frame's tiles.  The frame should END on its cheapest paths: when the counter runs out every lane
 * still holds a path, and whatever those paths have left is served by the drain, CTA by CTA, with nobody to share it with —
 * the CTA with the longest leftovers ends the frame (measured: 3 200 drain jobs against a median of 1 900).  So the cheapest
 * tiles of the previous frame (cost = summed path lengths of the tile's 32 pixels; as many as `tail`, taken bucket by bucket
 * from the cheap end) go LAST, costliest bucket first, and all other tiles keep the scan order.  Measured (tools/lpt_ab.py,
 * profiles/r2_lpt_ab.jsonl): whole 4K frames 0.3-1.9 % faster, one share of an 8-way split 3.5-4.2 % (the drain is 10 % of
 * such a frame).  Sorting the WHOLE frame by cost gains the same on a share but loses 1-2.4 % on whole frames: neighbouring
 * lanes then hold unrelated pixels.  Three small kernels (histogram per 4096-tile chunk, plan, scatter). */
__device__ __forceinline__ unsigned cost_bucket(unsigned c) { c >>= 2; return c > 255u ? 255u : c; }
constexpr int ORDER_CHUNK = 4096;                                  /* tiles per CTA: four consecutive ones per thread */
struct OrderPlan { unsigned split, quota, tail_start[256]; };      /* written by tile_plan_kernel, read by tile_scatter_kernel */

/* one warp-wide add per distinct value instead of one shared-memory atomic per lane (half the tiles of a frame can share one
 * bucket — sky); returns this lane's rank among the lanes with its key.  Every lane of the warp must call it. */
__device__ __forceinline__ unsigned warp_add_by_key(unsigned* counters, unsigned key, bool valid, unsigned lane) {
    const unsigned same = __match_any_sync(0xffffffffu, valid ? key : 0xffffffffu);
    unsigned base = 0;
    const int leader = __ffs(same) - 1;
    if (valid && (int)lane == leader) base = atomicAdd(&counters[key], (unsigned)__popc(same));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + (unsigned)__popc(same & ((1u << lane) - 1u));
}
/* 1: the cost histogram of every chunk */
__global__ void __launch_bounds__(1024) tile_hist_kernel(const unsigned* __restrict__ cost, unsigned* __restrict__ chunk_hist, int n) {
    __shared__ unsigned hist[256];
    const int t = threadIdx.x, i0 = blockIdx.x * ORDER_CHUNK + 4 * t;
    if (t < 256) hist[t] = 0;
    __syncthreads();
    const uint4 c = *(const uint4*)(cost + i0);                     /* the array is allocated and zeroed up to a whole chunk */
    const unsigned cs[4] = { c.x, c.y, c.z, c.w };
#pragma unroll
    for (int j = 0; j < 4; j++) warp_add_by_key(hist, cost_bucket(cs[j]), i0 + j < n, t & 31);
    __syncthreads();
    if (t < 256) chunk_hist[blockIdx.x * 256 + t] = hist[t];
}
/* 2: which buckets form the tail, and for every chunk and bucket the number of such tiles in the chunks before it (in place) */
__global__ void __launch_bounds__(256) tile_plan_kernel(unsigned* __restrict__ chunk_hist, OrderPlan* __restrict__ plan, int n, int n_chunks, int tail) {
    __shared__ unsigned total[256];
    const int k = threadIdx.x;
    unsigned sum = 0;
    for (int c = 0; c < n_chunks; c++) { const unsigned v = chunk_hist[c * 256 + k]; chunk_hist[c * 256 + k] = sum; sum += v; }
    total[k] = sum;
    __syncthreads();
    if (k == 0) {
        const unsigned want = min((unsigned)tail, (unsigned)n / 2u);
        unsigned cum = 0, b = 0;
        while (b < 255u && cum + total[b] <= want) cum += total[b++];            /* buckets 0..b-1 go last entirely ... */
        const unsigned q = min(total[b], want - cum);                            /* ... and the first q tiles of bucket b (a frame that is half sky has one huge cheapest bucket) */
        plan->split = b; plan->quota = q;
        unsigned o = (unsigned)n - (cum + q);                                    /* the tail: bucket b first, bucket 0 last */
        plan->tail_start[b] = o; o += q;
        for (int j = (int)b - 1; j >= 0; j--) { plan->tail_start[j] = o; o += total[j]; }
    }
}
/* 3: every chunk places its tiles: heads in scan order (a stable partition), tails by bucket */
__global__ void __launch_bounds__(1024) tile_scatter_kernel(const unsigned* __restrict__ cost, const unsigned* __restrict__ chunk_prefix,
                                                            const OrderPlan* __restrict__ plan, unsigned* __restrict__ perm, int n) {
    __shared__ unsigned cursor[256], warp_cnt[32], red[8];
    __shared__ unsigned taken;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5, i0 = blockIdx.x * ORDER_CHUNK + 4 * t;
    const unsigned sp = plan->split, quota = plan->quota;
    unsigned before_tail = 0;                                       /* tail tiles in the chunks before this one */
    if (t < 256) {
        const unsigned pre = chunk_prefix[blockIdx.x * 256 + t];
        cursor[t] = (t <= (int)sp ? plan->tail_start[t] : 0u) + ((unsigned)t == sp ? min(pre, quota) : pre);
        before_tail = (unsigned)t < sp ? pre : ((unsigned)t == sp ? min(pre, quota) : 0u);
#pragma unroll
        for (int d = 16; d; d >>= 1) before_tail += __shfl_xor_sync(0xffffffffu, before_tail, d);
        if (lane == 0) red[wid] = before_tail;
        if ((unsigned)t == sp) taken = min(pre, quota);
    }
    __syncthreads();
    before_tail = red[0] + red[1] + red[2] + red[3] + red[4] + red[5] + red[6] + red[7];
    const uint4 c = *(const uint4*)(cost + i0);
    const unsigned cs[4] = { c.x, c.y, c.z, c.w };
    bool is_head[4];
    unsigned mine = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const bool valid = i0 + j < n;
        const unsigned bk = cost_bucket(cs[j]);
        bool is_tail = valid && bk < sp;
        if (valid && bk == sp && *(volatile unsigned*)&taken < quota) is_tail = atomicAdd(&taken, 1u) < quota;   /* the quota runs out in ONE chunk */
        is_head[j] = valid && !is_tail;
        mine += is_head[j] ? 1u : 0u;
        const unsigned pos = warp_add_by_key(cursor, bk, is_tail, lane);
        if (is_tail) perm[pos] = (unsigned)(i0 + j);
    }
    unsigned incl = mine;                                           /* inclusive scan of the head counts over the warp */
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
    if (lane == 31) warp_cnt[wid] = incl;
    __syncthreads();
    unsigned before = 0;
    for (int w = 0; w < wid; w++) before += warp_cnt[w];
    unsigned pos = (unsigned)(blockIdx.x * ORDER_CHUNK) - before_tail + before + incl - mine;
#pragma unroll
    for (int j = 0; j < 4; j++) if (is_head[j]) perm[pos++] = (unsigned)(i0 + j);      /* stable: the scan order */
}

}  // namespace

int rtb_do_render(rtb_ctx* ctx, float* target, bool target_global_rows, cudaStream_t st, bool counted, bool timed) {
    if (!ctx) return fail(nullptr, RTB_ERR_INVALID, "null context");
    if (!ctx->have_defines) return fail(ctx, RTB_ERR_STATE, "rtb_render before rtb_set_defines (init_shaders)");
    const rtb_defines& d = ctx->defines;
    if (!ctx->have_scene) return fail(ctx, RTB_ERR_STATE, "scene_buf was never uploaded");
    const int counts[RTB_NUM_BINDINGS] = { 1, d.sphere_size, d.plane_size, d.surface_size, d.box_size, d.torus_size, d.ring_size,
                                           d.light_point_size, d.light_direct_size };
    for (int b = 0; b < RTB_NUM_BINDINGS; b++)
        if (ctx->raw_bytes[b] < (size_t)counts[b] * ELEM_SIZE[b])
            return fail(ctx, RTB_ERR_STATE, "binding %d holds %zu bytes, the defines need %zu", b, ctx->raw_bytes[b], (size_t)counts[b] * ELEM_SIZE[b]);
    CU(cudaSetDevice(ctx->device));

    FrameParams P;
    memset(&P, 0, sizeof P);
    P.n_sphere = d.sphere_size; P.n_plane = d.plane_size; P.n_surf = d.surface_size; P.n_box = d.box_size; P.n_torus = d.torus_size;
    P.n_ring = d.ring_size; P.n_lpoint = d.light_point_size; P.n_ldirect = d.light_direct_size; P.iterations = d.iterations;
    for (int i = 0; i < 3; i++) { P.ambient[i] = round_through_percent_f(d.ambient_color[i]); P.shadow_ambient[i] = round_through_percent_f(d.shadow_ambient[i]); }
    const rtb_scene* sc = &ctx->scene_host;
    memcpy(P.cam_q, sc->quat_camera_rotation, 16);
    memcpy(P.cam_pos, sc->camera_pos, 12);
    P.canvas_w = sc->canvas_width; P.canvas_h = sc->canvas_height;
    if (P.canvas_w != ctx->width || P.canvas_h != ctx->height)
        return fail(ctx, RTB_ERR_STATE, "scene canvas %dx%d differs from the context's %dx%d", P.canvas_w, P.canvas_h, ctx->width, ctx->height);
    P.spheres = (const rtb_sphere*)ctx->raw[RTB_BIND_SPHERES]; P.planes = (const rtb_plane*)ctx->raw[RTB_BIND_PLANES];
    P.surfaces = (const rtb_surface*)ctx->raw[RTB_BIND_SURFACES]; P.boxes = (const rtb_box*)ctx->raw[RTB_BIND_BOXES];
    P.toruses = (const rtb_torus*)ctx->raw[RTB_BIND_TORUSES]; P.rings = (const rtb_ring*)ctx->raw[RTB_BIND_RINGS];
    P.lights_point = (const rtb_light_point*)ctx->raw[RTB_BIND_LIGHTS_POINT]; P.lights_direct = (const rtb_light_direct*)ctx->raw[RTB_BIND_LIGHTS_DIRECT];
    const bool strict = ctx->opt_strict != 0;
    P.lay = make_layout(d, strict);
    if (P.lay.total_bytes > ctx->packed_cap) {
        if (ctx->packed) CU(cudaFreeAsync(ctx->packed, ctx->stream));          /* stream-ordered: frames in flight keep their block */
        ctx->packed_cap = P.lay.total_bytes + 4096;
        CU(cudaMallocAsync((void**)&ctx->packed, ctx->packed_cap, ctx->stream));
        ctx->dirty = true;
    }
    P.packed = ctx->packed;
    P.cube.base = ctx->cube; P.cube.w = ctx->cube_w; P.cube.h = ctx->cube_h;
    for (int u = 1; u <= 5; u++) {
        P.tex[u].base = ctx->tex[u].dev; P.tex[u].w = ctx->tex[u].w; P.tex[u].h = ctx->tex[u].h; P.tex[u].levels = ctx->tex[u].levels;
        memcpy(P.tex[u].level_off, ctx->tex[u].off, sizeof ctx->tex[u].off);
    }
    P.fb = target; P.fb_global = target_global_rows ? 1 : 0; P.rank = ctx->rank; P.world = ctx->world; P.block_rows = ctx->block_rows; P.local_rows = ctx->local_rows;
    P.tile_counter = ctx->tile_counter;
    P.n_tiles_x = (ctx->width + 7) / 8; P.n_tiles_y = (ctx->local_rows + 3) / 4;
    P.tile_cost = nullptr; P.tile_perm = nullptr;
    P.cull = ctx->opt_cull;
    P.coop = ctx->opt_coop;
    P.k_one = 1.0f; P.k_neg_zero = -0.0f; P.k_neg_one = -1.0f;
    P.counters = counted ? ctx->counters : nullptr;
    P.cta_times = nullptr;

    int kernel = ctx->opt_kernel;
    const bool textured = scene_uses_textures(ctx);
    if (kernel == RTB_KERNEL_AUTO) kernel = textured ? RTB_KERNEL_QUAD : RTB_KERNEL_PERSISTENT;
    if (kernel == RTB_KERNEL_PERSISTENT && textured)
        return fail(ctx, RTB_ERR_STATE, "the persistent kernel cannot render scenes that reference 2-D textures (textureNum != 0)");
    /* scenes without tori need no room for the Durand-Kerner state: the 24-warp variant (option "wide": -1 = automatic: fused build
     * only — spheres4k 17.85 -> 16.75 ms; the strict build spills at 80 registers and loses 4 %, profiles/r2_wide_ab.jsonl) */
    const bool wide = kernel == RTB_KERNEL_PERSISTENT && (ctx->opt_wide > 0 || (ctx->opt_wide < 0 && d.torus_size == 0 && !strict));
    const int launch_kernel = kernel == RTB_KERNEL_QUAD ? RTB_LAUNCH_QUAD : (wide ? RTB_LAUNCH_PERSISTENT_WIDE : RTB_LAUNCH_PERSISTENT);
    const int threads = kernel == RTB_KERNEL_QUAD ? QUAD_THREADS : (wide ? PERSIST_THREADS_WIDE : PERSIST_THREADS);
    const size_t smem = kernel == RTB_KERNEL_QUAD ? (size_t)P.lay.total_bytes : PERSIST_SMEM_BYTES_T(P.lay.total_bytes, threads);
    if (smem > 227 * 1024) return fail(ctx, RTB_ERR_INVALID, "packed scene (%zu bytes) exceeds the 227 KB shared-memory budget", smem);

    int per_sm = 0;
    int e = strict ? rtb_strict_occupancy(launch_kernel, smem, &per_sm) : rtb_fast_occupancy(launch_kernel, smem, &per_sm);
    if (e || per_sm < 1) return fail(ctx, RTB_ERR_CUDA, "occupancy query failed (%s), smem %zu", cudaGetErrorString((cudaError_t)e), smem);
    if (ctx->opt_ctas_per_sm > 0 && ctx->opt_ctas_per_sm < per_sm) per_sm = ctx->opt_ctas_per_sm;
    const int grid = ctx->n_sm * per_sm;                 /* persistent: a whole number of CTAs on every one of the SMs */

    if (getenv("RTB_DEBUG_TIMES") && kernel == RTB_KERNEL_PERSISTENT) {
        if (grid > ctx->cta_times_cap) {                  /* the grid depends on the build, the scene's shared-memory size and "ctas_per_sm" */
            if (ctx->cta_times) { CU(cudaStreamSynchronize(st)); cudaFree(ctx->cta_times); ctx->cta_times = nullptr; }
            CU(cudaMalloc(&ctx->cta_times, (size_t)grid * 5 * sizeof(unsigned long long)));
            ctx->cta_times_cap = grid;
        }
        CU(cudaMemsetAsync(ctx->cta_times, 0, (size_t)grid * 5 * sizeof(unsigned long long), st));
        ctx->cta_times_n = grid;
        P.cta_times = ctx->cta_times;
    }
    if (st != ctx->stream) {                             /* uploads (and allocations) ran on the context stream: order them before this frame */
        CU(cudaEventRecord(ctx->ev_order, ctx->stream));
        CU(cudaStreamWaitEvent(st, ctx->ev_order, 0));
    }
    if (ctx->dirty) {
        int pe = strict ? rtb_strict_launch_pack(&P, ctx->packed, st) : rtb_fast_launch_pack(&P, ctx->packed, st);
        if (pe) return fail(ctx, RTB_ERR_CUDA, "pack kernel launch: %s", cudaGetErrorString((cudaError_t)pe));
        ctx->dirty = false;
    }
    /* cost-ordered tiles (option "lpt": -1 = automatic, frames of at least 4096 tiles on the persistent kernel) */
    const int n_tiles = P.n_tiles_x * P.n_tiles_y;
    const bool lpt = kernel == RTB_KERNEL_PERSISTENT && n_tiles > 0 && (ctx->opt_lpt > 0 || (ctx->opt_lpt < 0 && n_tiles >= 4096));
    if (lpt) {
        if (ctx->lpt_tiles != n_tiles) {                   /* first frame, or the partition changed: no order yet */
            if (ctx->tile_cost) { CU(cudaFreeAsync(ctx->tile_cost, ctx->stream)); CU(cudaFreeAsync(ctx->tile_perm, ctx->stream)); CU(cudaFreeAsync(ctx->tile_hist, ctx->stream)); }
            const size_t chunks = ((size_t)n_tiles + ORDER_CHUNK - 1) / ORDER_CHUNK;                  /* costs are read a whole chunk at a time */
            CU(cudaMallocAsync((void**)&ctx->tile_cost, chunks * ORDER_CHUNK * sizeof(unsigned), ctx->stream));
            CU(cudaMallocAsync((void**)&ctx->tile_perm, chunks * ORDER_CHUNK * sizeof(unsigned), ctx->stream));
            CU(cudaMallocAsync((void**)&ctx->tile_hist, chunks * 256 * sizeof(unsigned) + sizeof(OrderPlan), ctx->stream));
            ctx->lpt_tiles = n_tiles; ctx->lpt_valid = false;
            if (st != ctx->stream) { CU(cudaEventRecord(ctx->ev_order, ctx->stream)); CU(cudaStreamWaitEvent(st, ctx->ev_order, 0)); }
        }
        CU(cudaMemsetAsync(ctx->tile_cost, 0, (size_t)((n_tiles + ORDER_CHUNK - 1) / ORDER_CHUNK) * ORDER_CHUNK * sizeof(unsigned), st));
        P.tile_cost = ctx->tile_cost;
        P.tile_perm = ctx->lpt_valid ? ctx->tile_perm : nullptr;
    }
    CU(cudaMemsetAsync(ctx->tile_counter, 0, sizeof(unsigned int), st));
    if (counted) CU(cudaMemsetAsync(ctx->counters, 0, CNT_NUM * sizeof(unsigned long long), st));
    if (timed) CU(cudaEventRecord(ctx->ev0, st));
    e = strict ? rtb_strict_launch(&P, launch_kernel, counted ? 1 : 0, grid, threads, smem, st)
               : rtb_fast_launch(&P, launch_kernel, counted ? 1 : 0, grid, threads, smem, st);
    if (e) return fail(ctx, RTB_ERR_CUDA, "kernel launch: %s", cudaGetErrorString((cudaError_t)e));
    if (timed) { CU(cudaEventRecord(ctx->ev1, st)); ctx->timed_pending = true; }
    if (lpt) {
        /* the tail: three times the tiles whose pixels are in flight when the counter runs out */
        int tail_factor = 3;
        if (const char* e = getenv("RTB_LPT_TAIL")) tail_factor = atoi(e) > 0 ? atoi(e) : 3;      /* development */
        const int chunks = (n_tiles + ORDER_CHUNK - 1) / ORDER_CHUNK;
        OrderPlan* plan = (OrderPlan*)(ctx->tile_hist + (size_t)chunks * 256);
        tile_hist_kernel<<<chunks, 1024, 0, st>>>(ctx->tile_cost, ctx->tile_hist, n_tiles);
        tile_plan_kernel<<<1, 256, 0, st>>>(ctx->tile_hist, plan, n_tiles, chunks, tail_factor * (grid * threads / 32));
        tile_scatter_kernel<<<chunks, 1024, 0, st>>>(ctx->tile_cost, ctx->tile_hist, plan, ctx->tile_perm, n_tiles);
        CU(cudaGetLastError());
        ctx->lpt_valid = true;
    }
    if (st != ctx->stream) {
        /* ... and order the context stream behind this frame: the next upload overwrites the arrays the kernels are reading,
         * the next frame reuses the packed block and the tile counter */
        CU(cudaEventRecord(ctx->ev_order, st));
        CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_order, 0));
    }
    { int rc_ = stage_rotate(ctx); if (rc_) return rc_; }
    ctx->stats.kernel_used = kernel; ctx->stats.grid = grid; ctx->stats.block = threads; ctx->stats.smem_bytes = (int)smem;
    return RTB_OK;
}

extern "C" {

const char* rtb_version(void) { return "rtb200 0.2 (sm_100a)"; }

const char* rtb_last_error(const rtb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

rtb_ctx* rtb_create(int width, int height, int device) {
    rtb_ctx* ctx = nullptr;
    if (width <= 0 || height <= 0) { fail(nullptr, RTB_ERR_INVALID, "bad size %dx%d", width, height); return nullptr; }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) { fail(nullptr, RTB_ERR_NO_DEVICE, "no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e)); return nullptr; }
    if (device < 0 || device >= n) { fail(nullptr, RTB_ERR_INVALID, "device %d out of range (%d devices)", device, n); return nullptr; }
    ctx = new rtb_ctx();
    ctx->device = device; ctx->width = width; ctx->height = height;
    auto bail = [&](const char* what, cudaError_t err) { fail(nullptr, RTB_ERR_CUDA, "%s: %s", what, cudaGetErrorString(err)); rtb_destroy(ctx); return (rtb_ctx*)nullptr; };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bail("cudaGetDeviceProperties", e);
    if (prop.major < 10) { fail(nullptr, RTB_ERR_NO_DEVICE, "device %s is sm_%d%d; this library carries sm_100a code only", prop.name, prop.major, prop.minor); rtb_destroy(ctx); return nullptr; }
    ctx->n_sm = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    if ((e = cudaEventCreate(&ctx->ev0)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreate(&ctx->ev1)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev_order, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    for (int i = 0; i < RTB_STAGE_SLOTS; i++)
        if ((e = cudaEventCreateWithFlags(&ctx->stage[i].done, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaMalloc(&ctx->tile_counter, 256)) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc(&ctx->counters, CNT_NUM * sizeof(unsigned long long))) != cudaSuccess) return bail("cudaMalloc", e);
    ctx->local_rows = height;
    ctx->fb_floats = (size_t)width * height * 4;
    if ((e = cudaMalloc(&ctx->fb, ctx->fb_floats * sizeof(float))) != cudaSuccess) return bail("cudaMalloc framebuffer", e);
    return ctx;
}

void rtb_destroy(rtb_ctx* ctx) {
    if (!ctx) return;
    for (rtb_ctx* p : ctx->peers) { p->root = nullptr; rtb_destroy(p); }
    ctx->peers.clear();
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    rtb_multi_release(ctx);
    rtb_smaa_release(ctx);
    for (int i = 0; i < RTB_STAGE_SLOTS; i++) { if (ctx->stage[i].host) cudaFreeHost(ctx->stage[i].host); if (ctx->stage[i].done) cudaEventDestroy(ctx->stage[i].done); }
    for (int b = 0; b < RTB_NUM_BINDINGS; b++) if (ctx->raw[b]) cudaFree(ctx->raw[b]);
    if (ctx->packed) cudaFree(ctx->packed);
    if (ctx->tile_counter) cudaFree(ctx->tile_counter);
    if (ctx->tile_cost) { cudaFree(ctx->tile_cost); cudaFree(ctx->tile_perm); cudaFree(ctx->tile_hist); }
    if (ctx->counters) cudaFree(ctx->counters);
    if (ctx->cta_times) cudaFree(ctx->cta_times);
    if (ctx->fb) cudaFree(ctx->fb);
    if (ctx->fb8) cudaFree(ctx->fb8);
    if (ctx->cube) cudaFree(ctx->cube);
    for (int u = 0; u < 6; u++) if (ctx->tex[u].dev) cudaFree(ctx->tex[u].dev);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev_order) cudaEventDestroy(ctx->ev_order);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int rtb_set_partition(rtb_ctx* ctx, int rank, int world, int block_rows) {
    if (!ctx) return fail(nullptr, RTB_ERR_INVALID, "null context");
    if (world < 1 || rank < 0 || rank >= world || block_rows < 4 || (block_rows & 3))
        return fail(ctx, RTB_ERR_INVALID, "bad partition rank %d / world %d / block_rows %d (block_rows must be a multiple of 4)", rank, world, block_rows);
    ctx->rank = rank; ctx->world = world; ctx->block_rows = block_rows;
    ctx->local_rows = rtb_compute_local_rows(ctx->height, rank, world, block_rows);
    ctx->lpt_valid = false;                               /* the tile order belongs to the previous share of the frame */
    return RTB_OK;
}

int rtb_local_rows(const rtb_ctx* ctx) { return ctx ? ctx->local_rows : 0; }

int rtb_set_defines(rtb_ctx* ctx, const rtb_defines* d) {
    if (!ctx || !d) return fail(ctx, RTB_ERR_INVALID, "null argument");
    const int* c = &d->sphere_size;
    for (int i = 0; i < 8; i++) if (c[i] < 0 || c[i] > (1 << 20)) return fail(ctx, RTB_ERR_INVALID, "count %d out of range: %d", i, c[i]);
    if (d->iterations < 0) return fail(ctx, RTB_ERR_INVALID, "negative iterations");
    for (rtb_ctx* p : ctx->peers) { int rc = rtb_set_defines(p, d); if (rc) return fail(ctx, rc, "%s", p->err.c_str()); }
    ctx->defines = *d;
    ctx->have_defines = true;
    ctx->dirty = true;
    return RTB_OK;
}

namespace {
/* stage the caller's bytes and queue the H2D copy; `replace` = glBufferData (the block holds exactly `bytes` afterwards),
 * otherwise glBufferSubData(0, bytes) */
int upload_block(rtb_ctx* ctx, int binding, const void* data, size_t bytes, bool replace) {
    if (!ctx) return fail(nullptr, RTB_ERR_INVALID, "null context");
    if (binding < 0 || binding >= RTB_NUM_BINDINGS) return fail(ctx, RTB_ERR_INVALID, "unknown uniform-block binding %d", binding);
    if (bytes % ELEM_SIZE[binding]) return fail(ctx, RTB_ERR_INVALID, "binding %d: %zu bytes is not a multiple of the %zu-byte element", binding, bytes, ELEM_SIZE[binding]);
    if (!replace && bytes > ctx->raw_bytes[binding])
        return fail(ctx, RTB_ERR_INVALID, "binding %d: update of %zu bytes exceeds the block's %zu bytes (glBufferSubData: GL_INVALID_VALUE)", binding, bytes, ctx->raw_bytes[binding]);
    for (rtb_ctx* p : ctx->peers) { int rc = upload_block(p, binding, data, bytes, replace); if (rc) return fail(ctx, rc, "%s", p->err.c_str()); }
    CU(cudaSetDevice(ctx->device));
    if (replace && (bytes > ctx->raw_cap[binding] || !ctx->raw[binding])) {
        /* stream-ordered free / allocate: frames already queued keep reading the old block, nothing synchronises */
        if (ctx->raw[binding]) { CU(cudaFreeAsync(ctx->raw[binding], ctx->stream)); ctx->raw[binding] = nullptr; }
        size_t cap = bytes < 256 ? 256 : bytes;
        CU(cudaMallocAsync(&ctx->raw[binding], cap, ctx->stream));
        ctx->raw_cap[binding] = cap;
    }
    if (data && bytes) {
        uint8_t* host = nullptr;
        int rc = stage_reserve(ctx, bytes, &host);
        if (rc) return rc;
        memcpy(host, data, bytes);                       /* the caller may reuse `data` as soon as we return */
        CU(cudaMemcpyAsync(ctx->raw[binding], host, bytes, cudaMemcpyHostToDevice, ctx->stream));
        if (binding == RTB_BIND_SCENE) { memcpy(&ctx->scene_host, data, sizeof(rtb_scene)); ctx->have_scene = true; }
        /* an update may cover a prefix only: a texture reference seen earlier in the tail stays valid */
        const bool refs = array_uses_textures(binding, data, bytes);
        ctx->uses_tex[binding] = replace ? refs : (refs || (ctx->uses_tex[binding] && bytes < ctx->raw_bytes[binding]));
    } else if (replace && bytes == 0) {
        ctx->uses_tex[binding] = false;
        if (binding == RTB_BIND_SCENE) ctx->have_scene = false;
    }   /* data == NULL with bytes > 0: allocation only (glBufferData(size, NULL)), contents arrive by a later update */
    if (replace) ctx->raw_bytes[binding] = bytes;
    ctx->dirty = true;
    return RTB_OK;
}
}  // namespace

int rtb_upload(rtb_ctx* ctx, int binding, const void* data, size_t bytes) { return upload_block(ctx, binding, data, bytes, true); }
int rtb_update(rtb_ctx* ctx, int binding, const void* data, size_t bytes) { return upload_block(ctx, binding, data, bytes, false); }

int rtb_set_cubemap(rtb_ctx* ctx, const uint8_t* const faces[6], int w, int h, int channels) {
    if (!ctx || !faces) return fail(ctx, RTB_ERR_INVALID, "null argument");
    if (w <= 0 || h <= 0 || channels < 1 || channels > 4) return fail(ctx, RTB_ERR_INVALID, "bad cubemap %dx%dx%d", w, h, channels);
    for (rtb_ctx* p : ctx->peers) { int rc = rtb_set_cubemap(p, faces, w, h, channels); if (rc) return fail(ctx, rc, "%s", p->err.c_str()); }
    CU(cudaSetDevice(ctx->device));
    size_t face_bytes = (size_t)w * h * 4;
    std::vector<uint8_t> rgba(face_bytes * 6);
    for (int f = 0; f < 6; f++) {
        if (!faces[f]) return fail(ctx, RTB_ERR_INVALID, "cubemap face %d is null", f);
        for (size_t i = 0; i < (size_t)w * h; i++) {
            const uint8_t* s = faces[f] + i * channels;
            uint8_t* o = rgba.data() + f * face_bytes + i * 4;
            o[0] = s[0]; o[1] = channels >= 2 ? s[1] : 0; o[2] = channels >= 3 ? s[2] : 0; o[3] = channels == 4 ? s[3] : 255;
        }
    }
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->cube) { cudaFree(ctx->cube); ctx->cube = nullptr; }
    CU(cudaMalloc(&ctx->cube, rgba.size()));
    CU(cudaMemcpy(ctx->cube, rgba.data(), rgba.size(), cudaMemcpyHostToDevice));
    ctx->cube_w = w; ctx->cube_h = h;
    return RTB_OK;
}

int rtb_set_texture2d(rtb_ctx* ctx, int unit, const uint8_t* pixels, int w, int h, int channels) {
    if (!ctx || !pixels) return fail(ctx, RTB_ERR_INVALID, "null argument");
    if (unit < 1 || unit > 5) return fail(ctx, RTB_ERR_INVALID, "texture unit %d out of range 1..5", unit);
    if (w <= 0 || h <= 0 || channels < 1 || channels > 4) return fail(ctx, RTB_ERR_INVALID, "bad texture %dx%dx%d", w, h, channels);
    for (rtb_ctx* p : ctx->peers) { int rc = rtb_set_texture2d(p, unit, pixels, w, h, channels); if (rc) return fail(ctx, rc, "%s", p->err.c_str()); }
    CU(cudaSetDevice(ctx->device));
    std::vector<uint8_t> chain;
    RtbTex2D t;
    build_mip_chain(pixels, w, h, channels, chain, t);
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->tex[unit].dev) { cudaFree(ctx->tex[unit].dev); ctx->tex[unit].dev = nullptr; }
    CU(cudaMalloc(&t.dev, chain.size()));
    CU(cudaMemcpy(t.dev, chain.data(), chain.size(), cudaMemcpyHostToDevice));
    ctx->tex[unit] = t;
    return RTB_OK;
}

int rtb_set_option(rtb_ctx* ctx, const char* key, int value) {
    if (!ctx || !key) return fail(ctx, RTB_ERR_INVALID, "null argument");
    if (strcmp(key, "gather")) for (rtb_ctx* p : ctx->peers) { int rc = rtb_set_option(p, key, value); if (rc) return fail(ctx, rc, "%s", p->err.c_str()); }
    if (!strcmp(key, "kernel")) { if (value < 0 || value > 2) return fail(ctx, RTB_ERR_INVALID, "kernel must be 0..2"); ctx->opt_kernel = value; }
    else if (!strcmp(key, "strict")) { if ((value ? 1 : 0) != ctx->opt_strict) ctx->dirty = true; ctx->opt_strict = value ? 1 : 0; }   /* the two builds stage different records */
    else if (!strcmp(key, "cull")) ctx->opt_cull = value ? 1 : 0;
    else if (!strcmp(key, "ctas_per_sm")) ctx->opt_ctas_per_sm = value;
    else if (!strcmp(key, "coop")) ctx->opt_coop = value ? 1 : 0;
    else if (!strcmp(key, "smaa_compact")) ctx->opt_smaa_compact = value ? 1 : 0;
    else if (!strcmp(key, "wide")) ctx->opt_wide = value < 0 ? -1 : (value ? 1 : 0);
    else if (!strcmp(key, "lpt")) { ctx->opt_lpt = value < 0 ? -1 : (value ? 1 : 0); ctx->lpt_valid = false; }
    else if (!strcmp(key, "gather")) { if (value != RTB_GATHER_NCCL && value != RTB_GATHER_P2P) return fail(ctx, RTB_ERR_INVALID, "gather must be 0 (NCCL) or 1 (P2P)");
        if (value == RTB_GATHER_P2P && (ctx->peers.empty() || !ctx->p2p_ok)) return fail(ctx, RTB_ERR_STATE, "P2P gather needs a multi-device context whose GPUs have peer access to device 0");
        ctx->opt_gather = value; }
    else return fail(ctx, RTB_ERR_INVALID, "unknown option '%s'", key);
    return RTB_OK;
}

int rtb_render(rtb_ctx* ctx) {
    if (!ctx) return fail(nullptr, RTB_ERR_INVALID, "null context");
    ctx->smaa_valid = false;
    if (!ctx->peers.empty()) return rtb_multi_render(ctx);
    int rc = rtb_do_render(ctx, ctx->fb, false, ctx->stream, false, true);
    if (rc == RTB_OK && ctx->smaa_preset >= 0 && ctx->world == 1) {      /* GLWrapper.cpp:173-204: the three SMAA draws follow the ray-trace draw */
        rc = rtb_smaa_after_frame(ctx, ctx->fb, ctx->stream);
        ctx->smaa_valid = rc == RTB_OK;
    }
    return rc;
}

int rtb_render_to(rtb_ctx* ctx, void* device_rgba32f, void* cuda_stream) {
    if (!ctx || !device_rgba32f) return fail(ctx, RTB_ERR_INVALID, "null argument");
    if (!ctx->peers.empty()) return fail(ctx, RTB_ERR_STATE, "rtb_render_to on a multi-device context: use rtb_render + rtb_device_framebuffer");
    return rtb_do_render(ctx, (float*)device_rgba32f, false, cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream, false, cuda_stream == nullptr);
}

int rtb_sync(rtb_ctx* ctx) {
    if (!ctx) return fail(nullptr, RTB_ERR_INVALID, "null context");
    for (rtb_ctx* p : ctx->peers) { int rc = rtb_sync(p); if (rc) return fail(ctx, rc, "%s", p->err.c_str()); }
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->cta_times && ctx->cta_times_n) {
        std::vector<unsigned long long> t((size_t)ctx->cta_times_n * 5);
        if (cudaMemcpy(t.data(), ctx->cta_times, t.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess) {
            unsigned long long t0 = ~0ull, dmin = ~0ull, dmax = 0, emin = ~0ull, emax = 0;
            for (int i = 0; i < ctx->cta_times_n; i++) if (t[i * 5] && t[i * 5] < t0) t0 = t[i * 5];
            std::vector<double> ends;
            for (int i = 0; i < ctx->cta_times_n; i++) {
                unsigned long long d = t[i * 5 + 1], e = t[i * 5 + 2];
                if (d) { if (d < dmin) dmin = d; if (d > dmax) dmax = d; }
                if (e) { if (e < emin) emin = e; if (e > emax) emax = e; ends.push_back((double)(e - t0) * 1e-6); }
            }
            std::sort(ends.begin(), ends.end());
            fprintf(stderr, "cta times (ms from first CTA start): drain start %.2f..%.2f, CTA end min %.2f p25 %.2f median %.2f p75 %.2f p95 %.2f max %.2f\n",
                    (double)(dmin - t0) * 1e-6, (double)(dmax - t0) * 1e-6, ends.empty() ? 0. : ends.front(), ends.empty() ? 0. : ends[ends.size() / 4],
                    ends.empty() ? 0. : ends[ends.size() / 2], ends.empty() ? 0. : ends[ends.size() * 3 / 4], ends.empty() ? 0. : ends[ends.size() * 95 / 100], ends.empty() ? 0. : ends.back());
            /* the five slowest CTAs: end time, drain trips, drain jobs, serial fallbacks */
            std::vector<int> order(ctx->cta_times_n);
            for (int i = 0; i < ctx->cta_times_n; i++) order[i] = i;
            std::sort(order.begin(), order.end(), [&](int a, int b) { return t[a * 5 + 2] > t[b * 5 + 2]; });
            unsigned long long tj = 0, tf = 0;
            for (int i = 0; i < ctx->cta_times_n; i++) { tj += t[i * 5 + 3] & 0xffffffffu; tf += t[i * 5 + 4]; }
            fprintf(stderr, "  drain jobs total %llu (mean %.0f per CTA), fallbacks %llu; slowest:", tj, (double)tj / ctx->cta_times_n, tf);
            for (int k = 0; k < 5 && k < ctx->cta_times_n; k++) {
                int i = order[k];
                fprintf(stderr, " [end %.2f trips %llu jobs %llu fb %llu]", (double)(t[i * 5 + 2] - t0) * 1e-6, t[i * 5 + 3] >> 32, t[i * 5 + 3] & 0xffffffffu, t[i * 5 + 4]);
            }
            int i = order[ctx->cta_times_n / 2];
            fprintf(stderr, " median [end %.2f trips %llu jobs %llu fb %llu]\n", (double)(t[i * 5 + 2] - t0) * 1e-6, t[i * 5 + 3] >> 32, t[i * 5 + 3] & 0xffffffffu, t[i * 5 + 4]);
        }
        ctx->cta_times_n = 0;
    }
    if (ctx->timed_pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess) ctx->stats.kernel_ms = ms;
        ctx->timed_pending = false;
    }
    if (ctx->frame_timed) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->ev_f0, ctx->ev_f1) == cudaSuccess) ctx->frame_ms = ms;
        ctx->frame_timed = false;
    }
    return RTB_OK;
}

int rtb_render_counted(rtb_ctx* ctx, rtb_stats* out) {
    if (!ctx) return fail(nullptr, RTB_ERR_INVALID, "null context");
    if (!ctx->peers.empty()) return fail(ctx, RTB_ERR_STATE, "rtb_render_counted on a multi-device context: count on a single-device one");
    int rc = rtb_do_render(ctx, ctx->fb, false, ctx->stream, true, false);
    if (rc) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    unsigned long long c[CNT_NUM];
    CU(cudaMemcpy(c, ctx->counters, sizeof c, cudaMemcpyDeviceToHost));
    rtb_stats& s = ctx->stats;
    const rtb_defines& d = ctx->defines;
    s.pixels = c[CNT_PIXELS]; s.rays_nearest = c[CNT_RAYS_NEAREST]; s.rays_shadow = c[CNT_RAYS_SHADOW];
    s.dk_iterations = c[CNT_DK]; s.light_evals = c[CNT_LIGHT_EVALS];
    for (int i = 0; i < 7; i++) s.shaded_hits[i] = c[CNT_SHADED0 + i];
    /* every scan tests every primitive of the classes it visits (rt.frag:587-658) */
    const uint64_t rn = s.rays_nearest, rs = s.rays_shadow;
    s.tests[RTB_TYPE_SPHERE] = (rn + rs) * d.sphere_size;   s.tests[RTB_TYPE_PLANE] = rn * d.plane_size;
    s.tests[RTB_TYPE_SURFACE] = (rn + rs) * d.surface_size; s.tests[RTB_TYPE_BOX] = (rn + rs) * d.box_size;
    s.tests[RTB_TYPE_TORUS] = (rn + rs) * d.torus_size;     s.tests[RTB_TYPE_RING] = (rn + rs) * d.ring_size;
    s.tests[RTB_TYPE_POINT_LIGHT] = rn * d.light_point_size;
    s.flops = algorithmic_flops(s);
    if (out) *out = s;
    return RTB_OK;
}

int rtb_get_stats(rtb_ctx* ctx, rtb_stats* out) {
    if (!ctx || !out) return fail(ctx, RTB_ERR_INVALID, "null argument");
    int rc = rtb_sync(ctx);
    if (rc) return rc;
    *out = ctx->stats;
    return RTB_OK;
}

int rtb_read_rgba32f(rtb_ctx* ctx, float* dst) {
    if (!ctx || !dst) return fail(ctx, RTB_ERR_INVALID, "null argument");
    int rc = rtb_sync(ctx);
    if (rc) return rc;
    if (ctx->fb_full) CU(cudaMemcpy(dst, ctx->fb_full, (size_t)ctx->height * ctx->width * 4 * sizeof(float), cudaMemcpyDeviceToHost));
    else CU(cudaMemcpy(dst, ctx->fb, (size_t)ctx->local_rows * ctx->width * 4 * sizeof(float), cudaMemcpyDeviceToHost));
    return RTB_OK;
}

int rtb_read_rgba8(rtb_ctx* ctx, uint8_t* dst) {
    if (!ctx || !dst) return fail(ctx, RTB_ERR_INVALID, "null argument");
    CU(cudaSetDevice(ctx->device));
    const size_t n = (size_t)(ctx->fb_full ? ctx->height : ctx->local_rows) * ctx->width;
    if (n == 0) return rtb_sync(ctx);
    if (ctx->smaa_valid) {                               /* the frame went through the SMAA passes: this is what the reference puts on screen */
        int rc_ = rtb_sync(ctx);
        if (rc_) return rc_;
        CU(cudaMemcpy(dst, ctx->smaa_out, n * 4, cudaMemcpyDeviceToHost));
        return RTB_OK;
    }
    if (!ctx->fb8) CU(cudaMalloc(&ctx->fb8, (size_t)ctx->width * ctx->height * 4));
    rgba8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((const float4*)(ctx->fb_full ? ctx->fb_full : ctx->fb), (uchar4*)ctx->fb8, n);   /* after the frame, same stream */
    CU(cudaGetLastError());
    int rc = rtb_sync(ctx);
    if (rc) return rc;
    CU(cudaMemcpy(dst, ctx->fb8, n * 4, cudaMemcpyDeviceToHost));
    return RTB_OK;
}

int rtb_tile_order(rtb_ctx* ctx, uint32_t* cost, uint32_t* order, int capacity) {
    if (!ctx || capacity < 0) { fail(ctx, RTB_ERR_INVALID, "bad argument"); return RTB_ERR_INVALID; }
    if (!ctx->lpt_valid || !ctx->tile_cost) return 0;
    int rc = rtb_sync(ctx);
    if (rc) return rc;
    const int n = ctx->lpt_tiles < capacity ? ctx->lpt_tiles : capacity;
    if (cost && cudaMemcpy(cost, ctx->tile_cost, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess) { fail(ctx, RTB_ERR_CUDA, "cudaMemcpy"); return RTB_ERR_CUDA; }
    if (order && cudaMemcpy(order, ctx->tile_perm, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess) { fail(ctx, RTB_ERR_CUDA, "cudaMemcpy"); return RTB_ERR_CUDA; }
    return ctx->lpt_tiles;
}

void* rtb_device_framebuffer(rtb_ctx* ctx) { return ctx ? (ctx->fb_full ? ctx->fb_full : ctx->fb) : nullptr; }

}  // extern "C"

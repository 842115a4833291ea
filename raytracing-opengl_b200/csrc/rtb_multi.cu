/* rtb_multi.cu — the multi-GPU part of the C-ABI (include/rtb200.h): the frame-end gather of the tile-partitioned
 * framebuffer over NCCL / NVLink, for one process per GPU (rtb_comm_init + rtb_gather) and for one process driving all
 * GPUs of the box (rtb_create_multi).  The reference is single-GPU (one glDrawArrays per frame, GLWrapper.cpp:155-165);
 * this is what stands behind the same draw() call when RT_GPUS > 1.
 *
 * NCCL is bound at run time (dlopen "libnccl.so.2"): inside a torch process that is the library torch already loaded, in
 * a plain C++ host the system one; a single-GPU user never touches it.
 */
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstring>
#include <mutex>

#include "rtb_ctx.h"

namespace {

struct NcclApi {
    void* lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    std::string why;
};

NcclApi* nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : { "libnccl.so.2", "libnccl.so" }) {
            api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) { api.why = std::string("dlopen libnccl.so.2: ") + dlerror(); return; }
#define BIND(sym) api.sym = (decltype(api.sym))dlsym(api.lib, "nccl" #sym); if (!api.sym) { api.why = "libnccl lacks nccl" #sym; api.lib = nullptr; return; }
        BIND(GetUniqueId) BIND(CommInitRank) BIND(CommInitAll) BIND(CommDestroy) BIND(GroupStart) BIND(GroupEnd) BIND(Send) BIND(Recv) BIND(GetErrorString)
#undef BIND
    });
    return &api;
}

#define NC(call)                                                                                                \
    do {                                                                                                        \
        ncclResult_t r_ = (call);                                                                               \
        if (r_ != ncclSuccess) return rtb_fail(ctx, RTB_ERR_CUDA, "%s: %s", #call, nccl()->GetErrorString(r_)); \
    } while (0)

/* where rank r's packed rows start inside the root's scratch (in floats); rank 0's rows come from its own buffer */
struct GatherMap { int world, block_rows, width, height; long long off[16]; };
constexpr int MAX_RANKS = 16;

/* packed per-rank rows -> the frame.  One float4 (pixel) per thread, both sides coalesced: a pure HBM stream. */
__global__ void deinterleave_kernel(const float4* __restrict__ own, const float4* __restrict__ scratch, float4* __restrict__ dst, GatherMap m) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)m.width * m.height) return;
    const int y = (int)(i / m.width), x = (int)(i - (size_t)y * m.width);
    const int b = y / m.block_rows, r = b % m.world, lb = b / m.world;
    const size_t ly = (size_t)lb * m.block_rows + (y - b * m.block_rows);
    const float4* src = r == 0 ? own : scratch + m.off[r] / 4;
    dst[i] = src[ly * m.width + x];
}

int make_map(rtb_ctx* ctx, GatherMap& m) {
    if (ctx->world > MAX_RANKS) return rtb_fail(ctx, RTB_ERR_INVALID, "at most %d ranks", MAX_RANKS);
    m.world = ctx->world; m.block_rows = ctx->block_rows; m.width = ctx->width; m.height = ctx->height;
    long long o = 0;
    for (int r = 0; r < ctx->world; r++) {
        m.off[r] = o;
        if (r > 0) o += (long long)rtb_compute_local_rows(ctx->height, r, ctx->world, ctx->block_rows) * ctx->width * 4;
    }
    m.off[0] = 0;
    if ((size_t)o > ctx->gather_scratch_floats) {
        if (ctx->gather_scratch) { CU(cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->gather_scratch); ctx->gather_scratch = nullptr; }
        CU(cudaMalloc(&ctx->gather_scratch, (size_t)o * sizeof(float)));
        ctx->gather_scratch_floats = (size_t)o;
    }
    return RTB_OK;
}

}  // namespace

void rtb_multi_release(rtb_ctx* ctx) {
    if (ctx->comm && nccl()->lib) nccl()->CommDestroy((ncclComm_t)ctx->comm);
    ctx->comm = nullptr;
    if (ctx->gather_scratch) cudaFree(ctx->gather_scratch);
    if (ctx->fb_full) cudaFree(ctx->fb_full);
    if (ctx->ev_f0) cudaEventDestroy(ctx->ev_f0);
    if (ctx->ev_f1) cudaEventDestroy(ctx->ev_f1);
    ctx->gather_scratch = nullptr; ctx->fb_full = nullptr; ctx->ev_f0 = ctx->ev_f1 = nullptr;
}

/* one frame of a single-process multi-device context: N kernels, one gather, everything queued asynchronously */
int rtb_multi_render(rtb_ctx* ctx) {
    const int n = (int)ctx->peers.size() + 1;
    const bool p2p = ctx->opt_gather == RTB_GATHER_P2P;
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventRecord(ctx->ev_f0, ctx->stream));
    GatherMap m;
    if (!p2p) { int rc = make_map(ctx, m); if (rc) return rc; }
    for (int r = 1; r < n; r++) {
        rtb_ctx* p = ctx->peers[r - 1];
        if (cudaSetDevice(p->device) != cudaSuccess) return rtb_fail(ctx, RTB_ERR_CUDA, "cudaSetDevice(%d)", p->device);
        if (cudaStreamWaitEvent(p->stream, ctx->ev_f0, 0) != cudaSuccess) return rtb_fail(ctx, RTB_ERR_CUDA, "cudaStreamWaitEvent");   /* the frame starts on the root's clock */
        int rc = p2p ? rtb_do_render(p, ctx->fb_full, true, p->stream, false, true) : rtb_do_render(p, p->fb, false, p->stream, false, true);
        if (rc) return rtb_fail(ctx, rc, "rank %d: %s", r, p->err.c_str());
    }
    CU(cudaSetDevice(ctx->device));
    int rc = p2p ? rtb_do_render(ctx, ctx->fb_full, true, ctx->stream, false, true) : rtb_do_render(ctx, ctx->fb, false, ctx->stream, false, true);
    if (rc) return rc;
    if (p2p) {
        /* the peers stored their scanlines straight into the root's frame (NVLink peer access): the frame is complete when all kernels are */
        for (int r = 1; r < n; r++) {
            rtb_ctx* p = ctx->peers[r - 1];
            if (cudaSetDevice(p->device) != cudaSuccess || cudaEventRecord(p->ev_order, p->stream) != cudaSuccess) return rtb_fail(ctx, RTB_ERR_CUDA, "cudaEventRecord on rank %d", r);
            CU(cudaSetDevice(ctx->device));
            CU(cudaStreamWaitEvent(ctx->stream, p->ev_order, 0));
        }
    } else {
        /* the single NCCL gather: every peer sends its packed rows, the root receives them side by side ... */
        NC(nccl()->GroupStart());
        for (int r = 1; r < n; r++) {
            rtb_ctx* p = ctx->peers[r - 1];
            const size_t count = (size_t)p->local_rows * p->width * 4;
            if (count == 0) continue;
            NC(nccl()->Send(p->fb, count, ncclFloat, 0, (ncclComm_t)p->comm, p->stream));
            NC(nccl()->Recv(ctx->gather_scratch + m.off[r], count, ncclFloat, r, (ncclComm_t)ctx->comm, ctx->stream));
        }
        NC(nccl()->GroupEnd());
        /* ... and one pass puts every 4-row block at its final place */
        CU(cudaSetDevice(ctx->device));
        const size_t px = (size_t)ctx->width * ctx->height;
        deinterleave_kernel<<<(unsigned)((px + 255) / 256), 256, 0, ctx->stream>>>((const float4*)ctx->fb, (const float4*)ctx->gather_scratch, (float4*)ctx->fb_full, m);
        CU(cudaGetLastError());
    }
    if (ctx->smaa_preset >= 0) {                          /* the post-pass runs on the assembled frame, on the root */
        rc = rtb_smaa_after_frame(ctx, ctx->fb_full, ctx->stream);
        if (rc) return rc;
        ctx->smaa_valid = true;
    }
    CU(cudaEventRecord(ctx->ev_f1, ctx->stream));
    ctx->frame_timed = true;
    return RTB_OK;
}

extern "C" {

int rtb_n_gpus(const rtb_ctx* ctx) { return ctx ? (int)ctx->peers.size() + 1 : 0; }

rtb_ctx* rtb_create_multi(int width, int height, int n_gpus, int block_rows) {
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have == 0) { rtb_fail(nullptr, RTB_ERR_NO_DEVICE, "no CUDA device (this library has no CPU fallback)"); return nullptr; }
    if (n_gpus < 1 || n_gpus > have || n_gpus > MAX_RANKS) { rtb_fail(nullptr, RTB_ERR_INVALID, "n_gpus %d out of range (%d devices visible)", n_gpus, have); return nullptr; }
    if (block_rows < 4 || (block_rows & 3)) { rtb_fail(nullptr, RTB_ERR_INVALID, "block_rows must be a positive multiple of 4"); return nullptr; }
    rtb_ctx* root = rtb_create(width, height, 0);
    if (!root) return nullptr;
    if (n_gpus == 1) return root;
    auto bail = [&](const char* what) { rtb_fail(nullptr, RTB_ERR_CUDA, "%s: %s", what, root->err.empty() ? cudaGetErrorString(cudaGetLastError()) : root->err.c_str()); rtb_destroy(root); return (rtb_ctx*)nullptr; };
    if (!nccl()->lib) { rtb_fail(nullptr, RTB_ERR_STATE, "NCCL is not available: %s", nccl()->why.c_str()); rtb_destroy(root); return nullptr; }
    for (int r = 1; r < n_gpus; r++) {
        rtb_ctx* p = rtb_create(width, height, r);
        if (!p) { rtb_destroy(root); return nullptr; }
        p->root = root;
        root->peers.push_back(p);
        /* P2P gather mode: rank r's kernel stores into device 0's frame */
        int can = 0;
        cudaSetDevice(r);
        if (cudaDeviceCanAccessPeer(&can, r, 0) == cudaSuccess && can) {
            cudaError_t e = cudaDeviceEnablePeerAccess(0, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
            if (e != cudaSuccess) can = 0;
        }
        if (!can) root->p2p_ok = false;
    }
    for (int r = 0; r < n_gpus; r++) {
        rtb_ctx* c = r == 0 ? root : root->peers[r - 1];
        if (rtb_set_partition(c, r, n_gpus, block_rows)) return bail("rtb_set_partition");
    }
    cudaSetDevice(0);
    if (cudaMalloc(&root->fb_full, (size_t)width * height * 4 * sizeof(float)) != cudaSuccess) return bail("cudaMalloc frame");
    if (cudaEventCreate(&root->ev_f0) != cudaSuccess || cudaEventCreate(&root->ev_f1) != cudaSuccess) return bail("cudaEventCreate");
    std::vector<ncclComm_t> comms(n_gpus);
    std::vector<int> devs(n_gpus);
    for (int r = 0; r < n_gpus; r++) devs[r] = r;
    ncclResult_t nr = nccl()->CommInitAll(comms.data(), n_gpus, devs.data());
    if (nr != ncclSuccess) { rtb_fail(nullptr, RTB_ERR_CUDA, "ncclCommInitAll: %s", nccl()->GetErrorString(nr)); rtb_destroy(root); return nullptr; }
    root->comm = comms[0];
    for (int r = 1; r < n_gpus; r++) root->peers[r - 1]->comm = comms[r];
    cudaSetDevice(0);
    return root;
}

int rtb_rank_times(rtb_ctx* ctx, float* kernel_ms, float* frame_ms) {
    if (!ctx) return rtb_fail(nullptr, RTB_ERR_INVALID, "null context");
    int rc = rtb_sync(ctx);
    if (rc) return rc;
    if (kernel_ms) {
        kernel_ms[0] = ctx->stats.kernel_ms;
        for (size_t r = 0; r < ctx->peers.size(); r++) kernel_ms[r + 1] = ctx->peers[r]->stats.kernel_ms;
    }
    if (frame_ms) *frame_ms = ctx->peers.empty() ? ctx->stats.kernel_ms : ctx->frame_ms;
    return RTB_OK;
}

int rtb_comm_unique_id(uint8_t id[RTB_COMM_ID_BYTES]) {
    rtb_ctx* ctx = nullptr;
    if (!id) return rtb_fail(nullptr, RTB_ERR_INVALID, "null argument");
    if (!nccl()->lib) return rtb_fail(nullptr, RTB_ERR_STATE, "NCCL is not available: %s", nccl()->why.c_str());
    static_assert(sizeof(ncclUniqueId) == RTB_COMM_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId u;
    NC(nccl()->GetUniqueId(&u));
    memcpy(id, &u, sizeof u);
    return RTB_OK;
}

int rtb_comm_init(rtb_ctx* ctx, const uint8_t id[RTB_COMM_ID_BYTES], int rank, int world, int block_rows) {
    if (!ctx || !id) return rtb_fail(ctx, RTB_ERR_INVALID, "null argument");
    if (!ctx->peers.empty()) return rtb_fail(ctx, RTB_ERR_STATE, "rtb_comm_init on a multi-device context");
    if (world > MAX_RANKS) return rtb_fail(ctx, RTB_ERR_INVALID, "at most %d ranks", MAX_RANKS);
    if (!nccl()->lib) return rtb_fail(ctx, RTB_ERR_STATE, "NCCL is not available: %s", nccl()->why.c_str());
    int rc = rtb_set_partition(ctx, rank, world, block_rows);
    if (rc) return rc;
    CU(cudaSetDevice(ctx->device));
    if (ctx->comm) { nccl()->CommDestroy((ncclComm_t)ctx->comm); ctx->comm = nullptr; }
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    ncclComm_t c;
    NC(nccl()->CommInitRank(&c, world, u, rank));
    ctx->comm = c;
    return RTB_OK;
}

int rtb_gather(rtb_ctx* ctx, const void* local_rows_device, void* full_frame_device, void* cuda_stream) {
    if (!ctx) return rtb_fail(nullptr, RTB_ERR_INVALID, "null context");
    if (!ctx->comm || !ctx->peers.empty()) return rtb_fail(ctx, RTB_ERR_STATE, "rtb_gather needs rtb_comm_init (one process per GPU)");
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    const float* own = local_rows_device ? (const float*)local_rows_device : ctx->fb;
    CU(cudaSetDevice(ctx->device));
    if (ctx->rank != 0) {
        const size_t count = (size_t)ctx->local_rows * ctx->width * 4;
        if (count) NC(nccl()->Send(own, count, ncclFloat, 0, (ncclComm_t)ctx->comm, st));
        return RTB_OK;
    }
    if (!full_frame_device) return rtb_fail(ctx, RTB_ERR_INVALID, "the root needs the destination of the frame");
    GatherMap m;
    int rc = make_map(ctx, m);
    if (rc) return rc;
    NC(nccl()->GroupStart());
    for (int r = 1; r < ctx->world; r++) {
        const size_t count = (size_t)rtb_compute_local_rows(ctx->height, r, ctx->world, ctx->block_rows) * ctx->width * 4;
        if (count) NC(nccl()->Recv(ctx->gather_scratch + m.off[r], count, ncclFloat, r, (ncclComm_t)ctx->comm, st));
    }
    NC(nccl()->GroupEnd());
    const size_t px = (size_t)ctx->width * ctx->height;
    deinterleave_kernel<<<(unsigned)((px + 255) / 256), 256, 0, st>>>((const float4*)own, (const float4*)ctx->gather_scratch, (float4*)full_frame_device, m);
    CU(cudaGetLastError());
    return RTB_OK;
}

}  // extern "C"

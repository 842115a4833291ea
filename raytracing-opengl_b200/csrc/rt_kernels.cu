/* rt_kernels.cu — the sm_100a kernels of the ray-trace pass.
 *
 *   quad_kernel        one warp = one 8x4 pixel tile, the four lanes of every 2x2
 *                      quad adjacent (lane^1 = x neighbour, lane^2 = y neighbour) and
 *                      in lock step, so fwidth()/implicit-LOD texture fetches see
 *                      their neighbours (rt.frag:326,396,433-435).  Persistent CTAs
 *                      pull tiles from an atomic counter.
 *   persistent_kernel  (rt_persistent.cuh) persistent threads with per-lane ray
 *                      refill for scenes without 2-D textures.
 *
 * Both stage the packed hot-geometry block (rt_params.h) into shared memory with
 * one TMA bulk copy per CTA (cp.async.bulk + mbarrier complete_tx).
 *
 * Compiled twice: -DRTB_STRICT=1 -DRTB_NS=rtb_strict and -DRTB_STRICT=0 -DRTB_NS=rtb_fast.
 */
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include "rt_launch.h"
#include "rt_scan.cuh"

namespace RTB_NS {

/* ------------------------------------------------------------------ TMA staging */
DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

DEV void stage_scene_tma(uint8_t* smem, uint64_t* mbar, const uint8_t* gsrc, uint32_t bytes) {
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(bytes) : "memory");
        /* 1-D bulk copy global -> shared, completion counted in bytes on the mbarrier (SASS: UBLKCP) */
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(smem)), "l"(gsrc), "r"(bytes), "r"(smem_u32(mbar)) : "memory");
    }
    /* every thread waits for phase 0 of the barrier */
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(mbar)) : "memory");
    }
}

DEV unsigned long long globaltimer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

DEV void flush_counters(const FrameParams& P, const Counters& c) {
    if (!P.counters) return;
    auto red = [&](unsigned v, int slot) {
        unsigned long long s = v;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
        if ((threadIdx.x & 31) == 0 && s) atomicAdd(P.counters + slot, s);
    };
    red(c.rays_n, CNT_RAYS_NEAREST); red(c.rays_s, CNT_RAYS_SHADOW); red(c.dk, CNT_DK); red(c.light_evals, CNT_LIGHT_EVALS);
    red(c.pixels, CNT_PIXELS);
    for (int i = 0; i < 7; i++) red(c.shaded[i], CNT_SHADED0 + i);
}

/* local (this rank's) scanline -> canvas scanline under the row-block partition */
DEV int global_row(const FrameParams& P, int ly) {
    int lb = ly / P.block_rows;
    return (lb * P.world + P.rank) * P.block_rows + (ly - lb * P.block_rows);
}

/* ------------------------------------------------------------------ quad kernel */
/* The quad kernel scans the scene from three places (main path, getReflectedColor, the light loop).  Inlined three times the
 * kernel is ~45 KB of SASS and, with 16 warps per SM at 16 different places of it, instruction-fetch bound: ncu shows
 * `stalled_no_instruction` 4.7 warps per issue and issue slots 46 % busy on the default scene
 * (profiles/r2_ncu_default1080_fused_quad_kernel.txt).  RTB_QUAD_SCAN_CALL=1 (default) makes the scan ONE function that is called. */
#ifndef RTB_QUAD_SCAN_CALL
#define RTB_QUAD_SCAN_CALL 1                    /* measured: default1080 fused 1.107 -> 0.936 ms, strict 2.43 -> 2.13 ms (profiles/r2_quad_scan_call_ab.jsonl) */
#endif
#ifndef QUAD_MIN_BLOCKS
#define QUAD_MIN_BLOCKS 1
#endif
template <bool COUNT>
#if RTB_QUAD_SCAN_CALL
__device__ __noinline__
#else
DEV
#endif
void scan_quad(const FrameParams& P, const SceneView& S, vec3 ro, vec3 rd, bool active, bool shadow_mode, float limit, int ctx,
               float& tmin_out, int& id_out, float& shadow_out, vec2& ring_uv_out, Counters& cnt) {
    scan_scene<COUNT, true, true>(P, S, ro, rd, active, shadow_mode, limit, ctx, tmin_out, id_out, shadow_out, ring_uv_out, cnt);
}

template <bool COUNT>
__global__ void __launch_bounds__(QUAD_THREADS, QUAD_MIN_BLOCKS) quad_kernel(const __grid_constant__ FrameParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t mbar;
    stage_scene_tma(smem, &mbar, P.packed, P.lay.total_bytes);
    const SceneView S = make_view(smem, P.lay);

    const int lane = threadIdx.x & 31;
    const int n_tiles = P.n_tiles_x * P.n_tiles_y;
    const int n_lights = P.n_lpoint + P.n_ldirect;
    Counters cnt = {};

    for (;;) {
        int tile = 0;
        if (lane == 0) tile = (int)atomicAdd(P.tile_counter, 1u);
        tile = __shfl_sync(FULL, tile, 0);
        if (tile >= n_tiles) break;
        const int tx = tile % P.n_tiles_x, ty = tile / P.n_tiles_x;
        /* lane -> pixel: quad q = lane>>2 (4 across, 2 down), inside the quad bit0 = x, bit1 = y */
        const int qd = lane >> 2;
        const int x = tx * 8 + (qd & 3) * 2 + (lane & 1);
        const int ly = ty * 4 + (qd >> 2) * 2 + ((lane >> 1) & 1);
        const int y = global_row(P, ly);
        const bool inside = x < P.canvas_w && ly < P.local_rows && y < P.canvas_h;   /* others are helper invocations */

        vec3 mask = mk3(1.f, 1.f, 1.f), color = mk3(0.f, 0.f, 0.f);
        vec3 ro = mk3(P.cam_pos[0], P.cam_pos[1], P.cam_pos[2]);
        vec3 rd = getRayDir(P, x, y);
        float absorbDistance = 0.f;
        int it = 0, glass = 0;
        bool alive = P.iterations > 0;
        if (COUNT && inside) cnt.pixels++;
        Counters* cp = &cnt;
        Counters dummy;
        if (COUNT && !inside) { dummy = Counters(); cp = &dummy; }   /* helper lanes are not counted */

        while (__any_sync(FULL, alive)) {                   /* one loop trip of rt.frag:821, all lanes together */
            float tm, sh_unused; int id; vec2 ruv;
            scan_quad<COUNT>(P, S, ro, rd, alive, false, MAX_DIST, 0, tm, id, sh_unused, ruv, *cp);
            bool hit = alive && tm < MAX_DIST;
            if (alive && !hit) {                            /* rt.frag:892-895 */
                color = color + texture_cube(P.cube, rd) * mask;
                alive = false;
            }
            vec3 pt = ro + rd * tm;
            Material mat = {}; vec3 n; float alpha, bias;
            hit_info<COUNT, true>(P, hit, id, ro, rd, pt, tm, ruv, mat, n, alpha, bias, *cp);
            if (hit && id_type(id) == RTB_TYPE_POINT_LIGHT) {   /* rt.frag:829-832 */
                const float* lc = P.lights_point[id_num(id)].color;
                color = color + mk3(lc[0], lc[1], lc[2]) * mask;
                alive = false; hit = false;
            }
            bool outside = dot(rd, n) < 0;
            n = outside ? n : -n;
            float reflectMultiplier = 0.f;
            if (hit) {
                if (mat.refraction > 0)
                    reflectMultiplier = FresnelReflectAmount(outside ? 1 : mat.refraction, outside ? mat.refraction : 1, rd, n, mat.reflection);
                else
                    reflectMultiplier = getFresnel(n, rd, mat.reflection);
            }
            float refractMultiplier = 1 - reflectMultiplier;
            const bool refractive = hit && mat.refraction > 0.0f;
            const bool reflective = hit && !refractive && mat.reflection > 0.0f;
            const bool diffuse_m = hit && !refractive && !reflective;

            /* getReflectedColor (rt.frag:787-802) for glass seen from outside */
            const bool sub = refractive && outside && mat.reflection > 0;
            vec3 sro = pt + n * bias, srd = reflect(rd, n);
            vec3 subcolor = mk3(0.f, 0.f, 0.f);
            bool sub_shade = false;
            Material mat2 = {}; vec3 n2 = mk3(0.f, 0.f, 0.f); vec3 spt2 = sro;
            if (__any_sync(FULL, sub)) {
                float t2; int id2; vec2 ruv2;
                scan_quad<COUNT>(P, S, sro, srd, sub, false, MAX_DIST, 1, t2, id2, sh_unused, ruv2, *cp);
                bool sub_light = sub && id2 >= 0 && id_type(id2) == RTB_TYPE_POINT_LIGHT;
                bool hit2 = sub && !sub_light && t2 < MAX_DIST;
                vec3 pt2 = sro + srd * t2;
                float alpha2, bias2;
                hit_info<COUNT, true>(P, hit2, id2, sro, srd, pt2, t2, ruv2, mat2, n2, alpha2, bias2, *cp);
                if (sub_light) { const float* lc = P.lights_point[id_num(id2)].color; subcolor = mk3(lc[0], lc[1], lc[2]); }
                if (hit2) { spt2 = dot(srd, n2) < 0 ? pt2 + n2 * bias2 : pt2 - n2 * bias2; sub_shade = true; }
            }

            /* calcShade (rt.frag:681-709): at most one per lane and trip */
            const bool do_shade = reflective || diffuse_m || sub_shade;
            vec3 s_pt = sub_shade ? spt2 : pt + n * bias;
            vec3 s_rd = sub_shade ? srd : rd;
            vec3 s_n = sub_shade ? n2 : n;
            vec3 s_col = sub_shade ? mat2.color : mat.color;
            float s_dif = sub_shade ? mat2.diffuse : mat.diffuse;
            int s_spec = sub_shade ? mat2.specular : mat.specular;
            float s_kd = sub_shade ? mat2.kd : mat.kd, s_ks = sub_shade ? mat2.ks : mat.ks;
            vec3 diffuse = mk3(0.f, 0.f, 0.f), specular = mk3(0.f, 0.f, 0.f);
            if (__any_sync(FULL, do_shade)) {
                for (int l = 0; l < n_lights; l++) {
                    LightSample L = light_sample(P, l, s_pt);
                    float tdummy, shadow; int iddummy; vec2 uvdummy;
                    scan_quad<COUNT>(P, S, s_pt, L.dir_n, do_shade, true, L.dist, sub_shade ? 1 : 0, tdummy, iddummy, shadow, uvdummy, *cp);
                    if (do_shade) {
                        if (COUNT) cp->light_evals++;
                        shade_light(P, L, shadow, s_rd, s_col, s_dif, s_spec, s_n, diffuse, specular);
                    }
                }
            }
            vec3 pixelColor = mk3(P.ambient[0], P.ambient[1], P.ambient[2]) * s_col;
            pixelColor = pixelColor + (diffuse * s_kd + specular * s_ks);

            if (refractive) {                               /* rt.frag:851-873 */
                if (sub) {
                    vec3 rc = sub_shade ? pixelColor : subcolor;
                    color = color + rc * reflectMultiplier * mask;
                    mask = mask * refractMultiplier;
                } else if (!outside) {
                    absorbDistance += tm;
                    vec3 a = -mat.absorb * absorbDistance;
                    mask = mask * mk3(expf(a.x), expf(a.y), expf(a.z));
                }
                if (reflectMultiplier >= 1) alive = false;
                else {
                    ro = pt - n * bias;
                    rd = refract(rd, n, outside ? 1 / mat.refraction : mat.refraction);
                    it--;
                    if (++glass >= MAX_GLASS_EVENTS) alive = false;
                }
            } else if (reflective) {                        /* rt.frag:874-880 */
                ro = pt + n * bias;
                color = color + pixelColor * refractMultiplier * mask;
                rd = reflect(rd, n);
                mask = mask * reflectMultiplier;
            } else if (diffuse_m) {                         /* rt.frag:881-890 */
                color = color + pixelColor * mask * alpha;
                if (alpha < 1) { ro = pt - n * bias; mask = mask * (1 - alpha); }
                else alive = false;
            }
            if (alive) { it++; if (it >= P.iterations) alive = false; }
        }
        if (inside) {
            float4 o = make_float4(color.x, color.y, color.z, 1.0f);
            *(float4*)(P.fb + ((size_t)(P.fb_global ? y : ly) * P.canvas_w + x) * 4) = o;
        }
    }
    if (COUNT) flush_counters(P, cnt);
}

/* ------------------------------------------------------------------ pack kernel */
/* raw std140 arrays -> hot geometry block (layout in rt_params.h).  One small block.  Strict build: exact copies plus the
 * squares r*r / R*R (single multiplies the shader would redo per test).  Fused build: the rotated primitives carry the 3x3
 * matrix of their quaternion sandwich (computed in fp64, rounded once) and a few folded constants. */
#if !RTB_STRICT
/* rotate(q, v) = q (v, 0) q* (rt.frag:305-311) as a matrix; no normalisation, exactly like the shader */
DEV void quat_matrix(const float* q, float* m) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    m[0] = (float)(w * w + x * x - y * y - z * z); m[1] = (float)(2.0 * (x * y - w * z)); m[2] = (float)(2.0 * (x * z + w * y));
    m[3] = (float)(2.0 * (x * y + w * z)); m[4] = (float)(w * w - x * x + y * y - z * z); m[5] = (float)(2.0 * (y * z - w * x));
    m[6] = (float)(2.0 * (x * z - w * y)); m[7] = (float)(2.0 * (y * z + w * x)); m[8] = (float)(w * w - x * x - y * y + z * z);
}
#endif
__global__ void pack_kernel(const FrameParams P, uint8_t* dst) {
    const int tid = threadIdx.x, nt = blockDim.x;
    PPlane* planes = (PPlane*)(dst + P.lay.off_plane);
    PSphere* spheres = (PSphere*)(dst + P.lay.off_sphere);
    HSurf* surfs = (HSurf*)(dst + P.lay.off_surf);
    HBox* boxes = (HBox*)(dst + P.lay.off_box);
    HTorus* tori = (HTorus*)(dst + P.lay.off_torus);
    HRing* rings = (HRing*)(dst + P.lay.off_ring);
    PLight* lights = (PLight*)(dst + P.lay.off_light);
    for (int i = tid; i < P.n_plane; i += nt) {
        const rtb_plane& s = P.planes[i];
        PPlane p = { s.normal[0], s.normal[1], s.normal[2], 0.f, s.pos[0], s.pos[1], s.pos[2], 0.f };
        planes[i] = p;
    }
    for (int i = tid; i < P.n_sphere; i += nt) {
        const rtb_sphere& s = P.spheres[i];
        float r2 = s.obj[3] * s.obj[3];                      /* >= +0 (or NaN): the sign bit is free to carry `hollow` */
        PSphere p = { s.obj[0], s.obj[1], s.obj[2], s.hollow != 0 ? __int_as_float(__float_as_int(r2) | 0x80000000) : r2 };
        spheres[i] = p;
    }
    for (int i = tid; i < P.n_lpoint; i += nt) {
        const rtb_light_point& s = P.lights_point[i];
        PLight p = { s.pos[0], s.pos[1], s.pos[2], s.pos[3] * s.pos[3] };
        lights[i] = p;
    }
#if RTB_STRICT
    for (int i = tid; i < P.n_surf; i += nt) {
        const rtb_surface& s = P.surfaces[i];
        PSurf p = { s.quat_rotation[0], s.quat_rotation[1], s.quat_rotation[2], s.quat_rotation[3], s.pos[0], s.pos[1], s.pos[2],
                    s.a, s.b, s.c, s.d, s.e, s.f, s.v_min[0], s.v_min[1], s.v_min[2], s.v_max[0], s.v_max[1], s.v_max[2], 0.f };
        surfs[i] = p;
    }
    for (int i = tid; i < P.n_box; i += nt) {
        const rtb_box& s = P.boxes[i];
        PBox p = { s.quat_rotation[0], s.quat_rotation[1], s.quat_rotation[2], s.quat_rotation[3], s.pos[0], s.pos[1], s.pos[2],
                   s.form[0], s.form[1], s.form[2], s.textureNum, 0 };
        boxes[i] = p;
    }
    for (int i = tid; i < P.n_torus; i += nt) {
        const rtb_torus& s = P.toruses[i];
        PTorus p = { s.quat_rotation[0], s.quat_rotation[1], s.quat_rotation[2], s.quat_rotation[3], s.pos[0], s.pos[1], s.pos[2],
                     s.form[0] * s.form[0], s.form[1] * s.form[1], 0.f, 0.f, 0.f };
        tori[i] = p;
    }
    for (int i = tid; i < P.n_ring; i += nt) {
        const rtb_ring& s = P.rings[i];
        PRing p = { s.quat_rotation[0], s.quat_rotation[1], s.quat_rotation[2], s.quat_rotation[3], s.pos[0], s.pos[1], s.pos[2],
                    s.r1, s.r2, s.textureNum, 0, 0 };
        rings[i] = p;
    }
#else
    for (int i = tid; i < P.n_surf; i += nt) {
        const rtb_surface& s = P.surfaces[i];
        PSurfM p;
        quat_matrix(s.quat_rotation, p.m);
        p.px = s.pos[0]; p.py = s.pos[1]; p.pz = s.pos[2];
        p.a = s.a; p.b = s.b; p.c = s.c; p.d = s.d; p.e = s.e; p.f = s.f;
        p.minx = s.v_min[0]; p.miny = s.v_min[1]; p.minz = s.v_min[2]; p.maxx = s.v_max[0]; p.maxy = s.v_max[1]; p.maxz = s.v_max[2];
        surfs[i] = p;
    }
    for (int i = tid; i < P.n_box; i += nt) {
        const rtb_box& s = P.boxes[i];
        PBoxM p;
        quat_matrix(s.quat_rotation, p.m);
        p.px = s.pos[0]; p.py = s.pos[1]; p.pz = s.pos[2]; p.fx = s.form[0]; p.fy = s.form[1]; p.fz = s.form[2]; p.tex = s.textureNum;
        boxes[i] = p;
    }
    for (int i = tid; i < P.n_torus; i += nt) {
        const rtb_torus& s = P.toruses[i];
        PTorusM p;
        quat_matrix(s.quat_rotation, p.m);
        p.px = s.pos[0]; p.py = s.pos[1]; p.pz = s.pos[2];
        p.R2 = s.form[0] * s.form[0]; p.r2 = s.form[1] * s.form[1]; p.k = p.R2 - p.r2; p.fourR2 = 4.f * p.R2;
        tori[i] = p;
    }
    for (int i = tid; i < P.n_ring; i += nt) {
        const rtb_ring& s = P.rings[i];
        PRingM p;
        quat_matrix(s.quat_rotation, p.m);
        p.px = s.pos[0]; p.py = s.pos[1]; p.pz = s.pos[2]; p.r1 = s.r1; p.r2 = s.r2; p.tex = s.textureNum; p._0 = 0;
        rings[i] = p;
    }
#endif
}

}  // namespace RTB_NS

#include "rt_persistent.cuh"

/* ------------------------------------------------------------------ host launchers (one set per build flavour) */
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

extern "C" int CAT(RTB_NS, _launch_pack)(const FrameParams* P, uint8_t* dst, cudaStream_t st) {
    RTB_NS::pack_kernel<<<1, 256, 0, st>>>(*P, dst);
    return (int)cudaGetLastError();
}

extern "C" int CAT(RTB_NS, _launch)(const FrameParams* P, int kernel, int counted, int grid, int threads, size_t smem, cudaStream_t st) {
    cudaError_t e;
    if (kernel == RTB_LAUNCH_QUAD) {
        auto k = counted ? RTB_NS::quad_kernel<true> : RTB_NS::quad_kernel<false>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        k<<<grid, QUAD_THREADS, smem, st>>>(*P);
    } else if (kernel == RTB_LAUNCH_PERSISTENT_WIDE) {
        auto k = counted ? RTB_NS::persistent_kernel<true, PERSIST_THREADS_WIDE> : RTB_NS::persistent_kernel<false, PERSIST_THREADS_WIDE>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        k<<<grid, PERSIST_THREADS_WIDE, smem, st>>>(*P);
    } else {
        auto k = counted ? RTB_NS::persistent_kernel<true, PERSIST_THREADS> : RTB_NS::persistent_kernel<false, PERSIST_THREADS>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        k<<<grid, PERSIST_THREADS, smem, st>>>(*P);
    }
    (void)threads;
    return (int)cudaGetLastError();
}

extern "C" int CAT(RTB_NS, _occupancy)(int kernel, size_t smem, int* blocks_per_sm) {
    cudaError_t e;
    if (kernel == RTB_LAUNCH_QUAD) {
        e = cudaFuncSetAttribute(RTB_NS::quad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, RTB_NS::quad_kernel<false>, QUAD_THREADS, smem);
    } else if (kernel == RTB_LAUNCH_PERSISTENT_WIDE) {
        e = cudaFuncSetAttribute(RTB_NS::persistent_kernel<false, PERSIST_THREADS_WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, RTB_NS::persistent_kernel<false, PERSIST_THREADS_WIDE>, PERSIST_THREADS_WIDE, smem);
    } else {
        e = cudaFuncSetAttribute(RTB_NS::persistent_kernel<false, PERSIST_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, RTB_NS::persistent_kernel<false, PERSIST_THREADS>, PERSIST_THREADS, smem);
    }
    return (int)e;
}

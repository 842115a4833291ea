/* rt_persistent.cuh — persistent-threads kernel for scenes without 2-D textures.
 *
 * ONE 20-warp CTA per SM lives for the whole frame.  Every LANE owns one
 * pixel's path at a time and runs a small job machine whose only expensive
 * state is "scan the whole scene with this ray":
 *
 *      MAIN    calcInter for the path's current ray            (rt.frag:823)
 *      SUB     calcInter inside getReflectedColor              (rt.frag:792)
 *      SHADOW  inShadow for light l of the pending calcShade   (rt.frag:667)
 *
 * Steady state: the warp executes ONE unified scan per loop trip in which each
 * lane has its own ray and its own mode (nearest / shadow); between scans each
 * lane post-processes its result (hit attributes, Fresnel, Phong accumulation)
 * and posts its next job.  A lane whose path ended takes a fresh pixel from the
 * frame's atomic counter immediately (ballot-compacted, one atomicAdd per
 * warp), so the scan — >95 % of the work — runs with full warps no matter how
 * differently deep neighbouring pixels bounce.
 *
 * Drain: once the counter has passed the last pixel, lanes fall idle one by
 * one.  From then on all warps of the CTA pool the scans of their live paths
 * in shared memory and serve them one ray per warp (coop_scan, rt_scan.cuh),
 * so the SM stays busy until its last path ends.
 *
 * Per-pixel arithmetic and its order are untouched in both phases: results
 * equal the quad kernel's bit for bit.
 */
#pragma once
#include "rt_scan.cuh"
/* GATE (rt_scan.cuh): skip the tests of lanes without a live path.  Idle lanes exist only in the last trips of a frame (the cooperative
 * drain takes over), so the 4-instruction branch region per test is dropped: idle lanes re-scan their last ray and the result is
 * dropped (fused build: mixed1024@4K 246.4 -> 240.1 ms, spheres4k 19.5 -> 18.4 ms; strict build, together with the shared-window
 * addressing: 388.5 -> 382.9 ms, 31.15 -> 28.85 ms; profiles/r2_strict_gate_saddr_ab.jsonl). */
#ifndef RTB_PERSIST_GATE
#define RTB_PERSIST_GATE false
#endif

namespace RTB_NS {

enum { JOB_MAIN = 0, JOB_SUB = 1, JOB_SHADOW = 2 };
enum { COMB_REFRACT_SUB = 0, COMB_REFLECT = 1, COMB_DIFFUSE = 2 };

/* Drain area in shared memory (behind the staged scene): one job and one result slot per thread. */
struct DrainJob { float ro[3], rd[3], limit; int shadow; };             /* 32 B */
struct DrainResult { float tm; int id; float shadow, u, v; int _pad[3]; };   /* 32 B */

/* THREADS: 640 (20 warps, 96 registers: what the Durand-Kerner solve needs without spilling into its loop) or PERSIST_THREADS_WIDE
 * (24 warps, 80 registers) for scenes without tori, which are bound by the sphere / box / quadric tests and gain from the extra warps. */
template <bool COUNT, int THREADS>
__global__ void __launch_bounds__(THREADS, PERSIST_MIN_BLOCKS) persistent_kernel(const __grid_constant__ FrameParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ int drain_flag;                                  /* 0 -> 1 once; read and written with atomics only */
    __shared__ int n_jobs[2], next_job[2];
    if (threadIdx.x == 0) { drain_flag = 0; n_jobs[0] = n_jobs[1] = 0; next_job[0] = next_job[1] = 0; }
    if (P.cta_times && threadIdx.x == 0) P.cta_times[blockIdx.x * 5] = globaltimer_ns();
    stage_scene_tma(smem, &mbar, P.packed, P.lay.total_bytes);          /* contains the __syncthreads that publishes the zeros above */
    const SceneView S = make_view(smem, P.lay);
    DrainJob* const jobs = (DrainJob*)(smem + ((P.lay.total_bytes + 127u) & ~127u));
    DrainResult* const results = (DrainResult*)(jobs + THREADS);

    const int lane = threadIdx.x & 31;
    const unsigned total = (unsigned)(P.n_tiles_x * P.n_tiles_y) * 32u;
    const int n_lights = P.n_lpoint + P.n_ldirect;
    Counters cnt = {};

    /* path state (px: the tile of the lane's pixel, -1 = no path) */
    int px = -1, fb_off = 0;
    vec3 ro = mk3(0, 0, 0), rd = mk3(0, 0, 1), mask = mk3(1, 1, 1), color = mk3(0, 0, 0);
    float absorbDistance = 0.f;
    int it = 0, glass = 0;
    /* job */
    int job = JOB_MAIN, jl = 0;
    vec3 jro = ro, jrd = rd;
    float jlimit = MAX_DIST;
    /* pending calcShade */
    vec3 s_pt = ro, s_rd = rd, s_n = rd, dif = color, spec = color, cmask = mask;
    int s_id = 0, cmode = 0;
    float cs = 0.f;
    bool cont = false;
    bool exhausted = false;
    int parity = 0;
    unsigned dbg_jobs = 0, dbg_trips = 0;

    for (;;) {
        /* ---- refill idle lanes with fresh pixels ---- */
        unsigned want = __ballot_sync(FULL, px < 0);
        if (want && !exhausted) {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(P.tile_counter, (unsigned)__popc(want));
            base = __shfl_sync(FULL, base, 0);
            if (px < 0) {
                unsigned idx = base + __popc(want & ((1u << lane) - 1u));
                if (idx < total) {
                    int tile = idx >> 5, l = idx & 31;
                    if (P.tile_perm) tile = (int)__ldg(P.tile_perm + tile);     /* costliest tiles of the previous frame first (rtb_api.cu) */
                    int tx = tile % P.n_tiles_x, ty = tile / P.n_tiles_x;
                    int qd = l >> 2;
                    int x = tx * 8 + (qd & 3) * 2 + (l & 1);
                    int ly = ty * 4 + (qd >> 2) * 2 + ((l >> 1) & 1);
                    int y = global_row(P, ly);
                    if (x < P.canvas_w && ly < P.local_rows && y < P.canvas_h) {
                        fb_off = (P.fb_global ? y : ly) * P.canvas_w + x;
                        mask = mk3(1.f, 1.f, 1.f); color = mk3(0.f, 0.f, 0.f);
                        ro = mk3(P.cam_pos[0], P.cam_pos[1], P.cam_pos[2]);
                        rd = getRayDir(P, x, y);
                        absorbDistance = 0.f; it = 0; glass = 0;
                        if (COUNT) cnt.pixels++;
                        if (P.iterations > 0) { px = tile; job = JOB_MAIN; jro = ro; jrd = rd; jlimit = MAX_DIST; }
                        else *(float4*)(P.fb + (size_t)fb_off * 4) = make_float4(0.f, 0.f, 0.f, 1.f);
                    }
                }
            }
            if (base + (unsigned)__popc(want) >= total) {       /* warp-uniform: the counter passed the last pixel */
                exhausted = true;
                if (lane == 0 && atomicExch(&drain_flag, 1) == 0 && P.cta_times) P.cta_times[blockIdx.x * 5 + 1] = globaltimer_ns();
            }
        }
        if (!exhausted) {                                       /* did some other warp of this CTA see the end of the frame?  then no more refills */
            int f = 0;
            if (lane == 0) f = atomicAdd(&drain_flag, 0);
            exhausted = __shfl_sync(FULL, f, 0) != 0;
        }
        const bool active = px >= 0;
        float tm, shadow; int id; vec2 ruv;

        if (!exhausted || !P.coop) {
            if (!__any_sync(FULL, active)) { if (exhausted) break; else continue; }     /* (break: only with the cooperative drain switched off) */
            /* ---- steady state: one unified scene scan per warp, each lane its own ray and mode ---- */
            scan_scene<COUNT, false, RTB_PERSIST_GATE>(P, S, jro, jrd, active, job == JOB_SHADOW, jlimit, 0, tm, id, shadow, ruv, cnt);
        } else {
            /* ---- drain: the frame has no fresh pixels left, lanes fall idle one by one.  All warps of the CTA pool
             * the scans of their live paths in shared memory and serve them one ray per warp (coop_scan), so the
             * SM stays busy until its last path ends instead of running full-length scans for a few lanes. ---- */
            int slot = -1;
            if (active) {
                slot = atomicAdd(&n_jobs[parity], 1);
                DrainJob j = { { jro.x, jro.y, jro.z }, { jrd.x, jrd.y, jrd.z }, jlimit, job == JOB_SHADOW ? 1 : 0 };
                jobs[slot] = j;
            }
            __syncthreads();
            const int n = n_jobs[parity];
            if (n == 0) break;                                  /* CTA-uniform: every path of this CTA has ended */
            dbg_jobs += n; dbg_trips++;
            if (threadIdx.x == 0) { n_jobs[parity ^ 1] = 0; next_job[parity ^ 1] = 0; }   /* nobody touches the other set before the next barrier */
            for (;;) {
                int j = 0;
                if (lane == 0) j = atomicAdd(&next_job[parity], 1);
                j = __shfl_sync(FULL, j, 0);
                if (j >= n) break;
                const DrainJob J = jobs[j];                     /* broadcast loads */
                const vec3 bro = mk3(J.ro[0], J.ro[1], J.ro[2]), brd = mk3(J.rd[0], J.rd[1], J.rd[2]);
                float r_tm, r_sh; int r_id; vec2 r_uv;
                coop_scan<COUNT>(P, S, bro, brd, J.shadow != 0, J.limit, r_tm, r_id, r_sh, r_uv, cnt);
#ifdef RTB_DEBUG_COOP_CHECK
                {   /* development: redo the ray with the serial scan on lane 0 and report any difference */
                    float s_tm, s_sh; int s_id; vec2 s_uv; Counters scratch = {};
                    scan_scene<false, false, true>(P, S, bro, brd, lane == 0, J.shadow != 0, J.limit, 0, s_tm, s_id, s_sh, s_uv, scratch);
                    if (lane == 0 && (s_tm != r_tm || s_id != r_id || s_sh != r_sh))
                        printf("COOP MISMATCH shadow=%d limit=%g ro=(%.9g %.9g %.9g) rd=(%.9g %.9g %.9g): coop tm=%.9g id=%x sh=%g | serial tm=%.9g id=%x sh=%g\n",
                               J.shadow, J.limit, bro.x, bro.y, bro.z, brd.x, brd.y, brd.z, r_tm, r_id, r_sh, s_tm, s_id, s_sh);
                }
#endif
                if (lane == 0) { DrainResult r = { r_tm, r_id, r_sh, r_uv.x, r_uv.y, { 0, 0, 0 } }; results[j] = r; }
            }
            __syncthreads();
            parity ^= 1;
            if (active) {
                const DrainResult r = results[slot];
                tm = r.tm; id = r.id; shadow = r.shadow; ruv = mk2(r.u, r.v);
                if (COUNT) { if (job == JOB_SHADOW) cnt.rays_s++; else cnt.rays_n++; }
            }
        }
        if (!active) continue;

        /* ---- per-lane post-processing ---- */
        bool start_shade = false, finish_shade = false, path_done = false;
        if (job == JOB_SHADOW) {
            LightSample L = light_sample(P, jl, s_pt);
            const rtb_material* m = (const rtb_material*)nullptr;
            {
                int type = id_type(s_id), num = id_num(s_id);
                m = type == RTB_TYPE_SPHERE ? &P.spheres[num].material : type == RTB_TYPE_PLANE ? &P.planes[num].material
                  : type == RTB_TYPE_SURFACE ? &P.surfaces[num].mat : type == RTB_TYPE_BOX ? &P.boxes[num].mat
                  : type == RTB_TYPE_TORUS ? &P.toruses[num].mat : &P.rings[num].mat;
            }
            Material sm = load_material(m);
            if (COUNT) cnt.light_evals++;
            shade_light(P, L, shadow, s_rd, sm.color, sm.diffuse, sm.specular, s_n, dif, spec);
            jl++;
            if (jl < n_lights) { LightSample Ln = light_sample(P, jl, s_pt); jrd = Ln.dir_n; jlimit = Ln.dist; }
            else {
                vec3 pixelColor = mk3(P.ambient[0], P.ambient[1], P.ambient[2]) * sm.color;
                pixelColor = pixelColor + (dif * sm.kd + spec * sm.ks);
                if (cmode == COMB_DIFFUSE) color = color + pixelColor * cmask * cs;
                else color = color + pixelColor * cs * cmask;
                finish_shade = true;
            }
        } else {
            const bool is_main = job == JOB_MAIN;
            const vec3 cro = jro, crd = jrd;
            const bool hit = tm < MAX_DIST;
            const bool light = hit && id_type(id) == RTB_TYPE_POINT_LIGHT;
            if (is_main && !hit) {                              /* rt.frag:892-895 */
                color = color + texture_cube(P.cube, crd) * mask;
                path_done = true;
            } else if (is_main && light) {                      /* rt.frag:829-832 (get_hit_info runs first: counted) */
                if (COUNT) cnt.shaded[RTB_TYPE_POINT_LIGHT]++;
                const float* lc = P.lights_point[id_num(id)].color;
                color = color + mk3(lc[0], lc[1], lc[2]) * mask;
                path_done = true;
            } else if (!is_main && (!hit || light)) {           /* getReflectedColor returned the light colour or 0 */
                vec3 rc = mk3(0.f, 0.f, 0.f);
                if (light) { const float* lc = P.lights_point[id_num(id)].color; rc = mk3(lc[0], lc[1], lc[2]); }
                color = color + rc * cs * cmask;
                finish_shade = true;
            } else {
                vec3 pt = cro + crd * tm;
                Material mat; vec3 n; float alpha, bias;
                hit_info<COUNT, false>(P, true, id, cro, crd, pt, tm, ruv, mat, n, alpha, bias, cnt);
                if (!is_main) {                                 /* rt.frag:796-799 */
                    s_pt = dot(crd, n) < 0 ? pt + n * bias : pt - n * bias;
                    s_rd = crd; s_n = n; s_id = id;
                    start_shade = true;
                } else {
                    bool outside = dot(rd, n) < 0;
                    n = outside ? n : -n;
                    float reflectMultiplier;
                    if (mat.refraction > 0)
                        reflectMultiplier = FresnelReflectAmount(outside ? 1 : mat.refraction, outside ? mat.refraction : 1, rd, n, mat.reflection);
                    else
                        reflectMultiplier = getFresnel(n, rd, mat.reflection);
                    float refractMultiplier = 1 - reflectMultiplier;
                    if (mat.refraction > 0.0f) {                /* rt.frag:851-873 */
                        bool sub = outside && mat.reflection > 0;
                        vec3 sro = pt + n * bias, srd = reflect(rd, n);
                        if (sub) { cmode = COMB_REFRACT_SUB; cs = reflectMultiplier; cmask = mask; mask = mask * refractMultiplier; }
                        else if (!outside) {
                            absorbDistance += tm;
                            vec3 a = -mat.absorb * absorbDistance;
                            mask = mask * mk3(expf(a.x), expf(a.y), expf(a.z));
                        }
                        cont = true;
                        if (reflectMultiplier >= 1) cont = false;
                        else {
                            ro = pt - n * bias;
                            rd = refract(rd, n, outside ? 1 / mat.refraction : mat.refraction);
                            it--;
                            if (++glass >= MAX_GLASS_EVENTS) cont = false;
                        }
                        if (sub) { job = JOB_SUB; jro = sro; jrd = srd; jlimit = MAX_DIST; }
                        else finish_shade = true;               /* nothing to shade: just continue / stop */
                    } else if (mat.reflection > 0.0f) {         /* rt.frag:874-880 */
                        ro = pt + n * bias;
                        s_pt = ro; s_rd = rd; s_n = n; s_id = id;
                        cmode = COMB_REFLECT; cs = refractMultiplier; cmask = mask;
                        rd = reflect(rd, n);
                        mask = mask * reflectMultiplier;
                        cont = true;
                        start_shade = true;
                    } else {                                    /* rt.frag:881-890 */
                        s_pt = pt + n * bias; s_rd = rd; s_n = n; s_id = id;
                        cmode = COMB_DIFFUSE; cs = alpha; cmask = mask;
                        if (alpha < 1) { ro = pt - n * bias; mask = mask * (1 - alpha); cont = true; }
                        else cont = false;
                        start_shade = true;
                    }
                }
            }
        }
        if (start_shade) {
            dif = mk3(0.f, 0.f, 0.f); spec = mk3(0.f, 0.f, 0.f);
            if (n_lights > 0) {
                job = JOB_SHADOW; jl = 0; jro = s_pt;
                LightSample L = light_sample(P, 0, s_pt);
                jrd = L.dir_n; jlimit = L.dist;
            } else {
                int type = id_type(s_id), num = id_num(s_id);
                const rtb_material* m = type == RTB_TYPE_SPHERE ? &P.spheres[num].material : type == RTB_TYPE_PLANE ? &P.planes[num].material
                  : type == RTB_TYPE_SURFACE ? &P.surfaces[num].mat : type == RTB_TYPE_BOX ? &P.boxes[num].mat
                  : type == RTB_TYPE_TORUS ? &P.toruses[num].mat : &P.rings[num].mat;
                vec3 pixelColor = mk3(P.ambient[0], P.ambient[1], P.ambient[2]) * load_material(m).color;
                pixelColor = pixelColor + (dif * load_material(m).kd + spec * load_material(m).ks);
                if (cmode == COMB_DIFFUSE) color = color + pixelColor * cmask * cs;
                else color = color + pixelColor * cs * cmask;
                finish_shade = true;
            }
        }
        if (finish_shade) {                                     /* end of one loop trip of rt.frag:821 */
            if (cont) { it++; if (it >= P.iterations) path_done = true; }
            else path_done = true;
            if (!path_done) { job = JOB_MAIN; jro = ro; jrd = rd; jlimit = MAX_DIST; }
        }
        if (path_done) {
            *(float4*)(P.fb + (size_t)fb_off * 4) = make_float4(color.x, color.y, color.z, 1.0f);
            if (P.tile_cost) atomicAdd(P.tile_cost + px, (unsigned)(it + glass + 1));   /* this path's length: the tile's cost for the next frame's order */
            px = -1;
        }
    }
    if (P.cta_times && threadIdx.x == 0) { P.cta_times[blockIdx.x * 5 + 2] = globaltimer_ns(); P.cta_times[blockIdx.x * 5 + 3] = ((unsigned long long)dbg_trips << 32) | dbg_jobs; }
    if (COUNT) flush_counters(P, cnt);
}

}  // namespace RTB_NS

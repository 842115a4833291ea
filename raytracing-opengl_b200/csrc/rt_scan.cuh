/* rt_scan.cuh — the linear scene scan and the hit-attribute stage, shared by the
 * quad kernel and the persistent kernel.
 *
 * scan_scene() is calcInter (rt.frag:587-628) and inShadow (rt.frag:630-658) in
 * one loop nest: every lane of the warp walks the same primitive (a broadcast
 * LDS.128 from the TMA-staged shared-memory block) with its own ray, and each
 * lane is independently in NEAREST mode (strict t<tmin update, type order
 * planes -> spheres -> quadrics -> boxes -> tori -> rings -> light spheres) or
 * in SHADOW mode (fixed limit `dist`, no early exit, hollow forced false,
 * planes and light spheres skipped because PLANE_ONESIDE==1).
 */
#pragma once
#include "rt_device.cuh"
#ifndef RTB_SCALAR_DK
#define RTB_SCALAR_DK 0
#endif
/* tests per trip of the quadric and box loops (A/B switch; `#pragma unroll` takes no macro, hence _Pragma) */
#ifndef RTB_SCAN_UNROLL
#define RTB_SCAN_UNROLL 1
#endif
#define RTB_STR_(x) #x
#define RTB_STR(x) RTB_STR_(x)
#define RTB_UNROLL_SCAN _Pragma(RTB_STR(unroll RTB_SCAN_UNROLL))

namespace RTB_NS {

struct Counters { unsigned rays_n, rays_s, dk, light_evals, pixels; unsigned shaded[7]; };

DEV int make_id(int type, int num) { return (type << 24) | num; }
DEV int id_type(int id) { return id >> 24; }
DEV int id_num(int id) { return id & 0xffffff; }

/* derivative exchange across the 2x2 quad (pin Q9): lanes l^1 / l^2 are the x / y
 * neighbours; a neighbour contributes only if it is at the same site (tag equal). */
struct Derivs { float dudx, dvdx, dudy, dvdy; };
DEV Derivs quad_derivs(bool site, int tag, vec2 uv) {
    int lane = threadIdx.x & 31;
    int mytag = site ? tag : -1;
    float ux = __shfl_xor_sync(FULL, uv.x, 1), vx = __shfl_xor_sync(FULL, uv.y, 1);
    int tx = __shfl_xor_sync(FULL, mytag, 1);
    float uy = __shfl_xor_sync(FULL, uv.x, 2), vy = __shfl_xor_sync(FULL, uv.y, 2);
    int ty = __shfl_xor_sync(FULL, mytag, 2);
    Derivs d = { 0.f, 0.f, 0.f, 0.f };
    if (site && tx == tag) {
        if (lane & 1) { d.dudx = uv.x - ux; d.dvdx = uv.y - vx; } else { d.dudx = ux - uv.x; d.dvdx = vx - uv.y; }
    }
    if (site && ty == tag) {
        if (lane & 2) { d.dudy = uv.x - uy; d.dvdy = uv.y - vy; } else { d.dudy = uy - uv.x; d.dvdy = vy - uv.y; }
    }
    return d;
}

/* TEX: 2-D textures may be referenced (quad kernel): textured rings add their alpha to
 * the shadow term (rt.frag:644-651) and need the quad exchange -> scan_scene must then be
 * called from warp-uniform control flow.  ctx = 0 main path / 1 getReflectedColor.
 * GATE: lanes with active == false skip the tests (quad kernel: dead paths and helper lanes are common; persistent kernel: see
 * RTB_PERSIST_GATE in rt_persistent.cuh). */
template <bool COUNT, bool TEX, bool GATE>
DEV void scan_scene(const FrameParams& P, const SceneView& S, vec3 ro, vec3 rd, bool active, bool shadow_mode, float limit, int ctx,
                    float& tmin_out, int& id_out, float& shadow_out, vec2& ring_uv_out, Counters& cnt) {
    float tmin = limit;
    int id = -1;
    float shadow = 0.f;
    vec2 ring_uv = mk2(0.f, 0.f);
    float t;
    const PackK K = make_packk(P);
    const bool on = GATE ? active : true;
    if (COUNT && active) { if (shadow_mode) cnt.rays_s++; else cnt.rays_n++; }

    for (int i = 0; i < P.n_plane; i++) {
        if (on && !shadow_mode) {
            float4 a = lds4(S.planes + i, 0), b = lds4(S.planes + i, 1);
            if (intersectPlane(ro, rd, mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), tmin, t)) { tmin = t; id = make_id(RTB_TYPE_PLANE, i); }
        }
    }
    {   /* spheres, four discriminants at a time: most rays miss most spheres, so the sqrt stage is rare */
        const bool may_hollow = !shadow_mode;                   /* inShadow passes hollow = false, rt.frag:636 */
        int i = 0;
        for (; i + 4 <= P.n_sphere; i += 4) {
            float4 o0 = lds4(S.spheres, i), o1 = lds4(S.spheres, i + 1), o2 = lds4(S.spheres, i + 2), o3 = lds4(S.spheres, i + 3);
            float b0, b1, b2, b3;
            float h0 = sphere_disc(ro, rd, o0, b0), h1 = sphere_disc(ro, rd, o1, b1);
            float h2 = sphere_disc(ro, rd, o2, b2), h3 = sphere_disc(ro, rd, o3, b3);
            if (on && !(h0 < 0.f && h1 < 0.f && h2 < 0.f && h3 < 0.f)) {
#define RTB_SPHERE_FINISH(K, O, B, H)                                                                                     \
                if (!(H < 0.f) && sphere_finish(B, H, may_hollow && __float_as_int(O.w) < 0, tmin, t)) {                 \
                    if (shadow_mode) shadow = 1.f; else { tmin = t; id = make_id(RTB_TYPE_SPHERE, i + K); }               \
                }
                RTB_SPHERE_FINISH(0, o0, b0, h0) RTB_SPHERE_FINISH(1, o1, b1, h1) RTB_SPHERE_FINISH(2, o2, b2, h2) RTB_SPHERE_FINISH(3, o3, b3, h3)
            }
        }
        for (; i < P.n_sphere; i++) {
            float4 o0 = lds4(S.spheres, i);
            float b0, h0 = sphere_disc(ro, rd, o0, b0);
            if (on) { RTB_SPHERE_FINISH(0, o0, b0, h0) }
        }
#undef RTB_SPHERE_FINISH
    }
    RTB_UNROLL_SCAN
    for (int i = 0; i < P.n_surf; i++) {
        if (on) {
            if (intersectSurface(K, ro, rd, S.surfs + i, tmin, t)) {
                if (shadow_mode) shadow = 1.f; else { tmin = t; id = make_id(RTB_TYPE_SURFACE, i); }
            }
        }
    }
    RTB_UNROLL_SCAN
    for (int i = 0; i < P.n_box; i++) {
        if (on) {
            if (intersectBox(K, ro, rd, S.boxes + i, tmin, t)) {
                if (shadow_mode) shadow = 1.f; else { tmin = t; id = make_id(RTB_TYPE_BOX, i); }
            }
        }
    }
    for (int i = 0; i < P.n_torus; i++) {
        if (on) {
            /* intersectTorus, rt.frag:462-487.  Lock-step over the warp keeps 26 of 32 lanes busy in the trips.  Tried: a capped loop +
             * straggler queue (measured slower), and a per-lane schedule through a shared-memory ring (every lane walks its own chain of
             * solves): its store-roots / load-next block and loop control cost ~80 issue cycles per trip in the SASS, more than the idle
             * lanes it removes in an issue-bound kernel (tests/dev/dk_schedule_sim.py) */
            TorusState st;
            if (torus_setup(K, ro, rd, S.tori + i, P.cull, st)) {
                int iters;
                t = RTB_SCALAR_DK ? torus_solve_scalar(st, iters) : torus_solve<!TEX>(K, st, iters);
                if (COUNT && active) cnt.dk += iters;
                if (t > 0 && t < 100 && t < tmin) {
                    if (shadow_mode) shadow = 1.f; else { tmin = t; id = make_id(RTB_TYPE_TORUS, i); }
                }
            }
        }
    }
    for (int i = 0; i < P.n_ring; i++) {
        vec2 uv = mk2(0.f, 0.f);
        bool hit = on && intersectRing(K, ro, rd, S.rings + i, tmin, t, uv);
        int tex = ldsi(S.rings + i, offsetof(HRing, tex));
        if (hit && !shadow_mode) { tmin = t; id = make_id(RTB_TYPE_RING, i); ring_uv = uv; }
        if (TEX && tex > 0) {
            bool site = hit && shadow_mode;
            if (__any_sync(FULL, site)) {
                Derivs d = quad_derivs(site, ctx, uv);
                if (site) {
                    float lod = implicit_lod(P.tex[4], d.dudx, d.dvdx, d.dudy, d.dvdy);
                    shadow += texture_lod(P.tex[4], uv.x, uv.y, lod).w;
                }
            }
        } else if (hit && shadow_mode) {
            shadow = 1.f;
        }
    }
    for (int i = 0; i < P.n_lpoint; i++) {
        if (on && !shadow_mode) {
            float4 o = lds4(S.lights, i);
            if (intersectSphere(ro, rd, o, false, tmin, t)) { tmin = t; id = make_id(RTB_TYPE_POINT_LIGHT, i); }
        }
    }
    tmin_out = tmin; id_out = id; shadow_out = gmin(shadow, 1.f); ring_uv_out = ring_uv;
}

/* coop_scan: the same calcInter / inShadow, for ONE ray held by the whole warp (ro, rd, mode, limit warp-uniform):
 * lane l tests primitives l, l+32, ... of every class and the warp reduces.  Used by the persistent kernel while
 * the frame drains (rt_persistent.cuh).  Equality with the serial scan: every accept test but one is `t < tmin`
 * with the running minimum, so the serial result is the lexicographic minimum of (t, scan position) over all
 * candidates, which is what the reductions compute (scan position = class rank planes < spheres < quadrics <
 * boxes < tori < rings < lights, then index).  Two exceptions, both resolved 32 primitives at a time against the
 * warp-uniform running minimum and replayed in index order (warp shuffles) when they occur: the degenerate-quadric
 * quirk (accepts t > tmin, rt.frag:541-545) and boxes whose slab test yields NaN (accepted by `!(tN >= tmin)`,
 * rt.frag:417-423; after that tmin is NaN and no later class can be accepted, exactly as in the serial scan).
 * Shadow mode is an any-hit against the fixed limit: OR over lanes.  No 2-D textures here (persistent kernel). */
DEV int scan_rank(int type) {       /* rtb_prim_type -> position of the class in calcInter's order */
    return type == RTB_TYPE_PLANE ? 0 : type == RTB_TYPE_SPHERE ? 1 : type;
}
DEV int scan_key(int id) { return id < 0 ? 0x7fffffff : (scan_rank(id_type(id)) << 24) | id_num(id); }
/* lexicographic minimum of (t, key) over the warp; every lane gets the winner's (t, id) */
DEV void warp_nearest(float& t, int& id) {
    float bt = t; int bk = scan_key(id), bid = id;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ot = __shfl_xor_sync(FULL, bt, o);
        int ok = __shfl_xor_sync(FULL, bk, o), oid = __shfl_xor_sync(FULL, bid, o);
        if (ok != 0x7fffffff && (bk == 0x7fffffff || ot < bt || (ot == bt && ok < bk))) { bt = ot; bk = ok; bid = oid; }
    }
    t = bt; id = bid;
}
template <bool COUNT>
DEV void coop_scan(const FrameParams& P, const SceneView& S, vec3 ro, vec3 rd, bool shadow_mode, float limit,
                   float& tmin_out, int& id_out, float& shadow_out, vec2& ring_uv_out, Counters& cnt) {
    const int lane = threadIdx.x & 31;
    const PackK K = make_packk(P);
    float tmin = limit, t;
    int id = -1;
    bool occluded = false;
    vec2 ring_uv = mk2(0.f, 0.f);
    if (!shadow_mode)
        for (int i = lane; i < P.n_plane; i += 32) {
            float4 a = lds4(S.planes + i, 0), b = lds4(S.planes + i, 1);
            if (intersectPlane(ro, rd, mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), tmin, t)) { tmin = t; id = make_id(RTB_TYPE_PLANE, i); }
        }
    for (int i = lane; i < P.n_sphere; i += 32) {
        float4 o = lds4(S.spheres, i);
        if (intersectSphere(ro, rd, o, !shadow_mode && __float_as_int(o.w) < 0, tmin, t)) {
            if (shadow_mode) occluded = true; else { tmin = t; id = make_id(RTB_TYPE_SPHERE, i); }
        }
    }
    if (shadow_mode) {
        for (int i = lane; i < P.n_surf; i += 32)
            if (intersectSurface(K, ro, rd, S.surfs + i, tmin, t)) occluded = true;
    } else if (P.n_surf > 0) {
        warp_nearest(tmin, id);                                 /* the running minimum after planes and spheres, on every lane */
        for (int base = 0; base < P.n_surf; base += 32) {
            const int i = base + lane;
            int kind = 0;
            t = 0.f;
            if (i < P.n_surf) kind = surface_candidate(K, ro, rd, S.surfs + i, t);
            if (!__any_sync(FULL, kind == 2)) {
                float ct = surface_accept(kind, t, tmin) ? t : tmin;
                int cid = surface_accept(kind, t, tmin) ? make_id(RTB_TYPE_SURFACE, i) : id;
                warp_nearest(ct, cid);
                tmin = ct; id = cid;
            } else {
                for (int l = 0; l < 32; l++) {                  /* replay the round in index order */
                    const int k = __shfl_sync(FULL, kind, l);
                    const float tt = __shfl_sync(FULL, t, l);
                    if (surface_accept(k, tt, tmin)) { tmin = tt; id = make_id(RTB_TYPE_SURFACE, base + l); }
                }
            }
        }
    }
    if (shadow_mode) {
        for (int i = lane; i < P.n_box; i += 32)
            if (intersectBox(K, ro, rd, S.boxes + i, tmin, t)) occluded = true;
    } else if (P.n_box > 0) {
        /* boxes accept with `!(tN >= tmin)`: NaN candidates and a NaN running minimum make the order matter
         * (box_accept above), so boxes are resolved like quadrics: 32 at a time against the warp-uniform running
         * minimum, by reduction when no NaN is involved, in index order otherwise */
        if (P.n_surf == 0) warp_nearest(tmin, id);
        for (int base = 0; base < P.n_box; base += 32) {
            const int i = base + lane;
            float tN = 0.f;
            const bool valid = i < P.n_box && box_candidate(K, ro, rd, S.boxes + i, tN);
            if (!__any_sync(FULL, (valid && tN != tN) || tmin != tmin)) {
                const bool acc = box_accept(valid, tN, tmin);
                float ct = acc ? tN : tmin;
                int cid = acc ? make_id(RTB_TYPE_BOX, i) : id;
                warp_nearest(ct, cid);
                tmin = ct; id = cid;
            } else {
                for (int l = 0; l < 32; l++) {
                    const bool v = __shfl_sync(FULL, valid ? 1 : 0, l) != 0;
                    const float tt = __shfl_sync(FULL, tN, l);
                    if (box_accept(v, tt, tmin)) { tmin = tt; id = make_id(RTB_TYPE_BOX, base + l); }
                }
            }
        }
    }
    for (int i = lane; i < P.n_torus; i += 32) {
        TorusState st;
        if (torus_setup(K, ro, rd, S.tori + i, P.cull, st)) {
            int iters;
            t = RTB_SCALAR_DK ? torus_solve_scalar(st, iters) : torus_solve<true>(K, st, iters);
            if (COUNT) cnt.dk += iters;
            if (t > 0 && t < 100 && t < tmin) {
                if (shadow_mode) occluded = true; else { tmin = t; id = make_id(RTB_TYPE_TORUS, i); }
            }
        }
    }
    for (int i = lane; i < P.n_ring; i += 32) {
        vec2 uv;
        if (intersectRing(K, ro, rd, S.rings + i, tmin, t, uv)) {
            if (shadow_mode) occluded = true; else { tmin = t; id = make_id(RTB_TYPE_RING, i); ring_uv = uv; }
        }
    }
    if (!shadow_mode)
        for (int i = lane; i < P.n_lpoint; i += 32) {
            float4 o = lds4(S.lights, i);
            if (intersectSphere(ro, rd, o, false, tmin, t)) { tmin = t; id = make_id(RTB_TYPE_POINT_LIGHT, i); }
        }
    const int my_id = id;
    warp_nearest(tmin, id);
    /* the winner's ring uv: the lowest lane that holds the winning id (a ring id is held by exactly one lane) */
    const unsigned win = __ballot_sync(FULL, id >= 0 && my_id == id);
    const int src = win ? __ffs(win) - 1 : 0;
    ring_uv_out = mk2(__shfl_sync(FULL, ring_uv.x, src), __shfl_sync(FULL, ring_uv.y, src));
    id_out = id;
    tmin_out = id >= 0 ? tmin : limit;
    shadow_out = __any_sync(FULL, occluded) ? 1.f : 0.f;
}

/* get_hit_info, rt.frag:744-784.  With TEX it must be called from warp-uniform control
 * flow (`valid` masks lanes without a hit). */
template <bool COUNT, bool TEX>
DEV void hit_info(const FrameParams& P, bool valid, int id, vec3 ro, vec3 rd, vec3 pt, float t, vec2 ring_uv,
                  Material& mat, vec3& normal, float& alpha, float& bias_mult, Counters& cnt) {
    int type = valid ? id_type(id) : -1, num = id_num(id);
    alpha = 1.f;
    normal = mk3(0.f, 0.f, 0.f);
    int site_kind = 0, texnum = 0;
    vec2 uv0 = mk2(0.f, 0.f), uv1 = uv0, uv2 = uv0;
    float w0 = 0.f, w1 = 0.f, w2 = 0.f;
    if (COUNT && valid) cnt.shaded[type]++;
    if (type == RTB_TYPE_SPHERE) {
        const rtb_sphere* s = P.spheres + num;
        mat = load_material(&s->material);
        float4 o = __ldg((const float4*)s->obj);
        normal = normalize(pt - mk3(o.x, o.y, o.z));
        if (TEX) {
            texnum = __ldg(&s->textureNum);
            if (texnum != 0) {                                  /* getSphereTexture, rt.frag:319-340 */
                float4 q = __ldg((const float4*)s->quat_rotation);
                vec3 sn = normal;
                if (q.x != 0.f || q.y != 0.f || q.z != 0.f || q.w != 1.f) sn = rotate(mk4(q.x, q.y, q.z, q.w), sn);
                uv0 = mk2(0.5f + atan2f(sn.z, sn.x) / (2.f * PI_F), 0.5f - asinf(sn.y) / PI_F);
                site_kind = 1;
            }
        }
    } else if (type == RTB_TYPE_PLANE) {
        const rtb_plane* s = P.planes + num;
        mat = load_material(&s->material);
        normal = normalize(mk3(__ldg(&s->normal[0]), __ldg(&s->normal[1]), __ldg(&s->normal[2])));
    } else if (type == RTB_TYPE_SURFACE) {
        mat = load_material(&P.surfaces[num].mat);
        normal = getSurfaceNormal(ro, rd, t, P.surfaces[num]);
    } else if (type == RTB_TYPE_BOX) {
        const rtb_box* b = P.boxes + num;
        mat = load_material(&b->mat);
        normal = boxNormal(ro, rd, *b);
        if (TEX) {
            texnum = __ldg(&b->textureNum);
            if (texnum != 0) {                                  /* getBoxTexture, rt.frag:428-436 */
                vec4 q = mk4(b->quat_rotation[0], b->quat_rotation[1], b->quat_rotation[2], b->quat_rotation[3]);
                vec3 pos = rotate(q, ld3(b->pos));
                vec3 p = rotate(q, pt);
                vec3 nn = rotate(q, normal);
                uv0 = mk2(0.5f * (p.z - pos.z) - 0.5f, 0.5f * (p.y - pos.y) - 0.5f);
                uv1 = mk2(0.5f * (p.z - pos.z) - 0.5f, 0.5f * (p.x - pos.x) - 0.5f);
                uv2 = mk2(0.5f * (p.x - pos.x) - 0.5f, 0.5f * (p.y - pos.y) - 0.5f);
                w0 = fabsf(nn.x); w1 = fabsf(nn.y); w2 = fabsf(nn.z);
                site_kind = 3;
            }
        }
    } else if (type == RTB_TYPE_TORUS) {
        mat = load_material(&P.toruses[num].mat);
        normal = getTorusNormal(ro, rd, t, P.toruses[num]);
    } else if (type == RTB_TYPE_RING) {
        const rtb_ring* r = P.rings + num;
        mat = load_material(&r->mat);
        vec4 q = mk4(r->quat_rotation[0], r->quat_rotation[1], r->quat_rotation[2], r->quat_rotation[3]);
        normal = rotate(quat_inv(q), mk3(0.f, 0.f, -1.f));      /* getRingNormal, rt.frag:391-394 */
        if (TEX) {
            texnum = __ldg(&r->textureNum);
            if (texnum != 0) { uv0 = ring_uv; site_kind = 2; }
        }
    }
    if (TEX) {
        if (__any_sync(FULL, site_kind != 0)) {
            Derivs d0 = quad_derivs(site_kind != 0, site_kind, uv0);
            if (site_kind == 1) {
                vec2 df = mk2(fabsf(d0.dudx) + fabsf(d0.dudy), fabsf(d0.dvdx) + fabsf(d0.dvdy));    /* fwidth */
                if (df.x > 0.5f) df.x = 0.f;
                vec4 c = mk4(0.f, 0.f, 0.f, 0.f);
                if (texnum >= 1 && texnum <= 3) c = texture_lod(P.tex[texnum], uv0.x, uv0.y, log2f(gmax(df.x, df.y) * 1024.f));
                mat.color = mk3(c.x, c.y, c.z);
                alpha = c.w;
            } else if (site_kind == 2) {
                vec4 c = texture_lod(P.tex[4], uv0.x, uv0.y, implicit_lod(P.tex[4], d0.dudx, d0.dvdx, d0.dudy, d0.dvdy));
                mat.color = mk3(c.x, c.y, c.z);
                alpha = c.w;
            }
            if (__any_sync(FULL, site_kind == 3)) {
                Derivs d1 = quad_derivs(site_kind == 3, 3, uv1);
                Derivs d2 = quad_derivs(site_kind == 3, 3, uv2);
                if (site_kind == 3) {
                    vec4 c0 = texture_lod(P.tex[5], uv0.x, uv0.y, implicit_lod(P.tex[5], d0.dudx, d0.dvdx, d0.dudy, d0.dvdy));
                    vec4 c1 = texture_lod(P.tex[5], uv1.x, uv1.y, implicit_lod(P.tex[5], d1.dudx, d1.dvdx, d1.dudy, d1.dvdy));
                    vec4 c2 = texture_lod(P.tex[5], uv2.x, uv2.y, implicit_lod(P.tex[5], d2.dudx, d2.dvdx, d2.dudy, d2.dvdy));
                    mat.color = mk3(w0 * c0.x + w1 * c1.x + w2 * c2.x, w0 * c0.y + w1 * c1.y + w2 * c2.y, w0 * c0.z + w1 * c1.z + w2 * c2.z);
                }
            }
        }
    }
    float distance = length(pt - ro);
    bias_mult = (9e-3f * distance + 35) / 35e3f;
}

}  // namespace RTB_NS

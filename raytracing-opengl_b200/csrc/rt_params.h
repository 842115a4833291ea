/* rt_params.h — host/device shared description of one frame's inputs.
 *
 * The std140 arrays uploaded through rtb_upload() (include/rtb200_types.h) stay
 * in HBM untouched ("raw"); a tiny pack kernel derives from them the HOT
 * geometry records below — everything the linear scan of rt.frag:587-658
 * touches, nothing it does not (the 64-byte material of every primitive is
 * cold: it is read once per shaded hit, from the raw arrays).  The packed
 * block is what each persistent CTA stages into shared memory with one TMA
 * bulk copy.  Every record is a multiple of 16 bytes.
 *
 * Squares that the shader recomputes per test (r*r, R*R) are stored squared:
 * the same single fp32 multiply, done once — results are bit-identical.
 */
#ifndef RT_PARAMS_H
#define RT_PARAMS_H

#include <stdint.h>
#include "../../include/rtb200_types.h"

struct PSphere { float cx, cy, cz, r2; };                                   /* 16 B; r2 = r*r, its SIGN BIT set for a hollow sphere */
struct PPlane  { float nx, ny, nz, _0, px, py, pz, _1; };                    /* 32 B */
struct PBox    { float qx, qy, qz, qw, px, py, pz, fx, fy, fz; int32_t tex; int32_t _0; };      /* 48 B */
struct PTorus  { float qx, qy, qz, qw, px, py, pz, R2, r2, _0, _1, _2; };    /* 48 B */
struct PRing   { float qx, qy, qz, qw, px, py, pz, r1, r2; int32_t tex; int32_t _0, _1; };      /* 48 B */
struct PSurf   { float qx, qy, qz, qw, px, py, pz, a, b, c, d, e, f, minx, miny, minz, maxx, maxy, maxz, _0; }; /* 80 B */
struct PLight  { float x, y, z, r2; };                                       /* 16 B */

/* The FUSED build (rtb_fast: FMA contraction, approximate reciprocals; parity by the envelope criterion) keeps
 * every rotated primitive as the 3x3 matrix of its quaternion sandwich  rotate(q, v) = q (v,0) q*  (rt.frag:305-311;
 * valid for non-unit q as well: the matrix carries |q|^2), row-major in m[9], followed by the position: the local ray
 * is  rd' = M rd,  ro' = M (ro - p)  — 21 FMA-pipe instructions instead of the 94 of the two Hamilton products. */
struct PBoxM   { float m[9], px, py, pz, fx, fy, fz; int32_t tex; };                          /* 64 B */
struct PTorusM { float m[9], px, py, pz, R2, r2, k, fourR2; };                                /* 64 B; k = R2 - r2 */
struct PRingM  { float m[9], px, py, pz, r1, r2; int32_t tex; int32_t _0; };                  /* 64 B */
struct PSurfM  { float m[9], px, py, pz, a, b, c, d, e, f, minx, miny, minz, maxx, maxy, maxz; };  /* 96 B */

/* byte offsets of each section inside the packed block (all multiples of 16) */
struct PackedLayout {
    uint32_t off_plane, off_sphere, off_surf, off_box, off_torus, off_ring, off_light;
    uint32_t total_bytes;
};

struct TexDesc {                      /* one mip-mapped RGBA8 2-D texture in HBM */
    const uint8_t* base;              /* NULL = unit not bound */
    int32_t w, h, levels;
    uint32_t level_off[16];           /* byte offset of each level */
};

struct CubeDesc {
    const uint8_t* base;              /* 6 RGBA8 faces, face f at base + f*w*h*4; NULL = unbound */
    int32_t w, h;
};

struct FrameParams {
    /* specialisation constants (rt.frag:122-132) */
    int32_t n_sphere, n_plane, n_surf, n_box, n_torus, n_ring, n_lpoint, n_ldirect, iterations;
    float ambient[3], shadow_ambient[3];
    /* scene uniform (rt.frag:104-113) */
    float cam_q[4], cam_pos[3];
    int32_t canvas_w, canvas_h;
    /* raw std140 arrays in HBM */
    const rtb_sphere* spheres; const rtb_plane* planes; const rtb_surface* surfaces; const rtb_box* boxes;
    const rtb_torus* toruses; const rtb_ring* rings; const rtb_light_point* lights_point; const rtb_light_direct* lights_direct;
    /* packed hot geometry */
    const uint8_t* packed; PackedLayout lay;
    /* samplers */
    CubeDesc cube; TexDesc tex[6];
    /* output: this rank's scanlines, packed in block order (rtb_set_partition) — or, fb_global != 0, the whole canvas (possibly
     * the root GPU's frame reached over NVLink peer access), of which this rank writes its own scanlines at their final place */
    float* fb; int32_t fb_global; int32_t rank, world, block_rows, local_rows;
    /* work distribution */
    unsigned int* tile_counter; int32_t n_tiles_x, n_tiles_y;
    /* persistent kernel, frame loops: tile_cost[t] accumulates the path lengths of tile t's pixels in this frame; tile_perm is
     * the hand-out order derived from the previous frame's costs (costliest first, so that the frame ends on its cheapest
     * paths and the drain has little left to do).  Either may be NULL. */
    unsigned int* tile_cost; const unsigned int* tile_perm;
    /* options */
    int32_t cull;
    int32_t coop;                     /* 1 = cooperative drain (default); 0 = every warp drains alone with serial scans (A/B, tests) */
    /* 1.0f, -0.0f, -1.0f as RUN-TIME values: the packed (f32x2) Durand-Kerner solver builds its separately rounded
     * multiplies and adds from FFMA2 with these operands; ptxas must not be able to fold them (rt_device.cuh, "packed") */
    float k_one, k_neg_zero, k_neg_one;
    unsigned long long* counters;     /* NULL unless the counting variant runs; layout = enum CounterSlot */
    unsigned long long* cta_times;    /* NULL, or 3 globaltimer stamps per CTA: start, drain start, end (RTB_DEBUG_TIMES) */
};

enum CounterSlot {
    CNT_RAYS_NEAREST = 0, CNT_RAYS_SHADOW = 1, CNT_TESTS0 = 2 /* ..8 */, CNT_DK = 9, CNT_SHADED0 = 10 /* ..16 */,
    CNT_LIGHT_EVALS = 17, CNT_PIXELS = 18, CNT_NUM = 19
};

#endif

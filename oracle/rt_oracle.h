/* rt_oracle.h — C interface of the CPU oracle (liboracle.so).
 *
 * TEST INFRASTRUCTURE.  The oracle is a CPU restatement of the reference's
 * fragment shader assets/shaders/rt.frag; it exists to CHECK the CUDA path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load it.  The product (librtb200.so and the
 * raytracing-opengl_b200 package) never links, imports or calls it.
 */
#ifndef RT_ORACLE_H
#define RT_ORACLE_H

#include <stddef.h>
#include <stdint.h>
#include "../include/rtb200_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_image { const uint8_t* px; int32_t w, h, ch; } orc_image;

/* Everything the shader sees: specialisation constants (rt.frag:122-132), the
 * nine uniform blocks (rt.frag:155-230) and the samplers (rt.frag:136-143). */
typedef struct orc_scene_desc {
    rtb_defines defines;                /* raw values; the %f rounding of GLWrapper.cpp:279-282 is applied inside */
    const rtb_scene* scene;
    const rtb_sphere* spheres;
    const rtb_plane* planes;
    const rtb_surface* surfaces;
    const rtb_box* boxes;
    const rtb_torus* toruses;
    const rtb_ring* rings;
    const rtb_light_point* lights_point;
    const rtb_light_direct* lights_direct;
    orc_image cube[6];                  /* +X,-X,+Y,-Y,+Z,-Z; px NULL => unbound (samples 0,0,0,1) */
    orc_image tex2d[6];                 /* index = texture unit 1..5 (0 unused); px NULL => unbound */
} orc_scene_desc;

typedef struct orc_stats {             /* same meaning as rtb_stats in include/rtb200.h */
    uint64_t pixels, rays_nearest, rays_shadow;
    uint64_t tests[7];
    uint64_t dk_iterations;
    uint64_t shaded_hits[7];
    uint64_t light_evals;
    uint64_t dk_hist[61];              /* histogram of Durand-Kerner iterations per (ray, torus) solve, index = k */
} orc_stats;

typedef struct orc_handle orc_handle;

/* derivative pairing rule at texture sites whose 2x2 quad diverged */
enum { ORC_PAIR_PROGRAM_ORDER = 0,     /* SIMT lock-step: same loop trip + same call site (default; what the CUDA quad kernel does) */
       ORC_PAIR_ORDINAL = 1 };         /* k-th call of the same sampler (what oracle/_ref can observe) */

orc_handle* orc_create(const orc_scene_desc* desc);   /* copies everything, builds mip chains */
void orc_destroy(orc_handle*);
void orc_set_pairing(orc_handle*, int rule);

/* Render the window [x0,x0+w) x [y0,y0+h) (all even) of the canvas into
 * out[h][w][4] (row 0 = y0 = bottom-most row; GL window coordinates).
 * n_threads <= 0 => hardware concurrency.  stats may be NULL. */
int orc_render(orc_handle*, int x0, int y0, int w, int h, float* out, orc_stats* stats, int n_threads);

/* Render n 2x2 quads whose lower-left pixels are (qx[i], qy[i]) (even);
 * out[n][4][4]: lanes in order (x,y) (x+1,y) (x,y+1) (x+1,y+1). */
int orc_render_quads(orc_handle*, int n, const int32_t* qx, const int32_t* qy, float* out, orc_stats* stats, int n_threads);

/* Known-answer-test entry points for single functions of rt.frag. */
/* calcInter (rt.frag:587): returns tmin; num/type are left at their inputs on a miss. */
float orc_calc_inter(orc_handle*, const float ro[3], const float rd[3], int32_t* num, int32_t* type);
/* inShadow (rt.frag:630) */
float orc_in_shadow(orc_handle*, const float ro[3], const float rd[3], float dist);
/* one intersector: type = rtb_prim_type, index into its array; returns 1 on hit and writes t */
int orc_intersect(orc_handle*, int type, int index, const float ro[3], const float rd[3], float tmin, float* t, int32_t* dk_iters);
/* getRayDir (rt.frag:313) for pixel (x,y) */
void orc_ray_dir(orc_handle*, int x, int y, float out[3]);
/* texture(skybox, dir) */
void orc_sample_cube(orc_handle*, const float dir[3], float out[4]);
/* textureLod(unit, uv, lod) */
void orc_sample_2d(orc_handle*, int unit, float u, float v, float lod, float out[4]);
/* mip level count and size/pixels of one level (for checking the CUDA library's chain) */
int orc_mip_levels(orc_handle*, int unit);
const uint8_t* orc_mip_level(orc_handle*, int unit, int level, int32_t* w, int32_t* h);

#ifdef __cplusplus
}
#endif
#endif

/* rtb200_types.h — layout contract of the hot path's inputs.
 *
 * Plain-C mirrors (no glm) of the reference's host structs in src/scene.h, which
 * are themselves byte-compatible with the std140 uniform blocks of
 * assets/shaders/rt.frag:24-113.  The kernel, the oracle and every host binding
 * read primitive arrays in exactly this layout, so a `std::vector<rt_sphere>`
 * built by the reference's SceneManager factories can be handed to
 * rtb_upload() unchanged.
 *
 *   reference struct      here            bytes   reference definition
 *   rt_material           rtb_material      64    scene.h:22-35   / rt.frag:24-34
 *   rt_sphere             rtb_sphere       112    scene.h:37-44   / rt.frag:36-42
 *   rt_plane              rtb_plane         96    scene.h:46-50   / rt.frag:44-48
 *   rt_box                rtb_box          112    scene.h:52-58   / rt.frag:50-56
 *   rt_torus              rtb_torus        112    scene.h:60-65   / rt.frag:81-86
 *   rt_ring               rtb_ring         112    scene.h:67-73   / rt.frag:58-65
 *   rt_surface            rtb_surface      160    scene.h:75-95   / rt.frag:67-79
 *   rt_light_direct       rtb_light_direct  32    scene.h:99-104  / rt.frag:88-93
 *   rt_light_point        rtb_light_point   48    scene.h:106-114 / rt.frag:95-102
 *   rt_scene              rtb_scene         64    scene.h:116-126 / rt.frag:104-113
 *
 * Quaternions are stored x,y,z,w (glm::quat memory order; rt.frag:285-288,320).
 * tests/test_layout.py checks these sizes/offsets against the real scene.h when
 * /root/reference is present.
 */
#ifndef RTB200_TYPES_H
#define RTB200_TYPES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rtb_material {
    float color[3];   float _p1;        /*  0 */
    float absorb[3];                    /* 16 */
    float diffuse;                      /* 28 */
    float reflect;                      /* 32 */
    float refract;                      /* 36 */
    int32_t specular;                   /* 40 */
    float kd;                           /* 44 */
    float ks;                           /* 48 */
    float _pad[3];                      /* 52 */
} rtb_material;                         /* 64 */

typedef struct rtb_sphere {
    rtb_material material;              /*  0 */
    float obj[4];                       /* 64  xyz = centre, w = radius */
    float quat_rotation[4];             /* 80  x,y,z,w */
    int32_t textureNum;                 /* 96 */
    uint32_t hollow;                    /* 100 C++ bool (1 byte) + zeroed padding, read as a 4-byte GLSL bool */
    float _pad[2];                      /* 104 */
} rtb_sphere;                           /* 112 */

typedef struct rtb_plane {
    rtb_material material;              /*  0 */
    float pos[3];    float _p1;         /* 64 */
    float normal[3]; float _p2;         /* 80 */
} rtb_plane;                            /* 96 */

typedef struct rtb_box {
    rtb_material mat;                   /*  0 */
    float quat_rotation[4];             /* 64 */
    float pos[3];  float _p1;           /* 80 */
    float form[3];                      /* 96  half extents */
    int32_t textureNum;                 /* 108 */
} rtb_box;                              /* 112 */

typedef struct rtb_torus {
    rtb_material mat;                   /*  0 */
    float quat_rotation[4];             /* 64 */
    float pos[3];  float _p1;           /* 80 */
    float form[2];                      /* 96  x = ring radius R, y = tube radius r */
    float _p2[2];                       /* 104 */
} rtb_torus;                            /* 112 */

typedef struct rtb_ring {
    rtb_material mat;                   /*  0 */
    float quat_rotation[4];             /* 64 */
    float pos[3];                       /* 80 */
    int32_t textureNum;                 /* 92 */
    float r1;                           /* 96  SQUARE of inner radius (SceneManager.cpp:195) */
    float r2;                           /* 100 SQUARE of outer radius */
    float _p2[2];                       /* 104 */
} rtb_ring;                             /* 112 */

typedef struct rtb_surface {
    rtb_material mat;                   /*  0 */
    float quat_rotation[4];             /* 64 */
    float v_min[3]; float _p0;          /* 80  world-space clip box, default -FLT_MAX */
    float v_max[3]; float _p1;          /* 96  default +FLT_MAX */
    float pos[3];                       /* 112 */
    float a, b, c, d, e, f;             /* 124 ax^2+by^2+cz^2+dz+ey+f = 0 */
    float _pad[3];                      /* 148 */
} rtb_surface;                          /* 160 */

typedef struct rtb_light_direct {
    float direction[3]; float _p1;      /*  0 */
    float color[3];                     /* 16 */
    float intensity;                    /* 28 */
} rtb_light_direct;                     /* 32 */

typedef struct rtb_light_point {
    float pos[4];                       /*  0 xyz + radius */
    float color[3];                     /* 16 */
    float intensity;                    /* 28 */
    float linear_k;                     /* 32 */
    float quadratic_k;                  /* 36 */
    float _pad[2];                      /* 40 */
} rtb_light_point;                      /* 48 */

typedef struct rtb_scene {
    float quat_camera_rotation[4];      /*  0 */
    float camera_pos[3]; float _p1;     /* 16 */
    float bg_color[3];                  /* 32 unused by the shader */
    int32_t canvas_width;               /* 44 */
    int32_t canvas_height;              /* 48 */
    int32_t reflect_depth;              /* 52 unused by the shader ({ITERATIONS} is) */
    float _pad[2];                      /* 56 */
} rtb_scene;                            /* 64 */

/* Uniform-block binding points, SceneManager.cpp:246-254. */
enum rtb_binding {
    RTB_BIND_SCENE = 0,
    RTB_BIND_SPHERES = 1,
    RTB_BIND_PLANES = 2,
    RTB_BIND_SURFACES = 3,
    RTB_BIND_BOXES = 4,
    RTB_BIND_TORUSES = 5,
    RTB_BIND_RINGS = 6,
    RTB_BIND_LIGHTS_POINT = 7,
    RTB_BIND_LIGHTS_DIRECT = 8,
    RTB_NUM_BINDINGS = 9
};

/* Primitive type tags, rt.frag:7-13. */
enum rtb_prim_type {
    RTB_TYPE_SPHERE = 0,
    RTB_TYPE_PLANE = 1,
    RTB_TYPE_SURFACE = 2,
    RTB_TYPE_BOX = 3,
    RTB_TYPE_TORUS = 4,
    RTB_TYPE_RING = 5,
    RTB_TYPE_POINT_LIGHT = 6
};

/* Shader specialisation constants, scene.h:7-20 -> rt.frag:122-132. */
typedef struct rtb_defines {
    int32_t sphere_size;
    int32_t plane_size;
    int32_t surface_size;
    int32_t box_size;
    int32_t torus_size;
    int32_t ring_size;
    int32_t light_point_size;
    int32_t light_direct_size;
    int32_t iterations;
    float ambient_color[3];
    float shadow_ambient[3];
} rtb_defines;                          /* 60; same field order as scene.h rt_defines */

#ifdef __cplusplus
}
#define RTB_SA(T, n) static_assert(sizeof(T) == n, #T " must be " #n " bytes (std140)")
RTB_SA(rtb_material, 64);  RTB_SA(rtb_sphere, 112);  RTB_SA(rtb_plane, 96);
RTB_SA(rtb_box, 112);      RTB_SA(rtb_torus, 112);   RTB_SA(rtb_ring, 112);
RTB_SA(rtb_surface, 160);  RTB_SA(rtb_light_direct, 32);
RTB_SA(rtb_light_point, 48); RTB_SA(rtb_scene, 64);  RTB_SA(rtb_defines, 60);
#undef RTB_SA
#endif

#endif /* RTB200_TYPES_H */

/* scene_gen.h — the synthetic benchmark scenes of BASELINE.json, generated in C++.
 *
 * The C++ twin of raytracing-opengl_b200/scenes.py `synthetic_scene()`: the generator IS the definition of the workloads
 * `spheres4k`, `tori1080`, `mixed1024` (and the small `mini<seed>` test scenes), so a C++ host can build them without
 * Python.  Same PCG32 stream (XSH-RR 64/32, stream 1), same 24-bit uniforms, same draw order, same fp32 roundings:
 * tests/test_scenes.py compares the bytes of every uniform block with the Python generator's.
 * Uses only include/rtb200_types.h (the std140 mirrors of src/scene.h); the factories restate SceneManager.cpp:137-236 and
 * Surface.h like scene.py does.
 */
#ifndef RTB_SCENE_GEN_H
#define RTB_SCENE_GEN_H

#include <string>
#include <vector>

#include "../../include/rtb200_types.h"

struct RtbSceneContainer {                 /* src/scene.h:128-154 */
    rtb_scene scene;
    float ambient_color[3], shadow_ambient[3];
    std::vector<rtb_sphere> spheres;
    std::vector<rtb_plane> planes;
    std::vector<rtb_surface> surfaces;
    std::vector<rtb_box> boxes;
    std::vector<rtb_torus> toruses;
    std::vector<rtb_ring> rings;
    std::vector<rtb_light_point> lights_point;
    std::vector<rtb_light_direct> lights_direct;
    rtb_defines defines() const;
};

/* name: "spheres4k" | "tori1080" | "mixed1024" | "mini<seed>"; returns false for an unknown name */
bool rtb_generate_scene(const std::string& name, int width, int height, int iterations, RtbSceneContainer& out);
/* "default256" ... "mixed1024_8k" -> scene name + canvas + bounces (scenes.py CONFIGS; the default main.cpp scene is not synthetic) */
bool rtb_config_lookup(const std::string& config, std::string& scene, int& width, int& height, int& iterations);

#endif

/* scene_gen.cpp — see scene_gen.h.  `rt_scene_gen <scene> <w> <h> <iterations> <outdir>` writes one raw file per uniform block. */
#include "scene_gen.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace {

/* PCG32 (O'Neill, XSH-RR 64/32), stream 1 — scenes.py PCG32 */
struct PCG32 {
    uint64_t state = 0, inc;
    explicit PCG32(uint64_t seed, uint64_t stream = 1) : inc((stream << 1) | 1) {
        next_u32();
        state += seed;
        next_u32();
    }
    uint32_t next_u32() {
        uint64_t old = state;
        state = old * 6364136223846793005ULL + inc;
        uint32_t xorshifted = (uint32_t)(((old >> 18) ^ old) >> 27);
        uint32_t rot = (uint32_t)(old >> 59);
        return (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
    }
    float uniform() { return (float)((double)(next_u32() >> 8) * (1.0 / 16777216.0)); }       /* 24 bits: exact in fp32 */
    float range(float lo, float hi) { float u = uniform(); float span = hi - lo; float prod = span * u; return lo + prod; }
    uint32_t index(uint32_t n) { return next_u32() % n; }
    void quat(float q[4]) {                   /* uniform unit quaternion: rejection-sample the 4-ball, normalise in fp64 */
        for (;;) {
            double v[4];
            for (int i = 0; i < 4; i++) v[i] = (double)range(-1.f, 1.f);
            double n2 = 0.0;
            for (int i = 0; i < 4; i++) n2 += v[i] * v[i];
            if (n2 > 1e-4 && n2 <= 1.0) {
                double n = std::sqrt(n2);
                for (int i = 0; i < 4; i++) q[i] = (float)(v[i] / n);
                return;
            }
        }
    }
};

void set3(float* d, float x, float y, float z) { d[0] = x; d[1] = y; d[2] = z; }
void ident(float* q) { q[0] = q[1] = q[2] = 0.f; q[3] = 1.f; }

/* SceneManager::create_material, SceneManager.cpp:137-152 */
rtb_material create_material(const float color[3], int specular, float reflect, float refract = 0.f) {
    rtb_material m;
    memset(&m, 0, sizeof m);
    set3(m.color, color[0], color[1], color[2]);
    m.specular = specular; m.reflect = reflect; m.refract = refract;
    m.diffuse = 0.7f; m.kd = 0.8f; m.ks = 0.2f;
    return m;
}
rtb_material rand_material(PCG32& rng) {
    static const int SPECULARS[5] = { 0, 10, 50, 100, 200 };
    float color[3];
    color[0] = rng.range(0.1f, 1.f); color[1] = rng.range(0.1f, 1.f); color[2] = rng.range(0.1f, 1.f);
    int specular = SPECULARS[rng.index(5)];
    float reflect = rng.range(0.f, 0.6f);
    if (rng.uniform() < 0.25f) reflect = 0.f;          /* 25 % purely diffuse */
    return create_material(color, specular, reflect);
}
void rand_centre(PCG32& rng, float c[3]) { c[0] = rng.range(-20.f, 20.f); c[1] = rng.range(0.3f, 10.f); c[2] = rng.range(0.f, 40.f); }

void add_spheres(RtbSceneContainer& sc, PCG32& rng, int n) {
    for (int i = 0; i < n; i++) {
        rtb_sphere s;
        memset(&s, 0, sizeof s);
        float c[3];
        rand_centre(rng, c);
        float r = rng.range(0.3f, 1.5f);
        s.material = rand_material(rng);
        s.obj[0] = c[0]; s.obj[1] = c[1]; s.obj[2] = c[2]; s.obj[3] = r;
        ident(s.quat_rotation);
        sc.spheres.push_back(s);
    }
}
void add_boxes(RtbSceneContainer& sc, PCG32& rng, int n) {
    for (int i = 0; i < n; i++) {
        rtb_box b;
        memset(&b, 0, sizeof b);
        float c[3], q[4];
        rand_centre(rng, c);
        float fx = rng.range(0.3f, 1.5f), fy = rng.range(0.3f, 1.5f), fz = rng.range(0.3f, 1.5f);
        rng.quat(q);
        b.mat = rand_material(rng);
        set3(b.pos, c[0], c[1], c[2]);
        set3(b.form, fx, fy, fz);
        memcpy(b.quat_rotation, q, sizeof q);
        sc.boxes.push_back(b);
    }
}
void add_tori(RtbSceneContainer& sc, PCG32& rng, int n) {
    for (int i = 0; i < n; i++) {
        rtb_torus t;
        memset(&t, 0, sizeof t);
        float c[3], q[4];
        rand_centre(rng, c);
        float R = rng.range(0.6f, 1.5f), r = rng.range(0.15f, 0.5f);
        rng.quat(q);
        t.mat = rand_material(rng);
        set3(t.pos, c[0], c[1], c[2]);
        t.form[0] = R; t.form[1] = r;
        memcpy(t.quat_rotation, q, sizeof q);
        sc.toruses.push_back(t);
    }
}
float inv2(float x) { return (float)std::pow((double)x, -2.0); }        /* Surface.h: powf(a, -2) */
void add_quadrics(RtbSceneContainer& sc, PCG32& rng, int n_each) {
    for (int kind = 0; kind < 3; kind++)                                  /* ellipsoid, cone, cylinder */
        for (int i = 0; i < n_each; i++) {
            rtb_surface s;
            memset(&s, 0, sizeof s);
            float c[3], q[4];
            rand_centre(rng, c);
            float a = rng.range(0.3f, 1.2f), b = rng.range(0.3f, 1.2f), cc = rng.range(0.3f, 1.2f);
            rng.quat(q);
            s.mat = rand_material(rng);
            if (kind == 0) { s.a = inv2(a); s.b = inv2(b); s.c = inv2(cc); s.f = -1.f; }          /* GetEllipsoid */
            else if (kind == 1) { s.a = inv2(a); s.b = inv2(b); s.c = -inv2(cc); }                /* GetEllipticCone */
            else { s.a = inv2(a); s.b = inv2(b); s.f = -1.f; }                                    /* GetEllipticCylinder */
            set3(s.pos, c[0], c[1], c[2]);
            memcpy(s.quat_rotation, q, sizeof q);
            for (int k = 0; k < 3; k++) { s.v_min[k] = c[k] - 2.f; s.v_max[k] = c[k] + 2.f; }     /* world-space clip box = centre +- 2 */
            sc.surfaces.push_back(s);
        }
}
void add_ground(RtbSceneContainer& sc, PCG32& rng) {
    rtb_plane p;
    memset(&p, 0, sizeof p);
    p.material = rand_material(rng);
    set3(p.normal, 0.f, 1.f, 0.f);
    set3(p.pos, 0.f, 0.f, 0.f);
    sc.planes.push_back(p);
}

void write_block(const std::string& dir, const char* name, const void* data, size_t bytes) {
    std::string path = dir + "/" + name + ".bin";
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { fprintf(stderr, "cannot write %s\n", path.c_str()); exit(1); }
    if (bytes) fwrite(data, 1, bytes, f);
    fclose(f);
}

}  // namespace

rtb_defines RtbSceneContainer::defines() const {
    rtb_defines d;
    memset(&d, 0, sizeof d);
    d.sphere_size = (int)spheres.size(); d.plane_size = (int)planes.size(); d.surface_size = (int)surfaces.size(); d.box_size = (int)boxes.size();
    d.torus_size = (int)toruses.size(); d.ring_size = (int)rings.size(); d.light_point_size = (int)lights_point.size();
    d.light_direct_size = (int)lights_direct.size(); d.iterations = scene.reflect_depth;
    memcpy(d.ambient_color, ambient_color, sizeof ambient_color);
    memcpy(d.shadow_ambient, shadow_ambient, sizeof shadow_ambient);
    return d;
}

bool rtb_generate_scene(const std::string& name, int width, int height, int iterations, RtbSceneContainer& sc) {
    sc = RtbSceneContainer();
    memset(&sc.scene, 0, sizeof sc.scene);
    sc.scene.canvas_width = width; sc.scene.canvas_height = height; sc.scene.reflect_depth = iterations;
    set3(sc.scene.camera_pos, 0.f, 2.f, -12.f);
    ident(sc.scene.quat_camera_rotation);
    set3(sc.shadow_ambient, 0.1f, 0.1f, 0.1f);                              /* main.cpp:47-52 */
    set3(sc.ambient_color, 0.025f, 0.025f, 0.025f);
    rtb_light_point lp;
    memset(&lp, 0, sizeof lp);
    lp.pos[0] = 3.f; lp.pos[1] = 5.f; lp.pos[2] = 0.f; lp.pos[3] = 0.1f; set3(lp.color, 1.f, 1.f, 1.f); lp.intensity = 25.5f; lp.linear_k = 0.22f; lp.quadratic_k = 0.2f;
    sc.lights_point.push_back(lp);
    rtb_light_direct ld;
    memset(&ld, 0, sizeof ld);
    set3(ld.direction, 3.f, -1.f, 1.f); set3(ld.color, 1.f, 1.f, 1.f); ld.intensity = 1.5f;
    sc.lights_direct.push_back(ld);
    if (name == "spheres4k") {
        PCG32 rng(3);
        add_spheres(sc, rng, 256); add_boxes(sc, rng, 64); add_ground(sc, rng);
    } else if (name == "tori1080") {
        PCG32 rng(4);
        add_tori(sc, rng, 128);
    } else if (name == "mixed1024") {
        PCG32 rng(5);
        add_spheres(sc, rng, 512); add_boxes(sc, rng, 256); add_quadrics(sc, rng, 64); add_tori(sc, rng, 64);
    } else if (name.compare(0, 4, "mini") == 0) {
        PCG32 rng((uint64_t)(name.size() > 4 ? atoi(name.c_str() + 4) : 1));
        add_spheres(sc, rng, 6); add_boxes(sc, rng, 4); add_quadrics(sc, rng, 1); add_tori(sc, rng, 2); add_ground(sc, rng);
    } else {
        return false;
    }
    return true;
}

bool rtb_config_lookup(const std::string& config, std::string& scene, int& w, int& h, int& it) {
    struct Row { const char* config; const char* scene; int w, h, it; };
    static const Row rows[] = { { "spheres4k", "spheres4k", 3840, 2160, 8 }, { "tori1080", "tori1080", 1920, 1080, 4 },
                                { "mixed1024_4k", "mixed1024", 3840, 2160, 8 }, { "mixed1024_8k", "mixed1024", 7680, 4320, 8 } };
    for (const Row& r : rows)
        if (config == r.config) { scene = r.scene; w = r.w; h = r.h; it = r.it; return true; }
    return false;
}

#ifdef RTB_SCENE_GEN_MAIN
int main(int argc, char** argv) {
    if (argc != 6) { fprintf(stderr, "usage: %s <scene|config> <width> <height> <iterations> <outdir>   (0 0 0 = the config's own size)\n", argv[0]); return 2; }
    std::string scene = argv[1];
    int w = atoi(argv[2]), h = atoi(argv[3]), it = atoi(argv[4]);
    std::string s2; int cw, ch, cit;
    if (rtb_config_lookup(scene, s2, cw, ch, cit)) { scene = s2; if (!w) { w = cw; h = ch; it = cit; } }
    RtbSceneContainer sc;
    if (!rtb_generate_scene(scene, w, h, it, sc)) { fprintf(stderr, "unknown scene '%s'\n", argv[1]); return 1; }
    const std::string dir = argv[5];
    rtb_defines d = sc.defines();
    write_block(dir, "defines", &d, sizeof d);
    write_block(dir, "scene_buf", &sc.scene, sizeof sc.scene);
    write_block(dir, "spheres_buf", sc.spheres.data(), sc.spheres.size() * sizeof(rtb_sphere));
    write_block(dir, "planes_buf", sc.planes.data(), sc.planes.size() * sizeof(rtb_plane));
    write_block(dir, "surfaces_buf", sc.surfaces.data(), sc.surfaces.size() * sizeof(rtb_surface));
    write_block(dir, "boxes_buf", sc.boxes.data(), sc.boxes.size() * sizeof(rtb_box));
    write_block(dir, "toruses_buf", sc.toruses.data(), sc.toruses.size() * sizeof(rtb_torus));
    write_block(dir, "rings_buf", sc.rings.data(), sc.rings.size() * sizeof(rtb_ring));
    write_block(dir, "lights_point_buf", sc.lights_point.data(), sc.lights_point.size() * sizeof(rtb_light_point));
    write_block(dir, "lights_direct_buf", sc.lights_direct.data(), sc.lights_direct.size() * sizeof(rtb_light_direct));
    return 0;
}
#endif

/* ref_harness.cpp — drives the reference's own rt.frag (compiled as C++ by
 * build_ref.py into _ref/rt_frag_gen.inc) one fragment at a time.
 *
 * TEST INFRASTRUCTURE: oracle/_ref/libref.so pins the restatement in
 * rt_oracle.cpp and generates tests/golden/.  What is the reference's and what
 * is ours here:
 *   reference : every statement of rt.frag (ray generation, intersectors,
 *               nearest-hit / shadow scans, shading, bounce loop)
 *   glm       : the GLSL built-ins (via glsl_prelude.h)
 *   ours      : the uniform/sampler plumbing below (the job of GLWrapper.cpp and
 *               the GL driver): UBO bytes -> GLSL structs, gl_FragCoord, texture
 *               filtering (gl_sampler.h, driver-defined in the reference) and
 *               the 2x2-quad derivative bookkeeping for fwidth/implicit LOD.
 */
#include "rt_oracle.h"
#include "gl_sampler.h"
#include "glsl_prelude.h"

#include <atomic>
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace glsl {

struct SiteRec { uint64_t key; float u, v; };
struct QuadCtx { std::vector<SiteRec> prev[4], cur[4]; };

struct Uniforms {                       /* the {TOKEN} specialisation constants */
    int sphere_size, plane_size, surface_size, box_size, torus_size, ring_size, light_point_size, light_direct_size, iterations;
    vec3 ambient_color, shadow_ambient;
};

struct RefShader {
    const Uniforms& U;
    const glsim::CubeMap* cube = nullptr;
    const glsim::Texture2D* tex2d = nullptr;     /* [6], index = unit */
    QuadCtx* quad = nullptr;
    int lane = 0;
    int ord[8] = { 0 };
    int glass_events__ = 0;
    vec4 gl_FragCoord;

    explicit RefShader(const Uniforms& u) : U(u) {}

    /* derivative provider: k-th call on the same sampler pairs with the neighbour's k-th call */
    void site(int kind, vec2 uv, float d[4]) {
        d[0] = d[1] = d[2] = d[3] = 0.f;
        if (!quad) return;
        uint64_t key = ((uint64_t)kind << 40) | (uint64_t)(ord[kind]++);
        quad->cur[lane].push_back({ key, uv.x, uv.y });
        auto find = [&](int other, float& u, float& v) {
            for (const SiteRec& r : quad->prev[other]) if (r.key == key) { u = r.u; v = r.v; return true; }
            return false;
        };
        float u, v;
        if (find(lane ^ 1, u, v)) { if (lane & 1) { d[0] = uv.x - u; d[1] = uv.y - v; } else { d[0] = u - uv.x; d[1] = v - uv.y; } }
        if (find(lane ^ 2, u, v)) { if (lane & 2) { d[2] = uv.x - u; d[3] = uv.y - v; } else { d[2] = u - uv.x; d[3] = v - uv.y; } }
    }
    vec2 fwidth(vec2 uv) {              /* only call site: rt.frag:326 (sphere texture) */
        float d[4];
        site(0, uv, d);
        return vec2(fabsf(d[0]) + fabsf(d[2]), fabsf(d[1]) + fabsf(d[3]));
    }
    vec4 textureLod(sampler2D s, vec2 uv, float lod) {
        glsim::rgba c = glsim::texture_lod(tex2d[s.unit], uv.x, uv.y, lod);
        return vec4(c.r, c.g, c.b, c.a);
    }
    vec4 texture(sampler2D s, vec2 uv) {   /* implicit LOD: ring (unit 4) and box (unit 5) */
        float d[4];
        site(s.unit == 4 ? 1 : 2, uv, d);
        float lod = glsim::implicit_lod(tex2d[s.unit], d[0], d[1], d[2], d[3]);
        glsim::rgba c = glsim::texture_lod(tex2d[s.unit], uv.x, uv.y, lod);
        return vec4(c.r, c.g, c.b, c.a);
    }
    vec4 texture(samplerCube, vec3 dir) {
        glsim::rgba c = glsim::texture_cube(*cube, dir.x, dir.y, dir.z);
        return vec4(c.r, c.g, c.b, c.a);
    }

#include "rt_frag_gen.inc"
};

}  // namespace glsl

namespace {

using glsl::RefShader;

struct RefScene {
    glsl::Uniforms U;
    RefShader::rt_scene scene;
    std::vector<RefShader::rt_sphere> spheres;
    std::vector<RefShader::rt_plane> planes;
    std::vector<RefShader::rt_surface> surfaces;
    std::vector<RefShader::rt_box> boxes;
    std::vector<RefShader::rt_torus> toruses;
    std::vector<RefShader::rt_ring> rings;
    std::vector<RefShader::rt_light_point> lights_point;
    std::vector<RefShader::rt_light_direct> lights_direct;
    glsim::CubeMap cube;
    glsim::Texture2D tex[6];
};

glsl::vec3 V3(const float* p) { return glsl::vec3(p[0], p[1], p[2]); }
glsl::vec4 V4(const float* p) { return glsl::vec4(p[0], p[1], p[2], p[3]); }

/* std140 bytes -> the GLSL struct (what the GL driver does when the shader reads a UBO) */
RefShader::rt_material M(const rtb_material& m) {
    RefShader::rt_material r;
    r.color = V3(m.color); r.absorb = V3(m.absorb); r.diffuse = m.diffuse; r.reflection = m.reflect;
    r.refraction = m.refract; r.specular = m.specular; r.kd = m.kd; r.ks = m.ks;
    return r;
}

float round_through_percent_f(float v) {      /* GLWrapper.cpp:279-282 */
    char buf[64];
    snprintf(buf, sizeof buf, "%f", v);
    return strtof(buf, nullptr);
}

void bind(RefShader& sh, const RefScene& S) {
    sh.scene = S.scene;
    sh.spheres = S.spheres.data(); sh.planes = S.planes.data(); sh.surfaces = S.surfaces.data();
    sh.boxes = S.boxes.data(); sh.toruses = S.toruses.data(); sh.rings = S.rings.data();
    sh.lights_point = S.lights_point.data(); sh.lights_direct = S.lights_direct.data();
    sh.cube = &S.cube; sh.tex2d = S.tex;
    /* sampler uniform -> texture unit, main.cpp:149-153 (texture_sphere_4 is never bound) */
    sh.texture_sphere_1.unit = 1; sh.texture_sphere_2.unit = 2; sh.texture_sphere_3.unit = 3;
    sh.texture_sphere_4.unit = 0; sh.texture_ring.unit = 4; sh.texture_box.unit = 5;
}

void render_quad(const RefScene& S, int qx, int qy, float* out) {
    glsl::QuadCtx quad;
    glsl::vec4 col[4];
    for (int pass = 0; pass < 8; pass++) {
        for (int l = 0; l < 4; l++) quad.cur[l].clear();
        for (int l = 0; l < 4; l++) {
            RefShader sh(S.U);
            bind(sh, S);
            sh.quad = &quad; sh.lane = l;
            sh.gl_FragCoord = glsl::vec4((float)(qx + (l & 1)) + 0.5f, (float)(qy + (l >> 1)) + 0.5f, 0.5f, 1.0f);
            sh.main_();
            col[l] = sh.FragColor;
        }
        bool any = false, same = true;
        for (int l = 0; l < 4; l++) {
            if (!quad.cur[l].empty()) any = true;
            if (quad.cur[l].size() != quad.prev[l].size()) same = false;
            else
                for (size_t i = 0; i < quad.cur[l].size(); i++) {
                    const glsl::SiteRec &a = quad.cur[l][i], &b = quad.prev[l][i];
                    if (a.key != b.key || memcmp(&a.u, &b.u, 4) || memcmp(&a.v, &b.v, 4)) { same = false; break; }
                }
        }
        if (!any || (pass > 0 && same)) break;
        for (int l = 0; l < 4; l++) quad.prev[l] = quad.cur[l];
    }
    for (int l = 0; l < 4; l++) { out[l * 4 + 0] = col[l].x; out[l * 4 + 1] = col[l].y; out[l * 4 + 2] = col[l].z; out[l * 4 + 3] = col[l].w; }
}

}  // namespace

struct ref_handle { RefScene S; };

extern "C" {

ref_handle* ref_create(const orc_scene_desc* d) {
    if (!d || !d->scene) return nullptr;
    ref_handle* h = new ref_handle();
    RefScene& S = h->S;
    const rtb_defines& D = d->defines;
    S.U.sphere_size = D.sphere_size; S.U.plane_size = D.plane_size; S.U.surface_size = D.surface_size; S.U.box_size = D.box_size;
    S.U.torus_size = D.torus_size; S.U.ring_size = D.ring_size; S.U.light_point_size = D.light_point_size;
    S.U.light_direct_size = D.light_direct_size; S.U.iterations = D.iterations;
    S.U.ambient_color = glsl::vec3(round_through_percent_f(D.ambient_color[0]), round_through_percent_f(D.ambient_color[1]), round_through_percent_f(D.ambient_color[2]));
    S.U.shadow_ambient = glsl::vec3(round_through_percent_f(D.shadow_ambient[0]), round_through_percent_f(D.shadow_ambient[1]), round_through_percent_f(D.shadow_ambient[2]));
    S.scene.quat_camera_rotation = V4(d->scene->quat_camera_rotation);
    S.scene.camera_pos = V3(d->scene->camera_pos);
    S.scene.bg_color = V3(d->scene->bg_color);
    S.scene.canvas_width = d->scene->canvas_width; S.scene.canvas_height = d->scene->canvas_height;
    S.scene.reflect_depth = d->scene->reflect_depth;
    for (int i = 0; i < D.sphere_size; i++) {
        const rtb_sphere& s = d->spheres[i]; RefShader::rt_sphere r;
        r.mat = M(s.material); r.obj = V4(s.obj); r.quat_rotation = V4(s.quat_rotation); r.textureNum = s.textureNum; r.hollow = s.hollow != 0;
        S.spheres.push_back(r);
    }
    for (int i = 0; i < D.plane_size; i++) {
        const rtb_plane& s = d->planes[i]; RefShader::rt_plane r;
        r.mat = M(s.material); r.pos = V3(s.pos); r.normal = V3(s.normal);
        S.planes.push_back(r);
    }
    for (int i = 0; i < D.surface_size; i++) {
        const rtb_surface& s = d->surfaces[i]; RefShader::rt_surface r;
        r.mat = M(s.mat); r.quat_rotation = V4(s.quat_rotation); r.v_min = V3(s.v_min); r.v_max = V3(s.v_max); r.pos = V3(s.pos);
        r.a = s.a; r.b = s.b; r.c = s.c; r.d = s.d; r.e = s.e; r.f = s.f;
        S.surfaces.push_back(r);
    }
    for (int i = 0; i < D.box_size; i++) {
        const rtb_box& s = d->boxes[i]; RefShader::rt_box r;
        r.mat = M(s.mat); r.quat_rotation = V4(s.quat_rotation); r.pos = V3(s.pos); r.form = V3(s.form); r.textureNum = s.textureNum;
        S.boxes.push_back(r);
    }
    for (int i = 0; i < D.torus_size; i++) {
        const rtb_torus& s = d->toruses[i]; RefShader::rt_torus r;
        r.mat = M(s.mat); r.quat_rotation = V4(s.quat_rotation); r.pos = V3(s.pos); r.form = glsl::vec2(s.form[0], s.form[1]);
        S.toruses.push_back(r);
    }
    for (int i = 0; i < D.ring_size; i++) {
        const rtb_ring& s = d->rings[i]; RefShader::rt_ring r;
        r.mat = M(s.mat); r.quat_rotation = V4(s.quat_rotation); r.pos = V3(s.pos); r.textureNum = s.textureNum; r.r1 = s.r1; r.r2 = s.r2;
        S.rings.push_back(r);
    }
    for (int i = 0; i < D.light_point_size; i++) {
        const rtb_light_point& s = d->lights_point[i]; RefShader::rt_light_point r;
        r.pos = V4(s.pos); r.color = V3(s.color); r.intensity = s.intensity; r.linear_k = s.linear_k; r.quadratic_k = s.quadratic_k;
        S.lights_point.push_back(r);
    }
    for (int i = 0; i < D.light_direct_size; i++) {
        const rtb_light_direct& s = d->lights_direct[i]; RefShader::rt_light_direct r;
        r.direction = V3(s.direction); r.color = V3(s.color); r.intensity = s.intensity;
        S.lights_direct.push_back(r);
    }
    if (d->cube[0].px) {
        S.cube.w = d->cube[0].w; S.cube.h = d->cube[0].h;
        for (int f = 0; f < 6; f++) glsim::expand_rgba8(d->cube[f].px, d->cube[f].w, d->cube[f].h, d->cube[f].ch, S.cube.face[f]);
    }
    for (int u = 1; u <= 5; u++)
        if (d->tex2d[u].px) glsim::build_mips(S.tex[u], d->tex2d[u].px, d->tex2d[u].w, d->tex2d[u].h, d->tex2d[u].ch);
    return h;
}

void ref_destroy(ref_handle* h) { delete h; }

int ref_render_quads(ref_handle* h, int n, const int32_t* qx, const int32_t* qy, float* out, int n_threads) {
    if (!h || n < 0) return -1;
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt > n) nt = n > 0 ? n : 1;
    std::atomic<int> next(0);
    auto work = [&]() {
        for (;;) {
            int b = next.fetch_add(16);
            if (b >= n) break;
            int e = b + 16 < n ? b + 16 : n;
            for (int i = b; i < e; i++) render_quad(h->S, qx[i], qy[i], out + (size_t)i * 16);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    return 0;
}

int ref_render(ref_handle* h, int x0, int y0, int w, int hgt, float* out, int n_threads) {
    if (!h || (x0 | y0 | w | hgt) & 1 || w <= 0 || hgt <= 0) return -1;
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    int qrows = hgt / 2, qcols = w / 2;
    std::atomic<int> next(0);
    auto work = [&]() {
        for (;;) {
            int r = next.fetch_add(1);
            if (r >= qrows) break;
            for (int c = 0; c < qcols; c++) {
                float px[16];
                render_quad(h->S, x0 + 2 * c, y0 + 2 * r, px);
                for (int l = 0; l < 4; l++)
                    memcpy(out + ((size_t)(2 * r + (l >> 1)) * w + (size_t)(2 * c + (l & 1))) * 4, px + l * 4, 16);
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    return 0;
}

/* single-function probes for KATs against the restatement */
float ref_calc_inter(ref_handle* h, const float ro[3], const float rd[3], int32_t* num, int32_t* type) {
    RefShader sh(h->S.U);
    bind(sh, h->S);
    int n = *num, t = *type;
    float tm = sh.calcInter(V3(ro), V3(rd), n, t);
    *num = n; *type = t;
    return tm;
}

float ref_in_shadow(ref_handle* h, const float ro[3], const float rd[3], float dist) {
    RefShader sh(h->S.U);
    bind(sh, h->S);
    return sh.inShadow(V3(ro), V3(rd), dist);
}

}  // extern "C"

/* fakegl.cpp — a minimal OpenGL 3.3 "driver" plus a headless GLFW, behind which the reference's OWN GLWrapper.cpp
 * (and main.cpp, SceneManager.cpp) run UNCHANGED: SURVEY.md 8f-4.
 *
 * The reference's GLWrapper.cpp talks to OpenGL through 45 glad function pointers and to the window system through
 * 22 GLFW calls (nm of the unchanged objects).  This file defines exactly those symbols.  It keeps the GL object
 * state the ray-trace pass depends on —
 *     program objects     the fragment source after GLWrapper::init_shaders' {TOKEN} substitution (GLWrapper.cpp:237-247):
 *                         the eleven `#define`s of rt.frag:122-132 are parsed back into an rtb_defines
 *     uniform blocks      glGetUniformBlockIndex / glUniformBlockBinding / glBindBufferBase (GLWrapper.cpp:365-379)
 *     buffer objects      the bytes of glBufferData / glBufferSubData (GLWrapper.cpp:369,381-386)
 *     sampler uniforms    glGetUniformLocation + glUniform1i (shader.h setInt): sampler name -> texture unit
 *     texture objects     the texels of glTexImage2D (cube faces GLWrapper.cpp:296-299, 2-D :336), per-unit bindings
 * — and when glDrawArrays runs with the ray-trace program current (GLWrapper.cpp:165) it hands that state to
 * librtb200.so through the C-ABI (include/rtb200.h) and launches the kernel.  Draws with any other program (the three
 * SMAA passes, GLWrapper.cpp:173-204) are accepted and ignored: the SMAA post-pass is out of scope (BASELINE.json).
 * glfwSwapBuffers presents the frame (RT_DUMP_DIR: written as .npy, like the replacement GLWrapper does).
 *
 * Environment (same meaning as in host/GLWrapper.cpp): RT_WIDTH, RT_HEIGHT (the size the "window system" grants,
 * GLWrapper.cpp:86), RT_ITERATIONS, RT_FRAMES, RT_DEVICE, RT_STRICT, RT_KERNEL, RT_DUMP_DIR, RT_VERBOSE.
 * RT_FAKEGL_CAPTURE_ONLY=1: at the first ray-trace draw dump the captured state to RT_DUMP_DIR and exit(0) WITHOUT
 * touching CUDA (used by the CPU test that compares the capture with what the replacement GLWrapper uploads).
 */
#include <glad/glad.h>
#ifndef GLFW_INCLUDE_NONE
#define GLFW_INCLUDE_NONE
#endif
#include <GLFW/glfw3.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../../include/rtb200.h"

struct GLFWwindow {
    int width = 0, height = 0;
    void* user = nullptr;
    int frames_left = 1;
    int should_close = 0;
};
struct GLFWmonitor { int dummy; };

namespace {

/* ------------------------------------------------------------------ GL objects */
struct Buffer { std::vector<uint8_t> bytes; unsigned version = 0; };
struct Image { std::vector<uint8_t> px; int w = 0, h = 0, ch = 0; };
struct Texture { GLenum target = 0; Image face[6]; unsigned version = 0; };       /* face[0] = the 2-D image */
struct ShaderObj { GLenum type = 0; std::string src; };
struct Program {
    std::vector<GLuint> shaders;
    std::string frag;
    bool is_rt = false;
    rtb_defines defines;
    int block_binding[RTB_NUM_BINDINGS];                /* uniform block index -> binding point (GL default 0) */
    std::map<GLint, int> uniform_i;                     /* location -> value */
};

const char* const kBlockNames[RTB_NUM_BINDINGS] = { "scene_buf", "spheres_buf", "planes_buf", "surfaces_buf", "boxes_buf",
                                                    "toruses_buf", "rings_buf", "lights_point_buf", "lights_direct_buf" };
/* sampler2D uniforms of rt.frag:138-143 -> the texture roles of the C-ABI (rtb_set_texture2d unit 1..5 = main.cpp:149-153);
 * texture_sphere_4 exists in the shader but nothing binds it (main.cpp): not mapped */
const struct { const char* name; int role; } kSamplers[] = { { "texture_sphere_1", 1 }, { "texture_sphere_2", 2 }, { "texture_sphere_3", 3 },
                                                              { "texture_ring", 4 }, { "texture_box", 5 } };

std::map<GLuint, Buffer> g_buffers;
std::map<GLuint, Texture> g_textures;
std::map<GLuint, ShaderObj> g_shaders;
std::map<GLuint, Program> g_programs;
std::map<std::pair<GLuint, std::string>, GLint> g_locations;    /* (program, uniform name) -> location */
std::map<GLint, GLuint> g_location_program;
GLuint g_next_id = 1;
GLint g_next_location = 1;

GLuint g_current_program = 0, g_bound_array = 0, g_bound_uniform = 0, g_bound_fbo = 0;
GLenum g_active_unit = 0;                                        /* index, not GL_TEXTUREi */
GLuint g_unit_2d[32] = { 0 }, g_unit_cube[32] = { 0 };
std::map<GLuint, GLuint> g_buffer_base;                          /* binding point -> buffer */

/* ------------------------------------------------------------------ the renderer behind the driver */
rtb_ctx* g_ctx = nullptr;
GLFWwindow* g_window = nullptr;
bool g_defines_sent = false;
unsigned g_sent_block_version[RTB_NUM_BINDINGS] = { 0 };
GLuint g_sent_block_buffer[RTB_NUM_BINDINGS] = { 0 };
GLuint g_sent_cube = 0; unsigned g_sent_cube_version = 0;
GLuint g_sent_tex[6] = { 0 }; unsigned g_sent_tex_version[6] = { 0 };
int g_frame_index = 0;
long g_frames_presented = 0;
std::chrono::steady_clock::time_point g_first_present;

int env_int(const char* name, int fallback) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : fallback;
}
std::string dump_dir() {
    const char* d = getenv("RT_DUMP_DIR");
    return d ? std::string(d) : std::string();
}
[[noreturn]] void die(const char* what) {
    fprintf(stderr, "fakegl: %s: %s\n", what, rtb_last_error(g_ctx));
    exit(1);                                                     /* the reference's error convention (utils.h, GLWrapper.cpp:374) */
}
void write_npy(const std::string& path, const char* descr, const std::vector<size_t>& shape, const void* data, size_t bytes) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { fprintf(stderr, "cannot write %s\n", path.c_str()); return; }
    std::string sh = "(";
    for (size_t d : shape) sh += std::to_string(d) + ",";
    sh += ")";
    std::string hdr = std::string("{'descr': '") + descr + "', 'fortran_order': False, 'shape': " + sh + ", }";
    size_t total = 10 + hdr.size() + 1;
    hdr += std::string((64 - total % 64) % 64, ' ') + "\n";
    unsigned short hl = (unsigned short)hdr.size();
    fwrite("\x93NUMPY\x01\x00", 1, 8, f);
    fwrite(&hl, 2, 1, f);
    fwrite(hdr.data(), 1, hdr.size(), f);
    if (bytes) fwrite(data, 1, bytes, f);
    fclose(f);
}

/* `#define NAME value` of the substituted fragment source -> text of value ("" if absent) */
std::string define_text(const std::string& src, const char* name) {
    const std::string key = std::string("#define ") + name + " ";
    size_t p = src.find(key);
    if (p == std::string::npos) return "";
    p += key.size();
    size_t e = src.find_first_of("\r\n", p);
    return src.substr(p, e == std::string::npos ? std::string::npos : e - p);
}
bool parse_vec3(const std::string& t, float out[3]) {           /* "vec3(0.025000,0.025000,0.025000)" (GLWrapper.cpp:279-282) */
    return sscanf(t.c_str(), " vec3( %f , %f , %f )", &out[0], &out[1], &out[2]) == 3;
}
/* the ray-trace program is the one whose fragment source carries the specialisation defines and the nine uniform blocks */
void classify_program(Program& p) {
    memset(&p.defines, 0, sizeof p.defines);
    const char* counts[8] = { "SPHERE_SIZE", "PLANE_SIZE", "SURFACE_SIZE", "BOX_SIZE", "TORUS_SIZE", "RING_SIZE", "LIGHT_POINT_SIZE", "LIGHT_DIRECT_SIZE" };
    int32_t* dst[8] = { &p.defines.sphere_size, &p.defines.plane_size, &p.defines.surface_size, &p.defines.box_size,
                        &p.defines.torus_size, &p.defines.ring_size, &p.defines.light_point_size, &p.defines.light_direct_size };
    bool ok = p.frag.find("uniform scene_buf") != std::string::npos;
    for (int i = 0; i < 8 && ok; i++) {
        const std::string t = define_text(p.frag, counts[i]);
        if (t.empty() || t[0] == '{') ok = false; else *dst[i] = atoi(t.c_str());
    }
    if (ok) {
        const std::string it = define_text(p.frag, "ITERATIONS");
        ok = !it.empty() && it[0] != '{' && parse_vec3(define_text(p.frag, "AMBIENT_COLOR"), p.defines.ambient_color) &&
             parse_vec3(define_text(p.frag, "SHADOW_AMBIENT"), p.defines.shadow_ambient);
        if (ok) p.defines.iterations = atoi(it.c_str());
    }
    p.is_rt = ok;
}

int channels_of(GLenum format) {
    switch (format) {
        case GL_RED: return 1;
        case GL_RG: return 2;
        case GL_RGB: return 3;
        case GL_RGBA: return 4;
        default: return 0;
    }
}

/* the state the current ray-trace draw reads, resolved the way GL resolves it */
struct Resolved {
    const Program* prog = nullptr;
    const Buffer* block[RTB_NUM_BINDINGS] = { nullptr };
    GLuint block_buffer[RTB_NUM_BINDINGS] = { 0 };
    GLuint cube = 0;
    GLuint tex[6] = { 0 };
};
Resolved resolve(const Program& p, GLuint program_id) {
    Resolved r;
    r.prog = &p;
    for (int b = 0; b < RTB_NUM_BINDINGS; b++) {
        auto base = g_buffer_base.find((GLuint)p.block_binding[b]);
        if (base == g_buffer_base.end()) continue;
        auto buf = g_buffers.find(base->second);
        if (buf != g_buffers.end()) { r.block[b] = &buf->second; r.block_buffer[b] = base->second; }
    }
    auto unit_of = [&](const char* sampler) {
        auto loc = g_locations.find(std::make_pair(program_id, std::string(sampler)));
        if (loc == g_locations.end()) return 0;                  /* never queried: GL's default value of a sampler is unit 0 */
        auto v = p.uniform_i.find(loc->second);
        return v == p.uniform_i.end() ? 0 : v->second;
    };
    const int cu = unit_of("skybox");
    if (cu >= 0 && cu < 32) r.cube = g_unit_cube[cu];
    for (const auto& s : kSamplers) {
        if (g_locations.find(std::make_pair(program_id, std::string(s.name))) == g_locations.end()) continue;   /* sampler never set by the application */
        const int u = unit_of(s.name);
        if (u >= 0 && u < 32) r.tex[s.role] = g_unit_2d[u];
    }
    return r;
}

void dump_capture(const Resolved& r) {
    const std::string dd = dump_dir();
    if (dd.empty()) return;
    write_npy(dd + "/defines.npy", "|u1", { sizeof(rtb_defines) }, &r.prog->defines, sizeof(rtb_defines));
    for (int b = 0; b < RTB_NUM_BINDINGS; b++)
        if (r.block[b] && !r.block[b]->bytes.empty())
            write_npy(dd + "/" + kBlockNames[b] + ".npy", "|u1", { r.block[b]->bytes.size() }, r.block[b]->bytes.data(), r.block[b]->bytes.size());
    auto cube = g_textures.find(r.cube);
    if (cube != g_textures.end())
        for (int f = 0; f < 6; f++) {
            const Image& im = cube->second.face[f];
            if (!im.px.empty()) write_npy(dd + "/cube_" + std::to_string(f) + ".npy", "|u1", { (size_t)im.h, (size_t)im.w, (size_t)im.ch }, im.px.data(), im.px.size());
        }
    for (int role = 1; role <= 5; role++) {
        auto t = g_textures.find(r.tex[role]);
        if (t == g_textures.end() || t->second.face[0].px.empty()) continue;
        const Image& im = t->second.face[0];
        write_npy(dd + "/tex_" + std::to_string(role) + ".npy", "|u1", { (size_t)im.h, (size_t)im.w, (size_t)im.ch }, im.px.data(), im.px.size());
    }
}

/* glDrawArrays with the ray-trace program: bring the renderer up to date with the GL state, launch */
void draw_raytrace(GLuint program_id) {
    Program& p = g_programs[program_id];
    const Resolved r = resolve(p, program_id);
    if (env_int("RT_FAKEGL_CAPTURE_ONLY", 0)) {
        dump_capture(r);
        printf("fakegl: captured the state of the first ray-trace draw (%dx%d), not rendering\n", g_window ? g_window->width : 0, g_window ? g_window->height : 0);
        exit(0);
    }
    if (!g_ctx) {
        if (!g_window) { fprintf(stderr, "fakegl: draw without a window\n"); exit(1); }
        g_ctx = rtb_create(g_window->width, g_window->height, env_int("RT_DEVICE", 0));
        if (!g_ctx) { fprintf(stderr, "rtb_create failed: %s\n", rtb_last_error(nullptr)); exit(1); }
        rtb_set_option(g_ctx, "strict", env_int("RT_STRICT", 1));
        rtb_set_option(g_ctx, "kernel", env_int("RT_KERNEL", 0));
        printf("rtb200 %s behind the fake GL driver, %dx%d\n", rtb_version(), g_window->width, g_window->height);
    }
    if (!g_defines_sent) {
        rtb_defines d = p.defines;
        d.iterations = env_int("RT_ITERATIONS", d.iterations);
        if (rtb_set_defines(g_ctx, &d)) die("Shader creation");
        g_defines_sent = true;
        if (g_frame_index == 0) dump_capture(r);
    }
    for (int b = 0; b < RTB_NUM_BINDINGS; b++) {
        if (!r.block[b]) continue;
        if (g_sent_block_buffer[b] == r.block_buffer[b] && g_sent_block_version[b] == r.block[b]->version) continue;
        if (rtb_upload(g_ctx, b, r.block[b]->bytes.empty() ? nullptr : r.block[b]->bytes.data(), r.block[b]->bytes.size())) die("uniform buffer upload");
        g_sent_block_buffer[b] = r.block_buffer[b];
        g_sent_block_version[b] = r.block[b]->version;
    }
    auto cube = g_textures.find(r.cube);
    if (cube != g_textures.end() && (g_sent_cube != r.cube || g_sent_cube_version != cube->second.version)) {
        const Texture& t = cube->second;
        bool complete = true;
        const uint8_t* faces[6];
        for (int f = 0; f < 6; f++) {
            complete = complete && !t.face[f].px.empty() && t.face[f].w == t.face[0].w && t.face[f].h == t.face[0].h && t.face[f].ch == t.face[0].ch;
            faces[f] = t.face[f].px.data();
        }
        if (complete && rtb_set_cubemap(g_ctx, faces, t.face[0].w, t.face[0].h, t.face[0].ch)) die("set_skybox");   /* an incomplete cube map samples as black in GL: left unbound */
        g_sent_cube = r.cube; g_sent_cube_version = t.version;
    }
    for (int role = 1; role <= 5; role++) {
        auto t = g_textures.find(r.tex[role]);
        if (t == g_textures.end() || t->second.face[0].px.empty()) continue;
        if (g_sent_tex[role] == r.tex[role] && g_sent_tex_version[role] == t->second.version) continue;
        const Image& im = t->second.face[0];
        if (rtb_set_texture2d(g_ctx, role, im.px.data(), im.w, im.h, im.ch)) die("load_texture");
        g_sent_tex[role] = r.tex[role]; g_sent_tex_version[role] = t->second.version;
    }
    if (rtb_render(g_ctx)) die("Draw raytraced image");
}

void present() {
    if (!g_ctx) return;
    const std::string dd = dump_dir();
    if (!dd.empty() && g_window) {
        std::vector<float> px((size_t)g_window->width * g_window->height * 4);
        if (rtb_read_rgba32f(g_ctx, px.data())) die("read frame");
        char name[64];
        snprintf(name, sizeof name, "/frame_%04d.npy", g_frame_index);
        write_npy(dd + name, "<f4", { (size_t)g_window->height, (size_t)g_window->width, 4 }, px.data(), px.size() * sizeof(float));
    } else {
        rtb_sync(g_ctx);
    }
    rtb_stats st;
    if (!rtb_get_stats(g_ctx, &st) && (g_frame_index < 3 || getenv("RT_VERBOSE"))) printf("frame %d: kernel %d, %.3f ms\n", g_frame_index, st.kernel_used, st.kernel_ms);
    g_frame_index++;
}

/* ------------------------------------------------------------------ the 45 GL entry points */
void APIENTRY fk_ActiveTexture(GLenum texture) { g_active_unit = texture >= GL_TEXTURE0 ? texture - GL_TEXTURE0 : 0; }
void APIENTRY fk_AttachShader(GLuint program, GLuint shader) { g_programs[program].shaders.push_back(shader); }
void APIENTRY fk_BindBuffer(GLenum target, GLuint buffer) {
    if (target == GL_UNIFORM_BUFFER) g_bound_uniform = buffer; else if (target == GL_ARRAY_BUFFER) g_bound_array = buffer;
    if (buffer) g_buffers[buffer];
}
void APIENTRY fk_BindBufferBase(GLenum target, GLuint index, GLuint buffer) {
    if (target != GL_UNIFORM_BUFFER) return;
    g_buffer_base[index] = buffer;
    g_bound_uniform = buffer;                                    /* glBindBufferBase also binds the generic target */
}
void APIENTRY fk_BindFramebuffer(GLenum, GLuint framebuffer) { g_bound_fbo = framebuffer; }
void APIENTRY fk_BindTexture(GLenum target, GLuint texture) {
    if (g_active_unit >= 32) return;
    if (target == GL_TEXTURE_2D) g_unit_2d[g_active_unit] = texture;
    else if (target == GL_TEXTURE_CUBE_MAP) g_unit_cube[g_active_unit] = texture;
    if (texture) { Texture& t = g_textures[texture]; if (!t.target) t.target = target; }
}
void APIENTRY fk_BindVertexArray(GLuint) {}
GLuint bound_buffer(GLenum target) { return target == GL_UNIFORM_BUFFER ? g_bound_uniform : target == GL_ARRAY_BUFFER ? g_bound_array : 0; }
void APIENTRY fk_BufferData(GLenum target, GLsizeiptr size, const void* data, GLenum) {
    const GLuint id = bound_buffer(target);
    if (!id) return;
    Buffer& b = g_buffers[id];
    b.bytes.assign((size_t)(size > 0 ? size : 0), 0);
    if (data && size > 0) memcpy(b.bytes.data(), data, (size_t)size);
    b.version++;
}
void APIENTRY fk_BufferSubData(GLenum target, GLintptr offset, GLsizeiptr size, const void* data) {
    const GLuint id = bound_buffer(target);
    if (!id || !data || size <= 0) return;
    Buffer& b = g_buffers[id];
    if (offset < 0 || (size_t)offset + (size_t)size > b.bytes.size()) return;      /* GL_INVALID_VALUE: no effect */
    memcpy(b.bytes.data() + offset, data, (size_t)size);
    b.version++;
}
GLenum APIENTRY fk_CheckFramebufferStatus(GLenum) { return GL_FRAMEBUFFER_COMPLETE; }
void APIENTRY fk_Clear(GLbitfield) {}
void APIENTRY fk_ClearColor(GLfloat, GLfloat, GLfloat, GLfloat) {}
void APIENTRY fk_CompileShader(GLuint) {}
GLuint APIENTRY fk_CreateProgram(void) {
    const GLuint id = g_next_id++;
    Program& p = g_programs[id];
    for (int b = 0; b < RTB_NUM_BINDINGS; b++) p.block_binding[b] = 0;
    return id;
}
GLuint APIENTRY fk_CreateShader(GLenum type) { const GLuint id = g_next_id++; g_shaders[id].type = type; return id; }
void APIENTRY fk_DeleteBuffers(GLsizei n, const GLuint* ids) { for (GLsizei i = 0; i < n; i++) g_buffers.erase(ids[i]); }
void APIENTRY fk_DeleteFramebuffers(GLsizei, const GLuint*) {}
void APIENTRY fk_DeleteProgram(GLuint program) { g_programs.erase(program); }
void APIENTRY fk_DeleteShader(GLuint) {}                         /* sources are kept: the program was linked from them */
void APIENTRY fk_DeleteTextures(GLsizei n, const GLuint* ids) { for (GLsizei i = 0; i < n; i++) g_textures.erase(ids[i]); }
void APIENTRY fk_DeleteVertexArrays(GLsizei, const GLuint*) {}
void APIENTRY fk_DrawArrays(GLenum, GLint, GLsizei) {
    auto p = g_programs.find(g_current_program);
    if (p != g_programs.end() && p->second.is_rt) draw_raytrace(g_current_program);
}
void APIENTRY fk_EnableVertexAttribArray(GLuint) {}
void APIENTRY fk_FramebufferTexture2D(GLenum, GLenum, GLenum, GLuint, GLint) {}
void gen_ids(GLsizei n, GLuint* ids) { for (GLsizei i = 0; i < n; i++) ids[i] = g_next_id++; }
void APIENTRY fk_GenBuffers(GLsizei n, GLuint* ids) { gen_ids(n, ids); }
void APIENTRY fk_GenFramebuffers(GLsizei n, GLuint* ids) { gen_ids(n, ids); }
void APIENTRY fk_GenTextures(GLsizei n, GLuint* ids) { gen_ids(n, ids); }
void APIENTRY fk_GenVertexArrays(GLsizei n, GLuint* ids) { gen_ids(n, ids); }
void APIENTRY fk_GenerateMipmap(GLenum) {}                       /* librtb200 builds the chain itself (rtb_set_texture2d) */
GLenum APIENTRY fk_GetError(void) { return GL_NO_ERROR; }
void APIENTRY fk_GetProgramInfoLog(GLuint, GLsizei bufSize, GLsizei* length, GLchar* infoLog) { if (bufSize > 0 && infoLog) infoLog[0] = 0; if (length) *length = 0; }
void APIENTRY fk_GetProgramiv(GLuint, GLenum pname, GLint* params) { if (params) *params = (pname == GL_INFO_LOG_LENGTH) ? 0 : GL_TRUE; }
void APIENTRY fk_GetShaderInfoLog(GLuint, GLsizei bufSize, GLsizei* length, GLchar* infoLog) { if (bufSize > 0 && infoLog) infoLog[0] = 0; if (length) *length = 0; }
void APIENTRY fk_GetShaderiv(GLuint, GLenum pname, GLint* params) { if (params) *params = (pname == GL_INFO_LOG_LENGTH) ? 0 : GL_TRUE; }
GLuint APIENTRY fk_GetUniformBlockIndex(GLuint program, const GLchar* name) {
    auto p = g_programs.find(program);
    if (p == g_programs.end() || !name) return GL_INVALID_INDEX;
    for (int b = 0; b < RTB_NUM_BINDINGS; b++)
        if (!strcmp(kBlockNames[b], name) && p->second.frag.find(std::string("uniform ") + name) != std::string::npos) return (GLuint)b;
    return GL_INVALID_INDEX;                                     /* GLWrapper.cpp:371-375 exits on this */
}
GLint APIENTRY fk_GetUniformLocation(GLuint program, const GLchar* name) {
    if (!name || g_programs.find(program) == g_programs.end()) return -1;
    const auto key = std::make_pair(program, std::string(name));
    auto it = g_locations.find(key);
    if (it != g_locations.end()) return it->second;
    const GLint loc = g_next_location++;
    g_locations[key] = loc;
    g_location_program[loc] = program;
    return loc;
}
void APIENTRY fk_LinkProgram(GLuint program) {
    Program& p = g_programs[program];
    for (GLuint s : p.shaders) {
        auto so = g_shaders.find(s);
        if (so != g_shaders.end() && so->second.type == GL_FRAGMENT_SHADER) p.frag = so->second.src;
    }
    classify_program(p);
}
void APIENTRY fk_ShaderSource(GLuint shader, GLsizei count, const GLchar* const* string, const GLint* length) {
    std::string src;
    for (GLsizei i = 0; i < count; i++) {
        if (!string[i]) continue;
        if (length && length[i] >= 0) src.append(string[i], (size_t)length[i]); else src.append(string[i]);
    }
    g_shaders[shader].src = src;
}
void APIENTRY fk_TexImage2D(GLenum target, GLint level, GLint, GLsizei width, GLsizei height, GLint, GLenum format, GLenum type, const void* pixels) {
    if (level != 0 || g_active_unit >= 32) return;
    int face = 0;
    GLuint id = 0;
    if (target == GL_TEXTURE_2D) id = g_unit_2d[g_active_unit];
    else if (target >= GL_TEXTURE_CUBE_MAP_POSITIVE_X && target < GL_TEXTURE_CUBE_MAP_POSITIVE_X + 6) { id = g_unit_cube[g_active_unit]; face = (int)(target - GL_TEXTURE_CUBE_MAP_POSITIVE_X); }
    if (!id) return;
    Texture& t = g_textures[id];
    Image& im = t.face[face];
    im.w = width; im.h = height; im.ch = channels_of(format);
    im.px.clear();
    if (pixels && type == GL_UNSIGNED_BYTE && im.ch > 0 && width > 0 && height > 0)
        im.px.assign((const uint8_t*)pixels, (const uint8_t*)pixels + (size_t)width * height * im.ch);       /* stb rows are tightly packed (GL_UNPACK_ALIGNMENT is left at 4 by the reference; its textures have 4-byte multiples per row) */
    t.version++;
}
void APIENTRY fk_TexParameteri(GLenum, GLenum, GLint) {}        /* the sampler state of the reference is fixed (GLWrapper.cpp:308-314,340-343) and modelled in the kernels */
void APIENTRY fk_Uniform1i(GLint location, GLint v0) {
    auto lp = g_location_program.find(location);
    if (lp == g_location_program.end() || lp->second != g_current_program) return;   /* GL_INVALID_OPERATION: not a location of the current program */
    g_programs[g_current_program].uniform_i[location] = v0;
}
void APIENTRY fk_UniformBlockBinding(GLuint program, GLuint blockIndex, GLuint binding) {
    auto p = g_programs.find(program);
    if (p != g_programs.end() && blockIndex < RTB_NUM_BINDINGS) p->second.block_binding[blockIndex] = (int)binding;
}
void APIENTRY fk_UseProgram(GLuint program) { g_current_program = program; }
void APIENTRY fk_VertexAttribPointer(GLuint, GLint, GLenum, GLboolean, GLsizei, const void*) {}
void APIENTRY fk_Viewport(GLint, GLint, GLsizei, GLsizei) {}

}  // namespace

extern "C" {

struct gladGLversionStruct GLVersion = { 0, 0 };

PFNGLACTIVETEXTUREPROC glad_glActiveTexture = fk_ActiveTexture;
PFNGLATTACHSHADERPROC glad_glAttachShader = fk_AttachShader;
PFNGLBINDBUFFERPROC glad_glBindBuffer = fk_BindBuffer;
PFNGLBINDBUFFERBASEPROC glad_glBindBufferBase = fk_BindBufferBase;
PFNGLBINDFRAMEBUFFERPROC glad_glBindFramebuffer = fk_BindFramebuffer;
PFNGLBINDTEXTUREPROC glad_glBindTexture = fk_BindTexture;
PFNGLBINDVERTEXARRAYPROC glad_glBindVertexArray = fk_BindVertexArray;
PFNGLBUFFERDATAPROC glad_glBufferData = fk_BufferData;
PFNGLBUFFERSUBDATAPROC glad_glBufferSubData = fk_BufferSubData;
PFNGLCHECKFRAMEBUFFERSTATUSPROC glad_glCheckFramebufferStatus = fk_CheckFramebufferStatus;
PFNGLCLEARPROC glad_glClear = fk_Clear;
PFNGLCLEARCOLORPROC glad_glClearColor = fk_ClearColor;
PFNGLCOMPILESHADERPROC glad_glCompileShader = fk_CompileShader;
PFNGLCREATEPROGRAMPROC glad_glCreateProgram = fk_CreateProgram;
PFNGLCREATESHADERPROC glad_glCreateShader = fk_CreateShader;
PFNGLDELETEBUFFERSPROC glad_glDeleteBuffers = fk_DeleteBuffers;
PFNGLDELETEFRAMEBUFFERSPROC glad_glDeleteFramebuffers = fk_DeleteFramebuffers;
PFNGLDELETEPROGRAMPROC glad_glDeleteProgram = fk_DeleteProgram;
PFNGLDELETESHADERPROC glad_glDeleteShader = fk_DeleteShader;
PFNGLDELETETEXTURESPROC glad_glDeleteTextures = fk_DeleteTextures;
PFNGLDELETEVERTEXARRAYSPROC glad_glDeleteVertexArrays = fk_DeleteVertexArrays;
PFNGLDRAWARRAYSPROC glad_glDrawArrays = fk_DrawArrays;
PFNGLENABLEVERTEXATTRIBARRAYPROC glad_glEnableVertexAttribArray = fk_EnableVertexAttribArray;
PFNGLFRAMEBUFFERTEXTURE2DPROC glad_glFramebufferTexture2D = fk_FramebufferTexture2D;
PFNGLGENBUFFERSPROC glad_glGenBuffers = fk_GenBuffers;
PFNGLGENFRAMEBUFFERSPROC glad_glGenFramebuffers = fk_GenFramebuffers;
PFNGLGENTEXTURESPROC glad_glGenTextures = fk_GenTextures;
PFNGLGENVERTEXARRAYSPROC glad_glGenVertexArrays = fk_GenVertexArrays;
PFNGLGENERATEMIPMAPPROC glad_glGenerateMipmap = fk_GenerateMipmap;
PFNGLGETERRORPROC glad_glGetError = fk_GetError;
PFNGLGETPROGRAMINFOLOGPROC glad_glGetProgramInfoLog = fk_GetProgramInfoLog;
PFNGLGETPROGRAMIVPROC glad_glGetProgramiv = fk_GetProgramiv;
PFNGLGETSHADERINFOLOGPROC glad_glGetShaderInfoLog = fk_GetShaderInfoLog;
PFNGLGETSHADERIVPROC glad_glGetShaderiv = fk_GetShaderiv;
PFNGLGETUNIFORMBLOCKINDEXPROC glad_glGetUniformBlockIndex = fk_GetUniformBlockIndex;
PFNGLGETUNIFORMLOCATIONPROC glad_glGetUniformLocation = fk_GetUniformLocation;
PFNGLLINKPROGRAMPROC glad_glLinkProgram = fk_LinkProgram;
PFNGLSHADERSOURCEPROC glad_glShaderSource = fk_ShaderSource;
PFNGLTEXIMAGE2DPROC glad_glTexImage2D = fk_TexImage2D;
PFNGLTEXPARAMETERIPROC glad_glTexParameteri = fk_TexParameteri;
PFNGLUNIFORM1IPROC glad_glUniform1i = fk_Uniform1i;
PFNGLUNIFORMBLOCKBINDINGPROC glad_glUniformBlockBinding = fk_UniformBlockBinding;
PFNGLUSEPROGRAMPROC glad_glUseProgram = fk_UseProgram;
PFNGLVERTEXATTRIBPOINTERPROC glad_glVertexAttribPointer = fk_VertexAttribPointer;
PFNGLVIEWPORTPROC glad_glViewport = fk_Viewport;

int gladLoadGL(void) {                                           /* GLWrapper.cpp:97-101 */
    GLVersion.major = 3;
    GLVersion.minor = 3;
    return 1;
}

/* ------------------------------------------------------------------ the 22 GLFW entry points (headless) */
int glfwInit(void) { return GLFW_TRUE; }
void glfwTerminate(void) {}
GLFWerrorfun glfwSetErrorCallback(GLFWerrorfun) { return nullptr; }
GLFWmonitor* glfwGetPrimaryMonitor(void) { static GLFWmonitor m; return &m; }
const GLFWvidmode* glfwGetVideoMode(GLFWmonitor*) {
    static GLFWvidmode mode;
    mode.width = env_int("RT_WIDTH", 1920); mode.height = env_int("RT_HEIGHT", 1080);
    mode.redBits = mode.greenBits = mode.blueBits = 8; mode.refreshRate = 60;
    return &mode;
}
void glfwWindowHint(int, int) {}
GLFWwindow* glfwCreateWindow(int width, int height, const char*, GLFWmonitor*, GLFWwindow*) {
    GLFWwindow* w = new GLFWwindow();
    /* the window system may grant another size than requested (GLWrapper.cpp:86 reads it back): RT_WIDTH x RT_HEIGHT */
    w->width = env_int("RT_WIDTH", width);
    w->height = env_int("RT_HEIGHT", height);
    const int frames = env_int("RT_FRAMES", 1);
    w->frames_left = frames < 1 ? 1 : frames;
    g_window = w;
    return w;
}
void glfwDestroyWindow(GLFWwindow* w) {
    if (g_frames_presented > 1) {
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - g_first_present).count();
        printf("frame loop: %ld frames in %.3f s after the first = %.3f ms per frame, %.1f frames/s\n", g_frames_presented - 1, s,
               s * 1e3 / (double)(g_frames_presented - 1), (double)(g_frames_presented - 1) / s);
    }
    if (g_ctx) { rtb_destroy(g_ctx); g_ctx = nullptr; }
    if (g_window == w) g_window = nullptr;
    delete w;
}
void glfwGetWindowSize(GLFWwindow* w, int* width, int* height) { if (w) { if (width) *width = w->width; if (height) *height = w->height; } }
void glfwMakeContextCurrent(GLFWwindow*) {}
double glfwGetTime(void) { return (double)g_frames_presented / 60.0; }      /* deterministic 60 Hz clock */
void glfwPollEvents(void) {}
void glfwSwapInterval(int) {}
void glfwSwapBuffers(GLFWwindow* w) {
    present();
    if (g_frames_presented == 0) g_first_present = std::chrono::steady_clock::now();
    g_frames_presented++;
    if (w && --w->frames_left <= 0) w->should_close = 1;
}
int glfwWindowShouldClose(GLFWwindow* w) { return w ? w->should_close : 1; }
void glfwSetWindowShouldClose(GLFWwindow* w, int v) { if (w) w->should_close = v; }
void glfwSetWindowUserPointer(GLFWwindow* w, void* p) { if (w) w->user = p; }
void* glfwGetWindowUserPointer(GLFWwindow* w) { return w ? w->user : nullptr; }
GLFWcursorposfun glfwSetCursorPosCallback(GLFWwindow*, GLFWcursorposfun) { return nullptr; }
GLFWkeyfun glfwSetKeyCallback(GLFWwindow*, GLFWkeyfun) { return nullptr; }
GLFWframebuffersizefun glfwSetFramebufferSizeCallback(GLFWwindow*, GLFWframebuffersizefun) { return nullptr; }
void glfwSetInputMode(GLFWwindow*, int, int) {}

}  // extern "C"

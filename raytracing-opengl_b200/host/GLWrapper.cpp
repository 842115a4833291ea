/* GLWrapper.cpp — the reference-side binding of librtb200.so: each method of the reference's GLWrapper
 * (src/GLWrapper.cpp, cited per method) re-implemented as a call into the C-ABI of include/rtb200.h. */
#include "GLWrapper.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/rtb200.h"
#include "shim_state.h"

#ifdef RTB_HAVE_STB
#include <stb_image.h>          /* the reference's own decoder (external_sources/stb_image), GLWrapper.cpp:293,325 */
#endif
#ifdef RTB_HAVE_SMAA_TABLES
#include <AreaTex.h>            /* the reference's own lookup tables, compiled in from where they lie (SMAA_Builder.h:6-7,45-79) */
#include <SearchTex.h>
#endif

namespace {

struct Image { std::vector<unsigned char> px; int w = 0, h = 0, ch = 0; };
struct CubeImages { Image face[6]; bool ok = true; };

/* load_cubemap() and update_buffer() are static in the reference: they have no `this`, so handles resolve
 * through these process-wide tables. */
std::map<GLuint, CubeImages> g_cubemaps;
std::map<GLuint, std::pair<rtb_ctx*, int>> g_ubos;     /* ubo handle -> (context, binding point) */
GLuint g_next_handle = 1;

int env_int(const char* name, int fallback) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : fallback;
}

[[noreturn]] void die(rtb_ctx* ctx, const char* what) {
    fprintf(stderr, "%s: %s\n", what, rtb_last_error(ctx));
    exit(1);                                            /* the reference's error convention (utils.h:25, GLWrapper.cpp:374) */
}

bool decode(const std::string& path, Image& img) {
#ifdef RTB_HAVE_STB
    unsigned char* data = stbi_load(path.c_str(), &img.w, &img.h, &img.ch, 0);
    if (!data) return false;
    img.px.assign(data, data + (size_t)img.w * img.h * img.ch);
    stbi_image_free(data);
    return true;
#else
    (void)path; (void)img;
    return false;
#endif
}

void write_npy(const std::string& path, const char* descr, const std::vector<size_t>& shape, const void* data, size_t bytes) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { fprintf(stderr, "cannot write %s\n", path.c_str()); return; }
    std::string sh = "(";
    for (size_t d : shape) sh += std::to_string(d) + ",";
    sh += ")";
    std::string hdr = std::string("{'descr': '") + descr + "', 'fortran_order': False, 'shape': " + sh + ", }";
    size_t total = 10 + hdr.size() + 1;
    size_t pad = (64 - total % 64) % 64;
    hdr += std::string(pad, ' ') + "\n";
    unsigned short hl = (unsigned short)hdr.size();
    fwrite("\x93NUMPY\x01\x00", 1, 8, f);
    fwrite(&hl, 2, 1, f);
    fwrite(hdr.data(), 1, hdr.size(), f);
    fwrite(data, 1, bytes, f);
    fclose(f);
}

std::string dump_dir() {
    const char* d = getenv("RT_DUMP_DIR");
    return d ? std::string(d) : std::string();
}

const char* const kBlockNames[RTB_NUM_BINDINGS] = { "scene_buf", "spheres_buf", "planes_buf", "surfaces_buf", "boxes_buf",
                                                    "toruses_buf", "rings_buf", "lights_point_buf", "lights_direct_buf" };

/* RT_DUMP_DIR: keep the latest bytes of every uniform block (what the next draw() will read) */
void dump_block(int binding, const void* data, size_t size) {
    std::string dd = dump_dir();
    if (!dd.empty() && data && size) write_npy(dd + "/" + kBlockNames[binding] + ".npy", "|u1", { size }, data, size);
}

}  // namespace

/* src/GLWrapper.cpp:12-18 */
GLWrapper::GLWrapper(int width, int height, bool fullScreen)
{
	this->width = env_int("RT_WIDTH", width);
	this->height = env_int("RT_HEIGHT", height);
	this->fullScreen = fullScreen;
	this->useCustomResolution = true;
	this->window = nullptr;
}

/* src/GLWrapper.cpp:20-23: "monitor resolution" — headless: RT_WIDTH x RT_HEIGHT, default 1920x1080 */
GLWrapper::GLWrapper(bool fullScreen)
{
	this->fullScreen = fullScreen;
	this->width = env_int("RT_WIDTH", 1920);
	this->height = env_int("RT_HEIGHT", 1080);
	this->window = nullptr;
}

GLWrapper::~GLWrapper()
{
	stop();
}

int GLWrapper::getWidth() { return width; }
int GLWrapper::getHeight() { return height; }
GLuint GLWrapper::getProgramId() { return 1; }

/* src/GLWrapper.cpp:61-133: window + GL context  ->  CUDA context, stream and RGBA32F framebuffer */
bool GLWrapper::init_window()
{
	/* RT_GPUS=n: the frame is tile-partitioned over n GPUs of the box and gathered on device 0 inside draw() (one NCCL fan-in
	 * per frame; RT_GATHER=1: the kernels store into device 0's frame over NVLink instead) */
	const int n_gpus = env_int("RT_GPUS", 1);
	ctx = n_gpus > 1 ? rtb_create_multi(width, height, n_gpus, env_int("RT_BLOCK_ROWS", 4)) : rtb_create(width, height, env_int("RT_DEVICE", 0));
	if (!ctx) {
		fprintf(stderr, "rtb_create failed: %s\n", rtb_last_error(nullptr));
		return false;
	}
	rtb_set_option(ctx, "strict", env_int("RT_STRICT", 1));
	rtb_set_option(ctx, "kernel", env_int("RT_KERNEL", 0));
	if (n_gpus > 1 && rtb_set_option(ctx, "gather", env_int("RT_GATHER", 0))) die(ctx, "RT_GATHER");
	/* src/GLWrapper.cpp:124-130 + SMAA_Builder: the SMAA render targets, shaders and lookup tables */
	if (SMAA_enabled && env_int("RT_SMAA", 1)) {
#ifdef RTB_HAVE_SMAA_TABLES
		if (rtb_smaa_set_tables(ctx, areaTexBytes, searchTexBytes) || rtb_enable_smaa(ctx, (int)SMAA_preset)) die(ctx, "SMAA");
#else
		fprintf(stderr, "SMAA requested but this host was built without the reference's AreaTex.h / SearchTex.h: post-pass off\n");
#endif
	}
	window = rtb_shim_create_window(this, env_int("RT_FRAMES", 1));
	printf("rtb200 %s, %dx%d\n", rtb_version(), width, height);
	return true;
}

/* src/GLWrapper.cpp:135-141 */
void GLWrapper::set_skybox(unsigned textureId)
{
	auto it = g_cubemaps.find(textureId);
	if (it == g_cubemaps.end() || !it->second.ok) return;       /* a failed load leaves the sampler unbound, as in GL */
	const CubeImages& c = it->second;
	const uint8_t* faces[6];
	for (int f = 0; f < 6; f++) faces[f] = c.face[f].px.data();
	if (rtb_set_cubemap(ctx, faces, c.face[0].w, c.face[0].h, c.face[0].ch)) die(ctx, "set_skybox");
	std::string dd = dump_dir();
	if (!dd.empty())
		for (int f = 0; f < 6; f++)
			write_npy(dd + "/cube_" + std::to_string(f) + ".npy", "|u1", { (size_t)c.face[f].h, (size_t)c.face[f].w, (size_t)c.face[f].ch },
			          c.face[f].px.data(), c.face[f].px.size());
}

/* src/GLWrapper.cpp:143-147 */
void GLWrapper::stop()
{
	if (ctx) { rtb_destroy(ctx); ctx = nullptr; }
	if (window) { rtb_shim_destroy_window(window); window = nullptr; }
}

/* src/GLWrapper.cpp:149-153 (called before init_window, main.cpp:32-34) */
void GLWrapper::enable_SMAA(SMAA_PRESET preset)
{
	SMAA_enabled = true;
	SMAA_preset = preset;
}

/* src/GLWrapper.cpp:155-165: glDrawArrays(GL_TRIANGLES, 0, 6) -> one launch of the ray-trace kernel */
void GLWrapper::draw()
{
	if (rtb_render(ctx)) die(ctx, "Draw raytraced image");
}

/* src/GLWrapper.cpp:232-247: the {TOKEN} specialisation of rt.frag */
void GLWrapper::init_shaders(rt_defines& defines)
{
	rtb_defines d;
	static_assert(sizeof(rtb_defines) == 60, "rt_defines layout");
	memcpy(&d, &defines, sizeof d);                     /* scene.h:7-20 has the same field order (tests/test_layout.py) */
	d.iterations = env_int("RT_ITERATIONS", d.iterations);
	if (rtb_set_defines(ctx, &d)) die(ctx, "Shader creation");
}

/* src/GLWrapper.cpp:284-317 */
GLuint GLWrapper::load_cubemap(std::vector<std::string> faces, bool genMipmap)
{
	(void)genMipmap;                                    /* main.cpp:147 passes false; mipmapped cubemaps are not used */
	GLuint id = g_next_handle++;
	CubeImages& c = g_cubemaps[id];
	for (unsigned int i = 0; i < faces.size() && i < 6; i++) {
		if (!decode(faces[i], c.face[i])) {
			printf("Cubemap tex failed to load at path: %s\n", faces[i].c_str());
			c.ok = false;
		}
	}
	if (faces.size() != 6) c.ok = false;
	return id;
}

/* src/GLWrapper.cpp:356-363 (+ :319-354) */
GLuint GLWrapper::load_texture(int texNum, const char* name, const char* uniformName, GLuint wrapMode)
{
	(void)uniformName; (void)wrapMode;                  /* sampler uniform -> unit is fixed by main.cpp:149-153; wrap is GL_REPEAT */
	const std::string path = ASSETS_DIR "/textures/" + std::string(name);
	GLuint id = g_next_handle++;
	Image img;
	if (!decode(path, img)) {
		printf("Texture failed to load at path: %s\n", path.c_str());
		return id;
	}
	if (rtb_set_texture2d(ctx, texNum, img.px.data(), img.w, img.h, img.ch)) die(ctx, "load_texture");
	std::string dd = dump_dir();
	if (!dd.empty())
		write_npy(dd + "/tex_" + std::to_string(texNum) + ".npy", "|u1", { (size_t)img.h, (size_t)img.w, (size_t)img.ch }, img.px.data(), img.px.size());
	return id;
}

/* src/GLWrapper.cpp:365-379 */
void GLWrapper::init_buffer(GLuint* ubo, const char* name, int bindingPoint, size_t size, void* data) const
{
	int binding = -1;
	for (int b = 0; b < RTB_NUM_BINDINGS; b++) if (!strcmp(kBlockNames[b], name)) binding = b;
	if (binding < 0 || binding != bindingPoint) {
		fprintf(stderr, "Invalid ubo block name '%s'", name);
		exit(1);
	}
	*ubo = g_next_handle++;
	g_ubos[*ubo] = std::make_pair(ctx, binding);
	if (rtb_upload(ctx, binding, data, size)) die(ctx, "init_buffer");
	dump_block(binding, data, size);
}

/* src/GLWrapper.cpp:381-386 */
void GLWrapper::update_buffer(GLuint ubo, size_t size, void* data)
{
	auto it = g_ubos.find(ubo);
	if (it == g_ubos.end()) { fprintf(stderr, "update_buffer: unknown ubo %u\n", ubo); exit(1); }
	if (rtb_update(it->second.first, it->second.second, data, size)) die(it->second.first, "update_buffer");
	dump_block(it->second.second, data, size);
}

/* glfwSwapBuffers: the frame becomes visible -> here: optionally written to RT_DUMP_DIR */
void GLWrapper::present()
{
	std::string dd = dump_dir();
	if (!dd.empty()) {
		std::vector<float> px((size_t)width * height * 4);
		if (rtb_read_rgba32f(ctx, px.data())) die(ctx, "read frame");
		char name[64];
		snprintf(name, sizeof name, "/frame_%04d.npy", frame_index);
		write_npy(dd + name, "<f4", { (size_t)height, (size_t)width, 4 }, px.data(), px.size() * sizeof(float));
		if (SMAA_enabled && env_int("RT_SMAA", 1)) {           /* what the reference puts on screen: the SMAA-filtered RGBA8 image */
			std::vector<unsigned char> px8((size_t)width * height * 4);
			if (rtb_read_rgba8(ctx, px8.data())) die(ctx, "read frame");
			snprintf(name, sizeof name, "/screen_%04d.npy", frame_index);
			write_npy(dd + name, "|u1", { (size_t)height, (size_t)width, 4 }, px8.data(), px8.size());
		}
	} else {
		rtb_sync(ctx);
	}
	rtb_stats st;
	if (!rtb_get_stats(ctx, &st) && (frame_index < 3 || getenv("RT_VERBOSE"))) printf("frame %d: kernel %d, %.3f ms\n", frame_index, st.kernel_used, st.kernel_ms);
	frame_index++;
}

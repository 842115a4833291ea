/* GLWrapper.h — drop-in replacement for the reference's render driver (src/GLWrapper.h).
 *
 * Same class name, same public methods with the same signatures (src/GLWrapper.h:17-38), so the
 * reference's main.cpp and SceneManager.{h,cpp} compile against it UNCHANGED.  Behind it there is no
 * OpenGL: every method forwards to the C-ABI of librtb200.so (include/rtb200.h), whose draw() is the
 * sm_100a ray-trace kernel.  Error behaviour follows the reference: print and exit (utils.h:25,62;
 * GLWrapper.cpp:371-375).
 *
 * Headless controls (environment), applied here because main.cpp hard-codes them:
 *   RT_WIDTH / RT_HEIGHT   canvas size reported by getWidth()/getHeight() (main.cpp:36-37)
 *   RT_ITERATIONS          overrides rt_defines::iterations in init_shaders()
 *   RT_FRAMES              frames to render before glfwWindowShouldClose() turns true (default 1)
 *   RT_DUMP_DIR            write frame_NNNN.npy (RGBA32F, row 0 = bottom), the uploaded uniform buffers and the
 *                          decoded textures there
 *   RT_STRICT / RT_KERNEL / RT_DEVICE   rtb_set_option("strict"/"kernel"), CUDA device
 *   RT_GPUS / RT_GATHER / RT_BLOCK_ROWS all GPUs of the box behind the same draw() (rtb_create_multi)
 *   RT_SMAA                0 switches the SMAA post-pass off although main.cpp:32 enables it (default 1)
 */
#pragma once

#include <glad/glad.h>
#include <GLFW/glfw3.h>
#include <string>
#include <vector>
#include <iostream>   /* the reference's GLWrapper.h provides these through utils.h / shader.h; main.cpp relies on it */
#include <cstdio>
#include <glm/glm.hpp>

enum SMAA_PRESET { LOW, MEDIUM, HIGH, ULTRA };     /* src/SMAA_Builder.h:9-12 */

struct rt_defines;
struct rtb_ctx;

class GLWrapper
{
public:
	GLWrapper(int width, int height, bool fullScreen);
	GLWrapper(bool fullScreen);
	~GLWrapper();

	int getWidth();
	int getHeight();
	GLuint getProgramId();

	bool init_window();
	void init_shaders(rt_defines& defines);
	void set_skybox(unsigned int textureId);

	void stop();
	void enable_SMAA(SMAA_PRESET preset);

	GLFWwindow* window;

	void draw();
	static GLuint load_cubemap(std::vector<std::string> faces, bool genMipmap = false);
	GLuint load_texture(int texNum, const char* name, const char* uniformName, GLuint wrapMode = GL_REPEAT);
	void init_buffer(GLuint* ubo, const char* name, int bindingPoint, size_t size, void* data) const;
	static void update_buffer(GLuint ubo, size_t size, void* data);

	/* not in the reference: called by the GLFW shim when the frame is "presented" */
	void present();

private:
	rtb_ctx* ctx = nullptr;
	int width = 0;
	int height = 0;
	bool fullScreen = true;
	bool useCustomResolution = false;
	bool SMAA_enabled = false;
	SMAA_PRESET SMAA_preset = ULTRA;
	int frame_index = 0;
};

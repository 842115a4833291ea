/* shim.cpp — headless GLFW + the three raw GL calls of the reference's application code. */
#include <glad/glad.h>
#include <GLFW/glfw3.h>
#include "GLWrapper.h"
#include "shim_state.h"

#include <chrono>
#include <cstdio>

struct GLFWwindow {
    GLWrapper* owner = nullptr;
    void* user = nullptr;
    int frames_left = 1;
    int should_close = 0;
    GLFWcursorposfun cursor_cb = nullptr;
    GLFWkeyfun key_cb = nullptr;
    GLFWframebuffersizefun fb_cb = nullptr;
};

namespace {
long g_frames_presented = 0;
std::chrono::steady_clock::time_point g_first_present;          /* wall clock of the frame loop (the reference prints FPS, main.cpp:158-174) */
std::chrono::steady_clock::time_point g_last_present;
double g_max_gap = 0.0, g_sum_small = 0.0;                      /* the longest frame (a stall of the box shows up here, not in the mean of the others) */
long g_max_gap_frame = 0, g_n_small = 0;
GLenum g_active_unit = GL_TEXTURE0;
GLuint g_bound_2d[8] = { 0 };
}

GLFWwindow* rtb_shim_create_window(GLWrapper* owner, int frames) {
    GLFWwindow* w = new GLFWwindow();
    w->owner = owner;
    w->frames_left = frames < 1 ? 1 : frames;
    return w;
}
void rtb_shim_destroy_window(GLFWwindow* w) {
    if (g_frames_presented > 1) {
        /* frames 2..N: update_scene + update_buffers (host -> device) + draw + present, as the unchanged loop runs them */
        /* up to the LAST present: what follows (stop(), the destructors, cudaFree) is teardown, not the loop — it took up to 0.7 s in some runs */
        const double s = std::chrono::duration<double>(g_last_present - g_first_present).count();
        printf("frame loop: %ld frames in %.3f s after the first = %.3f ms per frame, %.1f frames/s\n", g_frames_presented - 1, s,
               s * 1e3 / (double)(g_frames_presented - 1), (double)(g_frames_presented - 1) / s);
        printf("frame loop: longest frame %.3f ms (frame %ld); %.3f ms per frame without the frames longer than 20 ms (%ld of them)\n", g_max_gap * 1e3,
               g_max_gap_frame, g_n_small ? g_sum_small * 1e3 / (double)g_n_small : 0.0, g_frames_presented - 1 - g_n_small);
    }
    delete w;
}

extern "C" {

double glfwGetTime(void) { return (double)g_frames_presented / 60.0; }     /* deterministic 60 Hz clock */
void glfwPollEvents(void) {}
void glfwSwapInterval(int) {}
void glfwSwapBuffers(GLFWwindow* w) {
    if (w && w->owner) w->owner->present();
    const auto now = std::chrono::steady_clock::now();
    if (g_frames_presented == 0) g_first_present = now;
    else {
        const double gap = std::chrono::duration<double>(now - g_last_present).count();
        if (gap > g_max_gap) { g_max_gap = gap; g_max_gap_frame = g_frames_presented; }
        if (gap < 0.020) { g_sum_small += gap; g_n_small++; }
    }
    g_last_present = now;
    g_frames_presented++;
    if (w && --w->frames_left <= 0) w->should_close = 1;
}
int glfwWindowShouldClose(GLFWwindow* w) { return w ? w->should_close : 1; }
void glfwSetWindowShouldClose(GLFWwindow* w, int v) { if (w) w->should_close = v; }
void glfwSetWindowUserPointer(GLFWwindow* w, void* p) { if (w) w->user = p; }
void* glfwGetWindowUserPointer(GLFWwindow* w) { return w ? w->user : nullptr; }
GLFWcursorposfun glfwSetCursorPosCallback(GLFWwindow* w, GLFWcursorposfun cb) { GLFWcursorposfun o = w->cursor_cb; w->cursor_cb = cb; return o; }
GLFWkeyfun glfwSetKeyCallback(GLFWwindow* w, GLFWkeyfun cb) { GLFWkeyfun o = w->key_cb; w->key_cb = cb; return o; }
GLFWframebuffersizefun glfwSetFramebufferSizeCallback(GLFWwindow* w, GLFWframebuffersizefun cb) { GLFWframebuffersizefun o = w->fb_cb; w->fb_cb = cb; return o; }
void glfwSetInputMode(GLFWwindow*, int, int) {}

/* main.cpp:178-187 re-binds the five 2-D textures to units 1..5 every frame; the unit of each texture was
 * already fixed by load_texture(texNum, ...), so the calls only need to be accepted. */
void glActiveTexture(GLenum texture) { g_active_unit = texture; }
void glBindTexture(GLenum target, GLuint texture) {
    if (target == GL_TEXTURE_2D && g_active_unit >= GL_TEXTURE0 && g_active_unit < GL_TEXTURE0 + 8) g_bound_2d[g_active_unit - GL_TEXTURE0] = texture;
}
void glViewport(GLint, GLint, GLsizei, GLsizei) {}

}  // extern "C"

/* GLFW/glfw3.h (shim) — headless stand-in for the 12 GLFW entry points the reference's main.cpp and
 * SceneManager.cpp call.  Deterministic clock: glfwGetTime() = frames presented / 60;
 * glfwWindowShouldClose() turns true after RT_FRAMES frames (default 1); glfwSwapBuffers() presents
 * (= optionally dumps) the frame. */
#ifndef RTB_SHIM_GLFW3_H
#define RTB_SHIM_GLFW3_H
#ifdef __cplusplus
extern "C" {
#endif
typedef struct GLFWwindow GLFWwindow;
typedef void (*GLFWcursorposfun)(GLFWwindow*, double, double);
typedef void (*GLFWkeyfun)(GLFWwindow*, int, int, int, int);
typedef void (*GLFWframebuffersizefun)(GLFWwindow*, int, int);
#define GLFW_RELEASE 0
#define GLFW_PRESS 1
#define GLFW_REPEAT 2
#define GLFW_KEY_SPACE 32
#define GLFW_KEY_A 65
#define GLFW_KEY_D 68
#define GLFW_KEY_S 83
#define GLFW_KEY_W 87
#define GLFW_KEY_ESCAPE 256
#define GLFW_KEY_LEFT_SHIFT 340
#define GLFW_KEY_LEFT_CONTROL 341
#define GLFW_KEY_LEFT_ALT 342
#define GLFW_CURSOR 0x00033001
#define GLFW_CURSOR_DISABLED 0x00034003
double glfwGetTime(void);
void glfwPollEvents(void);
void glfwSwapBuffers(GLFWwindow* window);
void glfwSwapInterval(int interval);
int glfwWindowShouldClose(GLFWwindow* window);
void glfwSetWindowShouldClose(GLFWwindow* window, int value);
void glfwSetWindowUserPointer(GLFWwindow* window, void* pointer);
void* glfwGetWindowUserPointer(GLFWwindow* window);
GLFWcursorposfun glfwSetCursorPosCallback(GLFWwindow* window, GLFWcursorposfun cb);
GLFWkeyfun glfwSetKeyCallback(GLFWwindow* window, GLFWkeyfun cb);
GLFWframebuffersizefun glfwSetFramebufferSizeCallback(GLFWwindow* window, GLFWframebuffersizefun cb);
void glfwSetInputMode(GLFWwindow* window, int mode, int value);
#ifdef __cplusplus
}
#endif
#endif

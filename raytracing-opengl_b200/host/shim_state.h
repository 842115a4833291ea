/* shim_state.h — glue between the replacement GLWrapper and the headless GLFW/GL shim. */
#pragma once
#include <GLFW/glfw3.h>
class GLWrapper;
GLFWwindow* rtb_shim_create_window(GLWrapper* owner, int frames);
void rtb_shim_destroy_window(GLFWwindow* w);

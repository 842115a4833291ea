/* glad/glad.h (shim) — the handful of OpenGL names the reference's application code touches directly
 * (main.cpp:178-187 glActiveTexture/glBindTexture, SceneManager.cpp:105 glViewport, GLWrapper.h GLuint/GL_REPEAT).
 * There is no OpenGL here: the three functions record state for the CUDA render driver (shim.cpp). */
#ifndef RTB_SHIM_GLAD_H
#define RTB_SHIM_GLAD_H
typedef unsigned int GLuint;
typedef unsigned int GLenum;
typedef int GLint;
typedef int GLsizei;
typedef unsigned char GLboolean;
#define GL_FALSE 0
#define GL_TRUE 1
#define GL_TEXTURE_2D 0x0DE1
#define GL_TEXTURE_CUBE_MAP 0x8513
#define GL_TEXTURE0 0x84C0
#define GL_TEXTURE1 0x84C1
#define GL_TEXTURE2 0x84C2
#define GL_TEXTURE3 0x84C3
#define GL_TEXTURE4 0x84C4
#define GL_TEXTURE5 0x84C5
#define GL_TEXTURE6 0x84C6
#define GL_TEXTURE7 0x84C7
#define GL_REPEAT 0x2901
#define GL_CLAMP_TO_EDGE 0x812F
#ifdef __cplusplus
extern "C" {
#endif
void glActiveTexture(GLenum texture);
void glBindTexture(GLenum target, GLuint texture);
void glViewport(GLint x, GLint y, GLsizei width, GLsizei height);
#ifdef __cplusplus
}
#endif
#endif

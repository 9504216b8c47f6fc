/* No-op OpenGL surface for MeshPhysicsShape::Draw (reference engine/physics_shape.cpp:229-256).
 * Draw is never called by the oracle; these only let the translation unit compile. */
#pragma once
typedef float GLfloat; typedef unsigned int GLenum; typedef int GLsizei; typedef int GLint;
#define GL_TRIANGLES 4
#define GL_LINES 1
#define GL_LINE_STRIP 3
#define GL_MODELVIEW_MATRIX 0x0BA6
#define GL_MODELVIEW 0x1700
#define GL_COLOR_MATERIAL 0x0B57
#define GL_VERTEX_ARRAY 0x8074
#define GL_NORMAL_ARRAY 0x8075
#define GL_FLOAT 0x1406
#define GL_UNSIGNED_INT 0x1405
static inline void glGetFloatv(GLenum, GLfloat*) {}
static inline void glPushMatrix() {}
static inline void glPopMatrix() {}
static inline void glMatrixMode(GLenum) {}
static inline void glLoadMatrixf(const GLfloat*) {}
static inline void glEnable(GLenum) {}
static inline void glEnableClientState(GLenum) {}
static inline void glDisableClientState(GLenum) {}
static inline void glVertexPointer(GLint, GLenum, GLsizei, const void*) {}
static inline void glNormalPointer(GLenum, GLsizei, const void*) {}
static inline void glColor4fv(const GLfloat*) {}
static inline void glDrawElements(GLenum, GLsizei, GLenum, const void*) {}

// ORACLE BUILD STUB (test infrastructure): stand-in for glbinding so that reference sources which merely *mention*
// OpenGL (debug draw() helpers) compile without a GL stack. Every entry point is a no-op; nothing here computes.
#ifndef RR_REF_STUB_GL_H
#define RR_REF_STUB_GL_H
namespace gl {
enum GLenum : unsigned {
  GL_POINTS, GL_LINES, GL_LINE_LOOP, GL_QUADS, GL_TRIANGLE_FAN, GL_TRIANGLE_STRIP, GL_BLEND, GL_COLOR_MATERIAL,
  GL_DEPTH_TEST, GL_FLOAT, GL_LIGHTING, GL_LINE_SMOOTH, GL_ONE_MINUS_SRC_ALPHA, GL_SRC_ALPHA, GL_STATIC_DRAW,
  GL_UNSIGNED_INT, GL_ALL_ATTRIB_BITS
};
inline void glBegin(GLenum) {}
inline void glEnd() {}
inline void glVertex3f(float, float, float) {}
inline void glColor3f(float, float, float) {}
inline void glColor4f(float, float, float, float) {}
inline void glTexCoord2f(float, float) {}
inline void glPointSize(float) {}
inline void glLineWidth(float) {}
inline void glEnable(GLenum) {}
inline void glDisable(GLenum) {}
inline void glBlendFunc(GLenum, GLenum) {}
inline void glPushAttrib(GLenum) {}
inline void glPopAttrib() {}
}  // namespace gl
#endif

// ORACLE BUILD STUB (test infrastructure): draw calls are no-ops. No GL.
#ifndef RR_REF_STUB_VA_H
#define RR_REF_STUB_VA_H
#include <glbinding/gl/gl.h>
#include "VertexAttributeBinding.h"
namespace globjects {
class VertexArray {
 public:
  void enable(int) {}
  VertexAttributeBinding* binding(int) { return &m_b; }
  void drawArrays(gl::GLenum, int, unsigned) {}
  void drawElements(gl::GLenum, unsigned long, gl::GLenum, const void*) {}
  void drawElementsBaseVertex(gl::GLenum, unsigned long, gl::GLenum, const void*, unsigned) {}
  void drawArraysInstanced(gl::GLenum, int, unsigned, unsigned) {}
 private:
  VertexAttributeBinding m_b;
};
}  // namespace globjects
#endif

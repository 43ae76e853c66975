// ORACLE BUILD STUB (test infrastructure). No GL.
#ifndef RR_REF_STUB_VAB_H
#define RR_REF_STUB_VAB_H
#include <glbinding/gl/gl.h>
namespace globjects {
class Buffer;
class VertexAttributeBinding {
 public:
  void setAttribute(int) {}
  void setBuffer(Buffer*, int, int) {}
  void setFormat(int, gl::GLenum) {}
};
}  // namespace globjects
#endif

// ORACLE BUILD STUB (test infrastructure): globjects::Buffer that accepts and discards data. No GL.
#ifndef RR_REF_STUB_BUFFER_H
#define RR_REF_STUB_BUFFER_H
#include <glbinding/gl/gl.h>
#include <vector>
namespace globjects {
class Buffer {
 public:
  template <typename T> void setData(std::vector<T> const&, gl::GLenum) {}
};
}  // namespace globjects
#endif

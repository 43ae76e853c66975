// Stand-ins for the few glm / gloost value types that appear in the kept C++ signatures, so that the host layer builds
// without the reference's vendored third-party trees. Inside the reference tree define RR_USE_REFERENCE_MATH and the
// real <glm/...> and <gloost/BoundingBox.h> are used instead (identical member names for everything used here).
#ifndef RR_HOST_MINI_MATH_HPP
#define RR_HOST_MINI_MATH_HPP

#ifdef RR_USE_REFERENCE_MATH
#include <glm/gtc/type_precision.hpp>
#include "gloost/BoundingBox.h"
#else
#include <cstddef>
#include <cstdint>

namespace glm {
template <typename T> struct tvec2 {
  T x, y;
  tvec2() : x(0), y(0) {}
  tvec2(T a, T b) : x(a), y(b) {}
  explicit tvec2(T s) : x(s), y(s) {}
  T& operator[](int i) { return (&x)[i]; }
  T const& operator[](int i) const { return (&x)[i]; }
  bool operator==(tvec2 const& o) const { return x == o.x && y == o.y; }
};
template <typename T> struct tvec3 {
  T x, y, z;
  tvec3() : x(0), y(0), z(0) {}
  tvec3(T a, T b, T c) : x(a), y(b), z(c) {}
  explicit tvec3(T s) : x(s), y(s), z(s) {}
  T& operator[](int i) { return (&x)[i]; }
  T const& operator[](int i) const { return (&x)[i]; }
  bool operator==(tvec3 const& o) const { return x == o.x && y == o.y && z == o.z; }
  bool operator!=(tvec3 const& o) const { return !(*this == o); }
};
template <typename T> struct tvec4 {
  T x, y, z, w;
  tvec4() : x(0), y(0), z(0), w(0) {}
  tvec4(T a, T b, T c, T d) : x(a), y(b), z(c), w(d) {}
  explicit tvec4(T s) : x(s), y(s), z(s), w(s) {}
  T& operator[](int i) { return (&x)[i]; }
  T const& operator[](int i) const { return (&x)[i]; }
};
typedef tvec2<unsigned> uvec2;
typedef tvec3<unsigned> uvec3;
typedef tvec4<unsigned> uvec4;
typedef tvec2<float> fvec2;
typedef tvec3<float> fvec3;
typedef tvec4<float> fvec4;
}  // namespace glm

namespace gloost {
struct Point3 {
  float v[3];
  Point3() : v{0, 0, 0} {}
  Point3(float x, float y, float z) : v{x, y, z} {}
  float& operator[](int i) { return v[i]; }
  float const& operator[](int i) const { return v[i]; }
};
class BoundingBox {
 public:
  BoundingBox() {}
  BoundingBox(Point3 const& pMin, Point3 const& pMax) : m_min(pMin), m_max(pMax) {}
  void setPMin(Point3 const& p) { m_min = p; }
  void setPMax(Point3 const& p) { m_max = p; }
  Point3 const& getPMin() const { return m_min; }
  Point3 const& getPMax() const { return m_max; }
 private:
  Point3 m_min, m_max;
};
}  // namespace gloost
#endif
#endif

// ORACLE TOOLING (test infrastructure, NOT product code).
// A minimal GLSL 4.30 host environment: enough of the language's vector types, built-ins and texture/image objects to
// compile the REFERENCE'S OWN shader sources (glsl/pre_*.fs, inc_*.glsl, tsdf_integration.vs, read from /root/reference
// at build time by transpile.py, outputs only under oracle/_ref/) as C++ and run them on the CPU. What the shaders say
// is therefore executed as written; what a GL driver would add is stated here, from the OpenGL 4.4 core specification:
//   * texture(): section 8.14.2/8.14.3 - unnormalised coordinate u = s * size; NEAREST texel floor(u); LINEAR texels
//     i0 = floor(u - 0.5), i1 = i0 + 1 with weights a = frac(u - 0.5). The specification's weighted sum
//     (1-a)(1-b) t00 + a(1-b) t10 + (1-a) b t01 + a b t11 is exact arithmetic; evaluated term by term in binary32 its
//     weights do not sum to 1, so a constant texture would not filter to itself - which every hardware filter guarantees
//     (fixed-point weights) and the shaders rely on (`silhouette < 1.0`). It is therefore evaluated separably, x then y
//     then z, with GLSL's own mix() formula x*(1-a) + y*a, which keeps 0/1 textures exact. CLAMP_TO_EDGE clamps the texel
//     indices; array layer = clamp(round-half-even(r), 0, layers-1); RGB8 texels are c/255 (section 8.5).
//   * built-ins follow the GLSL 4.30 specification section 8 formulas in binary32 (length = sqrt(dot), distance =
//     length(a-b), normalize = v / length(v), min(x,y) = y < x ? y : x, sign, abs, floor, clamp); pow(x, y) is
//     undefined for x < 0 in GLSL - NaN here, which is what NVIDIA's exp2(y*log2(x)) lowering returns.
//   * float -> int/uint conversions of out-of-range values are undefined in GLSL; the shaders on this path only
//     convert in-range values.
// This is deliberately a SECOND formulation of the arithmetic (mix() without fma, libm pow, plain dot products) next to
// the oracle's (fma lerps, polynomial pow, fma dot chains), so agreement between the two is evidence, not tautology.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

namespace glsl {

typedef unsigned int uint;

struct vec2; struct vec3; struct vec4; struct ivec3; struct uvec3;

struct vec2 {
  union { float x, r, s; };
  union { float y, g, t; };
  vec2() : x(0.f), y(0.f) {}
  explicit vec2(float v) : x(v), y(v) {}
  template <typename A, typename B> vec2(A a, B b) : x((float)a), y((float)b) {}
  inline vec2(const struct ivec2& v);   // GLSL converts ivec2 / uvec2 to vec2 implicitly
  inline vec2(const struct uvec2& v);
  vec2 xy() const { return *this; }
  vec2 rg() const { return *this; }
  float& operator[](int i) { return i == 0 ? x : y; }
  float operator[](int i) const { return i == 0 ? x : y; }
};

struct vec3 {
  union { float x, r, s; };
  union { float y, g, t; };
  union { float z, b, p; };
  vec3() : x(0.f), y(0.f), z(0.f) {}
  explicit vec3(float v) : x(v), y(v), z(v) {}
  template <typename A, typename B, typename C> vec3(A a, B b_, C c) : x((float)a), y((float)b_), z((float)c) {}
  template <typename C> vec3(const vec2& v, C c) : x(v.x), y(v.y), z((float)c) {}
  inline vec3(const uvec3& v);          // GLSL converts uvec3 / ivec3 to vec3 implicitly
  inline vec3(const ivec3& v);
  vec2 xy() const { return vec2(x, y); }
  vec2 rg() const { return vec2(x, y); }
  vec2 xx() const { return vec2(x, x); }
  vec2 yz() const { return vec2(y, z); }
  vec3 xyz() const { return *this; }
  vec3 rgb() const { return *this; }
  float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
  float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};

struct vec4 {
  union { float x, r, s; };
  union { float y, g, t; };
  union { float z, b, p; };
  union { float w, a, q; };
  vec4() : x(0.f), y(0.f), z(0.f), w(0.f) {}
  explicit vec4(float v) : x(v), y(v), z(v), w(v) {}
  template <typename A, typename B, typename C, typename D> vec4(A a_, B b_, C c, D d) : x((float)a_), y((float)b_), z((float)c), w((float)d) {}
  template <typename D> vec4(const vec3& v, D d) : x(v.x), y(v.y), z(v.z), w((float)d) {}
  template <typename C, typename D> vec4(const vec2& v, C c, D d) : x(v.x), y(v.y), z((float)c), w((float)d) {}
  vec2 xy() const { return vec2(x, y); }
  vec2 rg() const { return vec2(x, y); }
  vec3 xyz() const { return vec3(x, y, z); }
  vec3 rgb() const { return vec3(x, y, z); }
  float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
  float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
};

struct ivec2 {
  int x, y;
  ivec2() : x(0), y(0) {}
  template <typename A, typename B> ivec2(A a, B b) : x((int)a), y((int)b) {}
  explicit ivec2(int v) : x(v), y(v) {}
  explicit ivec2(const vec2& v) : x((int)v.x), y((int)v.y) {}     // truncation toward zero
  inline explicit ivec2(const uvec2& v);
};

struct uvec2 {
  uint x, y;
  uvec2() : x(0u), y(0u) {}
  template <typename A, typename B> uvec2(A a, B b) : x((uint)a), y((uint)b) {}
};
inline uvec2 operator+(const uvec2& a, const uvec2& b) { return uvec2(a.x + b.x, a.y + b.y); }

// An unsized buffer array (SSBO member `T name[]`). Out-of-range accesses are undefined in GLSL and do happen on this path
// (index_3d + ivec3(-1) at the brick-grid border, bricks.gs:28; a filtered point that left the bbox, inc_bricks.glsl:52,57):
// reads return 0, writes are dropped - the robust-buffer-access behaviour.
template <typename T> struct ssbo_array {
  T* p = nullptr;
  size_t n = 0;
  T sink = T();
  T& operator[](size_t i) { if (i < n) return p[i]; sink = T(); return sink; }
};

// T[5] as a value (GLSL arrays are first-class: returned from functions, assigned)
template <typename T> struct arr5 {
  T v[5];
  arr5() : v{} {}
  arr5(const T& a, const T& b, const T& c, const T& d, const T& e) : v{a, b, c, d, e} {}
  T& operator[](uint i) { return v[i]; }
  const T& operator[](uint i) const { return v[i]; }
};

struct ivec3 {
  int x, y, z;
  ivec3() : x(0), y(0), z(0) {}
  explicit ivec3(int v) : x(v), y(v), z(v) {}
  template <typename A, typename B, typename C> ivec3(A a, B b, C c) : x((int)a), y((int)b), z((int)c) {}
  explicit ivec3(const vec3& v) : x((int)v.x), y((int)v.y), z((int)v.z) {}      // truncation toward zero (GLSL 5.4.1)
  inline explicit ivec3(const uvec3& v);
};

struct uvec3 {
  uint x, y, z;
  uvec3() : x(0u), y(0u), z(0u) {}
  explicit uvec3(uint v) : x(v), y(v), z(v) {}
  template <typename A, typename B, typename C> uvec3(A a, B b, C c) : x((uint)a), y((uint)b), z((uint)c) {}
  explicit uvec3(const vec3& v) : x((uint)v.x), y((uint)v.y), z((uint)v.z) {}
  explicit uvec3(const ivec3& v) : x((uint)v.x), y((uint)v.y), z((uint)v.z) {}
};

inline ivec2::ivec2(const uvec2& v) : x((int)v.x), y((int)v.y) {}
inline vec2::vec2(const ivec2& v) : x((float)v.x), y((float)v.y) {}
inline vec2::vec2(const uvec2& v) : x((float)v.x), y((float)v.y) {}
inline ivec2 operator+(const ivec2& a, const ivec2& b) { return ivec2(a.x + b.x, a.y + b.y); }
inline vec3::vec3(const uvec3& v) : x((float)v.x), y((float)v.y), z((float)v.z) {}
inline vec3::vec3(const ivec3& v) : x((float)v.x), y((float)v.y), z((float)v.z) {}
inline ivec3::ivec3(const uvec3& v) : x((int)v.x), y((int)v.y), z((int)v.z) {}

// ---- arithmetic (component-wise, binary32) ---------------------------------------------------------------------------
#define GLSL_VEC_OPS(V, ...)                                                                          \
  inline V operator+(const V& a, const V& b) { return V(__VA_ARGS__(+)); }                             \
  inline V operator-(const V& a, const V& b) { return V(__VA_ARGS__(-)); }                             \
  inline V operator*(const V& a, const V& b) { return V(__VA_ARGS__(*)); }                             \
  inline V operator/(const V& a, const V& b) { return V(__VA_ARGS__(/)); }
#define GLSL_C2(op) a.x op b.x, a.y op b.y
#define GLSL_C3(op) a.x op b.x, a.y op b.y, a.z op b.z
GLSL_VEC_OPS(vec2, GLSL_C2)
GLSL_VEC_OPS(vec3, GLSL_C3)
inline vec2 operator*(const vec2& a, float s) { return vec2(a.x * s, a.y * s); }
inline vec2 operator*(float s, const vec2& a) { return vec2(s * a.x, s * a.y); }
inline vec2 operator/(const vec2& a, float s) { return vec2(a.x / s, a.y / s); }
inline vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 operator+(const vec3& a, float s) { return vec3(a.x + s, a.y + s, a.z + s); }
inline vec3 operator-(const vec3& a, float s) { return vec3(a.x - s, a.y - s, a.z - s); }
inline vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
inline vec2 operator-(const vec2& a) { return vec2(-a.x, -a.y); }
inline vec3& operator+=(vec3& a, const vec3& b) { a = a + b; return a; }
inline vec3& operator-=(vec3& a, const vec3& b) { a = a - b; return a; }
inline vec3& operator*=(vec3& a, float s) { a = a * s; return a; }
inline vec3& operator/=(vec3& a, float s) { a = a / s; return a; }
inline vec2& operator+=(vec2& a, const vec2& b) { a = a + b; return a; }
inline vec2 operator+(const vec2& a, float s) { return vec2(a.x + s, a.y + s); }
inline vec2 operator-(const vec2& a, float s) { return vec2(a.x - s, a.y - s); }
inline vec3 operator/(float s, const vec3& a) { return vec3(s / a.x, s / a.y, s / a.z); }
inline vec4 operator+(const vec4& a, const vec4& b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline vec4 operator*(const vec4& a, float s) { return vec4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline vec4 operator/(const vec4& a, float s) { return vec4(a.x / s, a.y / s, a.z / s, a.w / s); }
inline ivec3 operator+(const ivec3& a, const ivec3& b) { return ivec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline ivec3 operator-(const ivec3& a, const ivec3& b) { return ivec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline uvec3 operator-(const uvec3& a, uint s) { return uvec3(a.x - s, a.y - s, a.z - s); }
inline uvec3 operator+(const uvec3& a, const uvec3& b) { return uvec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline uvec3 operator+(const uvec3& a, const ivec3& b) { return uvec3(a.x + (uint)b.x, a.y + (uint)b.y, a.z + (uint)b.z); }   // int -> uint, wraps

// ---- built-in functions (GLSL 4.30 section 8) --------------------------------------------------------------------------
inline float min(float x, float y) { return y < x ? y : x; }
inline float max(float x, float y) { return x < y ? y : x; }
inline int min(int x, int y) { return y < x ? y : x; }
inline int max(int x, int y) { return x < y ? y : x; }
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline int clamp(int x, int lo, int hi) { return min(max(x, lo), hi); }
inline ivec3 clamp(const ivec3& v, const ivec3& lo, const ivec3& hi) { return ivec3(clamp(v.x, lo.x, hi.x), clamp(v.y, lo.y, hi.y), clamp(v.z, lo.z, hi.z)); }
inline float abs(float x) { return std::fabs(x); }
inline int abs(int x) { return x < 0 ? -x : x; }
inline vec3 abs(const vec3& v) { return vec3(std::fabs(v.x), std::fabs(v.y), std::fabs(v.z)); }
inline float floor(float x) { return std::floor(x); }
inline vec3 floor(const vec3& v) { return vec3(std::floor(v.x), std::floor(v.y), std::floor(v.z)); }
inline float sign(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }
inline vec3 sign(const vec3& v) { return vec3(sign(v.x), sign(v.y), sign(v.z)); }
inline float sqrt(float x) { return std::sqrt(x); }
inline float ceil(float x) { return std::ceil(x); }
inline vec2 floor(const vec2& v) { return vec2(std::floor(v.x), std::floor(v.y)); }
inline vec2 min(const vec2& a, const vec2& b) { return vec2(min(a.x, b.x), min(a.y, b.y)); }
inline vec2 max(const vec2& a, const vec2& b) { return vec2(max(a.x, b.x), max(a.y, b.y)); }
inline vec2 clamp(const vec2& v, const vec2& lo, const vec2& hi) { return min(max(v, lo), hi); }   // GLSL 8.3: min(max(x, minVal), maxVal)
inline vec3 min(const vec3& a, const vec3& b) { return vec3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
inline vec3 max(const vec3& a, const vec3& b) { return vec3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
inline float dot(const vec2& a, const vec2& b) { return a.x * b.x + a.y * b.y; }
inline float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float length(const vec2& v) { return std::sqrt(dot(v, v)); }
inline float length(const vec3& v) { return std::sqrt(dot(v, v)); }
inline float distance(float a, float b) { return std::fabs(a - b); }
inline float distance(const vec2& a, const vec2& b) { return length(a - b); }
inline float distance(const vec3& a, const vec3& b) { return length(a - b); }
inline vec3 normalize(const vec3& v) { return v / length(v); }
inline vec3 cross(const vec3& a, const vec3& b) { return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
inline float pow(float x, float y) {
  if (x < 0.f) return std::numeric_limits<float>::quiet_NaN();   // undefined in GLSL; exp2(y*log2(x)) lowering gives NaN
  return std::pow(x, y);
}
inline float mix(float a, float b, float t) { return a * (1.f - t) + b * t; }
inline uint atomicAdd(uint& mem, uint v) {
  uint old;
#pragma omp atomic capture
  { old = mem; mem += v; }
  return old;
}

// ---- texture and image objects ---------------------------------------------------------------------------------------
inline int texel_clamp(float f, int n) {            // CLAMP_TO_EDGE on a floor()ed texel coordinate; NaN -> 0
  if (!(f >= 0.f)) return 0;
  if (f >= (float)(n - 1)) return n - 1;
  return (int)f;
}
struct LinearTap { int i0, i1; float a; };
inline LinearTap linear_tap(float s, int n) {
  const float u = s * (float)n - 0.5f;
  const float f = std::floor(u);
  LinearTap t;
  t.a = u - f;
  t.i0 = texel_clamp(f, n);
  t.i1 = texel_clamp(f + 1.f, n);
  return t;
}
inline int nearest_tap(float s, int n) { return texel_clamp(std::floor(s * (float)n), n); }
inline int array_layer(float r, int layers) {
  const float rn = std::nearbyint(r);               // round half to even (default rounding mode)
  return texel_clamp(rn, layers);
}

// A 2D array texture of `C` float channels per texel, or of RGB8 normalised bytes; LINEAR or NEAREST; CLAMP_TO_EDGE.
struct sampler2DArray {
  const float* f32 = nullptr;
  const uint8_t* u8 = nullptr;
  int W = 0, H = 0, L = 0, C = 1;
  bool linear = true;
  float texel(int l, int y, int x, int c) const {
    const size_t i = (((size_t)l * H + y) * W + x) * C + c;
    return u8 ? (float)u8[i] / 255.0f : f32[i];
  }
};
inline vec4 texture(const sampler2DArray& t, const vec3& p) {
  const int l = array_layer(p.z, t.L);
  float o[4] = {0.f, 0.f, 0.f, 1.f};
  if (!t.linear) {
    const int x = nearest_tap(p.x, t.W), y = nearest_tap(p.y, t.H);
    for (int c = 0; c < t.C; ++c) o[c] = t.texel(l, y, x, c);
  } else {
    const LinearTap tx = linear_tap(p.x, t.W), ty = linear_tap(p.y, t.H);
    const float a = tx.a, b = ty.a;
    for (int c = 0; c < t.C; ++c)
      o[c] = mix(mix(t.texel(l, ty.i0, tx.i0, c), t.texel(l, ty.i0, tx.i1, c), a),
                 mix(t.texel(l, ty.i1, tx.i0, c), t.texel(l, ty.i1, tx.i1, c), a), b);
  }
  return vec4(o[0], o[1], o[2], o[3]);
}

inline vec4 texture2DArray(const sampler2DArray& t, const vec3& p) { return texture(t, p); }     // GL_EXT_texture_array spelling

struct sampler3D {
  const float* f32 = nullptr;
  int X = 0, Y = 0, Z = 0, C = 1;
  bool linear = true;
  float texel(int z, int y, int x, int c) const { return f32[((((size_t)z * Y + y) * X) + x) * C + c]; }
};
inline vec4 texture(const sampler3D& t, const vec3& p) {
  float o[4] = {0.f, 0.f, 0.f, 1.f};
  if (!t.linear) {
    const int x = nearest_tap(p.x, t.X), y = nearest_tap(p.y, t.Y), z = nearest_tap(p.z, t.Z);
    for (int c = 0; c < t.C; ++c) o[c] = t.texel(z, y, x, c);
  } else {
    const LinearTap tx = linear_tap(p.x, t.X), ty = linear_tap(p.y, t.Y), tz = linear_tap(p.z, t.Z);
    const float a = tx.a, b = ty.a, g = tz.a;
    for (int c = 0; c < t.C; ++c)
      o[c] = mix(mix(mix(t.texel(tz.i0, ty.i0, tx.i0, c), t.texel(tz.i0, ty.i0, tx.i1, c), a),
                     mix(t.texel(tz.i0, ty.i1, tx.i0, c), t.texel(tz.i0, ty.i1, tx.i1, c), a), b),
                 mix(mix(t.texel(tz.i1, ty.i0, tx.i0, c), t.texel(tz.i1, ty.i0, tx.i1, c), a),
                     mix(t.texel(tz.i1, ty.i1, tx.i0, c), t.texel(tz.i1, ty.i1, tx.i1, c), a), b), g);
  }
  return vec4(o[0], o[1], o[2], o[3]);
}

// 2D texture of RGBA32F texels; only texelFetch() is used on it (tsdf_raymarch.fs getStartPos); pre_depth.fs declares one
// ("gauss") and never samples it
struct sampler2D {
  const float* f32 = nullptr;
  int W = 0, H = 0, C = 4;                            // C = 1: a depth texture (r = depth)
};
inline vec4 texelFetch(const sampler2D& t, const ivec2& p, int) {
  if (p.x < 0 || p.y < 0 || p.x >= t.W || p.y >= t.H) return vec4(0.f);          // undefined in GL; robust-access result (zeros)
  const float* q = t.f32 + ((size_t)p.y * t.W + p.x) * t.C;
  return t.C == 4 ? vec4(q[0], q[1], q[2], q[3]) : vec4(q[0], 0.f, 0.f, 1.f);
}
// LINEAR + MIRRORED_REPEAT (GL 4.4 table 8.20: mirror(a) = a >= 0 ? a : -(1 + a); i -> (size - 1) - mirror((i mod 2 size) - size))
inline int mirrored_repeat(int i, int size) {
  int m = i % (2 * size);
  if (m < 0) m += 2 * size;
  int a = m - size;
  if (a < 0) a = -(1 + a);
  return (size - 1) - a;
}
inline vec4 texture(const sampler2D& t, const vec2& p) {
  const float u = p.x * (float)t.W - 0.5f, v = p.y * (float)t.H - 0.5f;
  const float fu = std::floor(u), fv = std::floor(v);
  const float a = u - fu, b = v - fv;
  const int i0 = mirrored_repeat((int)fu, t.W), i1 = mirrored_repeat((int)fu + 1, t.W);
  const int j0 = mirrored_repeat((int)fv, t.H), j1 = mirrored_repeat((int)fv + 1, t.H);
  float o[4];
  for (int c = 0; c < 4; ++c) {
    auto tx = [&](int j, int i) { return t.f32[((size_t)j * t.W + i) * 4 + c]; };
    o[c] = mix(mix(tx(j0, i0), tx(j0, i1), a), mix(tx(j1, i0), tx(j1, i1), a), b);
  }
  return vec4(o[0], o[1], o[2], o[3]);
}

struct image2D {                                      // layout(r32f) image2D, write-only
  float* data = nullptr;
  int W = 0, H = 0;
};
inline void imageStore(image2D& img, const ivec2& p, const vec4& v) {
  if (p.x < 0 || p.y < 0 || p.x >= img.W || p.y >= img.H) return;
  img.data[(size_t)p.y * img.W + p.x] = v.x;
}

// ---- mat4 (column-major, m[c] is column c) -----------------------------------------------------------------------------
struct mat4 {
  vec4 c[4];
  mat4() { c[0] = vec4(1.f, 0.f, 0.f, 0.f); c[1] = vec4(0.f, 1.f, 0.f, 0.f); c[2] = vec4(0.f, 0.f, 1.f, 0.f); c[3] = vec4(0.f, 0.f, 0.f, 1.f); }
  explicit mat4(const float* m16) { for (int i = 0; i < 4; ++i) c[i] = vec4(m16[i * 4], m16[i * 4 + 1], m16[i * 4 + 2], m16[i * 4 + 3]); }
  vec4& operator[](int i) { return c[i]; }
  const vec4& operator[](int i) const { return c[i]; }
};
// GLSL 5.9: linear-algebraic products, written out as sums of column * component
inline vec4 operator*(const mat4& m, const vec4& v) { return m.c[0] * v.x + m.c[1] * v.y + m.c[2] * v.z + m.c[3] * v.w; }
inline mat4 operator*(const mat4& a, const mat4& b) {
  mat4 r;
  for (int i = 0; i < 4; ++i) r.c[i] = a * b.c[i];
  return r;
}
// inverse(): precision is implementation-defined in GLSL; cofactor expansion in binary32 here
inline mat4 inverse(const mat4& M) {
  float m[16], inv[16];
  for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) m[c * 4 + r] = M.c[c][r];
  inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
  inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
  inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
  inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
  inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
  inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
  inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
  inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
  inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
  inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
  inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
  inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
  inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
  inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
  inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
  inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
  const float det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
  const float id = 1.0f / det;
  for (int i = 0; i < 16; ++i) inv[i] *= id;
  return mat4(inv);
}

struct image3D {                                      // layout(r32f) image3D, write-only
  float* data = nullptr;
  int X = 0, Y = 0, Z = 0;
};
inline void imageStore(image3D& img, const ivec3& p, const vec4& v) {
  if (p.x < 0 || p.y < 0 || p.z < 0 || p.x >= img.X || p.y >= img.Y || p.z >= img.Z) return;   // out-of-bounds stores are discarded
  img.data[((size_t)p.z * img.Y + p.y) * img.X + p.x] = v.x;
}

}  // namespace glsl

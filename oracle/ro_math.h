// ORACLE (test infrastructure, NOT product code): scalar fp32 building blocks of the CPU restatement of
// steppobeck/rgbd-recon's volumetric-fusion path. Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may use anything under oracle/.
//
// PARITY UNPINNED BY THE REFERENCE'S TESTS: the reference ships no tests/golden vectors.
// What *is* pinned against real reference code compiled from /root/reference (oracle/_ref, see
// oracle/Makefile): Frustum planes/inside/camera position, CalibrationVolume<T> file I/O,
// CalibrationInverter::calculateInverseVolumes (+inverseDistance) and DataTypes getTrilinear.
//
// Arithmetic conventions ("the pin") — every kernel in rgbd-recon_b200/csrc follows the same rules:
//  * IEEE-754 binary32 everywhere, round-to-nearest-even, no fast-math, no implicit contraction
//    (this file is compiled with -ffp-contract=off; the CUDA side with --fmad=false). A fused
//    multiply-add happens ONLY where fmaf() is written.
//  * GLSL built-ins are restated as: min(a,b) = (b<a)?b:a, max(a,b) = (a<b)?b:a,
//    dot(a,b) = fma(a.z,b.z, fma(a.y,b.y, a.x*b.x)), length = sqrt(dot), normalize(v) = v * (1/sqrt(dot(v,v))),
//    pow(x,y) = exp2(y*log2(x)) with the deterministic exp2/log2 below (x<0 -> NaN, as NVIDIA GL does),
//    float->uint conversion saturating with NaN -> 0 (NVIDIA F2I semantics).
//  * Texture filtering (OpenGL 4.4 core spec §8.14) is evaluated as separable
//    lerps in x, then y, then z — the order the reference's own CPU getTrilinear uses
//    (framework/DataTypes.cpp:115-163) — with lerp(a,b,t) = fma(t, b, (1-t)*a).
#ifndef RO_MATH_H
#define RO_MATH_H

#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

namespace ro {

struct V2 { float x, y; };
struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };

static inline uint32_t f2bits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float bits2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

static inline float gl_min(float a, float b) { return (b < a) ? b : a; }
static inline float gl_max(float a, float b) { return (a < b) ? b : a; }
static inline float gl_clamp(float x, float lo, float hi) { return gl_min(gl_max(x, lo), hi); }
static inline float gl_sign(float x) { return (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f); }

static inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
static inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
static inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
static inline V3 operator/(V3 a, V3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }

static inline float dot3(V3 a, V3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
static inline float length3(V3 a) { return sqrtf(dot3(a, a)); }
static inline V3 normalize3(V3 a) { float r = 1.0f / sqrtf(dot3(a, a)); return a * r; }
static inline V3 cross3(V3 a, V3 b) {
  return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}

// float -> int with clamping into [lo,hi]; NaN -> lo. Used for texel indices after floor().
static inline int f2i_clamp(float f, int lo, int hi) {
  if (!(f >= (float)lo)) return lo;
  if (f >= (float)hi) return hi;
  return (int)f;
}
// float -> uint, saturating, NaN -> 0 (what NVIDIA's F2I.U32 does for GLSL uint(x)).
static inline uint32_t f2u_sat(float f) {
  if (!(f >= 0.0f)) return 0u;
  if (f >= 4294967296.0f) return 0xFFFFFFFFu;
  return (uint32_t)f;
}

// ---- deterministic log2 / exp2 / pow (pure IEEE fp32 ops; identical sequence in csrc/rr_math.cuh) ----
static inline float det_log2(float x) {
  if (x != x) return x;
  if (x < 0.0f) return std::numeric_limits<float>::quiet_NaN();
  if (x == 0.0f) return -std::numeric_limits<float>::infinity();
  if (x == std::numeric_limits<float>::infinity()) return x;
  int e = 0;
  uint32_t ix = f2bits(x);
  if (ix < 0x00800000u) { x = x * 8388608.0f; ix = f2bits(x); e = -23; }
  int32_t t = (int32_t)(ix - 0x3f3504f3u);             // signed distance to sqrt(1/2) in bit space
  e += (int)(t >> 23);                                 // arithmetic shift: unbiased exponent of x/m
  float m = bits2f(((uint32_t)t & 0x007fffffu) + 0x3f3504f3u);   // mantissa in [sqrt(1/2), sqrt(2))
  float f = m - 1.0f;
  float s = f / (2.0f + f);
  float z = s * s;
  float p = fmaf(z, 0.11111111f, 0.14285715f);
  p = fmaf(z, p, 0.2f);
  p = fmaf(z, p, 0.33333334f);
  p = fmaf(z, p, 1.0f);
  float ln_m = (2.0f * s) * p;
  return fmaf(ln_m, 1.4426950f, (float)e);
}

static inline float det_exp2(float x) {
  if (x != x) return x;
  if (x >= 128.0f) return std::numeric_limits<float>::infinity();
  if (x < -150.0f) return 0.0f;
  float n = floorf(x + 0.5f);
  float r = x - n;                                     // [-0.5, 0.5]
  float t = r * 0.69314718f;
  float p = fmaf(t, 1.9841270e-4f, 1.3888889e-3f);     // 1/5040, 1/720
  p = fmaf(t, p, 8.3333338e-3f);                       // 1/120
  p = fmaf(t, p, 4.1666668e-2f);                       // 1/24
  p = fmaf(t, p, 0.16666667f);
  p = fmaf(t, p, 0.5f);
  p = fmaf(t, p, 1.0f);
  p = fmaf(t, p, 1.0f);
  int ni = (int)n;
  int n1 = ni / 2;
  int n2 = ni - n1;
  float s1 = bits2f((uint32_t)(n1 + 127) << 23);
  float s2 = bits2f((uint32_t)(n2 + 127) << 23);
  return (p * s1) * s2;
}

static inline float gl_pow(float x, float y) { return det_exp2(y * det_log2(x)); }

static inline float lerpf(float a, float b, float t) { return fmaf(t, b, (1.0f - t) * a); }

// OpenGL LINEAR + CLAMP_TO_EDGE coordinate: texel pair and weight for normalised coordinate s, size W.
static inline void lin_coord(float s, int W, int& i0, int& i1, float& a) {
  float u = s * (float)W - 0.5f;
  float f = floorf(u);
  a = u - f;
  i0 = f2i_clamp(f, 0, W - 1);
  i1 = f2i_clamp(f + 1.0f, 0, W - 1);
}
// OpenGL NEAREST + CLAMP_TO_EDGE.
static inline int near_coord(float s, int W) { return f2i_clamp(floorf(s * (float)W), 0, W - 1); }

// Trilinear fetch of channel-interleaved volume T[z][y][x][C] (x fastest), C in {2,3,4}; returns up to 4 ch.
template <int C>
static inline void tex3d_linear(const float* T, int X, int Y, int Z, float s, float t, float r, float* out, int nout) {
  int x0, x1, y0, y1, z0, z1; float a, b, g;
  lin_coord(s, X, x0, x1, a);
  lin_coord(t, Y, y0, y1, b);
  lin_coord(r, Z, z0, z1, g);
  const size_t sy = (size_t)X, sz = (size_t)X * Y;
  const float* p000 = T + ((size_t)z0 * sz + (size_t)y0 * sy + x0) * C;
  const float* p100 = T + ((size_t)z0 * sz + (size_t)y0 * sy + x1) * C;
  const float* p010 = T + ((size_t)z0 * sz + (size_t)y1 * sy + x0) * C;
  const float* p110 = T + ((size_t)z0 * sz + (size_t)y1 * sy + x1) * C;
  const float* p001 = T + ((size_t)z1 * sz + (size_t)y0 * sy + x0) * C;
  const float* p101 = T + ((size_t)z1 * sz + (size_t)y0 * sy + x1) * C;
  const float* p011 = T + ((size_t)z1 * sz + (size_t)y1 * sy + x0) * C;
  const float* p111 = T + ((size_t)z1 * sz + (size_t)y1 * sy + x1) * C;
  for (int c = 0; c < nout; ++c) {
    float c00 = lerpf(p000[c], p100[c], a);
    float c10 = lerpf(p010[c], p110[c], a);
    float c01 = lerpf(p001[c], p101[c], a);
    float c11 = lerpf(p011[c], p111[c], a);
    float c0 = lerpf(c00, c10, b);
    float c1 = lerpf(c01, c11, b);
    out[c] = lerpf(c0, c1, g);
  }
}

// Bilinear fetch of one channel of a channel-interleaved image T[y][x][C].
static inline float tex2d_linear(const float* T, int W, int H, int C, int ch, float s, float t) {
  int x0, x1, y0, y1; float a, b;
  lin_coord(s, W, x0, x1, a);
  lin_coord(t, H, y0, y1, b);
  float v00 = T[((size_t)y0 * W + x0) * C + ch], v10 = T[((size_t)y0 * W + x1) * C + ch];
  float v01 = T[((size_t)y1 * W + x0) * C + ch], v11 = T[((size_t)y1 * W + x1) * C + ch];
  return lerpf(lerpf(v00, v10, a), lerpf(v01, v11, a), b);
}

static inline float tex2d_nearest(const float* T, int W, int H, int C, int ch, float s, float t) {
  int x = near_coord(s, W), y = near_coord(t, H);
  return T[((size_t)y * W + x) * C + ch];
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

}  // namespace ro
#endif

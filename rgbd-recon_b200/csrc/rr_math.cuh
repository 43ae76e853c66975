// Device-side fp32 building blocks for the fusion kernels (sm_100a).
//
// Arithmetic contract (DESIGN.md "Arithmetic pin"): binary32, round-to-nearest, compiled with --fmad=false so a
// fused multiply-add exists only where fmaf() is written; division and sqrt are the IEEE-exact CUDA defaults.
// GLSL built-ins used by the reference shaders are restated here:
//   min/max as comparisons (NaN falls through like the shader's), dot as an fma chain, normalize = v * (1/sqrt(dot)),
//   pow(x,y) = exp2(y*log2(x)) with the polynomial exp2/log2 below (x < 0 gives NaN, as NVIDIA GL does),
//   uint(x) saturating with NaN -> 0.
// Texture filtering follows the OpenGL 4.4 spec §8.14 equations (LINEAR / NEAREST, CLAMP_TO_EDGE), evaluated as
// separable lerps x -> y -> z with lerp(a,b,t) = fma(t, b, (1-t)*a): software filtering in full fp32, because the
// texture units' 8-bit interpolation weights cannot meet the 1e-5*limit tolerance.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace rr {

__device__ __forceinline__ float gmin(float a, float b) { return (b < a) ? b : a; }
__device__ __forceinline__ float gmax(float a, float b) { return (a < b) ? b : a; }
__device__ __forceinline__ float gsign(float x) { return (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f); }
__device__ __forceinline__ int iclamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator*(float3 a, float s) { return make_float3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float3 operator/(float3 a, float s) { return make_float3(a.x / s, a.y / s, a.z / s); }

__device__ __forceinline__ float dot3(float3 a, float3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
__device__ __forceinline__ float length3(float3 a) { return sqrtf(dot3(a, a)); }
__device__ __forceinline__ float3 normalize3(float3 a) { float r = 1.0f / sqrtf(dot3(a, a)); return a * r; }
__device__ __forceinline__ float3 cross3(float3 a, float3 b) {
  return make_float3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}

// floor()ed float -> texel index clamped to [lo, hi]; NaN -> lo
__device__ __forceinline__ int f2i_clamp(float f, int lo, int hi) {
  if (!(f >= (float)lo)) return lo;
  if (f >= (float)hi) return hi;
  return (int)f;
}
// GLSL uint(x) on NVIDIA: saturating, NaN -> 0 (F2I.U32.TRUNC)
__device__ __forceinline__ uint32_t f2u_sat(float f) { return __float2uint_rz(f); }

__device__ __forceinline__ float det_log2(float x) {
  if (x != x) return x;
  if (x < 0.0f) return __int_as_float(0x7fc00000);
  if (x == 0.0f) return __int_as_float(0xff800000);
  if (x == __int_as_float(0x7f800000)) return x;
  int e = 0;
  uint32_t ix = __float_as_uint(x);
  if (ix < 0x00800000u) { x = x * 8388608.0f; ix = __float_as_uint(x); e = -23; }
  int32_t t = (int32_t)(ix - 0x3f3504f3u);
  e += (t >> 23);
  float m = __uint_as_float(((uint32_t)t & 0x007fffffu) + 0x3f3504f3u);
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

__device__ __forceinline__ float det_exp2(float x) {
  if (x != x) return x;
  if (x >= 128.0f) return __int_as_float(0x7f800000);
  if (x < -150.0f) return 0.0f;
  float n = floorf(x + 0.5f);
  float r = x - n;
  float t = r * 0.69314718f;
  float p = fmaf(t, 1.9841270e-4f, 1.3888889e-3f);
  p = fmaf(t, p, 8.3333338e-3f);
  p = fmaf(t, p, 4.1666668e-2f);
  p = fmaf(t, p, 0.16666667f);
  p = fmaf(t, p, 0.5f);
  p = fmaf(t, p, 1.0f);
  p = fmaf(t, p, 1.0f);
  int ni = (int)n;
  int n1 = ni / 2;
  int n2 = ni - n1;
  float s1 = __uint_as_float((uint32_t)(n1 + 127) << 23);
  float s2 = __uint_as_float((uint32_t)(n2 + 127) << 23);
  return (p * s1) * s2;
}

__device__ __forceinline__ float gpow(float x, float y) { return det_exp2(y * det_log2(x)); }

__device__ __forceinline__ float lerpf(float a, float b, float t) { return fmaf(t, b, (1.0f - t) * a); }

// LINEAR + CLAMP_TO_EDGE texel pair and weight
__device__ __forceinline__ void lin_coord(float s, int W, int& i0, int& i1, float& a) {
  float u = s * (float)W - 0.5f;
  float f = floorf(u);
  a = u - f;
  i0 = f2i_clamp(f, 0, W - 1);
  i1 = f2i_clamp(f + 1.0f, 0, W - 1);
}
__device__ __forceinline__ int near_coord(float s, int W) { return f2i_clamp(floorf(s * (float)W), 0, W - 1); }

// Trilinear xyz fetch from a float4-padded volume [Z][Y][X] (forward cv_xyz repacked, or cv_xyz_inv).
__device__ __forceinline__ float3 tex3d_xyz(const float4* __restrict__ T, int X, int Y, int Z, float s, float t, float r) {
  int x0, x1, y0, y1, z0, z1; float a, b, g;
  lin_coord(s, X, x0, x1, a);
  lin_coord(t, Y, y0, y1, b);
  lin_coord(r, Z, z0, z1, g);
  const size_t sy = (size_t)X, sz = (size_t)X * Y;
  const float4 p000 = __ldg(T + z0 * sz + y0 * sy + x0), p100 = __ldg(T + z0 * sz + y0 * sy + x1);
  const float4 p010 = __ldg(T + z0 * sz + y1 * sy + x0), p110 = __ldg(T + z0 * sz + y1 * sy + x1);
  const float4 p001 = __ldg(T + z1 * sz + y0 * sy + x0), p101 = __ldg(T + z1 * sz + y0 * sy + x1);
  const float4 p011 = __ldg(T + z1 * sz + y1 * sy + x0), p111 = __ldg(T + z1 * sz + y1 * sy + x1);
  float3 o;
  o.x = lerpf(lerpf(lerpf(p000.x, p100.x, a), lerpf(p010.x, p110.x, a), b), lerpf(lerpf(p001.x, p101.x, a), lerpf(p011.x, p111.x, a), b), g);
  o.y = lerpf(lerpf(lerpf(p000.y, p100.y, a), lerpf(p010.y, p110.y, a), b), lerpf(lerpf(p001.y, p101.y, a), lerpf(p011.y, p111.y, a), b), g);
  o.z = lerpf(lerpf(lerpf(p000.z, p100.z, a), lerpf(p010.z, p110.z, a), b), lerpf(lerpf(p001.z, p101.z, a), lerpf(p011.z, p111.z, a), b), g);
  return o;
}

__device__ __forceinline__ float2 tex3d_uv(const float2* __restrict__ T, int X, int Y, int Z, float s, float t, float r) {
  int x0, x1, y0, y1, z0, z1; float a, b, g;
  lin_coord(s, X, x0, x1, a);
  lin_coord(t, Y, y0, y1, b);
  lin_coord(r, Z, z0, z1, g);
  const size_t sy = (size_t)X, sz = (size_t)X * Y;
  const float2 p000 = __ldg(T + z0 * sz + y0 * sy + x0), p100 = __ldg(T + z0 * sz + y0 * sy + x1);
  const float2 p010 = __ldg(T + z0 * sz + y1 * sy + x0), p110 = __ldg(T + z0 * sz + y1 * sy + x1);
  const float2 p001 = __ldg(T + z1 * sz + y0 * sy + x0), p101 = __ldg(T + z1 * sz + y0 * sy + x1);
  const float2 p011 = __ldg(T + z1 * sz + y1 * sy + x0), p111 = __ldg(T + z1 * sz + y1 * sy + x1);
  float2 o;
  o.x = lerpf(lerpf(lerpf(p000.x, p100.x, a), lerpf(p010.x, p110.x, a), b), lerpf(lerpf(p001.x, p101.x, a), lerpf(p011.x, p111.x, a), b), g);
  o.y = lerpf(lerpf(lerpf(p000.y, p100.y, a), lerpf(p010.y, p110.y, a), b), lerpf(lerpf(p001.y, p101.y, a), lerpf(p011.y, p111.y, a), b), g);
  return o;
}

// Bilinear RGB8 fetch: normalised fixed point c/255, LINEAR + CLAMP_TO_EDGE.
__device__ __forceinline__ float3 tex2d_rgb8(const uint8_t* __restrict__ img, int W, int H, float s, float t) {
  int x0, x1, y0, y1; float a, b;
  lin_coord(s, W, x0, x1, a);
  lin_coord(t, H, y0, y1, b);
  const uint8_t* p00 = img + ((size_t)y0 * W + x0) * 3;
  const uint8_t* p10 = img + ((size_t)y0 * W + x1) * 3;
  const uint8_t* p01 = img + ((size_t)y1 * W + x0) * 3;
  const uint8_t* p11 = img + ((size_t)y1 * W + x1) * 3;
  float o[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v00 = (float)__ldg(p00 + c) / 255.0f, v10 = (float)__ldg(p10 + c) / 255.0f;
    float v01 = (float)__ldg(p01 + c) / 255.0f, v11 = (float)__ldg(p11 + c) / 255.0f;
    o[c] = lerpf(lerpf(v00, v10, a), lerpf(v01, v11, a), b);
  }
  return make_float3(o[0], o[1], o[2]);
}

// The same fetch with the twelve c/255 divisions read from a 256-entry table the caller built with that very division
// (lut[c] = (float)c / 255.0f, IEEE: the quotient is the same number wherever it is computed): identical results, a
// shared-memory load instead of a ~10-instruction division per tap.
__device__ __forceinline__ float3 tex2d_rgb8_lut(const uint8_t* __restrict__ img, int W, int H, float s, float t, const float* __restrict__ lut) {
  int x0, x1, y0, y1; float a, b;
  lin_coord(s, W, x0, x1, a);
  lin_coord(t, H, y0, y1, b);
  const uint8_t* p00 = img + ((size_t)y0 * W + x0) * 3;
  const uint8_t* p10 = img + ((size_t)y0 * W + x1) * 3;
  const uint8_t* p01 = img + ((size_t)y1 * W + x0) * 3;
  const uint8_t* p11 = img + ((size_t)y1 * W + x1) * 3;
  float o[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v00 = lut[__ldg(p00 + c)], v10 = lut[__ldg(p10 + c)], v01 = lut[__ldg(p01 + c)], v11 = lut[__ldg(p11 + c)];
    o[c] = lerpf(lerpf(v00, v10, a), lerpf(v01, v11, a), b);
  }
  return make_float3(o[0], o[1], o[2]);
}

}  // namespace rr

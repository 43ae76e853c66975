// Shared by the drawing reconstructions (rr_points.cu, rr_trigrid.cu): column-major 4x4 products with explicit fma chains,
// shade() of glsl/shading.glsl:32-69, the sensors' debug colours (shading.glsl camera_colors) and the fp64 cofactor inverse the
// per-view matrices are derived with on the host.
#pragma once
#include "rr_math.cuh"

namespace rr {

__device__ __forceinline__ float4 pmulv(const float* m, float4 v) {
  float4 o;
  o.x = fmaf(m[12], v.w, fmaf(m[8], v.z, fmaf(m[4], v.y, m[0] * v.x)));
  o.y = fmaf(m[13], v.w, fmaf(m[9], v.z, fmaf(m[5], v.y, m[1] * v.x)));
  o.z = fmaf(m[14], v.w, fmaf(m[10], v.z, fmaf(m[6], v.y, m[2] * v.x)));
  o.w = fmaf(m[15], v.w, fmaf(m[11], v.z, fmaf(m[7], v.y, m[3] * v.x)));
  return o;
}

// shading.glsl:32-69 (the same arithmetic as rr_raymarch.cu's shade)
__device__ __forceinline__ float3 pshade(int shade_mode, const float* mvT3, float3 view_pos, float3 n, float3 diffuse) {
  if (shade_mode == 0) return diffuse;
  if (shade_mode == 1) {
    const float3 light_pos = make_float3(1.5f, 1.0f, 1.0f), light_diffuse = make_float3(1.0f, 0.9f, 0.7f);
    const float3 light_ambient = light_diffuse * 0.2f;
    float diff = 0.0f, spec = 0.0f;
    const float3 to_light = normalize3(light_pos - view_pos);
    const float light_angle = dot3(n, to_light);
    if (!(light_angle <= 0.0f)) {
      diff = gmax(light_angle, 0.0f);
      const float3 to_viewer = normalize3(make_float3(-view_pos.x, -view_pos.y, -view_pos.z));
      const float3 halfway = normalize3(to_light + to_viewer);
      const float reflected = dot3(halfway, n);
      spec = gpow(reflected, 20.0f);
      const float a = (1.0f - light_angle) * (1.0f - light_angle);
      spec *= 1.0f - a * a * a;
    }
    const float3 amb = light_ambient * 0.5f;
    const float3 dif = (light_diffuse * 0.5f) * diff;
    const float sp = (1.0f * 0.5f) * spec;
    return make_float3((amb.x + dif.x) + sp, (amb.y + dif.y) + sp, (amb.z + dif.z) + sp);
  }
  if (shade_mode == 2) {
    const float* t = mvT3;
    return make_float3(fmaf(t[6], n.z, fmaf(t[3], n.y, t[0] * n.x)), fmaf(t[7], n.z, fmaf(t[4], n.y, t[1] * n.x)),
                       fmaf(t[8], n.z, fmaf(t[5], n.y, t[2] * n.x)));
  }
  return make_float3(1.0f, 1.0f, 1.0f);
}

static __constant__ float kPointCameraColors[5][3] = {{228.f / 255.f, 26.f / 255.f, 28.f / 255.f}, {55.f / 255.f, 126.f / 255.f, 184.f / 255.f},
                                               {77.f / 255.f, 175.f / 255.f, 74.f / 255.f}, {152.f / 255.f, 78.f / 255.f, 163.f / 255.f},
                                               {255.f / 255.f, 127.f / 255.f, 0.f / 255.f}};

// 4x4 inverse (adjugate / determinant) in double, as rr_raymarch.cu derives its matrices
static inline bool pinvert4(const double* m, double* out) {
  double inv[16];
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
  double det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
  if (det == 0.0) return false;
  det = 1.0 / det;
  for (int i = 0; i < 16; ++i) out[i] = inv[i] * det;
  return true;
}

}  // namespace rr

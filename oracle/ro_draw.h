// ORACLE (test infrastructure, NOT product code): pieces shared by the restatements of the reference's drawing
// reconstructions (ro_points.cpp, ro_trigrid.cpp): column-major 4x4 products with the fma chains of ro_raymarch.cpp, the
// fp64 cofactor inverse the matrices of a view are derived with, the bilinear RGB8 fetch of kinect_colors, the sensors'
// debug colours and shade() of glsl/shading.glsl:32-69.
#pragma once
#include "ro_math.h"

#include <cstdint>

namespace ro {


static inline V4 mulv(const float* m, V4 v) {
  V4 o;
  o.x = fmaf(m[12], v.w, fmaf(m[8], v.z, fmaf(m[4], v.y, m[0] * v.x)));
  o.y = fmaf(m[13], v.w, fmaf(m[9], v.z, fmaf(m[5], v.y, m[1] * v.x)));
  o.z = fmaf(m[14], v.w, fmaf(m[10], v.z, fmaf(m[6], v.y, m[2] * v.x)));
  o.w = fmaf(m[15], v.w, fmaf(m[11], v.z, fmaf(m[7], v.y, m[3] * v.x)));
  return o;
}

static inline bool inverse4(const double* m, double* out) {
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

static inline V3 fetch_rgb8(const uint8_t* img, int W, int H, float s, float t) {
  int x0, x1, y0, y1; float a, b;
  lin_coord(s, W, x0, x1, a);
  lin_coord(t, H, y0, y1, b);
  float o[3];
  for (int c = 0; c < 3; ++c) {
    float v00 = (float)img[((size_t)y0 * W + x0) * 3 + c] / 255.0f, v10 = (float)img[((size_t)y0 * W + x1) * 3 + c] / 255.0f;
    float v01 = (float)img[((size_t)y1 * W + x0) * 3 + c] / 255.0f, v11 = (float)img[((size_t)y1 * W + x1) * 3 + c] / 255.0f;
    o[c] = lerpf(lerpf(v00, v10, a), lerpf(v01, v11, a), b);
  }
  return {o[0], o[1], o[2]};
}

static const float kCameraColors[5][3] = {{228.f / 255.f, 26.f / 255.f, 28.f / 255.f}, {55.f / 255.f, 126.f / 255.f, 184.f / 255.f},
                                   {77.f / 255.f, 175.f / 255.f, 74.f / 255.f}, {152.f / 255.f, 78.f / 255.f, 163.f / 255.f},
                                   {255.f / 255.f, 127.f / 255.f, 0.f / 255.f}};

// shading.glsl:32-69
static inline V3 shade(int shade_mode, const float* mvT3, V3 view_pos, V3 view_normal, V3 diffuse) {
  if (shade_mode == 0) return diffuse;
  if (shade_mode == 1) {
    const V3 light_pos{1.5f, 1.0f, 1.0f}, light_diffuse{1.0f, 0.9f, 0.7f};
    const V3 light_ambient = light_diffuse * 0.2f;
    float diff = 0.0f, spec = 0.0f;
    V3 to_light = normalize3(light_pos - view_pos);
    float light_angle = dot3(view_normal, to_light);
    if (!(light_angle <= 0.0f)) {
      diff = gl_max(light_angle, 0.0f);
      V3 to_viewer = normalize3(V3{-view_pos.x, -view_pos.y, -view_pos.z});
      V3 halfway = normalize3(to_light + to_viewer);
      float reflected = dot3(halfway, view_normal);
      spec = gl_pow(reflected, 20.0f);
      float a = (1.0f - light_angle) * (1.0f - light_angle);
      spec *= 1.0f - a * a * a;
    }
    V3 amb = light_ambient * 0.5f;
    V3 dif = (light_diffuse * 0.5f) * diff;
    float sp = (1.0f * 0.5f) * spec;
    return V3{(amb.x + dif.x) + sp, (amb.y + dif.y) + sp, (amb.z + dif.z) + sp};
  }
  if (shade_mode == 2)       // (inverse(gl_NormalMatrix) * vec4(n, 0)).xyz = transpose(mat3(modelview)) * n
    return V3{fmaf(mvT3[6], view_normal.z, fmaf(mvT3[3], view_normal.y, mvT3[0] * view_normal.x)),
              fmaf(mvT3[7], view_normal.z, fmaf(mvT3[4], view_normal.y, mvT3[1] * view_normal.x)),
              fmaf(mvT3[8], view_normal.z, fmaf(mvT3[5], view_normal.y, mvT3[2] * view_normal.x))};
  return V3{1.0f, 1.0f, 1.0f};
}

}  // namespace ro

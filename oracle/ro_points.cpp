// ORACLE (test infrastructure, NOT product code). CPU restatement of the reference's two point-drawing reconstructions
// (SURVEY.md §8f-4), serial and in the draw order OpenGL prescribes:
//   * ReconPoints::draw (framework/reconstruction/recon_points.cpp:46-52 vertex buffer, :71-111 draw) with glsl/points.vs
//     (:24-37), points.gs (:39-60: bbox / depth cull, gl_PointSize = max_size / |pos_eye|), points.fs (:36-75: colour-view
//     border cull, shade() of shading.glsl or the sensor's camera colour), depth test GL_LESS;
//   * ReconCalibs::draw (recon_calibs.cpp:39-46,56-66) over VolumeSampler::sample (rendering/volume_sampler.cpp:14-23,71-73)
//     with glsl/calib_vis.vs (:25-38) and calib_vis.fs (:17-29).
// Rasterisation follows OpenGL 4.4: a point whose centre is outside the clip volume is culled (§13.5); window coordinates
// by the viewport transform with depth range [0, 1] (§13.6.1); a point sprite produces a fragment for every pixel whose
// centre lies inside the square of side gl_PointSize (clamped to the implementation's range, taken as [1, 256]; 1 where the
// shader does not write it) centred at
// the point (§14.4.1; the square is taken half-open, [c - s/2, c + s/2)); fragments are depth-tested in draw order, an
// equal depth fails GL_LESS, so the first drawn fragment keeps the pixel. Depth is kept in binary32.
// PARITY: the shader stages are pinned against the reference's own shader sources run on the CPU where
// tests/test_oracle_cpu.py says so; the rasterisation rule is the specification's.
#include "ro_draw.h"
#include "ro_math.h"
#include "rr_oracle.h"

#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

using namespace ro;

namespace {

// clip test + perspective divide + viewport transform
inline bool to_window(V4 clip, int vw, int vh, float& xw, float& yw, float& zw) {
  if (!(clip.w > 0.0f) || !(fabsf(clip.x) <= clip.w) || !(fabsf(clip.y) <= clip.w) || !(fabsf(clip.z) <= clip.w)) return false;
  const float nx = clip.x / clip.w, ny = clip.y / clip.w, nz = clip.z / clip.w;
  xw = (nx * 0.5f + 0.5f) * (float)vw;
  yw = (ny * 0.5f + 0.5f) * (float)vh;
  zw = nz * 0.5f + 0.5f;
  return true;
}

// the depth-tested square of one point, in draw order
inline void raster_point(float xw, float yw, float zw, float size, const float* rgba, int vw, int vh, float* out_rgba, float* out_depth) {
  const float h = size * 0.5f;
  int x0 = (int)ceilf(xw - h - 0.5f), x1 = (int)ceilf(xw + h - 0.5f);
  int y0 = (int)ceilf(yw - h - 0.5f), y1 = (int)ceilf(yw + h - 0.5f);
  if (x0 < 0) x0 = 0;
  if (y0 < 0) y0 = 0;
  if (x1 > vw) x1 = vw;
  if (y1 > vh) y1 = vh;
  for (int y = y0; y < y1; ++y)
    for (int x = x0; x < x1; ++x) {
      const size_t o = (size_t)y * vw + x;
      if (zw < out_depth[o]) {                       // GL_LESS
        out_depth[o] = zw;
        std::memcpy(out_rgba + o * 4, rgba, 4 * sizeof(float));
      }
    }
}

}  // namespace

extern "C" {

// depth_b float32 [N][H][W][2], normal float32 [N][H][W][3], colour uint8 [N][CH][CW][3], cv_xyz float32 [N][Z][Y][X][3],
// cv_uv float32 [N][Z][Y][X][2] (one resolution for all sensors). out_rgba [vh][vw][4], out_depth [vh][vw].
void ro_draw_points(int N, int W, int H, const float* depth_b, const float* normal, const uint8_t* color, int CW, int CH,
                    const float* cv_xyz, const float* cv_uv, const int32_t* cv_res, const float* bmin, const float* bmax,
                    const float* mv, const float* proj, int vw, int vh, int shade_mode, float* out_rgba, float* out_depth) {
  const int CX = cv_res[0], CY = cv_res[1], CZ = cv_res[2];
  const size_t cvn = (size_t)CX * CY * CZ, px = (size_t)W * H;
  double MV[16], inv[16];
  for (int i = 0; i < 16; ++i) MV[i] = mv[i];
  float normal_matrix[16] = {0}, mvT3[9];
  if (inverse4(MV, inv))
    for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) normal_matrix[c * 4 + r] = (float)inv[r * 4 + c];
  for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) mvT3[c * 3 + r] = mv[r * 4 + c];
  for (size_t i = 0; i < (size_t)vw * vh; ++i) { out_depth[i] = 1.0f; out_rgba[4 * i] = out_rgba[4 * i + 1] = out_rgba[4 * i + 2] = out_rgba[4 * i + 3] = 0.0f; }
  const float stepX = 1.0f / (float)W, stepY = 1.0f / (float)H;
  for (int layer = 0; layer < N; ++layer)                                   // recon_points.cpp:105-109: one draw call per sensor
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        const float sx = (float)(((double)x + 0.5) * (double)stepX), sy = (float)(((double)y + 0.5) * (double)stepY);
        const float depth = depth_b[((size_t)layer * px + (size_t)y * W + x) * 2];
        float pc[3], tc[2];
        tex3d_linear<3>(cv_xyz + (size_t)layer * cvn * 3, CX, CY, CZ, sx, sy, depth, pc, 3);
        const bool in_box = pc[0] >= bmin[0] && pc[1] >= bmin[1] && pc[2] >= bmin[2] && pc[0] <= bmax[0] && pc[1] <= bmax[1] && pc[2] <= bmax[2];
        if (!in_box || depth <= 0.0f) continue;                            // points.gs:39-41
        tex3d_linear<2>(cv_uv + (size_t)layer * cvn * 2, CX, CY, CZ, sx, sy, depth, tc, 2);
        if (tc[0] > 0.99f || tc[0] < 0.01f || tc[1] > 0.99f || tc[1] < 0.01f) continue;      // points.fs:38-41 (flat: the whole point)
        const V4 es = mulv(mv, V4{pc[0], pc[1], pc[2], 1.0f});
        const V3 pos_es{es.x, es.y, es.z};
        float xw, yw, zw;
        if (!to_window(mulv(proj, es), vw, vh, xw, yw, zw)) continue;
        const float dist = sqrtf(dot3(pos_es, pos_es));
        const float max_size = shade_mode == 3 ? 4.0f : 10.0f;
        const float size = gl_min(gl_max(max_size / dist, 1.0f), 256.0f);     // ALIASED_POINT_SIZE_RANGE taken as [1, 256]
        V3 c;
        if (shade_mode == 3) {
          const float* cc = kCameraColors[layer < 5 ? layer : 4];
          c = V3{cc[0], cc[1], cc[2]};
        } else {
          const V3 col = fetch_rgb8(color + (size_t)CW * CH * 3 * layer, CW, CH, tc[0], tc[1]);
          const float* n3 = normal + ((size_t)layer * px + (size_t)y * W + x) * 3;
          const V4 vn = mulv(normal_matrix, V4{n3[0], n3[1], n3[2], 0.0f});
          c = shade(shade_mode, mvT3, pos_es, V3{vn.x, vn.y, vn.z}, col);
        }
        const float rgba[4] = {c.x, c.y, c.z, 1.0f};
        raster_point(xw, yw, zw, size, rgba, vw, vh, out_rgba, out_depth);
      }
}

// tsdf float32 [Z][Y][X] (the TSDF half of half2 voxels already widened); the sample grid has the inverse volumes' resolution.
void ro_draw_calibs(const float* tsdf, const uint32_t* res, const int32_t* inv_res, float limit, const float* bmin, const float* bmax,
                    const float* mv, const float* proj, int vw, int vh, float* out_rgba, float* out_depth) {
  const int X = (int)res[0], Y = (int)res[1], Z = (int)res[2], IX = inv_res[0], IY = inv_res[1], IZ = inv_res[2];
  float v2w[16] = {0};
  v2w[0] = bmax[0] - bmin[0]; v2w[5] = bmax[1] - bmin[1]; v2w[10] = bmax[2] - bmin[2];
  v2w[12] = bmin[0]; v2w[13] = bmin[1]; v2w[14] = bmin[2]; v2w[15] = 1.0f;
  for (size_t i = 0; i < (size_t)vw * vh; ++i) { out_depth[i] = 1.0f; out_rgba[4 * i] = out_rgba[4 * i + 1] = out_rgba[4 * i + 2] = out_rgba[4 * i + 3] = 0.0f; }
  for (int z = 0; z < IZ; ++z)
    for (int y = 0; y < IY; ++y)
      for (int x = 0; x < IX; ++x) {
        const float qx = ((float)x + 0.5f) * (1.0f / (float)IX), qy = ((float)y + 0.5f) * (1.0f / (float)IY), qz = ((float)z + 0.5f) * (1.0f / (float)IZ);
        // texture(volume_tsdf, q).r
        int x0, x1, y0, y1, z0, z1; float a, b, g;
        lin_coord(qx, X, x0, x1, a);
        lin_coord(qy, Y, y0, y1, b);
        lin_coord(qz, Z, z0, z1, g);
        const size_t sy = (size_t)X, sz = (size_t)X * Y;
        const float c00 = lerpf(tsdf[z0 * sz + y0 * sy + x0], tsdf[z0 * sz + y0 * sy + x1], a);
        const float c10 = lerpf(tsdf[z0 * sz + y1 * sy + x0], tsdf[z0 * sz + y1 * sy + x1], a);
        const float c01 = lerpf(tsdf[z1 * sz + y0 * sy + x0], tsdf[z1 * sz + y0 * sy + x1], a);
        const float c11 = lerpf(tsdf[z1 * sz + y1 * sy + x0], tsdf[z1 * sz + y1 * sy + x1], a);
        const float distance = lerpf(lerpf(c00, c10, b), lerpf(c01, c11, b), g);
        if (distance <= -limit) continue;                                   // calib_vis.fs:29
        const V4 world = mulv(v2w, V4{qx, qy, qz, 1.0f});
        const V4 view = mulv(mv, V4{world.x, world.y, world.z, 1.0f});
        float xw, yw, zw;
        if (!to_window(mulv(proj, V4{view.x, view.y, view.z, 1.0f}), vw, vh, xw, yw, zw)) continue;
        const float inverted = fabsf(distance) / limit;
        float rgba[4] = {0.0f, 0.0f, 0.0f, 1.0f};
        if (distance > 0.0f) rgba[0] = 1.0f - inverted; else rgba[1] = 1.0f - inverted;
        if (distance >= limit) { rgba[0] = 0.0f; rgba[1] = 0.0f; rgba[2] = 1.0f; }
        raster_point(xw, yw, zw, 1.0f, rgba, vw, vh, out_rgba, out_depth);
      }
}

}  // extern "C"

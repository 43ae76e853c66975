// ORACLE (test infrastructure, NOT product code). CPU restatement of ReconTrigrid::draw, serial and in
// the draw order OpenGL prescribes (framework/reconstruction/recon_trigrid.cpp):
//   * :48-61 the triangle grid: two triangles per cell, six vec2 per cell at ((x + 0.5 | 1.5) * stepX, (y + 0.5 | 1.5) * stepY)
//     with stepX = 1 / tex_width, stepY = 1 / tex_height - and, as the reference writes it, y running to tex_WIDTH and x to
//     tex_HEIGHT (:51-52): cells x < H, y < W. The rows beyond the image clamp (CLAMP_TO_EDGE) into zero-area triangles, the
//     columns x >= H are never drawn; restated as written;
//   * :82-149 draw(): pass 1 (stage 0) depth only, GL_LESS against a depth buffer cleared to 1; pass 2 (stage 1) without depth
//     test, additive blending (ONE, ONE) into a cleared RGBA32F target; pass 3 trigrid_normalize.fs;
//   * glsl/trigrid_accum.vs:22-35 (depth NEAREST, quality LINEAR, cv_xyz / cv_uv trilinear), trigrid_accum.gs:27-73
//     (validSurface: no invalid depth, every edge shorter than min_length * avg_depth * 4; flat eye-space normal),
//     trigrid_accum.fs:41-80 (bbox, colour-view border, back face; stage 1: within epsilon of the pass-1 surface, then
//     shade() * quality, quality), trigrid_normalize.fs:13-31 (colour / alpha where alpha > 0, depth of pass 1).
// Fixed-function stages: ro_raster.h (OpenGL 4.4 clipping, viewport transform, coverage, interpolation; fp64). The sums of the
// additive blend are binary32 additions in draw order (layer, then triangle), which is what in-order blending gives. Depth is
// kept in binary32 (the reference's attachment is DEPTH_COMPONENT32).
// PARITY: pinned against the reference's own trigrid shaders run on the CPU through the same fixed-function stages
// (tests/test_oracle_cpu.py, golden ref_glsl_trigrid.npz).
#include "ro_draw.h"
#include "ro_math.h"
#include "ro_raster.h"
#include "rr_oracle.h"

#include <cmath>
#include <cstdint>
#include <vector>

using namespace ro;

namespace {

struct TVert {
  float clip[4];
  V3 pos_es, pos_cs;
  V2 tc;
  float depth, quality;
};

}  // namespace

extern "C" {

// depth_b float32 [N][H][W][2] (kinect_depths), quality float32 [N][H][W] (kinect_qualities), colour uint8 [N][CH][CW][3],
// cv_xyz float32 [N][Z][Y][X][3], cv_uv float32 [N][Z][Y][X][2]. out_rgba [vh][vw][4], out_depth [vh][vw]; out_accum
// (optional) [vh][vw][4] receives the accumulation target of pass 2, out_depth1 (optional) [vh][vw] the depth of pass 1.
void ro_draw_trigrid(int N, int W, int H, const float* depth_b, const float* quality, const uint8_t* color, int CW, int CH,
                     const float* cv_xyz, const float* cv_uv, const int32_t* cv_res, const float* bmin, const float* bmax,
                     const float* mv, const float* proj, int vw, int vh, int shade_mode, float min_length, float epsilon,
                     float* out_rgba, float* out_depth, float* out_accum, float* out_depth1) {
  const int CX = cv_res[0], CY = cv_res[1], CZ = cv_res[2];
  const size_t cvn = (size_t)CX * CY * CZ, px = (size_t)W * H, npx = (size_t)vw * vh;
  float mvT3[9], img_to_eye[16] = {0};
  for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) mvT3[c * 3 + r] = mv[r * 4 + c];
  {
    // image_to_eye = inverse(viewport_scale * viewport_translate * projection)  (recon_trigrid.cpp:84-95)
    double P[16], t[16], t2[16], inv[16];
    for (int i = 0; i < 16; ++i) P[i] = proj[i];
    const double Tr[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 1, 1, 1, 1};
    const double Sc[16] = {vw * 0.5, 0, 0, 0, 0, vh * 0.5, 0, 0, 0, 0, 0.5, 0, 0, 0, 0, 1};
    auto mul = [](const double* a, const double* b, double* o) {
      for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) { double acc = 0.0; for (int k = 0; k < 4; ++k) acc += a[k * 4 + r] * b[c * 4 + k]; o[c * 4 + r] = acc; }
    };
    mul(Tr, P, t); mul(Sc, t, t2);
    if (inverse4(t2, inv)) for (int i = 0; i < 16; ++i) img_to_eye[i] = (float)inv[i];
  }
  std::vector<float> depth1(npx, 1.0f), accum(npx * 4, 0.0f);

  // the vertex stage once per grid vertex: (H + 1) columns, (W + 1) rows per sensor
  const int GW = H + 1, GH = W + 1;
  const float stepX = 1.0f / (float)W, stepY = 1.0f / (float)H;
  std::vector<TVert> verts((size_t)N * GW * GH);
  for (int layer = 0; layer < N; ++layer)
    for (int j = 0; j < GH; ++j)
      for (int i = 0; i < GW; ++i) {
        TVert& v = verts[((size_t)layer * GH + j) * GW + i];
        const float sx = (float)(((double)i + 0.5) * (double)stepX), sy = (float)(((double)j + 0.5) * (double)stepY);
        v.depth = tex2d_nearest(depth_b + (size_t)layer * px * 2, W, H, 2, 0, sx, sy);
        float pc[3], tc[2];
        tex3d_linear<3>(cv_xyz + (size_t)layer * cvn * 3, CX, CY, CZ, sx, sy, v.depth, pc, 3);
        tex3d_linear<2>(cv_uv + (size_t)layer * cvn * 2, CX, CY, CZ, sx, sy, v.depth, tc, 2);
        v.pos_cs = V3{pc[0], pc[1], pc[2]};
        v.tc = V2{tc[0], tc[1]};
        const V4 es = mulv(mv, V4{pc[0], pc[1], pc[2], 1.0f});
        v.pos_es = V3{es.x, es.y, es.z};
        const V4 clip = mulv(proj, es);
        v.clip[0] = clip.x; v.clip[1] = clip.y; v.clip[2] = clip.z; v.clip[3] = clip.w;
        v.quality = tex2d_linear(quality + (size_t)layer * px, W, H, 1, 0, sx, sy);
      }

  for (int stage = 0; stage < 2; ++stage)
    for (int layer = 0; layer < N; ++layer)
      for (int y = 0; y < W; ++y)                       // recon_trigrid.cpp:51-52, as written
        for (int x = 0; x < H; ++x)
          for (int k = 0; k < 2; ++k) {
            const TVert* g = verts.data() + (size_t)layer * GH * GW;
            const TVert& v0 = k == 0 ? g[(size_t)y * GW + x] : g[(size_t)y * GW + x + 1];
            const TVert& v1 = k == 0 ? g[(size_t)y * GW + x + 1] : g[(size_t)(y + 1) * GW + x + 1];
            const TVert& v2 = g[(size_t)(y + 1) * GW + x];
            // trigrid_accum.gs:27-37,44-55
            if (v0.depth < 0.0f || v1.depth < 0.0f || v2.depth < 0.0f) continue;
            const float avg_depth = (v0.depth + v1.depth + v2.depth) / 3.0f;
            const float l = min_length * avg_depth * 4.0f;
            if (!(length3(v1.pos_cs - v0.pos_cs) < l) || !(length3(v2.pos_cs - v0.pos_cs) < l) || !(length3(v2.pos_cs - v1.pos_cs) < l)) continue;
            const V3 tri_normal = normalize3(cross3(v1.pos_es - v0.pos_es, v2.pos_es - v0.pos_es));
            const float clip[3][4] = {{v0.clip[0], v0.clip[1], v0.clip[2], v0.clip[3]}, {v1.clip[0], v1.clip[1], v1.clip[2], v1.clip[3]},
                                      {v2.clip[0], v2.clip[1], v2.clip[2], v2.clip[3]}};
            raster_triangle(clip, vw, vh, [&](int fx, int fy, float zw, const double* B) {
              const V3 pos_cs{rinterp(B, v0.pos_cs.x, v1.pos_cs.x, v2.pos_cs.x), rinterp(B, v0.pos_cs.y, v1.pos_cs.y, v2.pos_cs.y),
                              rinterp(B, v0.pos_cs.z, v1.pos_cs.z, v2.pos_cs.z)};
              // trigrid_accum.fs:43-45
              if (!(pos_cs.x >= bmin[0] && pos_cs.y >= bmin[1] && pos_cs.z >= bmin[2] && pos_cs.x <= bmax[0] && pos_cs.y <= bmax[1] && pos_cs.z <= bmax[2])) return;
              const float s = rinterp(B, v0.tc.x, v1.tc.x, v2.tc.x), t = rinterp(B, v0.tc.y, v1.tc.y, v2.tc.y);
              if (s > 0.99f || s < 0.01f || t > 0.99f || t < 0.01f) return;                 // :48-51
              const V3 pos_es{rinterp(B, v0.pos_es.x, v1.pos_es.x, v2.pos_es.x), rinterp(B, v0.pos_es.y, v1.pos_es.y, v2.pos_es.y),
                              rinterp(B, v0.pos_es.z, v1.pos_es.z, v2.pos_es.z)};
              const V3 nn = normalize3(tri_normal);
              const V3 normal{-nn.x, -nn.y, -nn.z};
              if (dot3(normal, normalize3(pos_es)) > 0.0f) return;                          // :55-57 (a NaN normal passes, as in GLSL)
              const size_t o = (size_t)fy * vw + fx;
              if (stage == 0) {
                if (zw < depth1[o]) depth1[o] = zw;                                         // GL_LESS
                return;
              }
              // :60-69
              const float depth_curr = depth1[o];
              const V4 pc = mulv(img_to_eye, V4{((float)fx + 0.5f) + 0.5f, ((float)fy + 0.5f) + 0.5f, depth_curr, 1.0f});
              const V3 cur{pc.x / pc.w, pc.y / pc.w, pc.z / pc.w};
              if (epsilon < length3(cur - pos_es)) return;
              const float q = rinterp(B, v0.quality, v1.quality, v2.quality);
              V3 c;
              if (shade_mode == 3) {
                const float* cc = kCameraColors[layer < 5 ? layer : 4];
                c = V3{cc[0], cc[1], cc[2]};
              } else {
                c = shade(shade_mode, mvT3, pos_es, normal, fetch_rgb8(color + (size_t)CW * CH * 3 * layer, CW, CH, s, t));
              }
              float* a = accum.data() + o * 4;
              a[0] += c.x * q; a[1] += c.y * q; a[2] += c.z * q; a[3] += q;
            });
          }

  // trigrid_normalize.fs:13-31
  for (size_t o = 0; o < npx; ++o) {
    const float* a = accum.data() + o * 4;
    if (a[3] > 0.0f) {
      out_rgba[o * 4] = a[0] / a[3]; out_rgba[o * 4 + 1] = a[1] / a[3]; out_rgba[o * 4 + 2] = a[2] / a[3]; out_rgba[o * 4 + 3] = a[3] / a[3];
      out_depth[o] = depth1[o];
    } else {
      out_rgba[o * 4] = out_rgba[o * 4 + 1] = out_rgba[o * 4 + 2] = out_rgba[o * 4 + 3] = 0.0f;
      out_depth[o] = 1.0f;
    }
    if (out_accum) for (int c = 0; c < 4; ++c) out_accum[o * 4 + c] = a[c];
    if (out_depth1) out_depth1[o] = depth1[o];
  }
}

}  // extern "C"

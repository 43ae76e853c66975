// ORACLE TOOLING (test infrastructure, NOT product code).
// Runs the reference's own shader sources on the CPU: every .inc included below is generated at build time by
// transpile.py from /root/reference/glsl/<name> (see that script for the exact textual changes) and compiled against the
// GLSL host environment of glsl_compat.hpp. The entry points take the same arrays as the oracle's ro_pre_* / ro_integrate
// (oracle/rr_oracle.h), so tests can put the two side by side on identical inputs. The pass order, bindings, sampler
// filters and uniforms set here are the ones the reference's host code sets:
//   pre_morph / pre_depth / pre_boundary / pre_normal / pre_quality: NetKinectArray::processDepth + processTextures
//     (framework/NetKinectArray.cpp:251-290, 311-428), filters from :163-197 (depth arrays NEAREST, everything else LINEAR),
//     pass_TexCoord at pixel centres (framework/rendering/screen_quad.cpp:11-15), texSizeInv = 1/resolution (:197);
//   tsdf_integration: ReconIntegration::integrate (framework/reconstruction/recon_integration.cpp:243-270), one vertex
//     per voxel centre (framework/rendering/volume_sampler.cpp:33-48).
// Built only where the reference tree is present; output oracle/_ref/libref_glsl.so (git-ignored).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "glsl_compat.hpp"
#include "../ro_raster.h"

namespace glsl {
struct S_pre_morph {
#include "pre_morph.inc"
};
struct S_pre_depth {
#include "pre_depth.inc"
};
struct S_pre_boundary {
#include "pre_boundary.inc"
};
struct S_pre_normal {
#include "pre_normal.inc"
};
struct S_pre_quality {
#include "pre_quality.inc"
};
struct S_tsdf_integration {
#include "tsdf_integration.inc"
};
struct S_tsdf_raymarch {
  // fragment-shader built-ins the source refers to
  vec4 gl_FragCoord;
  struct { float near, far, diff; } gl_DepthRange = {0.0f, 1.0f, 1.0f};       // glDepthRange defaults
  bool discarded = false;
#include "tsdf_raymarch.inc"
};
struct S_bricks_vs {
  int gl_InstanceID = 0;
  vec4 gl_Position;
#include "bricks_vs.inc"
};
struct S_bricks_gs {
  struct { vec4 gl_Position; } gl_in[3];
  vec4 gl_Position;
  std::vector<vec4> emitted;
  void EmitVertex() { emitted.push_back(gl_Position); }
#include "bricks_gs.inc"
};
struct S_bricks_fs {
  vec4 gl_FragCoord;
  bool gl_FrontFacing = true;
#include "bricks_fs.inc"
};
struct S_points_vs {
  vec4 gl_Position;
#include "points_vs.inc"
};
struct S_points_gs {
  struct { vec4 gl_Position; } gl_in[1];
  vec4 gl_Position;
  float gl_PointSize = 1.0f;
  int emitted = 0;
  void EmitVertex() { ++emitted; }
#include "points_gs.inc"
};
struct S_points_fs {
  vec4 gl_FragCoord;
  vec2 gl_PointCoord;
  bool discarded = false;
#include "points_fs.inc"
};
struct S_calib_vis_vs {
  vec4 gl_Position;
#include "calib_vis_vs.inc"
};
struct S_calib_vis_fs {
  bool discarded = false;
#include "calib_vis_fs.inc"
};
struct S_trigrid_vs {
  vec4 gl_Position;
#include "trigrid_vs.inc"
};
struct S_trigrid_gs {
  struct { vec4 gl_Position; } gl_in[3];
  vec4 gl_Position;
  struct Emitted { vec4 pos; vec2 texcoord; vec3 pos_es, pos_cs, normal_es; float depth, quality; };
  std::vector<Emitted> emitted;
  void EmitVertex() { emitted.push_back(Emitted{gl_Position, pass_texcoord, pass_pos_es, pass_pos_cs, pass_normal_es, pass_depth, pass_quality}); }
  void EndPrimitive() {}
#include "trigrid_gs.inc"
};
struct S_trigrid_fs {
  vec4 gl_FragCoord;
  bool discarded = false;
#include "trigrid_fs.inc"
};
struct S_trigrid_norm_fs {
  vec4 gl_FragCoord, gl_FragColor;
  float gl_FragDepth = 0.0f;
  bool discarded = false;
#include "trigrid_norm_fs.inc"
};
struct S_framebuffer_transfer {
#include "framebuffer_transfer.inc"
};
struct S_tsdf_inpaint {
#include "tsdf_inpaint.inc"
};
struct S_tsdf_colorfill {
#include "tsdf_colorfill.inc"
};
}  // namespace glsl

using namespace glsl;

static sampler2DArray tex2d(const float* p, int W, int H, int C, bool linear) {
  sampler2DArray t;
  t.f32 = p; t.W = W; t.H = H; t.L = 1; t.C = C; t.linear = linear;
  return t;
}
static sampler3D tex3d(const float* p, int X, int Y, int Z, int C) {
  sampler3D t;
  t.f32 = p; t.X = X; t.Y = Y; t.Z = Z; t.C = C; t.linear = true;
  return t;
}
static const float kOne[4] = {1.f, 1.f, 1.f, 1.f};

// Every pass renders ONE layer; the harness presents that layer as a one-layer array and sets `layer` to 0, so
// vec3(coords, layer) selects it and cv_*[layer] is the sensor's volume.
extern "C" {

void rg_pre_morph(const float* depth_in, int W, int H, float* depth_out) {
  S_pre_morph proto{};
  proto.layer = 0u;
  proto.mode = 0u;                                                    // processDepth: mode 0 = dilate (:258-266); mode 1 copies
  proto.kinect_depths = tex2d(depth_in, W, H, 1, false);
  proto.texSizeInv = vec2(1.0f / (float)W, 1.0f / (float)H);
  proto.cv_xyz[0] = tex3d(kOne, 1, 1, 1, 3);                          // sampled by in_bbox(vec2, float), result unused ("return true")
  proto.bbox_min = vec3(0.f); proto.bbox_max = vec3(0.f);
#pragma omp parallel for schedule(dynamic, 4)
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      S_pre_morph s(proto);
      s.pass_TexCoord = vec2(((float)x + 0.5f) / (float)W, ((float)y + 0.5f) / (float)H);
      s.main();
      depth_out[(size_t)y * W + x] = s.out_Depth;
    }
}

void rg_pre_depth(const float* depth_in, int W, int H, const float* cv_xyz, const float* cv_uv, int CX, int CY, int CZ,
                  const uint8_t* color, int CW, int CH, const float* bbox_min, const float* bbox_max, float cv_min_ds,
                  float cv_max_ds, int filter_textures, int compress, float scale, float near_, float scaled_near,
                  float* out_depth, float* out_lab) {
  S_pre_depth proto{};
  proto.layer = 0u;
  proto.kinect_depths = tex2d(depth_in, W, H, 1, false);
  sampler2DArray col;
  col.u8 = color; col.W = CW; col.H = CH; col.L = 1; col.C = 3; col.linear = true;
  proto.kinect_colors = col;
  proto.texSizeInv = vec2(1.0f / (float)W, 1.0f / (float)H);
  proto.filter_textures = filter_textures != 0;
  proto.cv_xyz[0] = tex3d(cv_xyz, CX, CY, CZ, 3);
  proto.cv_uv[0] = tex3d(cv_uv, CX, CY, CZ, 2);
  proto.compress = compress != 0; proto.scale = scale; proto.near = near_; proto.scaled_near = scaled_near;
  proto.cv_min_ds = cv_min_ds; proto.cv_max_ds = cv_max_ds;
  proto.mode = 0; proto.processed_depth = true;
  proto.bbox_min = vec3(bbox_min[0], bbox_min[1], bbox_min[2]);
  proto.bbox_max = vec3(bbox_max[0], bbox_max[1], bbox_max[2]);
#pragma omp parallel for schedule(dynamic, 4)
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      S_pre_depth s(proto);
      s.pass_TexCoord = vec2(((float)x + 0.5f) / (float)W, ((float)y + 0.5f) / (float)H);
      s.main();
      const size_t o = (size_t)y * W + x;
      out_depth[o * 2] = s.out_Depth.x; out_depth[o * 2 + 1] = s.out_Depth.y;
      out_lab[o * 3] = s.out_Color.x; out_lab[o * 3 + 1] = s.out_Color.y; out_lab[o * 3 + 2] = s.out_Color.z;
    }
}

void rg_pre_boundary(const float* depth_rg, const float* lab, int W, int H, int refine, float* out_depth_b, float* out_sil) {
  S_pre_boundary proto{};
  proto.layer = 0u;
  proto.kinect_depths = tex2d(depth_rg, W, H, 2, false);
  proto.kinect_colors_lab = tex2d(lab, W, H, 3, true);
  proto.kinect_colors = tex2d(kOne, 1, 1, 3, true);                   // only recompute_depth() reads these; main() never calls it
  proto.cv_uv[0] = tex3d(kOne, 1, 1, 1, 2);
  proto.refine = refine != 0;
  proto.texSizeInv = vec2(1.0f / (float)W, 1.0f / (float)H);
#pragma omp parallel for schedule(dynamic, 4)
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      S_pre_boundary s(proto);
      s.pass_TexCoord = vec2(((float)x + 0.5f) / (float)W, ((float)y + 0.5f) / (float)H);
      s.main();
      const size_t o = (size_t)y * W + x;
      out_depth_b[o * 2] = s.out_Depth.x; out_depth_b[o * 2 + 1] = s.out_Depth.y;
      out_sil[o] = s.out_Silhouette;
    }
}

void rg_pre_normal(const float* depth_b, int W, int H, const float* cv_xyz, int CX, int CY, int CZ, const float* bbox_min,
                   float brick_size, const uint32_t* brick_res, uint32_t num_bricks, uint32_t* bricks, float* out_normal) {
  S_pre_normal proto{};
  proto.layer = 0u;
  proto.kinect_depths = tex2d(depth_b, W, H, 2, false);
  proto.texSizeInv = vec2(1.0f / (float)W, 1.0f / (float)H);
  proto.cv_xyz[0] = tex3d(cv_xyz, CX, CY, CZ, 3);
  proto.cv_uv[0] = tex3d(kOne, 1, 1, 1, 2);
  proto.bbox_min = vec3(bbox_min[0], bbox_min[1], bbox_min[2]);
  proto.bbox_max = vec3(0.f);
  proto.brick_size = brick_size;
  proto.resolution = uvec3(brick_res[0], brick_res[1], brick_res[2]);
  proto.bricks.p = bricks; proto.bricks.n = num_bricks;
#pragma omp parallel for schedule(dynamic, 4)
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      S_pre_normal s(proto);
      s.pass_TexCoord = vec2(((float)x + 0.5f) / (float)W, ((float)y + 0.5f) / (float)H);
      s.main();
      const size_t o = (size_t)y * W + x;
      out_normal[o * 3] = s.out_Normal.x; out_normal[o * 3 + 1] = s.out_Normal.y; out_normal[o * 3 + 2] = s.out_Normal.z;
    }
}

void rg_pre_quality(const float* depth_b, const float* normals, const float* lab, int W, int H, const float* cv_xyz, int CX, int CY,
                    int CZ, const float* camera_pos, float* out_quality) {
  S_pre_quality proto{};
  proto.layer = 0u;
  proto.kinect_depths = tex2d(depth_b, W, H, 2, false);
  proto.kinect_normals = tex2d(normals, W, H, 3, true);
  proto.kinect_colors_lab = tex2d(lab, W, H, 3, true);                // read by the dead get_color_diff() loop (:115)
  proto.texSizeInv = vec2(1.0f / (float)W, 1.0f / (float)H);
  proto.camera_positions[0] = vec3(camera_pos[0], camera_pos[1], camera_pos[2]);
  proto.processed_depth = true;
  proto.cv_xyz[0] = tex3d(cv_xyz, CX, CY, CZ, 3);
#pragma omp parallel for schedule(dynamic, 4)
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      S_pre_quality s(proto);
      s.pass_TexCoord = vec2(((float)x + 0.5f) / (float)W, ((float)y + 0.5f) / (float)H);
      s.main();
      out_quality[(size_t)y * W + x] = s.out_Quality;
    }
}

// ReconIntegration::integrate: clear to -limit, then one vertex per voxel (all of them, or those of the occupied bricks).
void rg_integrate(int N, const float* inv, const int32_t* inv_res, const float* sil, const float* depth_b, const float* quality,
                  int W, int H, float limit, const uint32_t* res, int use_bricks, const int32_t* brick_ranges,
                  const uint32_t* occupied, uint32_t num_occupied, float* tsdf) {
  if (N > 5) return;                                                  // uniform sampler3D[5] cv_xyz_inv
  const int X = (int)res[0], Y = (int)res[1], Z = (int)res[2];
  const size_t nvox = (size_t)X * Y * Z;
  for (size_t i = 0; i < nvox; ++i) tsdf[i] = -limit;                 // glClearTexImage(-limit), recon_integration.cpp:250-251
  S_tsdf_integration proto{};
  sampler2DArray s_sil = tex2d(sil, W, H, 1, true), s_depth = tex2d(depth_b, W, H, 2, false), s_q = tex2d(quality, W, H, 1, true);
  s_sil.L = s_depth.L = s_q.L = N;
  proto.kinect_silhouettes = s_sil; proto.kinect_depths = s_depth; proto.kinect_qualities = s_q;
  const size_t inv_vox = (size_t)inv_res[0] * inv_res[1] * inv_res[2];
  for (int i = 0; i < N; ++i) {
    sampler3D t = tex3d(inv + (size_t)i * inv_vox * 4, inv_res[0], inv_res[1], inv_res[2], 4);
    proto.cv_xyz_inv[i] = t;
  }
  image3D img;
  img.data = tsdf; img.X = X; img.Y = Y; img.Z = Z;
  proto.volume_tsdf = img;
  proto.limit = limit;
  proto.num_kinects = (uint)N;
  proto.res_tsdf = uvec3(res[0], res[1], res[2]);
  auto run = [&](int x0, int x1, int y0, int y1, int z0, int z1) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int z = z0; z < z1; ++z)
      for (int y = y0; y < y1; ++y)
        for (int x = x0; x < x1; ++x) {
          S_tsdf_integration s(proto);
          // VolumeSampler::sample: (pos + 0.5) * step, step = 1 / res (volume_sampler.cpp:36-42)
          s.in_Position = vec3(((float)x + 0.5f) * (1.0f / (float)X), ((float)y + 0.5f) * (1.0f / (float)Y), ((float)z + 0.5f) * (1.0f / (float)Z));
          s.main();
        }
  };
  if (!use_bricks) {
    run(0, X, 0, Y, 0, Z);
  } else {
    for (uint32_t b = 0; b < num_occupied; ++b) {
      const int32_t* r = brick_ranges + (size_t)occupied[b] * 6;
      run(r[0], r[1], r[2], r[3], r[4], r[5]);
    }
  }
}

// ReconIntegration::draw (framework/reconstruction/recon_integration.cpp:177-241): one fragment per covered pixel of the cube
// proxy. pass_Position is a point of the pixel's ray in volume space (the rasteriser would interpolate the cube's surface
// point; the shader only uses normalize(pass_Position - CameraPos)). uniforms16 = gl_ModelViewMatrix, gl_ProjectionMatrix,
// gl_NormalMatrix, NormalMatrix, vol_to_world, img_to_eye_curr (column-major, 16 floats each). depth_peels may be null
// (skipSpace off). Outputs: rgba, gl_FragDepth, the num-samples image, and a hit flag (0 = discarded / not covered).
void rg_raymarch(const float* tsdf, const uint32_t* res, float limit, int N, const float* inv, const int32_t* inv_res,
                 const float* cv_uv, const int32_t* cv_res, const uint8_t* color, int CW, int CH, const float* depth_b,
                 const float* quality, const float* normals, int W, int H, const float* bbox_min, const float* bbox_max,
                 const float* uniforms16, const float* camera_pos, int vw, int vh, int shade_mode, const float* pass_position,
                 const uint8_t* covered, const float* depth_peels, float* out_rgba, float* out_depth, float* out_samples,
                 uint8_t* out_hit) {
  if (N > 5) return;
  S_tsdf_raymarch proto{};
  sampler2DArray col;
  col.u8 = color; col.W = CW; col.H = CH; col.L = N; col.C = 3; col.linear = true;
  proto.kinect_colors = col;
  sampler2DArray d = tex2d(depth_b, W, H, 2, false), q = tex2d(quality, W, H, 1, true), nr = tex2d(normals, W, H, 3, true);
  d.L = q.L = nr.L = N;
  proto.kinect_depths = d; proto.kinect_qualities = q; proto.kinect_normals = nr;
  const size_t inv_vox = (size_t)inv_res[0] * inv_res[1] * inv_res[2], cv_vox = (size_t)cv_res[0] * cv_res[1] * cv_res[2];
  for (int i = 0; i < N; ++i) {
    proto.cv_xyz_inv[i] = tex3d(inv + (size_t)i * inv_vox * 4, inv_res[0], inv_res[1], inv_res[2], 4);
    proto.cv_uv[i] = tex3d(cv_uv + (size_t)i * cv_vox * 2, cv_res[0], cv_res[1], cv_res[2], 2);
  }
  proto.num_kinects = (uint)N;
  proto.limit = limit;
  proto.sampleDistance = limit * 0.5f;             // the global's initialiser reads the uniform (tsdf_raymarch.fs:33)
  proto.gl_ModelViewMatrix = mat4(uniforms16); proto.gl_ProjectionMatrix = mat4(uniforms16 + 16);
  proto.gl_NormalMatrix = mat4(uniforms16 + 32); proto.NormalMatrix = mat4(uniforms16 + 48);
  proto.vol_to_world = mat4(uniforms16 + 64); proto.img_to_eye_curr = mat4(uniforms16 + 80);
  proto.volume_tsdf = tex3d(tsdf, (int)res[0], (int)res[1], (int)res[2], 1);
  proto.CameraPos = vec3(camera_pos[0], camera_pos[1], camera_pos[2]);
  proto.skipSpace = depth_peels != nullptr;
  sampler2D peels;
  peels.f32 = depth_peels; peels.W = vw; peels.H = vh;
  proto.depth_peels = peels;
  proto.viewport_offset = vec2(0.0f, 0.0f);
  image2D ns;
  ns.data = out_samples; ns.W = vw; ns.H = vh;
  proto.tex_num_samples = ns;
  proto.g_shade_mode = shade_mode;
  proto.bbox_min = vec3(bbox_min[0], bbox_min[1], bbox_min[2]);
  proto.bbox_max = vec3(bbox_max[0], bbox_max[1], bbox_max[2]);
#pragma omp parallel for schedule(dynamic, 4)
  for (int y = 0; y < vh; ++y)
    for (int x = 0; x < vw; ++x) {
      const size_t o = (size_t)y * vw + x;
      out_rgba[o * 4] = out_rgba[o * 4 + 1] = out_rgba[o * 4 + 2] = out_rgba[o * 4 + 3] = 0.0f;   // glClearColor(0,0,0,0)
      out_depth[o] = 1.0f;                                                                          // glClearDepth(1)
      out_samples[o] = 0.0f;
      out_hit[o] = 0;
      if (!covered[o]) continue;
      S_tsdf_raymarch s(proto);
      s.gl_FragCoord = vec4((float)x + 0.5f, (float)y + 0.5f, 0.0f, 1.0f);
      s.pass_Position = vec3(pass_position[o * 3], pass_position[o * 3 + 1], pass_position[o * 3 + 2]);
      s.gl_FragDepth = 1.0f;
      s.main();
      if (s.discarded) continue;
      out_rgba[o * 4] = s.out_Color.x; out_rgba[o * 4 + 1] = s.out_Color.y; out_rgba[o * 4 + 2] = s.out_Color.z; out_rgba[o * 4 + 3] = s.out_Color.w;
      out_depth[o] = s.gl_FragDepth;
      out_hit[o] = 1;
    }
}

// ReconIntegration::fillColors (recon_integration.cpp:280-339) with the reference's framebuffer_transfer.fs, tsdf_inpaint.fs
// and tsdf_colorfill.fs. The harness plays the host: the two ViewLod atlases (1.5 W x H, RGBA32F + depth; view_lod.cpp:24-52
// for the lod viewports), ViewLod::enable's viewport and clears (:61-81), the ping-pong of :282-312, glDepthFunc(GL_ALWAYS)
// inside and GL_LESS against a cleared depth buffer for the final pass. Same I/O convention as ro_fill_colors.
void rg_fill_colors(const float* rgba, const float* depth, int W, int H, float* out_rgba, float* atlas_rgba, float* atlas_depth) {
  const int FW = (int)((float)W * 1.5f);
  int n = 1 + (int)std::floor(std::log2((float)(W < H ? W : H)));
  if (n > 20) n = 20;
  uvec2 off[20], res[20];
  int oy = H;
  for (int i = 0; i < n; ++i) {
    res[i] = uvec2((uint)std::floor((float)W / std::pow(2.0f, (float)i)), (uint)std::floor((float)H / std::pow(2.0f, (float)i)));
    if (i > 0) { oy -= (int)res[i].y; off[i] = uvec2((uint)W, (uint)oy); }
  }
  const size_t npx = (size_t)FW * H;
  std::vector<float> col[2], dep[2];
  for (int k = 0; k < 2; ++k) { col[k].assign(npx * 4, 0.0f); dep[k].assign(npx, 1.0f); }
  auto clear = [&](int k) {                                          // glClearColor(0,1,0,0), glClearDepth(1)
    for (size_t i = 0; i < npx; ++i) { col[k][i * 4] = 0.f; col[k][i * 4 + 1] = 1.f; col[k][i * 4 + 2] = 0.f; col[k][i * 4 + 3] = 0.f; dep[k][i] = 1.0f; }
  };
  // draw(): atlas 0 cleared, raymarch fragments that were not discarded and pass GL_LESS
  clear(0);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const float* p = rgba + ((size_t)y * W + x) * 4;
      const float d = depth[(size_t)y * W + x];
      if (p[3] != 0.0f && d < 1.0f) {
        for (int c = 0; c < 4; ++c) col[0][((size_t)y * FW + x) * 4 + c] = p[c];
        dep[0][(size_t)y * FW + x] = d;
      }
    }
  auto samplers = [&](int k, sampler2D& c, sampler2D& d) {
    c.f32 = col[k].data(); c.W = FW; c.H = H; c.C = 4;
    d.f32 = dep[k].data(); d.W = FW; d.H = H; d.C = 1;
  };
  int cur = 0;                                                        // m_view_inpaint; 1 - cur = m_view_inpaint2
  auto transfer = [&]() {
    const int dst = 1 - cur;
    clear(dst);                                                       // enable(0): clears the whole attachment
    S_framebuffer_transfer proto{};
    samplers(cur, proto.texture_color, proto.texture_depth);
    proto.resolution_tex = uvec2((uint)FW, (uint)H);                  // resolution_full (:511)
    proto.lod = 0;
    for (int y = 0; y < H; ++y)                                       // viewport (0, 0, W, H)
      for (int x = 0; x < W; ++x) {
        S_framebuffer_transfer s(proto);
        s.pass_TexCoord = vec2(((float)x + 0.5f) / (float)W, ((float)y + 0.5f) / (float)H);
        s.main();
        float* o = col[dst].data() + ((size_t)y * FW + x) * 4;
        o[0] = s.out_FragColor.x; o[1] = s.out_FragColor.y; o[2] = s.out_FragColor.z; o[3] = s.out_FragColor.w;
        dep[dst][(size_t)y * FW + x] = s.gl_FragDepth;
      }
    cur = dst;                                                        // std::swap(m_view_inpaint, m_view_inpaint2)
  };
  transfer();
  for (int i = 1; i < n; ++i) {
    const int dst = 1 - cur;
    S_tsdf_inpaint proto{};
    samplers(cur, proto.texture_color, proto.texture_depth);
    proto.resolution_inv = vec2(1.0f / (float)FW, 1.0f / (float)H);
    proto.lod = i - 1;
    for (int k = 0; k < n; ++k) { proto.texture_offsets[k] = off[k]; proto.texture_resolutions[k] = res[k]; }
    proto.viewport_offset = vec2(0.0f, 0.0f);
    const int vx = (int)off[i].x, vy = (int)off[i].y, vw = (int)res[i].x, vh = (int)res[i].y;   // enable(i, false, false)
    for (int y = vy; y < vy + vh; ++y)
      for (int x = vx; x < vx + vw; ++x) {
        if (x < 0 || y < 0 || x >= FW || y >= H) continue;
        S_tsdf_inpaint s(proto);
        s.gl_FragCoord = vec4((float)x, (float)y, 0.0f, 1.0f);      // layout(pixel_center_integer)
        s.pass_TexCoord = vec2(((float)(x - vx) + 0.5f) / (float)vw, ((float)(y - vy) + 0.5f) / (float)vh);
        s.main();
        float* o = col[dst].data() + ((size_t)y * FW + x) * 4;
        o[0] = s.out_FragColor.x; o[1] = s.out_FragColor.y; o[2] = s.out_FragColor.z; o[3] = s.out_FragColor.w;
        dep[dst][(size_t)y * FW + x] = s.gl_FragDepth;                // GL_ALWAYS
      }
    cur = dst;
    transfer();
  }
  const int F = 1 - cur;                                              // m_view_inpaint2 after the last swap: the atlas with every lod
  if (atlas_rgba) std::memcpy(atlas_rgba, col[F].data(), npx * 4 * sizeof(float));
  if (atlas_depth) std::memcpy(atlas_depth, dep[F].data(), npx * sizeof(float));
  S_tsdf_colorfill proto{};
  samplers(F, proto.texture_color, proto.texture_depth);
  proto.resolution_inv = vec2(1.0f / (float)FW, 1.0f / (float)H);
  proto.num_lods = n;
  for (int k = 0; k < n; ++k) { proto.texture_offsets[k] = off[k]; proto.texture_resolutions[k] = res[k]; }
  proto.viewport_offset = vec2(0.0f, 0.0f);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      S_tsdf_colorfill s(proto);
      s.gl_FragCoord = vec4((float)x, (float)y, 0.0f, 1.0f);
      s.pass_TexCoord = vec2(((float)x + 0.5f) / (float)W, ((float)y + 0.5f) / (float)H);
      s.main();
      float* o = out_rgba + ((size_t)y * W + x) * 4;
      const float* in = rgba + ((size_t)y * W + x) * 4;
      if (s.gl_FragDepth < 1.0f) { o[0] = s.out_FragColor.x; o[1] = s.out_FragColor.y; o[2] = s.out_FragColor.z; o[3] = s.out_FragColor.w; }
      else { o[0] = in[0]; o[1] = in[1]; o[2] = in[2]; o[3] = in[3]; }   // GL_LESS against the cleared depth buffer
    }
}

// ReconIntegration::drawDepthLimits (recon_integration.cpp:409-429): UnitCube::drawInstanced (one triangle strip per occupied
// brick, the reference's vertices and strip order passed in) through bricks.vs -> bricks.gs -> a rasteriser -> bricks.fs, GL_MIN
// blending over the clear colour (1, 0, 1, 0) (:144), face culling off. The rasteriser is the OpenGL 4.4 pipeline in fp64:
// clipping against the near and far planes (section 13.5), perspective divide and viewport transform (13.6), fragments where
// the pixel centre lies inside the projected polygon (14.6.1), window z interpolated affinely, facing from the sign of the
// window-space area (14.6.1, CCW = front). out_peels [vh][vw][4].
void rg_depth_peels(const float* modelview, const float* projection, const float* bbox_min, float brick_size, const uint32_t* brick_res,
                    uint32_t* bricks, uint32_t num_bricks, uint32_t* occupied, uint32_t n_occ, const float* cube_vertices,
                    const uint8_t* strip, int strip_len, int vw, int vh, float* out_peels) {
  for (size_t i = 0; i < (size_t)vw * vh; ++i) { out_peels[i * 4] = 1.f; out_peels[i * 4 + 1] = 0.f; out_peels[i * 4 + 2] = 1.f; out_peels[i * 4 + 3] = 0.f; }
  S_bricks_vs vs{};
  vs.gl_ModelViewMatrix = mat4(modelview); vs.gl_ProjectionMatrix = mat4(projection);
  vs.bbox_min = vec3(bbox_min[0], bbox_min[1], bbox_min[2]); vs.bbox_max = vec3(0.f);
  vs.brick_size = brick_size; vs.resolution = uvec3(brick_res[0], brick_res[1], brick_res[2]);
  vs.bricks.p = bricks; vs.bricks.n = num_bricks; vs.bricks_occupied.p = occupied; vs.bricks_occupied.n = n_occ;
  S_bricks_gs gs_proto{};
  gs_proto.bbox_min = vs.bbox_min; gs_proto.bbox_max = vs.bbox_max; gs_proto.brick_size = brick_size; gs_proto.resolution = vs.resolution;
  gs_proto.bricks.p = bricks; gs_proto.bricks.n = num_bricks; gs_proto.bricks_occupied.p = occupied; gs_proto.bricks_occupied.n = n_occ;
  struct H4 { double x, y, z, w; };
  auto lerp4 = [](const H4& a, const H4& b, double t) { return H4{a.x + (b.x - a.x) * t, a.y + (b.y - a.y) * t, a.z + (b.z - a.z) * t, a.w + (b.w - a.w) * t}; };
  for (uint32_t inst = 0; inst < n_occ; ++inst) {
    vec4 pos[8]; vec3 gpos[8]; uint gid[8];
    for (int v = 0; v < 8; ++v) {
      S_bricks_vs s(vs);
      s.gl_InstanceID = (int)inst;
      s.in_Position = vec3(cube_vertices[v * 3], cube_vertices[v * 3 + 1], cube_vertices[v * 3 + 2]);
      s.main();
      pos[v] = s.gl_Position; gpos[v] = s.geo_Position; gid[v] = s.geo_Id;
    }
    for (int t = 0; t + 2 < strip_len; ++t) {
      // GL_TRIANGLE_STRIP: odd triangles swap their first two vertices so that every triangle keeps the strip's winding
      const int i0 = (t & 1) ? strip[t + 1] : strip[t], i1 = (t & 1) ? strip[t] : strip[t + 1], i2 = strip[t + 2];
      if (i0 == i1 || i1 == i2 || i0 == i2) continue;
      S_bricks_gs g(gs_proto);
      const int idx[3] = {i0, i1, i2};
      for (int k = 0; k < 3; ++k) { g.geo_Position[k] = gpos[idx[k]]; g.geo_Id[k] = gid[idx[k]]; g.gl_in[k].gl_Position = pos[idx[k]]; }
      g.main();
      if (g.emitted.size() < 3) continue;
      std::vector<H4> poly;
      for (int k = 0; k < 3; ++k) poly.push_back(H4{g.emitted[k].x, g.emitted[k].y, g.emitted[k].z, g.emitted[k].w});
      for (int plane = 0; plane < 2; ++plane) {                      // near: z >= -w, far: z <= w
        std::vector<H4> outp;
        auto dist = [&](const H4& p) { return plane == 0 ? p.z + p.w : p.w - p.z; };
        for (size_t a = 0; a < poly.size(); ++a) {
          const H4& A = poly[a]; const H4& B = poly[(a + 1) % poly.size()];
          const double da = dist(A), db = dist(B);
          if (da >= 0) outp.push_back(A);
          if ((da >= 0) != (db >= 0)) outp.push_back(lerp4(A, B, da / (da - db)));
        }
        poly.swap(outp);
        if (poly.size() < 3) break;
      }
      if (poly.size() < 3) continue;
      std::vector<double> wx(poly.size()), wy(poly.size()), wz(poly.size());
      for (size_t a = 0; a < poly.size(); ++a) {
        wx[a] = (poly[a].x / poly[a].w + 1.0) * 0.5 * vw; wy[a] = (poly[a].y / poly[a].w + 1.0) * 0.5 * vh; wz[a] = (poly[a].z / poly[a].w + 1.0) * 0.5;
      }
      double area2 = 0.0;
      for (size_t a = 0; a < poly.size(); ++a) { const size_t b = (a + 1) % poly.size(); area2 += wx[a] * wy[b] - wx[b] * wy[a]; }
      if (area2 == 0.0) continue;
      const bool front = area2 > 0.0;
      for (size_t f = 1; f + 1 < poly.size(); ++f) {                 // fan of the clipped polygon
        const double x0 = wx[0], y0 = wy[0], x1 = wx[f], y1 = wy[f], x2 = wx[f + 1], y2 = wy[f + 1];
        const double den = (x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0);
        if (den == 0.0) continue;
        const int px0 = std::max(0, (int)std::floor(std::min({x0, x1, x2}) - 0.5)), px1 = std::min(vw - 1, (int)std::ceil(std::max({x0, x1, x2}) - 0.5));
        const int py0 = std::max(0, (int)std::floor(std::min({y0, y1, y2}) - 0.5)), py1 = std::min(vh - 1, (int)std::ceil(std::max({y0, y1, y2}) - 0.5));
        for (int py = py0; py <= py1; ++py)
          for (int px = px0; px <= px1; ++px) {
            const double cx = px + 0.5, cy = py + 0.5;
            const double b1 = ((cx - x0) * (y2 - y0) - (x2 - x0) * (cy - y0)) / den;
            const double b2 = ((x1 - x0) * (cy - y0) - (cx - x0) * (y1 - y0)) / den;
            const double b0 = 1.0 - b1 - b2;
            if (b0 < 0.0 || b1 < 0.0 || b2 < 0.0) continue;
            S_bricks_fs fs{};
            fs.gl_FragCoord = vec4((float)cx, (float)cy, (float)(b0 * wz[0] + b1 * wz[f] + b2 * wz[f + 1]), 1.0f);
            fs.gl_FrontFacing = front;
            fs.main();
            float* o = out_peels + ((size_t)py * vw + px) * 4;       // glBlendEquation(GL_MIN)
            o[0] = std::min(o[0], fs.out_Color.x); o[1] = std::min(o[1], fs.out_Color.y);
            o[2] = std::min(o[2], fs.out_Color.z); o[3] = std::min(o[3], fs.out_Color.w);
          }
      }
    }
  }
}


// ---- the point-drawing reconstructions -------------------------------------------------------------------
// One point through the fixed-function stages between the last vertex-processing stage and the fragment shader, as OpenGL 4.4
// prescribes them: clip-volume cull of the point's centre (13.5), perspective divide and viewport transform with depth range
// [0, 1] (13.6), a fragment for every pixel whose centre lies in the half-open square of side max(size, 1) centred at the
// window position (14.4.1, point sprites; the shaders run with GL_PROGRAM_POINT_SIZE, kinect_client.cpp:260-261). fp64 for
// the fixed-function arithmetic. `frag(px, py, zw, inv_w, s, t)` runs the fragment shader and returns false on discard;
// the depth test is GL_LESS in draw order against out_depth.
}  // extern "C" (a template cannot have C linkage)
template <typename Frag>
static void raster_point_gl(const vec4& clip, double size, int vw, int vh, float* out_rgba, float* out_depth, Frag frag) {
  if (!(clip.w > 0.0f) || !(std::fabs(clip.x) <= clip.w) || !(std::fabs(clip.y) <= clip.w) || !(std::fabs(clip.z) <= clip.w)) return;
  const double xw = ((double)clip.x / clip.w * 0.5 + 0.5) * vw, yw = ((double)clip.y / clip.w * 0.5 + 0.5) * vh;
  const float zw = (float)((double)clip.z / clip.w * 0.5 + 0.5);
  if (!(size >= 1.0)) size = 1.0;
  if (size > 256.0) size = 256.0;                                   // the implementation's point-size range, taken as [1, 256]
  const double h = size * 0.5;
  const int x0 = std::max(0, (int)std::ceil(xw - h - 0.5)), x1 = std::min(vw, (int)std::ceil(xw + h - 0.5));
  const int y0 = std::max(0, (int)std::ceil(yw - h - 0.5)), y1 = std::min(vh, (int)std::ceil(yw + h - 0.5));
  for (int y = y0; y < y1; ++y)
    for (int x = x0; x < x1; ++x) {
      const size_t o = (size_t)y * vw + x;
      // point sprite coordinates (14.4.1, upper-left origin): s = 1/2 + (xf + 1/2 - xw) / size, t = 1/2 - (yf + 1/2 - yw) / size
      const float ps = (float)(0.5 + ((double)x + 0.5 - xw) / size), pt = (float)(0.5 - ((double)y + 0.5 - yw) / size);
      vec4 color;
      if (!frag(x, y, zw, 1.0f / clip.w, ps, pt, color)) continue;
      if (!(zw < out_depth[o])) continue;                           // GL_LESS
      out_depth[o] = zw;
      out_rgba[o * 4] = color.x; out_rgba[o * 4 + 1] = color.y; out_rgba[o * 4 + 2] = color.z; out_rgba[o * 4 + 3] = color.w;
    }
}

extern "C" {
// ReconPoints::draw (recon_points.cpp:46-52,71-111) with the reference's points.vs -> points.gs -> rasteriser -> points.fs.
// uniforms16: gl_ModelViewMatrix, gl_ProjectionMatrix, gl_NormalMatrix, img_to_eye_curr, projection_inv, modelview_inv.
void rg_draw_points(int N, int W, int H, const float* depth_b, const float* normals, const uint8_t* color, int CW, int CH,
                    const float* cv_xyz, const float* cv_uv, const int32_t* cv_res, const float* bbox_min, const float* bbox_max,
                    const float* uniforms16, int vw, int vh, int shade_mode, float* out_rgba, float* out_depth) {
  for (size_t i = 0; i < (size_t)vw * vh; ++i) { out_rgba[i * 4] = out_rgba[i * 4 + 1] = out_rgba[i * 4 + 2] = out_rgba[i * 4 + 3] = 0.f; out_depth[i] = 1.f; }
  if (N > 5) return;
  sampler2DArray col;
  col.u8 = color; col.W = CW; col.H = CH; col.L = N; col.C = 3; col.linear = true;
  sampler2DArray d = tex2d(depth_b, W, H, 2, false), nr = tex2d(normals, W, H, 3, true);
  d.L = nr.L = N;
  const size_t cv_vox = (size_t)cv_res[0] * cv_res[1] * cv_res[2];
  S_points_vs vs{};
  S_points_gs gs{};
  S_points_fs fs{};
  vs.kinect_depths = d; vs.kinect_normals = nr;
  for (int i = 0; i < N; ++i) {
    vs.cv_xyz[i] = tex3d(cv_xyz + (size_t)i * cv_vox * 3, cv_res[0], cv_res[1], cv_res[2], 3);
    vs.cv_uv[i] = tex3d(cv_uv + (size_t)i * cv_vox * 2, cv_res[0], cv_res[1], cv_res[2], 2);
  }
  vs.gl_ModelViewMatrix = mat4(uniforms16); vs.gl_ProjectionMatrix = mat4(uniforms16 + 16);
  gs.g_shade_mode = shade_mode;
  gs.bbox_min = vec3(bbox_min[0], bbox_min[1], bbox_min[2]); gs.bbox_max = vec3(bbox_max[0], bbox_max[1], bbox_max[2]);
  fs.kinect_colors = col; fs.kinect_normals = nr;
  fs.gl_ModelViewMatrix = mat4(uniforms16); fs.gl_ProjectionMatrix = mat4(uniforms16 + 16); fs.gl_NormalMatrix = mat4(uniforms16 + 32);
  fs.img_to_eye_curr = mat4(uniforms16 + 48); fs.projection_inv = mat4(uniforms16 + 64); fs.modelview_inv = mat4(uniforms16 + 80);
  fs.viewportSizeInv = vec2(1.0f / (float)vw, 1.0f / (float)vh);
  fs.epsilon = 0.075f;                                               // recon_points.cpp:40
  fs.g_shade_mode = shade_mode;
  const float stepX = 1.0f / (float)W, stepY = 1.0f / (float)H;
  for (int layer = 0; layer < N; ++layer)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        S_points_vs v(vs);
        v.layer = (uint)layer;
        v.in_position = vec2((float)(((double)x + 0.5) * (double)stepX), (float)(((double)y + 0.5) * (double)stepY));
        v.main();
        S_points_gs g(gs);
        g.geo_texcoord[0] = v.geo_texcoord; g.geo_pos_norm[0] = v.geo_pos_norm; g.geo_pos_es[0] = v.geo_pos_es; g.geo_pos_cs[0] = v.geo_pos_cs;
        g.geo_depth[0] = v.geo_depth; g.geo_quality[0] = 0.0f;         // points.vs never writes geo_quality; points.fs never uses it
        g.gl_in[0].gl_Position = v.gl_Position;
        g.main();
        if (!g.emitted) continue;
        raster_point_gl(g.gl_Position, (double)g.gl_PointSize, vw, vh, out_rgba, out_depth,
                        [&](int px, int py, float zw, float inv_w, float ps, float pt, vec4& out) {
                          S_points_fs f(fs);
                          f.layer = (uint)layer;
                          f.pass_texcoord = g.pass_texcoord; f.pass_pos_norm = g.pass_pos_norm; f.pass_pos_es = g.pass_pos_es; f.pass_pos_cs = g.pass_pos_cs;
                          f.pass_depth = g.pass_depth; f.pass_quality = g.pass_quality; f.pass_glpos = g.pass_glpos;
                          f.gl_FragCoord = vec4((float)px + 0.5f, (float)py + 0.5f, zw, inv_w);
                          f.gl_PointCoord = vec2(ps, pt);
                          f.main();
                          out = f.gl_FragColor;
                          return !f.discarded;
                        });
      }
}

// ReconCalibs::draw (recon_calibs.cpp:39-46,56-66) over VolumeSampler's voxel centres (volume_sampler.cpp:14-23) with the
// reference's calib_vis.vs -> rasteriser -> calib_vis.fs. The vertex shader does not write gl_PointSize: size 1.
void rg_draw_calibs(const float* tsdf, const uint32_t* res, int N, const float* inv, const int32_t* inv_res, const float* cv_xyz, const int32_t* cv_res,
                    int layer, float limit, const float* bbox_min, const float* bbox_max, const float* modelview, const float* projection,
                    int vw, int vh, float* out_rgba, float* out_depth) {
  for (size_t i = 0; i < (size_t)vw * vh; ++i) { out_rgba[i * 4] = out_rgba[i * 4 + 1] = out_rgba[i * 4 + 2] = out_rgba[i * 4 + 3] = 0.f; out_depth[i] = 1.f; }
  if (N > 5) return;
  const size_t inv_vox = (size_t)inv_res[0] * inv_res[1] * inv_res[2], cv_vox = (size_t)cv_res[0] * cv_res[1] * cv_res[2];
  S_calib_vis_vs vs{};
  S_calib_vis_fs fs{};
  for (int i = 0; i < N; ++i) {
    vs.cv_xyz_inv[i] = tex3d(inv + (size_t)i * inv_vox * 4, inv_res[0], inv_res[1], inv_res[2], 4);
    fs.cv_xyz_inv[i] = vs.cv_xyz_inv[i];
    vs.cv_xyz[i] = tex3d(cv_xyz + (size_t)i * cv_vox * 3, cv_res[0], cv_res[1], cv_res[2], 3);
  }
  vs.volume_tsdf = tex3d(tsdf, (int)res[0], (int)res[1], (int)res[2], 1);
  vs.layer = fs.layer = (uint)layer;
  vs.limit = fs.limit = limit;
  vs.gl_ModelViewMatrix = mat4(modelview); vs.gl_ProjectionMatrix = mat4(projection);
  float v2w[16] = {0};
  v2w[0] = bbox_max[0] - bbox_min[0]; v2w[5] = bbox_max[1] - bbox_min[1]; v2w[10] = bbox_max[2] - bbox_min[2];
  v2w[12] = bbox_min[0]; v2w[13] = bbox_min[1]; v2w[14] = bbox_min[2]; v2w[15] = 1.0f;
  vs.vol_to_world = mat4(v2w);
  const float stepX = 1.0f / (float)inv_res[0], stepY = 1.0f / (float)inv_res[1], stepZ = 1.0f / (float)inv_res[2];
  for (int z = 0; z < inv_res[2]; ++z)
    for (int y = 0; y < inv_res[1]; ++y)
      for (int x = 0; x < inv_res[0]; ++x) {
        S_calib_vis_vs v(vs);
        v.in_Position = vec3(((float)x + 0.5f) * stepX, ((float)y + 0.5f) * stepY, ((float)z + 0.5f) * stepZ);
        v.main();
        raster_point_gl(v.gl_Position, 1.0, vw, vh, out_rgba, out_depth, [&](int, int, float, float, float, float, vec4& out) {
          S_calib_vis_fs f(fs);
          f.geo_pos_view = v.geo_pos_view; f.geo_pos_world = v.geo_pos_world; f.geo_pos_volume = v.geo_pos_volume; f.geo_distance = v.geo_distance;
          f.main();
          out = f.gl_FragColor;
          return !f.discarded;
        });
      }
}
// ReconTrigrid::draw (recon_trigrid.cpp:48-61 grid, :82-149 passes) with the reference's trigrid_accum.vs -> trigrid_accum.gs ->
// fixed-function stages (oracle/ro_raster.h, the OpenGL 4.4 pipeline in fp64) -> trigrid_accum.fs, then trigrid_normalize.fs.
// The vertex shader runs once per grid vertex (GL runs it per triangle corner on identical inputs). Pass 1: depth only, GL_LESS
// against 1; pass 2: no depth test, glBlendFunc(ONE, ONE) into a cleared RGBA32F target, in draw order; pass 3: a screen quad.
// uniforms16: gl_ModelViewMatrix, gl_ProjectionMatrix, gl_NormalMatrix, img_to_eye_curr.
void rg_draw_trigrid(int N, int W, int H, const float* depth_b, const float* quality, const uint8_t* color, int CW, int CH,
                     const float* cv_xyz, const float* cv_uv, const int32_t* cv_res, const float* bbox_min, const float* bbox_max,
                     const float* uniforms16, int vw, int vh, int shade_mode, float min_length, float* out_rgba, float* out_depth) {
  const size_t npx = (size_t)vw * vh;
  for (size_t i = 0; i < npx; ++i) { out_rgba[i * 4] = out_rgba[i * 4 + 1] = out_rgba[i * 4 + 2] = out_rgba[i * 4 + 3] = 0.f; out_depth[i] = 1.f; }
  if (N > 5) return;
  sampler2DArray col;
  col.u8 = color; col.W = CW; col.H = CH; col.L = N; col.C = 3; col.linear = true;
  sampler2DArray d = tex2d(depth_b, W, H, 2, false), q = tex2d(quality, W, H, 1, true);
  d.L = q.L = N;
  const size_t cv_vox = (size_t)cv_res[0] * cv_res[1] * cv_res[2];
  S_trigrid_vs vs{};
  vs.kinect_depths = d; vs.kinect_qualities = q;
  for (int i = 0; i < N; ++i) {
    vs.cv_xyz[i] = tex3d(cv_xyz + (size_t)i * cv_vox * 3, cv_res[0], cv_res[1], cv_res[2], 3);
    vs.cv_uv[i] = tex3d(cv_uv + (size_t)i * cv_vox * 2, cv_res[0], cv_res[1], cv_res[2], 2);
  }
  vs.gl_ModelViewMatrix = mat4(uniforms16); vs.gl_ProjectionMatrix = mat4(uniforms16 + 16);
  S_trigrid_gs gs{};
  gs.min_length = min_length;
  std::vector<float> depth1(npx, 1.0f), accum(npx * 4, 0.0f);
  S_trigrid_fs fs{};
  fs.kinect_colors = col;
  fs.depth_map_curr = tex2d(depth1.data(), vw, vh, 1, false);         // ViewArray's depth array is NEAREST (ViewArray.cpp:22)
  fs.gl_NormalMatrix = mat4(uniforms16 + 32); fs.img_to_eye_curr = mat4(uniforms16 + 48);
  fs.viewportSizeInv = vec2(1.0f / (float)vw, 1.0f / (float)vh);
  fs.epsilon = 0.075f;                                               // recon_trigrid.cpp:35
  fs.g_shade_mode = shade_mode;
  fs.bbox_min = vec3(bbox_min[0], bbox_min[1], bbox_min[2]); fs.bbox_max = vec3(bbox_max[0], bbox_max[1], bbox_max[2]);

  const int GW = H + 1, GH = W + 1;                                   // cells x < tex_height, y < tex_width, as :51-52 loop
  const float stepX = 1.0f / (float)W, stepY = 1.0f / (float)H;
  std::vector<S_trigrid_vs> verts;
  verts.reserve((size_t)N * GW * GH);
  for (int layer = 0; layer < N; ++layer)
    for (int j = 0; j < GH; ++j)
      for (int i = 0; i < GW; ++i) {
        S_trigrid_vs v(vs);
        v.layer = (uint)layer;
        v.in_Position = vec2((float)(((double)i + 0.5) * (double)stepX), (float)(((double)j + 0.5) * (double)stepY));
        v.main();
        verts.push_back(v);
      }
  for (int stage = 0; stage < 2; ++stage)
    for (int layer = 0; layer < N; ++layer)
      for (int y = 0; y < W; ++y)
        for (int x = 0; x < H; ++x)
          for (int k = 0; k < 2; ++k) {
            const S_trigrid_vs* g0 = verts.data() + (size_t)layer * GW * GH;
            const S_trigrid_vs* tv[3] = {k == 0 ? g0 + (size_t)y * GW + x : g0 + (size_t)y * GW + x + 1,
                                         k == 0 ? g0 + (size_t)y * GW + x + 1 : g0 + (size_t)(y + 1) * GW + x + 1, g0 + (size_t)(y + 1) * GW + x};
            S_trigrid_gs g(gs);
            for (int c = 0; c < 3; ++c) {
              g.geo_texcoord[c] = tv[c]->geo_texcoord; g.geo_pos_es[c] = tv[c]->geo_pos_es; g.geo_pos_cs[c] = tv[c]->geo_pos_cs;
              g.geo_depth[c] = tv[c]->geo_depth; g.geo_quality[c] = tv[c]->geo_quality; g.gl_in[c].gl_Position = tv[c]->gl_Position;
            }
            g.main();
            if (g.emitted.size() < 3) continue;
            const S_trigrid_gs::Emitted* e = g.emitted.data();
            const float clip[3][4] = {{e[0].pos.x, e[0].pos.y, e[0].pos.z, e[0].pos.w}, {e[1].pos.x, e[1].pos.y, e[1].pos.z, e[1].pos.w},
                                      {e[2].pos.x, e[2].pos.y, e[2].pos.z, e[2].pos.w}};
            ro::raster_triangle(clip, vw, vh, [&](int fx, int fy, float zw, const double* B) {
              S_trigrid_fs f(fs);
              f.stage = (uint)stage; f.layer = (uint)layer;
              f.gl_FragCoord = vec4((float)fx + 0.5f, (float)fy + 0.5f, zw, 1.0f);
              f.pass_texcoord = vec2(ro::rinterp(B, e[0].texcoord.x, e[1].texcoord.x, e[2].texcoord.x), ro::rinterp(B, e[0].texcoord.y, e[1].texcoord.y, e[2].texcoord.y));
              f.pass_pos_es = vec3(ro::rinterp(B, e[0].pos_es.x, e[1].pos_es.x, e[2].pos_es.x), ro::rinterp(B, e[0].pos_es.y, e[1].pos_es.y, e[2].pos_es.y),
                                   ro::rinterp(B, e[0].pos_es.z, e[1].pos_es.z, e[2].pos_es.z));
              f.pass_pos_cs = vec3(ro::rinterp(B, e[0].pos_cs.x, e[1].pos_cs.x, e[2].pos_cs.x), ro::rinterp(B, e[0].pos_cs.y, e[1].pos_cs.y, e[2].pos_cs.y),
                                   ro::rinterp(B, e[0].pos_cs.z, e[1].pos_cs.z, e[2].pos_cs.z));
              f.pass_normal_es = vec3(ro::rinterp(B, e[0].normal_es.x, e[1].normal_es.x, e[2].normal_es.x), ro::rinterp(B, e[0].normal_es.y, e[1].normal_es.y, e[2].normal_es.y),
                                      ro::rinterp(B, e[0].normal_es.z, e[1].normal_es.z, e[2].normal_es.z));
              f.pass_depth = ro::rinterp(B, e[0].depth, e[1].depth, e[2].depth);
              f.pass_quality = ro::rinterp(B, e[0].quality, e[1].quality, e[2].quality);
              f.main();
              if (f.discarded) return;
              const size_t o = (size_t)fy * vw + fx;
              if (stage == 0) {
                if (zw < depth1[o]) depth1[o] = zw;                      // GL_LESS, depth writes on
              } else {
                float* a = accum.data() + o * 4;                         // GL_FUNC_ADD, ONE, ONE
                a[0] += f.gl_FragColor.x; a[1] += f.gl_FragColor.y; a[2] += f.gl_FragColor.z; a[3] += f.gl_FragColor.w;
              }
            });
          }
  S_trigrid_norm_fs nf{};
  nf.color_map = tex2d(accum.data(), vw, vh, 4, true);
  nf.depth_map = tex2d(depth1.data(), vw, vh, 1, false);
  nf.texSizeInv = vec2(1.0f / (float)vw, 1.0f / (float)vh);
  nf.offset = vec2(0.0f, 0.0f);
  for (int y = 0; y < vh; ++y)
    for (int x = 0; x < vw; ++x) {
      S_trigrid_norm_fs f(nf);
      f.gl_FragCoord = vec4((float)x + 0.5f, (float)y + 0.5f, 0.0f, 1.0f);
      f.main();
      if (f.discarded) continue;
      const size_t o = (size_t)y * vw + x;
      out_rgba[o * 4] = f.gl_FragColor.x; out_rgba[o * 4 + 1] = f.gl_FragColor.y; out_rgba[o * 4 + 2] = f.gl_FragColor.z; out_rgba[o * 4 + 3] = f.gl_FragColor.w;
      out_depth[o] = f.gl_FragDepth;
    }
}
}  // extern "C"

// ORACLE (test infrastructure, NOT product code). The reference ships no tests; this file is pinned against the reference's
// own shaders compiled as C++ and run on the CPU (oracle/glsl_host/, oracle/_ref/libref_glsl.so, tests/golden/ref_glsl_stages.npz).
// Scalar restatement of the five depth pre-processing passes that NetKinectArray::processTextures drives
// (framework/NetKinectArray.cpp:251-290, 311-428). One call = one pass over one sensor layer.
// Arithmetic rules: see ro_math.h. Image layout: row-major [y][x][channels], texel (0,0) first.
#include "ro_math.h"
#include "rr_oracle.h"

#include <vector>

using namespace ro;

namespace {

struct FwdCalib {           // one sensor's forward calibration volumes (CalibVolumes.cpp:132-144)
  const float* xyz;         // [Z][Y][X][3] world position of (u, v, depth)
  const float* uv;          // [Z][Y][X][2] colour-image coordinate of (u, v, depth)
  int X, Y, Z;
};

inline V3 fetch_xyz(const FwdCalib& c, float s, float t, float r) {
  float o[3];
  tex3d_linear<3>(c.xyz, c.X, c.Y, c.Z, s, t, r, o, 3);
  return {o[0], o[1], o[2]};
}
inline V2 fetch_uv(const FwdCalib& c, float s, float t, float r) {
  float o[2];
  tex3d_linear<2>(c.uv, c.X, c.Y, c.Z, s, t, r, o, 2);
  return {o[0], o[1]};
}

// glsl/inc_bbox_test.glsl:11-21
inline bool in_bbox(V3 p, const float* bmin, const float* bmax) {
  return p.x >= bmin[0] && p.y >= bmin[1] && p.z >= bmin[2] && p.x <= bmax[0] && p.y <= bmax[1] && p.z <= bmax[2];
}

// glsl/inc_color.glsl:8-46 (note the extra /255 although the sampler already returns [0,1]: as written)
inline float pivot_rgb(float n) {
  return ((n > 0.04045f) ? gl_pow((n + 0.055f) / 1.055f, 2.4f) : n / 12.92f) * 100.0f;
}
inline float pivot_xyz(float n) {
  return (n > 0.008856f) ? gl_pow(n, 1.0f / 3.0f) : (903.3f * n + 16.0f) / 116.0f;
}
inline V3 rgb_to_lab(V3 rgb) {
  float r = pivot_rgb(rgb.x / 255.0f);
  float g = pivot_rgb(rgb.y / 255.0f);
  float b = pivot_rgb(rgb.z / 255.0f);
  float X = (r * 0.4124f + g * 0.3576f) + b * 0.1805f;
  float Y = (r * 0.2126f + g * 0.7152f) + b * 0.0722f;
  float Z = (r * 0.0193f + g * 0.1192f) + b * 0.9505f;
  float x = pivot_xyz(X / 95.047f);
  float y = pivot_xyz(Y / 100.000f);
  float z = pivot_xyz(Z / 108.883f);
  return {gl_max(0.0f, 116.0f * y - 16.0f), 500.0f * (x - y), 200.0f * (y - z)};
}

// bilinear RGB8 fetch, normalised fixed point -> float c/255 (GL spec §8.5 / 2.3.5), LINEAR + CLAMP_TO_EDGE
inline V3 fetch_rgb8(const uint8_t* img, int W, int H, float s, float t) {
  int x0, x1, y0, y1; float a, b;
  lin_coord(s, W, x0, x1, a);
  lin_coord(t, H, y0, y1, b);
  float o[3];
  for (int c = 0; c < 3; ++c) {
    float v00 = (float)img[((size_t)y0 * W + x0) * 3 + c] / 255.0f;
    float v10 = (float)img[((size_t)y0 * W + x1) * 3 + c] / 255.0f;
    float v01 = (float)img[((size_t)y1 * W + x0) * 3 + c] / 255.0f;
    float v11 = (float)img[((size_t)y1 * W + x1) * 3 + c] / 255.0f;
    o[c] = lerpf(lerpf(v00, v10, a), lerpf(v01, v11, a), b);
  }
  return {o[0], o[1], o[2]};
}

}  // namespace

extern "C" {

void ro_kat_rgb_to_lab(const float* rgb, float* lab) {
  V3 l = rgb_to_lab(V3{rgb[0], rgb[1], rgb[2]});
  lab[0] = l.x; lab[1] = l.y; lab[2] = l.z;
}

// glsl/pre_morph.fs:73-112 (dilate, kernel 1) + main :114-140 mode 0. Mode 1 is a plain copy (:130-134).
// in_bbox(texcoord, depth) there returns true unconditionally (:44-50) and is therefore dropped.
void ro_pre_morph(const float* depth_in, int W, int H, float* depth_out) {
  const float min_depth = 0.5f, max_depth = 4.5f, max_dist = 0.2f;
  auto valid = [&](float d) { return d > min_depth && d < max_depth; };
#pragma omp parallel for schedule(static)
  for (int py = 0; py < H; ++py) {
    for (int px = 0; px < W; ++px) {
      float depth = depth_in[(size_t)py * W + px];
      float result;
      if (valid(depth)) {
        result = depth;
      } else {
        float average_depth = 0.0f, num_samples = 0.0f;
        bool any = false;
        for (int y = -1; y < 2; ++y)
          for (int x = -1; x < 2; ++x) {
            float ds = depth_in[(size_t)clampi(py + y, 0, H - 1) * W + clampi(px + x, 0, W - 1)];
            if (valid(ds)) { any = true; average_depth += ds; num_samples += 1.0f; }
          }
        if (!any) {
          result = 0.0f;
        } else {
          average_depth /= num_samples;
          float new_depth = 0.0f;
          num_samples = 0.0f;
          any = false;
          for (int y = -1; y < 2; ++y)
            for (int x = -1; x < 2; ++x) {
              float ds = depth_in[(size_t)clampi(py + y, 0, H - 1) * W + clampi(px + x, 0, W - 1)];
              if (valid(ds) && fabsf(average_depth - ds) < max_dist) { any = true; new_depth += ds; num_samples += 1.0f; }
            }
          result = any ? new_depth / num_samples : 0.0f;
        }
      }
      depth_out[(size_t)py * W + px] = result;
    }
  }
}

// glsl/pre_depth.fs:129-154 main, :85-127 bilateral_filter, :51-61 uncompress, inc_color.glsl.
// depth_in: metres (or, with compress != 0, the normalised 8-bit value byte/255) — already the morph output when
// NetKinectArray::m_use_processed_depth is set (NetKinectArray.cpp:287-289).
void ro_pre_depth(const float* depth_in, int W, int H,
                  const float* cv_xyz, const float* cv_uv, int CX, int CY, int CZ,
                  const uint8_t* color, int CW, int CH,
                  const float* bbox_min, const float* bbox_max,
                  float cv_min_ds, float cv_max_ds, int filter_textures,
                  int compress, float scale, float near_, float scaled_near,
                  float* out_depth /*[H][W][2]*/, float* out_lab /*[H][W][3]*/) {
  FwdCalib cal{cv_xyz, cv_uv, CX, CY, CZ};
  const int ks = 6;
  auto sample = [&](int x, int y) -> float {
    float d = depth_in[(size_t)clampi(y, 0, H - 1) * W + clampi(x, 0, W - 1)];
    if (compress) {
      if (d < scaled_near) return 0.0f;
      return (d * d + 0.15f * scaled_near) * scale + near_;
    }
    return d;
  };
  auto normalize_depth = [&](float d) { return (d - cv_min_ds) / (cv_max_ds - cv_min_ds); };
  auto is_outside = [&](float d) { return (d < cv_min_ds) || (d > cv_max_ds); };
  const float dist_space_max_inv = 1.0f / 6.0f;
  const float len06 = sqrtf(36.0f);  // length(vec2(0,6))
#pragma omp parallel for schedule(dynamic, 4)
  for (int py = 0; py < H; ++py) {
    for (int px = 0; px < W; ++px) {
      const float tcx = ((float)px + 0.5f) / (float)W;
      const float tcy = ((float)py + 0.5f) / (float)H;
      float depth = sample(px, py);
      float depth_norm = normalize_depth(depth);
      V3 pos_world = fetch_xyz(cal, tcx, tcy, depth_norm);
      bool is_in_box = in_bbox(pos_world, bbox_min, bbox_max);
      float zc = (depth_norm <= 0.0f || depth_norm >= 1.0f) ? 1.0f : depth_norm;
      V2 cc = fetch_uv(cal, tcx, tcy, zc);
      V3 lab = rgb_to_lab(fetch_rgb8(color, CW, CH, cc.x, cc.y));
      float* ol = out_lab + ((size_t)py * W + px) * 3;
      ol[0] = lab.x; ol[1] = lab.y; ol[2] = lab.z;
      float* od = out_depth + ((size_t)py * W + px) * 2;
      if (!is_in_box) { od[0] = 0.0f; od[1] = 0.0f; continue; }
      if (!filter_textures) { od[0] = depth_norm; od[1] = 1.0f; continue; }
      // bilateral_filter(vec3(pass_TexCoord, depth))
      const float max_depth = 4.5f;
      float d_dmax = depth / max_depth;
      float dist_range_max = 0.35f * d_dmax;
      float dist_range_max_inv = 1.0f / dist_range_max;
      float depth_bf = 0.0f, w = 0.0f, w_range = 0.0f, num_samples = 0.0f;
      for (int y = -ks; y <= ks; ++y) {
        for (int x = -ks; x <= ks; ++x) {
          num_samples += 1.0f;
          float depth_s = sample(px + x, py + y);
          float depth_range = fabsf(depth_s - depth);
          if (is_outside(depth_s) || (depth_range > dist_range_max)) continue;
          float gauss_space = 1.0f - sqrtf((float)(x * x + y * y)) * dist_space_max_inv;
          float gauss_range = 1.0f - gl_min(depth_range, dist_range_max) * dist_range_max_inv;
          float w_s = gauss_space * gauss_range;
          depth_bf = fmaf(w_s, depth_s, depth_bf);
          w += w_s;
          w_range += gauss_range;
        }
      }
      (void)len06;
      float filtered_depth = depth_bf / w;
      od[0] = normalize_depth(filtered_depth);
      od[1] = w_range / num_samples;
    }
  }
}

// glsl/pre_boundary.fs:86-118 main, :37-55 get_color_diff. depth: RG32F NEAREST, lab: RGB32F LINEAR sampled at
// pixel centres + integer pixel offsets => exact texel fetches with edge clamp.
void ro_pre_boundary(const float* depth_rg, const float* lab, int W, int H, int refine,
                     float* out_depth_b /*[H][W][2]*/, float* out_sil /*[H][W]*/) {
  const float max_color_dist = 0.5f, min_range = 0.65f;
  const int ks = 2;
  const float total_samples = 16.0f;  // (kernel_size*2)^2
#pragma omp parallel for schedule(static)
  for (int py = 0; py < H; ++py) {
    for (int px = 0; px < W; ++px) {
      const size_t idx = (size_t)py * W + px;
      float dx = depth_rg[idx * 2 + 0], dy = depth_rg[idx * 2 + 1];
      float sil = 1.0f;
      if (dx <= 0.0f) {
        dy = 0.0f;
        sil = 0.0f;
      } else if (!(dy > min_range)) {
        sil = 0.0f;
        // get_color_diff
        V3 color{lab[idx * 3], lab[idx * 3 + 1], lab[idx * 3 + 2]};
        float total_dist = 0.0f, num_samples = 0.0f;
        for (int y = -ks; y <= ks; ++y)
          for (int x = -ks; x <= ks; ++x) {
            size_t si = (size_t)clampi(py + y, 0, H - 1) * W + clampi(px + x, 0, W - 1);
            float sx = depth_rg[si * 2], sy = depth_rg[si * 2 + 1];
            if (sx > 0.0f && sy > min_range) {
              num_samples += 1.0f;
              V3 cs{lab[si * 3], lab[si * 3 + 1], lab[si * 3 + 2]};
              total_dist += length3(color - cs);
            }
          }
        float color_dist = (num_samples < total_samples * 0.5f) ? 1.0f : total_dist / num_samples;
        if (color_dist > max_color_dist || !refine) {
          dx = -1.0f; dy = 0.1f; sil = 0.0f;
        } else {
          dy = 1.0f;
        }
      } else {
        dy = 0.0f;
      }
      out_depth_b[idx * 2] = dx; out_depth_b[idx * 2 + 1] = dy;
      out_sil[idx] = sil;
    }
  }
}

// glsl/pre_normal.fs:26-56 calculate_normal + glsl/inc_bricks.glsl:22-58 (to_world, get_id, mark_brick).
// bricks: uint32 counters [num_bricks] (the SSBO payload after the 8-uint header), accumulated atomically.
// Pinned undefined behaviour: uvec3(floor(v)) saturates (negative/NaN -> 0); a brick id >= num_bricks is dropped
// (robust buffer access).
void ro_pre_normal(const float* depth_b, int W, int H,
                   const float* cv_xyz, int CX, int CY, int CZ,
                   const float* bbox_min, float brick_size, const uint32_t* brick_res, uint32_t num_bricks,
                   uint32_t* bricks, float* out_normal /*[H][W][3]*/) {
  FwdCalib cal{cv_xyz, nullptr, CX, CY, CZ};
  const float tsx = 1.0f / (float)W, tsy = 1.0f / (float)H;
  auto is_outside = [](float d) { return (d <= 0.0f) || (d >= 1.0f); };
  auto dep = [&](int x, int y) { return depth_b[((size_t)clampi(y, 0, H - 1) * W + clampi(x, 0, W - 1)) * 2]; };
  const V3 bmin{bbox_min[0], bbox_min[1], bbox_min[2]};
  auto get_id = [&](uint32_t ix, uint32_t iy, uint32_t iz) -> uint32_t {
    return iz * brick_res[1] * brick_res[0] + iy * brick_res[0] + ix;
  };
#pragma omp parallel for schedule(static)
  for (int py = 0; py < H; ++py) {
    for (int px = 0; px < W; ++px) {
      float* on = out_normal + ((size_t)py * W + px) * 3;
      const float tcx = ((float)px + 0.5f) / (float)W;
      const float tcy = ((float)py + 0.5f) / (float)H;
      float depth = dep(px, py);
      if (is_outside(depth)) { on[0] = on[1] = on[2] = 0.0f; continue; }
      V3 world = fetch_xyz(cal, tcx, tcy, depth);
      {  // mark_brick(world)
        V3 rel = (world - bmin) / brick_size;
        uint32_t ix = f2u_sat(floorf(rel.x)), iy = f2u_sat(floorf(rel.y)), iz = f2u_sat(floorf(rel.z));
        // to_world(vec3(0.5), index) = vec3(index) * brick_size + bbox_min + position * brick_size
        V3 fidx{(float)ix, (float)iy, (float)iz};
        V3 half{0.5f * brick_size, 0.5f * brick_size, 0.5f * brick_size};
        V3 brick_center = (fidx * brick_size + bmin) + half;
        V3 diff = world - brick_center;
        V3 d_abs{fabsf(diff.x), fabsf(diff.y), fabsf(diff.z)};
        float min_v = gl_max(d_abs.x, gl_max(d_abs.y, d_abs.z));
        float cx = (d_abs.x < min_v) ? 0.0f : 1.0f;
        float cy = (d_abs.y < min_v) ? 0.0f : 1.0f;
        float cz = (d_abs.z < min_v) ? 0.0f : 1.0f;
        int ox = (int)gl_sign(diff.x * cx), oy = (int)gl_sign(diff.y * cy), oz = (int)gl_sign(diff.z * cz);
        // ivec3(index): uint -> int reinterpretation (two's complement)
        int nx = clampi((int)ix + ox, 0, (int)(brick_res[0] - 1u));
        int ny = clampi((int)iy + oy, 0, (int)(brick_res[1] - 1u));
        int nz = clampi((int)iz + oz, 0, (int)(brick_res[2] - 1u));
        uint32_t nid = get_id((uint32_t)nx, (uint32_t)ny, (uint32_t)nz);
        uint32_t inc = (d_abs.x > brick_size * 0.1f) ? 1u : 0u;
        if (nid < num_bricks && inc) {
#pragma omp atomic
          bricks[nid] += inc;
        }
        uint32_t id = get_id(ix, iy, iz);
        if (id < num_bricks) {
#pragma omp atomic
          bricks[id] += 1u;
        }
      }
      float tty = tcy + tsy, tby = tcy - tsy, tlx = tcx - tsx, trx = tcx + tsx;
      float depth_t = dep(px, py + 1), depth_bb = dep(px, py - 1), depth_l = dep(px - 1, py), depth_r = dep(px + 1, py);
      depth_t = is_outside(depth_t) ? depth : depth_t;
      depth_bb = is_outside(depth_bb) ? depth : depth_bb;
      depth_l = is_outside(depth_l) ? depth : depth_l;
      depth_r = is_outside(depth_r) ? depth : depth_r;
      V3 world_t = fetch_xyz(cal, tcx, tty, depth_t);
      V3 world_b = fetch_xyz(cal, tcx, tby, depth_bb);
      V3 world_l = fetch_xyz(cal, tlx, tcy, depth_l);
      V3 world_r = fetch_xyz(cal, trx, tcy, depth_r);
      V3 n = normalize3(cross3(world_b - world_t, world_l - world_r));
      on[0] = n.x; on[1] = n.y; on[2] = n.z;
    }
  }
}

// glsl/pre_quality.fs:65-119 bilateral_filter, :43-48 normal_angle. The 169-tap get_color_diff (:115) is dead code.
void ro_pre_quality(const float* depth_b, const float* normals, int W, int H,
                    const float* cv_xyz, int CX, int CY, int CZ,
                    const float* camera_pos, float* out_quality /*[H][W]*/) {
  FwdCalib cal{cv_xyz, nullptr, CX, CY, CZ};
  const int ks = 6;
  auto is_outside = [](float d) { return (d <= 0.0f) || (d >= 1.0f); };
  auto dep = [&](int x, int y) { return depth_b[((size_t)clampi(y, 0, H - 1) * W + clampi(x, 0, W - 1)) * 2]; };
  const V3 cam{camera_pos[0], camera_pos[1], camera_pos[2]};
#pragma omp parallel for schedule(dynamic, 4)
  for (int py = 0; py < H; ++py) {
    for (int px = 0; px < W; ++px) {
      const size_t idx = (size_t)py * W + px;
      float depth = dep(px, py);
      if (is_outside(depth)) { out_quality[idx] = 0.0f; continue; }
      float dist_range_max = 0.35f * (depth / 1.0f);
      float dist_range_max_inv = 1.0f / dist_range_max;
      float w_range = 0.0f, border_samples = 0.0f, num_samples = 0.0f;
      for (int y = -ks; y <= ks; ++y)
        for (int x = -ks; x <= ks; ++x) {
          num_samples += 1.0f;
          float depth_s = dep(px + x, py + y);
          float depth_range = fabsf(depth_s - depth);
          if (is_outside(depth_s) || (depth_range > dist_range_max)) { border_samples += 1.0f; continue; }
          float gauss_range = 1.0f - gl_min(depth_range, dist_range_max) * dist_range_max_inv;
          w_range += gauss_range;
        }
      float lateral_quality = 1.0f - border_samples / num_samples;
      float quality_strong = gl_pow(lateral_quality, 6.0f);
      quality_strong *= gl_pow(w_range / num_samples, 6.0f);
      quality_strong /= depth * 6.5f;
      // normal_angle(vec3(coords.xy, depth), layer)
      const float tcx = ((float)px + 0.5f) / (float)W;
      const float tcy = ((float)py + 0.5f) / (float)H;
      V3 wn{normals[idx * 3], normals[idx * 3 + 1], normals[idx * 3 + 2]};
      V3 world_pos = fetch_xyz(cal, tcx, tcy, depth);
      float angle = dot3(normalize3(cam - world_pos), wn);
      quality_strong *= gl_pow(angle, 2.0f);
      out_quality[idx] = quality_strong;
    }
  }
}

}  // extern "C"

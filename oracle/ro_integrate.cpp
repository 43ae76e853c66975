// ORACLE (test infrastructure, NOT product code). The reference ships no tests; this file is pinned against the reference's
// own glsl/tsdf_integration.vs compiled as C++ and run on the CPU (oracle/glsl_host/, oracle/_ref/libref_glsl.so).
// Scalar restatement of the brick bookkeeping and of the per-voxel TSDF integration:
//   ReconIntegration::setVoxelSize / setBrickSize / divideBox / updateOccupiedBricks
//     (framework/reconstruction/recon_integration.cpp:341-354, 474-484, 361-407, 431-446),
//   VolumeSampler::resize / containedVoxels (framework/rendering/volume_sampler.cpp:33-62),
//   glsl/tsdf_integration.vs:23-59 driven by ReconIntegration::integrate (recon_integration.cpp:243-270).
#include "ro_math.h"
#include "rr_oracle.h"

#include <vector>

using namespace ro;

extern "C" {

// recon_integration.cpp:341-345: m_res_volume = ceil(bbox_size / voxel_size)
void ro_volume_res(const float* bbox_min, const float* bbox_max, float voxel_size, uint32_t* res_out) {
  for (int c = 0; c < 3; ++c) res_out[c] = (uint32_t)std::ceil((bbox_max[c] - bbox_min[c]) / voxel_size);
}

// recon_integration.cpp:475: m_brick_size = m_voxel_size * glm::round(size / m_voxel_size)
// glm-0.9.5.3 round(x) = float(int(x + 0.5)) for x >= 0 (glm/detail/func_common.inl:204)
float ro_adjust_brick_size(float voxel_size, float size) {
  float q = size / voxel_size;
  float r = (q < 0.0f) ? (float)(int)(q - 0.5f) : (float)(int)(q + 0.5f);
  return voxel_size * r;
}

// divideBox (recon_integration.cpp:361-388) + containedVoxels (volume_sampler.cpp:50-62).
// Pass ranges == nullptr to query the brick count/res only. ranges: int32 [num_bricks][6] = x0,x1,y0,y1,z0,z1
// (half-open voxel index ranges, clamped to the volume). Returns the number of bricks.
uint32_t ro_divide_box(const float* bbox_min, const float* bbox_max, float brick_size, const uint32_t* res_volume,
                       uint32_t* res_bricks_out, int32_t* ranges) {
  const V3 mn{bbox_min[0], bbox_min[1], bbox_min[2]};
  const V3 size{bbox_max[0] - mn.x, bbox_max[1] - mn.y, bbox_max[2] - mn.z};
  V3 start = mn;
  uint32_t rb[3] = {0, 0, 0};
  uint32_t count = 0;
  const V3 step{1.0f / (float)res_volume[0], 1.0f / (float)res_volume[1], 1.0f / (float)res_volume[2]};
  auto axis_range = [](float pos, float sz, float st, uint32_t dim, int32_t& lo, int32_t& hi) {
    uint32_t a = (uint32_t)(pos / st);
    uint32_t b = a;
    const float lim = (pos + sz) / st;
    while ((float)b < lim) ++b;
    lo = (int32_t)(a > dim ? dim : a);
    hi = (int32_t)(b > dim ? dim : b);
  };
  while (size.z - start.z + mn.z > 0.0f) {
    while (size.y - start.y + mn.y > 0.0f) {
      while (size.x - start.x + mn.x > 0.0f) {
        V3 rem{size.x - start.x + mn.x, size.y - start.y + mn.y, size.z - start.z + mn.z};
        V3 bs{gl_min(brick_size, rem.x), gl_min(brick_size, rem.y), gl_min(brick_size, rem.z)};
        if (ranges) {
          V3 pn = (start - mn) / size;
          V3 sn = bs / size;
          int32_t* r = ranges + (size_t)count * 6;
          axis_range(pn.x, sn.x, step.x, res_volume[0], r[0], r[1]);
          axis_range(pn.y, sn.y, step.y, res_volume[1], r[2], r[3]);
          axis_range(pn.z, sn.z, step.z, res_volume[2], r[4], r[5]);
        }
        ++count;
        start.x += brick_size;
        if (rb[2] == 0 && rb[1] == 0) ++rb[0];
      }
      start.x = mn.x;
      start.y += brick_size;
      if (rb[2] == 0) ++rb[1];
    }
    start.y = mn.y;
    start.z += brick_size;
    ++rb[2];
  }
  res_bricks_out[0] = rb[0]; res_bricks_out[1] = rb[1]; res_bricks_out[2] = rb[2];
  return count;
}

// The (pos, size) arguments divideBox passes to VolumeSampler::containedVoxels for every brick, in brick order
// (recon_integration.cpp:371-373): args float[num_bricks][6] = pos_n.xyz, size_n.xyz. Also returns the UNCLAMPED
// per-axis ranges (the raw loop bounds of volume_sampler.cpp:53-55) in raw_ranges int32[num_bricks][6] (nullable).
uint32_t ro_divide_box_args(const float* bbox_min, const float* bbox_max, float brick_size, const uint32_t* res_volume,
                            float* args, int32_t* raw_ranges) {
  const V3 mn{bbox_min[0], bbox_min[1], bbox_min[2]};
  const V3 size{bbox_max[0] - mn.x, bbox_max[1] - mn.y, bbox_max[2] - mn.z};
  const V3 step{1.0f / (float)res_volume[0], 1.0f / (float)res_volume[1], 1.0f / (float)res_volume[2]};
  V3 start = mn;
  uint32_t count = 0;
  auto raw = [](float pos, float sz, float st, int32_t& lo, int32_t& hi) {
    uint32_t a = (uint32_t)(pos / st), b = a;
    const float lim = (pos + sz) / st;
    while ((float)b < lim) ++b;
    lo = (int32_t)a; hi = (int32_t)b;
  };
  while (size.z - start.z + mn.z > 0.0f) {
    while (size.y - start.y + mn.y > 0.0f) {
      while (size.x - start.x + mn.x > 0.0f) {
        V3 rem{size.x - start.x + mn.x, size.y - start.y + mn.y, size.z - start.z + mn.z};
        V3 bs{gl_min(brick_size, rem.x), gl_min(brick_size, rem.y), gl_min(brick_size, rem.z)};
        V3 pn = (start - mn) / size;
        V3 sn = bs / size;
        if (args) { float* a = args + (size_t)count * 6; a[0] = pn.x; a[1] = pn.y; a[2] = pn.z; a[3] = sn.x; a[4] = sn.y; a[5] = sn.z; }
        if (raw_ranges) {
          int32_t* r = raw_ranges + (size_t)count * 6;
          raw(pn.x, sn.x, step.x, r[0], r[1]); raw(pn.y, sn.y, step.y, r[2], r[3]); raw(pn.z, sn.z, step.z, r[4], r[5]);
        }
        ++count;
        start.x += brick_size;
      }
      start.x = mn.x;
      start.y += brick_size;
    }
    start.y = mn.y;
    start.z += brick_size;
  }
  return count;
}

// updateOccupiedBricks (recon_integration.cpp:436-441): ascending ids with counter >= min_voxels.
uint32_t ro_occupied_bricks(const uint32_t* counters, uint32_t num_bricks, uint32_t min_voxels, uint32_t* occupied_out) {
  uint32_t n = 0;
  for (uint32_t i = 0; i < num_bricks; ++i)
    if (counters[i] >= min_voxels) occupied_out[n++] = i;
  return n;
}

namespace {
struct IntegrateArgs {
  int N; const float* inv; int IX, IY, IZ;
  const float* sil; const float* depth_b; const float* quality; int W, H;
  float limit; int X, Y, Z;
};

// tsdf_integration.vs:23-59 for the voxel (x, y, z); position from volume_sampler.cpp:36-42.
inline void integrate_voxel(const IntegrateArgs& a, int x, int y, int z, float* tsdf, float* weight) {
  const float stepX = 1.0f / (float)a.X, stepY = 1.0f / (float)a.Y, stepZ = 1.0f / (float)a.Z;
  const float px = ((float)x + 0.5f) * stepX, py = ((float)y + 0.5f) * stepY, pz = ((float)z + 0.5f) * stepZ;
  const float limit = a.limit;
  float weighted_tsd = limit;
  float total_weight = 0.0f;
  const size_t inv_stride = (size_t)a.IX * a.IY * a.IZ * 4;
  const size_t img = (size_t)a.W * a.H;
  for (int i = 0; i < a.N; ++i) {
    float pc[3];
    tex3d_linear<4>(a.inv + inv_stride * i, a.IX, a.IY, a.IZ, px, py, pz, pc, 3);
    float silhouette = tex2d_linear(a.sil + img * i, a.W, a.H, 1, 0, pc[0], pc[1]);
    if (silhouette < 1.0f) {
      if (weighted_tsd >= limit) { weighted_tsd = -limit; continue; }
    }
    float depth = tex2d_nearest(a.depth_b + img * 2 * i, a.W, a.H, 2, 0, pc[0], pc[1]);
    float sdist = pc[2] - depth;
    if (sdist <= -limit) {
      weighted_tsd = -limit;
    } else if (sdist >= limit) {
    } else {
      float w = tex2d_linear(a.quality + img * i, a.W, a.H, 1, 0, pc[0], pc[1]);
      weighted_tsd = (weighted_tsd * total_weight + w * sdist) / (total_weight + w);
      total_weight += w;
    }
  }
  // ivec3(position * res_tsdf)
  int sx = (int)(px * (float)a.X), sy = (int)(py * (float)a.Y), sz = (int)(pz * (float)a.Z);
  size_t o = ((size_t)sz * a.Y + sy) * a.X + sx;
  tsdf[o] = weighted_tsd;
  if (weight) weight[o] = total_weight;
}
}  // namespace

// ReconIntegration::integrate: clear to -limit, then every voxel (dense) or the voxels of every occupied brick.
// inv: [N][IZ][IY][IX][4]; sil: [N][H][W]; depth_b: [N][H][W][2]; quality: [N][H][W]; tsdf/weight: [Z][Y][X].
// weight (nullable) receives the shader-local total_weight (extension), cleared to 0.
void ro_integrate(int N, const float* inv, const int32_t* inv_res,
                  const float* sil, const float* depth_b, const float* quality, int W, int H,
                  float limit, const uint32_t* res, int use_bricks, const int32_t* brick_ranges,
                  const uint32_t* occupied, uint32_t num_occupied, float* tsdf, float* weight) {
  IntegrateArgs a{N, inv, inv_res[0], inv_res[1], inv_res[2], sil, depth_b, quality, W, H,
                  limit, (int)res[0], (int)res[1], (int)res[2]};
  const size_t nvox = (size_t)a.X * a.Y * a.Z;
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < nvox; ++i) { tsdf[i] = -limit; if (weight) weight[i] = 0.0f; }
  if (!use_bricks) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int z = 0; z < a.Z; ++z)
      for (int y = 0; y < a.Y; ++y)
        for (int x = 0; x < a.X; ++x) integrate_voxel(a, x, y, z, tsdf, weight);
  } else {
#pragma omp parallel for schedule(dynamic, 1)
    for (uint32_t b = 0; b < num_occupied; ++b) {
      const int32_t* r = brick_ranges + (size_t)occupied[b] * 6;
      for (int y = r[2]; y < r[3]; ++y)
        for (int x = r[0]; x < r[1]; ++x)
          for (int z = r[4]; z < r[5]; ++z) integrate_voxel(a, x, y, z, tsdf, weight);
    }
  }
}

}  // extern "C"

// Host-side geometry that stays on the CPU, as in the reference: sensor frusta / camera positions
// (framework/calibration/frustum.cpp:16-43, 97-176, built from the 8 corner voxels of cv_xyz as in
// CalibVolumes.cpp:98-113), volume resolution and the brick table (recon_integration.cpp:341-407 with
// VolumeSampler::containedVoxels, volume_sampler.cpp:50-62). Float-driven discrete decisions are evaluated with the
// same single-precision operation order as the reference's glm code (no fused multiply-adds on this path).
#include "rr_context.h"

#include <cmath>

namespace rr {
namespace {

struct f3 { float x, y, z; };
inline f3 add(f3 a, f3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline f3 sub(f3 a, f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline f3 mul(f3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline f3 divs(f3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline float dotg(f3 a, f3 b) { float tx = a.x * b.x, ty = a.y * b.y, tz = a.z * b.z; return tx + ty + tz; }   // glm compute_dot<tvec3>
inline f3 crossg(f3 a, f3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline f3 unit(f3 a) { float sqr = a.x * a.x + a.y * a.y + a.z * a.z; return mul(a, 1.0f / std::sqrt(sqr)); }
inline f3 avg4(f3 a, f3 b, f3 c, f3 d) { return divs(add(add(add(a, b), c), d), 4.0f); }
inline f3 mid(f3 a, f3 b) { return mul(add(a, b), 0.5f); }

// closest approach of the lines p + s*u and q + t*v, midpoint (frustum.cpp:97-111)
f3 line_line_midpoint(f3 p, f3 u, f3 q, f3 v) {
  const f3 w0 = sub(p, q);
  const float a = dotg(u, u), b = dotg(u, v), c = dotg(v, v), d = dotg(u, w0), e = dotg(v, w0);
  const float sc = (b * e - c * d) / (a * c - b * b);
  const float tc = (a * e - b * d) / (a * c - b * b);
  return mul(add(add(p, mul(u, sc)), add(q, mul(v, tc))), 0.5f);
}

}  // namespace

void host_frustum(const float* cv_xyz, const uint32_t res[3], float planes[6][4], float cam[3]) {
  const uint32_t X = res[0], Y = res[1], Z = res[2];
  auto voxel = [&](uint32_t x, uint32_t y, uint32_t z) {
    const float* p = cv_xyz + (((size_t)z * Y + y) * X + x) * 3;
    return f3{p[0], p[1], p[2]};
  };
  const uint32_t ex = X - 1, ey = Y - 1, ez = Z - 1;
  // corner order of getCornerPoints (CalibVolumes.cpp:98-113): near quad 0..3, far quad 4..7
  const f3 k[8] = {voxel(0, 0, 0), voxel(0, ey, 0), voxel(ex, ey, 0), voxel(ex, 0, 0),
                   voxel(0, 0, ez), voxel(0, ey, ez), voxel(ex, ey, ez), voxel(ex, 0, ez)};
  const f3 side_c[6] = {avg4(k[0], k[1], k[2], k[3]), avg4(k[4], k[5], k[6], k[7]), avg4(k[0], k[1], k[4], k[5]),
                        avg4(k[2], k[3], k[6], k[7]), avg4(k[1], k[2], k[5], k[6]), avg4(k[0], k[3], k[4], k[7])};
  const f3 e[12] = {mid(k[0], k[1]), mid(k[1], k[2]), mid(k[2], k[3]), mid(k[3], k[0]),
                    mid(k[4], k[5]), mid(k[5], k[6]), mid(k[6], k[7]), mid(k[7], k[4]),
                    mid(k[0], k[4]), mid(k[1], k[5]), mid(k[2], k[6]), mid(k[3], k[7])};
  const f3 nrm[6] = {unit(crossg(sub(e[0], e[2]), sub(e[3], e[2]))),    // near
                     unit(crossg(sub(e[4], e[6]), sub(e[5], e[7]))),    // far
                     unit(crossg(sub(e[0], e[4]), sub(e[9], e[8]))),    // left
                     unit(crossg(sub(e[2], e[6]), sub(e[11], e[10]))),  // right
                     unit(crossg(sub(e[9], e[10]), sub(e[1], e[5]))),   // top
                     unit(crossg(sub(e[8], e[11]), sub(e[7], e[3])))};  // bottom
  for (int i = 0; i < 6; ++i) {
    planes[i][0] = nrm[i].x; planes[i][1] = nrm[i].y; planes[i][2] = nrm[i].z;
    planes[i][3] = -dotg(nrm[i], side_c[i]);
  }
  const f3 cn = side_c[0], cf = side_c[1];
  const f3 view = sub(cf, cn);
  const f3 q0 = line_line_midpoint(k[0], sub(k[0], k[4]), cn, view);
  const f3 q1 = line_line_midpoint(k[1], sub(k[1], k[5]), cn, view);
  const f3 q2 = line_line_midpoint(k[2], sub(k[2], k[6]), cn, view);
  const f3 q3 = line_line_midpoint(k[3], sub(k[3], k[7]), cn, view);
  const f3 cp = avg4(q0, q1, q2, q3);
  cam[0] = cp.x; cam[1] = cp.y; cam[2] = cp.z;
}

void host_volume_res(const float bmin[3], const float bmax[3], float voxel_size, uint32_t res[3]) {
  for (int a = 0; a < 3; ++a) res[a] = (uint32_t)std::ceil((bmax[a] - bmin[a]) / voxel_size);
}

float host_adjust_brick_size(float voxel_size, float size) {
  const float ratio = size / voxel_size;
  const float rounded = ratio < 0.0f ? float(int(ratio - 0.5f)) : float(int(ratio + 0.5f));   // glm::round
  return voxel_size * rounded;
}

// Walks the box exactly like ReconIntegration::divideBox: float accumulation of the brick origin, the last brick of
// a row truncated to the box, and per-axis voxel ranges from the float comparisons of containedVoxels.
uint32_t host_divide_box(const float bmin[3], const float bmax[3], float brick_size, const uint32_t res[3],
                         uint32_t res_bricks[3], std::vector<int32_t>* ranges) {
  const float ext[3] = {bmax[0] - bmin[0], bmax[1] - bmin[1], bmax[2] - bmin[2]};
  const float step[3] = {1.0f / (float)res[0], 1.0f / (float)res[1], 1.0f / (float)res[2]};
  // per-axis brick intervals: the x interval of a brick depends only on its x origin, etc.
  std::vector<int32_t> lo[3], hi[3];
  for (int a = 0; a < 3; ++a) {
    float start = bmin[a];
    while (ext[a] - start + bmin[a] > 0.0f) {
      const float remaining = ext[a] - start + bmin[a];
      const float bsz = (remaining < brick_size) ? remaining : brick_size;   // glm::min(brick, remaining)
      const float pos_n = (start - bmin[a]) / ext[a];
      const float size_n = bsz / ext[a];
      uint32_t first = (uint32_t)(pos_n / step[a]);
      const float bound = (pos_n + size_n) / step[a];
      uint32_t last = first;
      while ((float)last < bound) ++last;
      if (first > res[a]) first = res[a];
      if (last > res[a]) last = res[a];
      lo[a].push_back((int32_t)first);
      hi[a].push_back((int32_t)last);
      start += brick_size;
    }
    res_bricks[a] = (uint32_t)lo[a].size();
  }
  const uint32_t count = res_bricks[0] * res_bricks[1] * res_bricks[2];
  if (ranges) {
    ranges->resize((size_t)count * 6);
    size_t o = 0;
    for (uint32_t z = 0; z < res_bricks[2]; ++z)
      for (uint32_t y = 0; y < res_bricks[1]; ++y)
        for (uint32_t x = 0; x < res_bricks[0]; ++x) {
          int32_t* r = ranges->data() + o;
          r[0] = lo[0][x]; r[1] = hi[0][x]; r[2] = lo[1][y]; r[3] = hi[1][y]; r[4] = lo[2][z]; r[5] = hi[2][z];
          o += 6;
        }
  }
  return count;
}

}  // namespace rr

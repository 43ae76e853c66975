// ORACLE (test infrastructure, NOT product code).
// Scalar restatement of the offline calibration-volume inversion:
//   Frustum (framework/calibration/frustum.cpp:16-43, 97-176)                      -- pinned against oracle/_ref
//   CalibrationInverter::getXyzSamples / inverseDistance / calculateInverseVolumes
//     (framework/calibration/calibration_inverter.cpp:38-69, 99-155)                -- pinned against oracle/_ref
//   NearestNeighbourSearch::search (nearest_neighbour_search.cpp:32-43) = CGAL Orthogonal_k_neighbor_search,
//     a third-party dependency ABSENT from /root/reference (system libcgal, no version pinned,
//     utils/dependencies.txt:4). Restated from its published contract: exact k nearest neighbours under the
//     Euclidean metric evaluated in double on float-promoted coordinates, reported in ascending distance.
//     Ties (unspecified in CGAL) are broken by the linear sample index of getXyzSamples (x-outer, z-inner).
// The glm-0.9.5.3 functions used by this C++ path do not fuse: dot = (x*x' + y*y') + z*z'.
#include "ro_math.h"
#include "rr_oracle.h"

#include <algorithm>
#include <array>
#include <cstdio>
#include <vector>

using namespace ro;

namespace {

inline float gdot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 gnormalize(V3 a) { float sqr = (a.x * a.x + a.y * a.y) + a.z * a.z; return a * (1.0f / sqrtf(sqr)); }

// frustum.cpp:97-111
V3 closest_point(V3 p, V3 u, V3 q, V3 v) {
  V3 w0 = p - q;
  float a = gdot3(u, u), b = gdot3(u, v), c = gdot3(v, v), d = gdot3(u, w0), e = gdot3(v, w0);
  float sc = (b * e - c * d) / (a * c - b * b);
  float tc = (a * e - b * d) / (a * c - b * b);
  V3 pc = p + u * sc;
  V3 qc = q + v * tc;
  return (pc + qc) * 0.5f;
}

void corner_points(const float* xyz, int X, int Y, int Z, V3* c) {
  auto at = [&](int x, int y, int z) { const float* p = xyz + (((size_t)z * Y + y) * X + x) * 3; return V3{p[0], p[1], p[2]}; };
  int ex = X - 1, ey = Y - 1, ez = Z - 1;
  c[0] = at(0, 0, 0);  c[1] = at(0, ey, 0);  c[2] = at(ex, ey, 0);  c[3] = at(ex, 0, 0);
  c[4] = at(0, 0, ez); c[5] = at(0, ey, ez); c[6] = at(ex, ey, ez); c[7] = at(ex, 0, ez);
}

// frustum.cpp:113-176 getSideCenters / getEdgeCenters / getSideNormals / getPlanes
void frustum_planes(const V3* pc, V4* planes) {
  V3 sc[6], ec[12], n[6];
  sc[0] = (pc[0] + pc[1] + pc[2] + pc[3]) / 4.0f;
  sc[1] = (pc[4] + pc[5] + pc[6] + pc[7]) / 4.0f;
  sc[2] = (pc[0] + pc[1] + pc[4] + pc[5]) / 4.0f;
  sc[3] = (pc[2] + pc[3] + pc[6] + pc[7]) / 4.0f;
  sc[4] = (pc[1] + pc[2] + pc[5] + pc[6]) / 4.0f;
  sc[5] = (pc[0] + pc[3] + pc[4] + pc[7]) / 4.0f;
  ec[0] = (pc[0] + pc[1]) * 0.5f; ec[1] = (pc[1] + pc[2]) * 0.5f; ec[2] = (pc[2] + pc[3]) * 0.5f; ec[3] = (pc[3] + pc[0]) * 0.5f;
  ec[4] = (pc[4] + pc[5]) * 0.5f; ec[5] = (pc[5] + pc[6]) * 0.5f; ec[6] = (pc[6] + pc[7]) * 0.5f; ec[7] = (pc[7] + pc[4]) * 0.5f;
  ec[8] = (pc[0] + pc[4]) * 0.5f; ec[9] = (pc[1] + pc[5]) * 0.5f; ec[10] = (pc[2] + pc[6]) * 0.5f; ec[11] = (pc[3] + pc[7]) * 0.5f;
  n[0] = gnormalize(cross3(ec[0] - ec[2], ec[3] - ec[2]));
  n[1] = gnormalize(cross3(ec[4] - ec[6], ec[5] - ec[7]));
  n[2] = gnormalize(cross3(ec[0] - ec[4], ec[9] - ec[8]));
  n[3] = gnormalize(cross3(ec[2] - ec[6], ec[11] - ec[10]));
  n[4] = gnormalize(cross3(ec[9] - ec[10], ec[1] - ec[5]));
  n[5] = gnormalize(cross3(ec[8] - ec[11], ec[7] - ec[3]));
  for (int i = 0; i < 6; ++i) planes[i] = V4{n[i].x, n[i].y, n[i].z, -gdot3(n[i], sc[i])};
}

inline bool frustum_inside(const V4* planes, V3 p) {
  for (int i = 0; i < 6; ++i) {
    const V4& pl = planes[i];
    float d = (pl.x * p.x + pl.y * p.y) + (pl.z * p.z + pl.w * 1.0f);   // glm dot(vec4, vec4)
    if (d < 0.0f) return false;
  }
  return true;
}

struct Neigh { double d2; uint32_t idx; };
inline bool neigh_less(const Neigh& a, const Neigh& b) { return a.d2 < b.d2 || (a.d2 == b.d2 && a.idx < b.idx); }

// Exact kNN over a uniform grid (search structure is an implementation detail; the result is the exact k-set).
struct KnnGrid {
  std::vector<V3> pos;            // sample positions in getXyzSamples order
  double gmin[3], cell;
  int dim[3];
  std::vector<uint32_t> cell_start, sorted;

  void build(const float* xyz, int X, int Y, int Z) {
    const size_t n = (size_t)X * Y * Z;
    pos.resize(n);
    size_t k = 0;
    for (int x = 0; x < X; ++x) for (int y = 0; y < Y; ++y) for (int z = 0; z < Z; ++z) {
      const float* p = xyz + (((size_t)z * Y + y) * X + x) * 3;
      pos[k++] = V3{p[0], p[1], p[2]};
    }
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    for (const V3& p : pos) {
      const double c[3] = {p.x, p.y, p.z};
      for (int a = 0; a < 3; ++a) { mn[a] = std::min(mn[a], c[a]); mx[a] = std::max(mx[a], c[a]); }
    }
    double vol = 1.0;
    for (int a = 0; a < 3; ++a) vol *= std::max(mx[a] - mn[a], 1e-9);
    cell = std::cbrt(vol / std::max<double>(1.0, (double)n / 4.0));
    for (int a = 0; a < 3; ++a) {
      gmin[a] = mn[a];
      dim[a] = (int)std::min(1024.0, std::max(1.0, std::ceil((mx[a] - mn[a]) / cell + 1e-9)));
    }
    const size_t ncell = (size_t)dim[0] * dim[1] * dim[2];
    cell_start.assign(ncell + 1, 0);
    std::vector<uint32_t> cid(n);
    for (size_t i = 0; i < n; ++i) { cid[i] = (uint32_t)cell_of(pos[i]); ++cell_start[cid[i] + 1]; }
    for (size_t c = 0; c < ncell; ++c) cell_start[c + 1] += cell_start[c];
    sorted.resize(n);
    std::vector<uint32_t> fill(cell_start.begin(), cell_start.end() - 1);
    for (size_t i = 0; i < n; ++i) sorted[fill[cid[i]]++] = (uint32_t)i;
  }
  int axis_cell(double v, int a) const {
    int c = (int)std::floor((v - gmin[a]) / cell);
    return c < 0 ? 0 : (c >= dim[a] ? dim[a] - 1 : c);
  }
  size_t cell_of(V3 p) const {
    return ((size_t)axis_cell(p.z, 2) * dim[1] + axis_cell(p.y, 1)) * dim[0] + axis_cell(p.x, 0);
  }
  // k <= 8
  int search(V3 q, int k, Neigh* best) const {
    const double qd[3] = {q.x, q.y, q.z};
    const int c[3] = {axis_cell(qd[0], 0), axis_cell(qd[1], 1), axis_cell(qd[2], 2)};
    int found = 0;
    const int rmax = std::max(dim[0], std::max(dim[1], dim[2]));
    for (int r = 0; r <= rmax; ++r) {
      const int z0 = std::max(c[2] - r, 0), z1 = std::min(c[2] + r, dim[2] - 1);
      const int y0 = std::max(c[1] - r, 0), y1 = std::min(c[1] + r, dim[1] - 1);
      const int x0 = std::max(c[0] - r, 0), x1 = std::min(c[0] + r, dim[0] - 1);
      for (int z = z0; z <= z1; ++z) for (int y = y0; y <= y1; ++y) {
        const bool yz_shell = (std::abs(z - c[2]) == r) || (std::abs(y - c[1]) == r);
        for (int x = x0; x <= x1; ++x) {
          if (!yz_shell && std::abs(x - c[0]) != r) continue;   // interior cell: visited in an earlier shell
          const size_t ci = ((size_t)z * dim[1] + y) * dim[0] + x;
          for (uint32_t s = cell_start[ci]; s < cell_start[ci + 1]; ++s) {
            const uint32_t i = sorted[s];
            const double dx = qd[0] - (double)pos[i].x, dy = qd[1] - (double)pos[i].y, dz = qd[2] - (double)pos[i].z;
            Neigh nb{(dx * dx + dy * dy) + dz * dz, i};
            if (found < k) {
              int j = found++;
              while (j > 0 && neigh_less(nb, best[j - 1])) { best[j] = best[j - 1]; --j; }
              best[j] = nb;
            } else if (neigh_less(nb, best[k - 1])) {
              int j = k - 1;
              while (j > 0 && neigh_less(nb, best[j - 1])) { best[j] = best[j - 1]; --j; }
              best[j] = nb;
            }
          }
        }
      }
      if (found == k) {
        // distance from q to the nearest face of the covered cell cube that still has grid cells behind it
        double dout = 1e300;
        for (int a = 0; a < 3; ++a) {
          if (c[a] - r > 0) dout = std::min(dout, qd[a] - (gmin[a] + (double)(c[a] - r) * cell));
          if (c[a] + r < dim[a] - 1) dout = std::min(dout, (gmin[a] + (double)(c[a] + r + 1) * cell) - qd[a]);
        }
        if (dout == 1e300) break;                       // whole grid covered
        if (dout > 0.0 && best[k - 1].d2 < dout * dout) break;
      }
    }
    return found;
  }
};

}  // namespace

extern "C" {

// planes_out: float[6][4]; campos_out: float[3] (Frustum::getCameraPos, frustum.cpp:21-33)
void ro_frustum(const float* cv_xyz, int X, int Y, int Z, float* planes_out, float* campos_out) {
  V3 c[8]; V4 pl[6];
  corner_points(cv_xyz, X, Y, Z, c);
  frustum_planes(c, pl);
  for (int i = 0; i < 6; ++i) { planes_out[i * 4] = pl[i].x; planes_out[i * 4 + 1] = pl[i].y; planes_out[i * 4 + 2] = pl[i].z; planes_out[i * 4 + 3] = pl[i].w; }
  V3 center_near = (c[0] + c[1] + c[2] + c[3]) / 4.0f;
  V3 center_far = (c[4] + c[5] + c[6] + c[7]) / 4.0f;
  V3 view_dir = center_far - center_near;
  V3 p3 = closest_point(c[0], c[0] - c[4], center_near, view_dir);
  V3 p4 = closest_point(c[1], c[1] - c[5], center_near, view_dir);
  V3 p5 = closest_point(c[2], c[2] - c[6], center_near, view_dir);
  V3 p6 = closest_point(c[3], c[3] - c[7], center_near, view_dir);
  V3 cam = (p3 + p4 + p5 + p6) / 4.0f;
  campos_out[0] = cam.x; campos_out[1] = cam.y; campos_out[2] = cam.z;
}

int ro_frustum_inside(const float* planes, const float* p) {
  V4 pl[6];
  for (int i = 0; i < 6; ++i) pl[i] = V4{planes[i * 4], planes[i * 4 + 1], planes[i * 4 + 2], planes[i * 4 + 3]};
  return frustum_inside(pl, V3{p[0], p[1], p[2]}) ? 1 : 0;
}

// CalibrationInverter::calculateInverseVolumes for one sensor. out: float[oz][oy][ox][4].
// neigh_out (nullable): uint32 [oz][oy][ox][8] linear sample ids (x*Y*Z + y*Z + z) of the 8-NN, 0xFFFFFFFF if culled.
// brute != 0 uses an O(n) scan per voxel instead of the grid (cross-check of the search structure).
void ro_calib_invert(const float* cv_xyz, int X, int Y, int Z, const float* bbox_min, const float* bbox_max,
                     const uint32_t* out_res, float* out, uint32_t* neigh_out, int brute) {
  V3 c[8]; V4 pl[6];
  corner_points(cv_xyz, X, Y, Z, c);
  frustum_planes(c, pl);
  KnnGrid grid;
  grid.build(cv_xyz, X, Y, Z);
  const V3 dims{bbox_max[0] - bbox_min[0], bbox_max[1] - bbox_min[1], bbox_max[2] - bbox_min[2]};
  const V3 trans{bbox_min[0], bbox_min[1], bbox_min[2]};
  const V3 volume_step{1.0f / (float)out_res[0], 1.0f / (float)out_res[1], 1.0f / (float)out_res[2]};
  const V3 sample_step = dims * volume_step;
  const V3 sample_start = trans + sample_step * 0.5f;
  const V3 calib_dims{(float)X, (float)Y, (float)Z};
  const int ox = (int)out_res[0], oy = (int)out_res[1], oz = (int)out_res[2];
  const size_t n = grid.pos.size();
#pragma omp parallel for schedule(dynamic, 1)
  for (int x = 0; x < ox; ++x) {
    for (int y = 0; y < oy; ++y) {
      for (int z = 0; z < oz; ++z) {
        const size_t o = ((size_t)z * oy + y) * ox + x;
        V3 sp = sample_start + V3{(float)x, (float)y, (float)z} * sample_step;
        if (!frustum_inside(pl, sp)) {
          out[o * 4] = out[o * 4 + 1] = out[o * 4 + 2] = out[o * 4 + 3] = -1.0f;
          if (neigh_out) for (int k = 0; k < 8; ++k) neigh_out[o * 8 + k] = 0xFFFFFFFFu;
          continue;
        }
        Neigh best[8];
        int found;
        if (brute) {
          found = 0;
          for (size_t i = 0; i < n; ++i) {
            const double dx = (double)sp.x - (double)grid.pos[i].x, dy = (double)sp.y - (double)grid.pos[i].y, dz = (double)sp.z - (double)grid.pos[i].z;
            Neigh nb{(dx * dx + dy * dy) + dz * dz, (uint32_t)i};
            if (found < 8) { int j = found++; while (j > 0 && neigh_less(nb, best[j - 1])) { best[j] = best[j - 1]; --j; } best[j] = nb; }
            else if (neigh_less(nb, best[7])) { int j = 7; while (j > 0 && neigh_less(nb, best[j - 1])) { best[j] = best[j - 1]; --j; } best[j] = nb; }
          }
        } else {
          found = grid.search(sp, 8, best);
        }
        // inverseDistance (calibration_inverter.cpp:55-69)
        float total_weight = 0.0f;
        V3 wi{0.0f, 0.0f, 0.0f};
        for (int k = 0; k < found; ++k) {
          const uint32_t i = best[k].idx;
          V3 d = grid.pos[i] - sp;
          float weight = 1.0f / sqrtf(gdot3(d, d));
          const uint32_t iz = i % (uint32_t)Z, iy = (i / (uint32_t)Z) % (uint32_t)Y, ix = i / ((uint32_t)Z * (uint32_t)Y);
          wi = wi + V3{(float)ix, (float)iy, (float)iz} * weight;
          total_weight += weight;
          if (neigh_out) neigh_out[o * 8 + k] = i;
        }
        wi = wi / total_weight;
        V3 r = (wi + V3{0.5f, 0.5f, 0.5f}) / calib_dims;
        out[o * 4] = r.x; out[o * 4 + 1] = r.y; out[o * 4 + 2] = r.z; out[o * 4 + 3] = 1.0f;
      }
    }
  }
}

}  // extern "C"

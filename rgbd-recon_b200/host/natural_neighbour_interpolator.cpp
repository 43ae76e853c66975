// See natural_neighbour_interpolator.hpp. Host C++ only (no CUDA): the reference never calls this class on the per-frame path.
#include "natural_neighbour_interpolator.hpp"

#include <algorithm>
#include <cmath>
#include <ostream>
#include <unordered_set>

namespace kinect {

std::ostream& operator<<(std::ostream& o, const nniSample& s) {        // NaturalNeighbourInterpolator.cpp:7-13
  o << "s_pos: (" << s.s_pos.x << "," << s.s_pos.y << "," << s.s_pos.z << ") "
    << "s_pos_off: (" << s.s_pos_off.x << "," << s.s_pos_off.y << "," << s.s_pos_off.z << ") "
    << "s_tex_off: (" << s.s_tex_off.u << "," << s.s_tex_off.v << ")"
    << "s_quality: " << s.quality;
  return o;
}

namespace {

struct P3 { double x, y, z; };
inline P3 operator+(P3 a, P3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline P3 operator-(P3 a, P3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline P3 operator*(P3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline double dot(P3 a, P3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline P3 cross(P3 a, P3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double norm(P3 a) { return std::sqrt(dot(a, a)); }

// A convex polyhedron as its faces (convex polygons, vertices in order around the face); `site` = the sample whose
// bisector plane carries the face, or -1 for the faces of the initial box.
struct Face { std::vector<P3> v; int site; };
struct Poly {
  std::vector<Face> faces;
  bool empty() const { return faces.size() < 4; }
};

Poly box(P3 c, double h) {
  Poly p;
  const P3 v[8] = {{c.x - h, c.y - h, c.z - h}, {c.x + h, c.y - h, c.z - h}, {c.x + h, c.y + h, c.z - h}, {c.x - h, c.y + h, c.z - h},
                   {c.x - h, c.y - h, c.z + h}, {c.x + h, c.y - h, c.z + h}, {c.x + h, c.y + h, c.z + h}, {c.x - h, c.y + h, c.z + h}};
  const int f[6][4] = {{0, 1, 2, 3}, {4, 5, 6, 7}, {0, 1, 5, 4}, {2, 3, 7, 6}, {1, 2, 6, 5}, {0, 3, 7, 4}};
  for (auto& q : f) p.faces.push_back(Face{{v[q[0]], v[q[1]], v[q[2]], v[q[3]]}, -1});
  return p;
}

// Keep the part of `p` with dot(n, x) <= d. The cut, if any, becomes a new face tagged `site`. eps is an absolute distance.
void clip(Poly& p, P3 n, double d, int site, double eps) {
  const double nl = norm(n);
  if (nl == 0.0) return;
  n = n * (1.0 / nl);
  d /= nl;
  std::vector<P3> cap;
  std::vector<Face> kept;
  bool any_out = false;
  for (Face& f : p.faces) {
    const size_t m = f.v.size();
    std::vector<double> s(m);
    bool in = false, out = false;
    for (size_t i = 0; i < m; ++i) {
      s[i] = dot(n, f.v[i]) - d;
      if (s[i] > eps) out = true;
      else if (s[i] < -eps) in = true;
    }
    if (!out) { kept.push_back(std::move(f)); continue; }     // entirely inside or on the plane
    any_out = true;
    // a face that is cut (or dropped): its vertices on the plane are corners of the cut
    for (size_t i = 0; i < m; ++i) if (std::fabs(s[i]) <= eps) cap.push_back(f.v[i]);
    if (!in) continue;                                        // nothing of it lies strictly inside
    Face g;
    g.site = f.site;
    for (size_t i = 0; i < m; ++i) {                          // Sutherland-Hodgman, "on the plane" counts as inside
      const size_t j = (i + 1) % m;
      if (s[i] <= eps) g.v.push_back(f.v[i]);
      if ((s[i] > eps && s[j] < -eps) || (s[i] < -eps && s[j] > eps)) {
        const P3 x = f.v[i] + (f.v[j] - f.v[i]) * (s[i] / (s[i] - s[j]));
        g.v.push_back(x);
        cap.push_back(x);
      }
    }
    if (g.v.size() >= 3) kept.push_back(std::move(g));
  }
  p.faces.swap(kept);                                         // (faces were moved into `kept`, cut or not)
  if (!any_out) return;
  // the cap: the cut points are the corners of a convex polygon in the plane; order them by angle around their centroid
  if (cap.size() >= 3) {
    P3 c{0, 0, 0};
    for (auto& x : cap) c = c + x;
    c = c * (1.0 / (double)cap.size());
    P3 u = cap[0] - c;
    for (size_t i = 1; i < cap.size() && norm(u) <= eps; ++i) u = cap[i] - c;
    if (norm(u) > eps) {
      u = u * (1.0 / norm(u));
      const P3 w = cross(n, u);
      std::vector<std::pair<double, P3>> ang;
      for (auto& x : cap) ang.push_back({std::atan2(dot(x - c, w), dot(x - c, u)), x});
      std::sort(ang.begin(), ang.end(), [](const std::pair<double, P3>& a, const std::pair<double, P3>& b) { return a.first < b.first; });
      Face g;
      g.site = site;
      for (auto& a : ang) {
        if (!g.v.empty() && norm(a.second - g.v.back()) <= eps) continue;     // duplicates (each cut point is met by two faces)
        g.v.push_back(a.second);
      }
      while (g.v.size() > 1 && norm(g.v.front() - g.v.back()) <= eps) g.v.pop_back();
      if (g.v.size() >= 3) p.faces.push_back(std::move(g));
    }
  }
}

double face_area(const Face& f, P3* normal_out = nullptr) {
  P3 a{0, 0, 0};
  for (size_t i = 1; i + 1 < f.v.size(); ++i) a = a + cross(f.v[i] - f.v[0], f.v[i + 1] - f.v[0]);
  if (normal_out) *normal_out = a;
  return 0.5 * norm(a);
}

// Volume of a convex polyhedron: sum over faces of area * distance(interior point, face plane) / 3.
double volume(const Poly& p) {
  if (p.empty()) return 0.0;
  P3 c{0, 0, 0};
  size_t n = 0;
  for (auto& f : p.faces) for (auto& x : f.v) { c = c + x; ++n; }
  if (!n) return 0.0;
  c = c * (1.0 / (double)n);
  double vol = 0.0;
  for (auto& f : p.faces) {
    P3 a;
    const double area = face_area(f, &a);
    const double al = norm(a);
    if (al == 0.0) continue;
    vol += area * std::fabs(dot(a * (1.0 / al), f.v[0] - c)) / 3.0;
  }
  return vol;
}

}  // namespace

NaturalNeighbourInterpolator::NaturalNeighbourInterpolator(const std::vector<nniSample>& samples) : m_samples(samples) {
  for (int a = 0; a < 3; ++a) { m_min[a] = 0.0; m_max[a] = 1.0; m_dim[a] = 1; }
  m_cell = 1.0;
  if (m_samples.empty()) return;
  for (int a = 0; a < 3; ++a) { m_min[a] = 1e300; m_max[a] = -1e300; }
  for (auto& s : m_samples) {
    const double p[3] = {s.s_pos.x, s.s_pos.y, s.s_pos.z};
    for (int a = 0; a < 3; ++a) { m_min[a] = std::min(m_min[a], p[a]); m_max[a] = std::max(m_max[a], p[a]); }
  }
  double ext[3], vol = 1.0;
  for (int a = 0; a < 3; ++a) { ext[a] = std::max(m_max[a] - m_min[a], 1e-9); vol *= ext[a]; }
  m_cell = std::cbrt(vol * 2.0 / (double)m_samples.size());             // about two samples per bucket
  m_cell = std::max(m_cell, 1e-9);
  size_t cells = 1;
  for (int a = 0; a < 3; ++a) { m_dim[a] = std::max(1, std::min(256, (int)std::ceil(ext[a] / m_cell))); cells *= (size_t)m_dim[a]; }
  auto cell_of = [&](const nniSample& s) {
    const double p[3] = {s.s_pos.x, s.s_pos.y, s.s_pos.z};
    size_t idx = 0, stride = 1;
    for (int a = 0; a < 3; ++a) {
      int c = (int)std::floor((p[a] - m_min[a]) / ext[a] * m_dim[a]);
      c = std::max(0, std::min(m_dim[a] - 1, c));
      idx += (size_t)c * stride;
      stride *= (size_t)m_dim[a];
    }
    return idx;
  };
  m_cell_start.assign(cells + 1, 0);
  for (auto& s : m_samples) ++m_cell_start[cell_of(s) + 1];
  for (size_t i = 0; i < cells; ++i) m_cell_start[i + 1] += m_cell_start[i];
  m_cell_items.resize(m_samples.size());
  std::vector<uint32_t> fill(m_cell_start.begin(), m_cell_start.end() - 1);
  for (uint32_t i = 0; i < m_samples.size(); ++i) m_cell_items[fill[cell_of(m_samples[i])]++] = i;
}

NaturalNeighbourInterpolator::~NaturalNeighbourInterpolator() {}

void NaturalNeighbourInterpolator::sitesWithin(const double q[3], double radius, std::vector<uint32_t>& out) const {
  out.clear();
  int lo[3], hi[3];
  for (int a = 0; a < 3; ++a) {
    const double ext = std::max(m_max[a] - m_min[a], 1e-9);
    lo[a] = std::max(0, std::min(m_dim[a] - 1, (int)std::floor((q[a] - radius - m_min[a]) / ext * m_dim[a])));
    hi[a] = std::max(0, std::min(m_dim[a] - 1, (int)std::floor((q[a] + radius - m_min[a]) / ext * m_dim[a])));
  }
  const double r2 = radius * radius;
  for (int z = lo[2]; z <= hi[2]; ++z)
    for (int y = lo[1]; y <= hi[1]; ++y)
      for (int x = lo[0]; x <= hi[0]; ++x) {
        const size_t c = ((size_t)z * m_dim[1] + y) * m_dim[0] + x;
        for (uint32_t k = m_cell_start[c]; k < m_cell_start[c + 1]; ++k) {
          const nniSample& s = m_samples[m_cell_items[k]];
          const double dx = s.s_pos.x - q[0], dy = s.s_pos.y - q[1], dz = s.s_pos.z - q[2];
          if (dx * dx + dy * dy + dz * dz <= r2) out.push_back(m_cell_items[k]);
        }
      }
}

bool NaturalNeighbourInterpolator::coordinates(double qx, double qy, double qz, std::vector<std::pair<uint32_t, double>>& coords,
                                               double& norm_out) const {
  coords.clear();
  norm_out = 0.0;
  if (m_samples.size() < 4) return false;
  const double q[3] = {qx, qy, qz};
  const P3 Q{qx, qy, qz};
  double diag = 0.0;
  for (int a = 0; a < 3; ++a) diag += (m_max[a] - m_min[a]) * (m_max[a] - m_min[a]);
  diag = std::sqrt(diag);
  if (!(diag > 0.0)) return false;
  const double eps = 1e-11 * diag;
  // outside the samples' bounding box there is no bounded cell; inside, start from a box that any bounded cell fits in
  for (int a = 0; a < 3; ++a) if (q[a] < m_min[a] || q[a] > m_max[a]) return false;
  const double H = 4.0 * diag;
  Poly cell = box(Q, H);
  std::unordered_set<uint32_t> used;                          // samples already clipped against (a few hundred at most)
  std::vector<uint32_t> near;
  double radius = 2.0 * m_cell;
  auto site = [&](uint32_t i) { return P3{m_samples[i].s_pos.x, m_samples[i].s_pos.y, m_samples[i].s_pos.z}; };
  for (int round = 0; round < 64; ++round) {
    sitesWithin(q, radius, near);
    std::vector<std::pair<double, uint32_t>> order;
    for (uint32_t i : near) if (!used.count(i)) { const P3 d = site(i) - Q; order.push_back({dot(d, d), i}); }
    std::sort(order.begin(), order.end());
    for (auto& o : order) {
      used.insert(o.second);
      const P3 p = site(o.second);
      if (o.first <= eps * eps) {                              // q is a sample: its value, with weight 1
        coords.assign(1, {o.second, 1.0});
        norm_out = 1.0;
        return true;
      }
      // bisector: points closer to q than to p satisfy (p - q) . x <= (|p|^2 - |q|^2) / 2
      clip(cell, p - Q, 0.5 * (dot(p, p) - dot(Q, Q)), (int)o.second, eps);
      if (cell.empty()) return false;
    }
    double rmax = 0.0;
    for (auto& f : cell.faces) for (auto& x : f.v) rmax = std::max(rmax, norm(x - Q));
    const double need = 2.0 * rmax * (1.0 + 1e-9);            // security radius: farther samples cannot cut the cell
    if (need <= radius) break;
    if (radius > 4.0 * H) break;                              // every sample has been seen
    radius = std::min(need, 2.0 * radius);                    // grow geometrically, never past what is needed
  }
  for (auto& f : cell.faces)
    if (f.site < 0 && face_area(f) > 0.0) return false;        // still bounded by the start box: unbounded Voronoi cell
  // natural neighbours = the samples that own a face of the cell
  std::vector<int> nb;
  for (auto& f : cell.faces) if (f.site >= 0 && std::find(nb.begin(), nb.end(), f.site) == nb.end()) nb.push_back(f.site);
  if (nb.empty()) return false;
  for (int i : nb) {
    // the part of the cell whose nearest old sample is p_i (for x in the cell that sample is always a natural neighbour)
    Poly part = cell;
    const P3 pi = site((uint32_t)i);
    for (int j : nb) {
      if (j == i || part.empty()) continue;
      const P3 pj = site((uint32_t)j);
      clip(part, pj - pi, 0.5 * (dot(pj, pj) - dot(pi, pi)), -2, eps);
    }
    const double v = volume(part);
    if (v > 0.0) { coords.push_back({(uint32_t)i, v}); norm_out += v; }
  }
  return norm_out > 0.0 && !coords.empty();
}

bool NaturalNeighbourInterpolator::interpolate(nniSample& ipolant) {
  std::vector<std::pair<uint32_t, double>> coor_sibson;
  double norm_coeff_sibson = 0.0;
  if (!coordinates(ipolant.s_pos.x, ipolant.s_pos.y, ipolant.s_pos.z, coor_sibson, norm_coeff_sibson)) return false;
  if (coor_sibson.empty()) return false;                                 // NaturalNeighbourInterpolator.cpp:49-51
  double pos_off[3] = {0.0, 0.0, 0.0}, tex_off[2] = {0.0, 0.0};          // xyz_d / uv_d accumulators (:59-62)
  for (auto& c : coor_sibson) {
    const double contribution_i = c.second;
    const nniSample& s = m_samples[c.first];
    pos_off[0] += contribution_i * s.s_pos_off.x;
    pos_off[1] += contribution_i * s.s_pos_off.y;
    pos_off[2] += contribution_i * s.s_pos_off.z;
    tex_off[0] += contribution_i * s.s_tex_off.u;
    tex_off[1] += contribution_i * s.s_tex_off.v;
  }
  ipolant.s_pos_off.x = (float)(pos_off[0] / norm_coeff_sibson);         // :80-85
  ipolant.s_pos_off.y = (float)(pos_off[1] / norm_coeff_sibson);
  ipolant.s_pos_off.z = (float)(pos_off[2] / norm_coeff_sibson);
  ipolant.s_tex_off.u = (float)(tex_off[0] / norm_coeff_sibson);
  ipolant.s_tex_off.v = (float)(tex_off[1] / norm_coeff_sibson);
  return true;
}

}  // namespace kinect

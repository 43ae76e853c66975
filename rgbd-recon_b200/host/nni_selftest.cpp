// nni_selftest — device-free checks of kinect::NaturalNeighbourInterpolator: the properties that define
// Sibson coordinates, on scattered samples and on the fully degenerate regular grids calibration samples sit on.
// Exit code 0 = all passed. Run by tests/test_host_cpp.py.
#include <cmath>
#include <algorithm>
#include <cstdio>
#include <iostream>
#include <random>
#include <string>
#include <utility>
#include <vector>

#include "natural_neighbour_interpolator.hpp"

using kinect::nniSample;

static int g_failed = 0;
#define EXPECT(cond)                                                                            \
  do {                                                                                          \
    if (!(cond)) { std::cerr << "FAILED " << __LINE__ << ": " #cond << std::endl; ++g_failed; } \
  } while (0)

// an affine field: natural-neighbour interpolation reproduces it exactly (linear precision)
static void field(double x, double y, double z, float off[3], float tex[2]) {
  off[0] = (float)(0.3 * x - 0.2 * y + 0.5 * z + 0.1);
  off[1] = (float)(-0.7 * x + 0.1 * y + 0.2 * z - 0.3);
  off[2] = (float)(0.05 * x + 0.9 * y - 0.4 * z + 0.7);
  tex[0] = (float)(0.6 * x + 0.3 * y - 0.1 * z);
  tex[1] = (float)(-0.2 * x + 0.8 * y + 0.25 * z + 0.05);
}
static nniSample sample_at(double x, double y, double z) {
  nniSample s{};
  s.s_pos = {(float)x, (float)y, (float)z};
  float off[3], tex[2];
  field(s.s_pos.x, s.s_pos.y, s.s_pos.z, off, tex);
  s.s_pos_off = {off[0], off[1], off[2]};
  s.s_tex_off = {tex[0], tex[1]};
  s.quality = 1.0f;
  return s;
}

static double check_queries(kinect::NaturalNeighbourInterpolator& nni, std::mt19937& rng, int n, double lo, double hi, const std::vector<nniSample>& samples) {
  std::uniform_real_distribution<double> U(lo, hi);
  double worst = 0.0;
  for (int i = 0; i < n; ++i) {
    nniSample q{};
    q.s_pos = {(float)U(rng), (float)U(rng), (float)U(rng)};
    const bool ok = nni.interpolate(q);
    EXPECT(ok);
    if (!ok) continue;
    float off[3], tex[2];
    field(q.s_pos.x, q.s_pos.y, q.s_pos.z, off, tex);
    const double errs[5] = {std::fabs(q.s_pos_off.x - off[0]), std::fabs(q.s_pos_off.y - off[1]), std::fabs(q.s_pos_off.z - off[2]),
                            std::fabs(q.s_tex_off.u - tex[0]), std::fabs(q.s_tex_off.v - tex[1])};
    const double e = *std::max_element(errs, errs + 5);
    worst = std::max(worst, e);
    // the coordinates: positive, and they reproduce the query position itself (local coordinates property)
    std::vector<std::pair<uint32_t, double>> c;
    double norm = 0.0;
    EXPECT(nni.coordinates(q.s_pos.x, q.s_pos.y, q.s_pos.z, c, norm));
    double px = 0, py = 0, pz = 0, sum = 0;
    for (auto& w : c) {
      EXPECT(w.second > 0.0);
      px += w.second * samples[w.first].s_pos.x; py += w.second * samples[w.first].s_pos.y; pz += w.second * samples[w.first].s_pos.z;
      sum += w.second;
    }
    EXPECT(std::fabs(sum - norm) <= 1e-12 * norm);
    EXPECT(std::fabs(px / norm - q.s_pos.x) < 1e-7 && std::fabs(py / norm - q.s_pos.y) < 1e-7 && std::fabs(pz / norm - q.s_pos.z) < 1e-7);
    EXPECT(c.size() >= 4);
  }
  return worst;
}

// --coords sites.bin queries.bin out.bin: float32 [n][3] sites and float32 [m][3] queries in, per query one row of float64 [n]
// normalised Sibson coordinates out (all zeros where the query has no natural neighbours). Used by tests/test_nni_cpu.py to put
// the coordinates next to an independent computation (Voronoi cell volumes from Qhull).
static int dump_coordinates(const char* sites_path, const char* queries_path, const char* out_path) {
  auto slurp = [](const char* path, std::vector<float>& v) {
    std::FILE* f = std::fopen(path, "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    const long bytes = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    v.resize((size_t)bytes / sizeof(float));
    const size_t got = std::fread(v.data(), sizeof(float), v.size(), f);
    std::fclose(f);
    return got == v.size();
  };
  std::vector<float> sites, queries;
  if (!slurp(sites_path, sites) || !slurp(queries_path, queries) || sites.size() % 3 || queries.size() % 3) {
    std::cerr << "nni_selftest --coords: cannot read the inputs" << std::endl;
    return 1;
  }
  std::vector<nniSample> s(sites.size() / 3);
  for (size_t i = 0; i < s.size(); ++i) { s[i] = nniSample{}; s[i].s_pos = {sites[3 * i], sites[3 * i + 1], sites[3 * i + 2]}; s[i].quality = 1.0f; }
  kinect::NaturalNeighbourInterpolator nni(s);
  std::FILE* o = std::fopen(out_path, "wb");
  if (!o) return 1;
  std::vector<double> row(s.size());
  for (size_t q = 0; q < queries.size() / 3; ++q) {
    std::fill(row.begin(), row.end(), 0.0);
    std::vector<std::pair<uint32_t, double>> c;
    double norm = 0.0;
    if (nni.coordinates(queries[3 * q], queries[3 * q + 1], queries[3 * q + 2], c, norm) && norm > 0.0)
      for (const auto& e : c) row[e.first] += e.second / norm;
    std::fwrite(row.data(), sizeof(double), row.size(), o);
  }
  std::fclose(o);
  return 0;
}

int main(int argc, char** argv) {
  if (argc == 5 && std::string(argv[1]) == "--coords") return dump_coordinates(argv[2], argv[3], argv[4]);
  std::mt19937 rng(12345);
  std::uniform_real_distribution<double> U(0.0, 1.0);

  // ---- scattered samples
  {
    std::vector<nniSample> s;
    for (int i = 0; i < 3000; ++i) s.push_back(sample_at(U(rng), U(rng), U(rng)));
    kinect::NaturalNeighbourInterpolator nni(s);
    const double worst = check_queries(nni, rng, 300, 0.2, 0.8, s);
    EXPECT(worst < 2e-6);
    std::printf("scattered: worst |error| of an affine field %.3g\n", worst);
    // at a sample: that sample
    nniSample q = s[17];
    q.s_pos_off = {0, 0, 0};
    EXPECT(nni.interpolate(q) && q.s_pos_off.x == s[17].s_pos_off.x && q.s_tex_off.v == s[17].s_tex_off.v);
    // outside the convex hull: no natural neighbours (NaturalNeighbourInterpolator.cpp:49-51 returns false)
    nniSample o{};
    o.s_pos = {1.5f, 0.5f, 0.5f};
    EXPECT(!nni.interpolate(o));
    o.s_pos = {0.5f, 0.5f, -0.01f};
    EXPECT(!nni.interpolate(o));
  }

  // ---- a regular grid: every Delaunay cell is degenerate (eight cospherical corners)
  {
    std::vector<nniSample> s;
    const int n = 10;
    for (int z = 0; z < n; ++z) for (int y = 0; y < n; ++y) for (int x = 0; x < n; ++x) s.push_back(sample_at(x / double(n - 1), y / double(n - 1), z / double(n - 1)));
    kinect::NaturalNeighbourInterpolator nni(s);
    const double worst = check_queries(nni, rng, 200, 0.15, 0.85, s);
    EXPECT(worst < 2e-6);
    std::printf("grid: worst |error| of an affine field %.3g\n", worst);
    // the centre of a grid cell: its eight corners, equal weights
    std::vector<std::pair<uint32_t, double>> c;
    double norm = 0.0;
    const double h = 1.0 / (n - 1);
    EXPECT(nni.coordinates(4.5 * h, 4.5 * h, 4.5 * h, c, norm));
    double wmin = 1e300, wmax = 0.0;
    int big = 0;
    for (auto& w : c) if (w.second / norm > 1e-6) { ++big; wmin = std::min(wmin, w.second / norm); wmax = std::max(wmax, w.second / norm); }
    EXPECT(big == 8 && std::fabs(wmin - 0.125) < 1e-6 && std::fabs(wmax - 0.125) < 1e-6);
    // a grid point itself
    nniSample q = s[555];
    q.s_tex_off = {0, 0};
    EXPECT(nni.interpolate(q) && q.s_tex_off.u == s[555].s_tex_off.u);
  }

  // ---- the definition, by Monte Carlo: weight_i = volume of the points that are nearest to q and whose nearest sample is i
  {
    std::vector<nniSample> s;
    for (int i = 0; i < 40; ++i) s.push_back(sample_at(U(rng), U(rng), U(rng)));
    kinect::NaturalNeighbourInterpolator nni(s);
    const double q[3] = {0.5, 0.45, 0.55};
    std::vector<std::pair<uint32_t, double>> c;
    double norm = 0.0;
    EXPECT(nni.coordinates(q[0], q[1], q[2], c, norm));
    std::vector<double> mc(s.size(), 0.0);
    double total = 0.0;
    std::uniform_real_distribution<double> B(-0.5, 0.5);
    const int trials = 400000;
    for (int t = 0; t < trials; ++t) {
      const double x = q[0] + B(rng), y = q[1] + B(rng), z = q[2] + B(rng);
      const double dq = (x - q[0]) * (x - q[0]) + (y - q[1]) * (y - q[1]) + (z - q[2]) * (z - q[2]);
      double best = 1e300;
      int bi = -1;
      for (size_t i = 0; i < s.size(); ++i) {
        const double d = (x - s[i].s_pos.x) * (x - s[i].s_pos.x) + (y - s[i].s_pos.y) * (y - s[i].s_pos.y) + (z - s[i].s_pos.z) * (z - s[i].s_pos.z);
        if (d < best) { best = d; bi = (int)i; }
      }
      if (dq < best) { mc[bi] += 1.0; total += 1.0; }
    }
    EXPECT(total > 2000);
    EXPECT(std::fabs(total / trials - norm) < 0.05 * norm);             // the box has volume 1: hit fraction = vol(V_q)
    double worst = 0.0;
    for (auto& w : c) worst = std::max(worst, std::fabs(w.second / norm - mc[w.first] / total));
    EXPECT(worst < 0.03);
    std::printf("monte carlo: cell volume %.5f vs %.5f, worst coordinate difference %.4f\n", norm, total / trials, worst);
  }

  if (g_failed) { std::cerr << g_failed << " check(s) failed" << std::endl; return 1; }
  std::cout << "nni_selftest ok" << std::endl;
  return 0;
}

// kinect::NaturalNeighbourInterpolator (framework/NaturalNeighbourInterpolator.h:15-56, .cpp:16-92) without CGAL.
//
// The reference wraps CGAL's Delaunay_triangulation_3 + sibson_natural_neighbor_coordinates_3: the interpolant at q is
// sum_i vol(V_q ∩ V_i) * sample_i / vol(V_q), V_q being q's Voronoi cell after inserting q among the samples and V_i the
// samples' cells before. It has no call sites in the reference; it is kept as host C++ because
// BASELINE.json's north_star names it. This version computes the same Sibson coordinates directly from their definition
// - the cell V_q by clipping a box with the bisector planes of the surrounding samples (found through a uniform grid, with
// the usual security radius 2 * max vertex distance), then for every face neighbour i the part of V_q whose nearest old
// sample is p_i - so no global triangulation and no exact predicates are needed: the coordinates are continuous in the
// sample positions, cospherical samples (calibration samples sit on regular grids) only produce zero-volume parts.
// fp64 throughout, like the reference's accumulation (NaturalNeighbourInterpolator.cpp:59-85).
#ifndef RR_NATURAL_NEIGHBOUR_INTERPOLATOR_HPP
#define RR_NATURAL_NEIGHBOUR_INTERPOLATOR_HPP

#include <cstdint>
#include <iosfwd>
#include <utility>
#include <vector>

#include "rr_host.hpp"      // kinect::xyz, kinect::uv (framework/DataTypes.h:12-34)

namespace kinect {

struct nniSample {       // NaturalNeighbourInterpolator.h:15-20
  xyz s_pos;
  xyz s_pos_off;
  uv s_tex_off;
  float quality;
};

std::ostream& operator<<(std::ostream& o, const nniSample& s);

class NaturalNeighbourInterpolator {
 public:
  explicit NaturalNeighbourInterpolator(const std::vector<nniSample>& samples);
  ~NaturalNeighbourInterpolator();

  // Fills ipolant.s_pos_off / s_tex_off from ipolant.s_pos. false if q has no natural neighbours, i.e. lies outside (or on
  // the boundary of) the convex hull of the samples, where its Voronoi cell is unbounded; ipolant is left untouched then.
  bool interpolate(nniSample& ipolant);

  // The coordinates themselves: (sample index, vol(V_q ∩ V_i)) pairs and their sum, as CGAL's
  // sibson_natural_neighbor_coordinates_3 reports them. Returns false like interpolate().
  bool coordinates(double qx, double qy, double qz, std::vector<std::pair<uint32_t, double>>& coords, double& norm) const;

  std::size_t size() const { return m_samples.size(); }

 private:
  void sitesWithin(const double q[3], double radius, std::vector<uint32_t>& out) const;
  std::vector<nniSample> m_samples;
  double m_min[3], m_max[3], m_cell;
  int m_dim[3];
  std::vector<uint32_t> m_cell_start, m_cell_items;   // grid buckets (counting sort)
};

}  // namespace kinect
#endif

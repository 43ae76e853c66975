// ORACLE BUILD STUB (test infrastructure): stand-in for the CGAL spatial-searching types that the reference's
// framework/calibration/nearest_neighbour_search.{hpp,cpp} instantiate. CGAL is a system dependency of the reference
// (utils/dependencies.txt:4, no version pinned) that is absent from /root/reference and from this image.
// Restated from CGAL's published contract for Orthogonal_k_neighbor_search with the default Euclidean_distance over
// an EPICK (double) kernel: the EXACT k nearest neighbours of the query, distances = squared Euclidean distances
// evaluated in double, reported in ascending distance. CGAL leaves the order of equidistant points unspecified;
// this stub breaks ties by insertion index. The search structure (a median-split kd-tree) is an implementation
// detail: results are the exact k-set.
#ifndef RR_REF_STUB_CGAL_KNN_H
#define RR_REF_STUB_CGAL_KNN_H
#include <algorithm>
#include <cstddef>
#include <tuple>
#include <utility>
#include <vector>

namespace boost {
template <class... T> using tuple = std::tuple<T...>;
using std::get;
using std::make_tuple;
template <class IterTuple>
struct zip_iterator {
  IterTuple its;
  auto operator*() const -> decltype(std::make_tuple(*std::get<0>(its), *std::get<1>(its))) { return std::make_tuple(*std::get<0>(its), *std::get<1>(its)); }
  zip_iterator& operator++() { ++std::get<0>(its); ++std::get<1>(its); return *this; }
  bool operator!=(zip_iterator const& o) const { return std::get<0>(its) != std::get<0>(o.its); }
};
template <class IterTuple>
zip_iterator<IterTuple> make_zip_iterator(IterTuple t) { return zip_iterator<IterTuple>{t}; }
}  // namespace boost

namespace CGAL {

struct Point_3_stub {
  double c[3];
  Point_3_stub() : c{0, 0, 0} {}
  Point_3_stub(double x, double y, double z) : c{x, y, z} {}
};

struct Exact_predicates_inexact_constructions_kernel { typedef Point_3_stub Point_3; };

template <class K> struct Search_traits_3 { typedef typename K::Point_3 Point_d; };
template <int N, class Tuple> struct Nth_of_tuple_property_map {};
template <class Value, class PMap, class Base> struct Search_traits_adapter { typedef Value Point_d; typedef typename Base::Point_d Query; };

template <class Traits>
class Orthogonal_k_neighbor_search {
 public:
  typedef typename Traits::Point_d Point_d;      // tuple<Point_3, int>
  typedef typename Traits::Query Query;          // Point_3
  struct Distance {};
  typedef std::pair<Point_d, double> Point_with_transformed_distance;
  typedef typename std::vector<Point_with_transformed_distance>::iterator iterator;

  class Tree {
   public:
    template <class It>
    Tree(It first, It last) {
      for (; first != last; ++first) m_pts.push_back(Point_d(*first));
      m_order.resize(m_pts.size());
      for (std::size_t i = 0; i < m_order.size(); ++i) m_order[i] = i;
      if (!m_pts.empty()) build(0, m_order.size());
    }
    std::vector<Point_d> m_pts;
    std::vector<std::size_t> m_order;             // permutation; nodes are implicit (median at the middle of a range)
    struct Node { std::size_t lo, hi; int axis; double split; };
    static const std::size_t kLeaf = 12;
    const double* coords(std::size_t i) const { return std::get<0>(m_pts[i]).c; }
    void build(std::size_t lo, std::size_t hi) {
      if (hi - lo <= kLeaf) return;
      double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
      for (std::size_t i = lo; i < hi; ++i) {
        const double* c = coords(m_order[i]);
        for (int a = 0; a < 3; ++a) { mn[a] = std::min(mn[a], c[a]); mx[a] = std::max(mx[a], c[a]); }
      }
      int axis = 0;
      for (int a = 1; a < 3; ++a) if (mx[a] - mn[a] > mx[axis] - mn[axis]) axis = a;
      const std::size_t mid = lo + (hi - lo) / 2;
      std::nth_element(m_order.begin() + lo, m_order.begin() + mid, m_order.begin() + hi,
                       [&](std::size_t a, std::size_t b) { return coords(a)[axis] < coords(b)[axis]; });
      m_axis.resize(std::max(m_axis.size(), mid + 1), -1);
      m_axis[mid] = axis;
      build(lo, mid);
      build(mid + 1, hi);
    }
    std::vector<int> m_axis;                      // split axis of the node whose median sits at this position
  };

  Orthogonal_k_neighbor_search(Tree const& tree, Query const& q, unsigned k) {
    m_k = k;
    if (!tree.m_pts.empty() && k > 0) descend(tree, q.c, 0, tree.m_order.size());
    std::sort(m_best.begin(), m_best.end(), less);
    for (auto const& b : m_best) m_result.push_back(Point_with_transformed_distance(tree.m_pts[b.second], b.first));
  }
  iterator begin() { return m_result.begin(); }
  iterator end() { return m_result.end(); }

 private:
  typedef std::pair<double, std::size_t> Cand;   // (squared distance, insertion index)
  static bool less(Cand const& a, Cand const& b) { return a.first < b.first || (a.first == b.first && a.second < b.second); }
  void offer(Tree const& t, const double* q, std::size_t idx) {
    const double* c = t.coords(idx);
    const double dx = q[0] - c[0], dy = q[1] - c[1], dz = q[2] - c[2];
    Cand cand((dx * dx + dy * dy) + dz * dz, idx);
    if (m_best.size() < m_k) {
      m_best.push_back(cand);
      std::push_heap(m_best.begin(), m_best.end(), less);
    } else if (less(cand, m_best.front())) {
      std::pop_heap(m_best.begin(), m_best.end(), less);
      m_best.back() = cand;
      std::push_heap(m_best.begin(), m_best.end(), less);
    }
  }
  void descend(Tree const& t, const double* q, std::size_t lo, std::size_t hi) {
    if (hi - lo <= Tree::kLeaf) {
      for (std::size_t i = lo; i < hi; ++i) offer(t, q, t.m_order[i]);
      return;
    }
    const std::size_t mid = lo + (hi - lo) / 2;
    const int axis = t.m_axis[mid];
    const double split = t.coords(t.m_order[mid])[axis];
    const double d = q[axis] - split;
    offer(t, q, t.m_order[mid]);
    if (d < 0.0) {
      descend(t, q, lo, mid);
      if (m_best.size() < m_k || d * d <= m_best.front().first) descend(t, q, mid + 1, hi);
    } else {
      descend(t, q, mid + 1, hi);
      if (m_best.size() < m_k || d * d <= m_best.front().first) descend(t, q, lo, mid);
    }
  }
  unsigned m_k;
  std::vector<Cand> m_best;                       // max-heap on (distance, index)
  std::vector<Point_with_transformed_distance> m_result;
};

}  // namespace CGAL
#endif

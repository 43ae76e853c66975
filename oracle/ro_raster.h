// ORACLE (test infrastructure, NOT product code): the fixed-function stages of OpenGL 4.4 between the last vertex-processing
// stage and the fragment shader for ONE triangle, in fp64 from the binary32 clip coordinates, shared by the oracle's
// restatement of ReconTrigrid::draw (ro_trigrid.cpp) and by the harness that runs the reference's own trigrid shaders
// (glsl_host/glsl_harness.cpp). rgbd-recon_b200/csrc/rr_trigrid.cu states the same arithmetic for the device.
//   * primitive clipping against the near and far planes (§13.5): Sutherland-Hodgman in clip space, new vertices carry the
//     barycentric coordinates of the original triangle; x / y clipping is left to the viewport scissor (same fragments);
//   * perspective divide and viewport transform, depth range [0, 1] (§13.6.1);
//   * rasterisation (§14.6.1): a fragment for every pixel whose centre lies inside the polygon (a fan of the clipped polygon).
//     Every edge function is evaluated with its end points in one canonical order (lexicographic in window x, y), so the two
//     triangles that share an edge compute bit-identical values of opposite sign: a pixel centre belongs to exactly one of them
//     (a centre exactly on an edge goes to the triangle on the edge's positive side, which is what the top-left rule is for);
//   * window z interpolated affinely (eq. 14.10), every other attribute perspective-correct (eq. 14.9): the callback receives
//     the pixel, z_w as binary32 and the three perspective-correct weights of the ORIGINAL triangle's vertices.
#pragma once
#include <cmath>

namespace ro {

struct RVert { double x, y, z, w; double b[3]; };

inline RVert rlerp(const RVert& A, const RVert& B, double t) {
  RVert o;
  o.x = A.x + (B.x - A.x) * t; o.y = A.y + (B.y - A.y) * t; o.z = A.z + (B.z - A.z) * t; o.w = A.w + (B.w - A.w) * t;
  for (int k = 0; k < 3; ++k) o.b[k] = A.b[k] + (B.b[k] - A.b[k]) * t;
  return o;
}

// near: z >= -w, far: z <= w. Returns the vertex count of the clipped polygon (0: nothing left), at most 5.
inline int clip_near_far(const float clip[3][4], RVert* poly) {
  RVert a[8], b[8];
  int n = 3;
  for (int k = 0; k < 3; ++k) {
    a[k].x = clip[k][0]; a[k].y = clip[k][1]; a[k].z = clip[k][2]; a[k].w = clip[k][3];
    a[k].b[0] = k == 0 ? 1.0 : 0.0; a[k].b[1] = k == 1 ? 1.0 : 0.0; a[k].b[2] = k == 2 ? 1.0 : 0.0;
  }
  for (int plane = 0; plane < 2; ++plane) {
    int m = 0;
    for (int i = 0; i < n; ++i) {
      const RVert& A = a[i]; const RVert& B = a[(i + 1) % n];
      const double da = plane == 0 ? A.z + A.w : A.w - A.z, db = plane == 0 ? B.z + B.w : B.w - B.z;
      const bool ia = da >= 0.0, ib = db >= 0.0;
      if (ia) b[m++] = A;
      if (ia != ib) b[m++] = rlerp(A, B, da / (da - db));
    }
    n = m;
    for (int i = 0; i < n; ++i) a[i] = b[i];
    if (n < 3) return 0;
  }
  for (int i = 0; i < n; ++i) poly[i] = a[i];
  return n;
}

// the edge A -> B evaluated at P with the end points in canonical order; `flip` tells whether they were exchanged
inline double edge_canon(double ax, double ay, double bx, double by, double px, double py, bool& flip) {
  flip = (bx < ax) || (bx == ax && by < ay);
  if (flip) { double t = ax; ax = bx; bx = t; t = ay; ay = by; by = t; }
  return (bx - ax) * (py - ay) - (by - ay) * (px - ax);
}
inline bool edge_inside(double ax, double ay, double bx, double by, double px, double py, double s) {
  bool flip;
  const double e = edge_canon(ax, ay, bx, by, px, py, flip);
  const double sigma = flip ? -s : s;
  return sigma > 0.0 ? e >= 0.0 : e < 0.0;
}

// frag(px, py, zw, B): B = perspective-correct weights of the original triangle's three vertices
template <typename Frag>
inline void raster_triangle(const float clip[3][4], int vw, int vh, Frag frag) {
  RVert poly[8];
  const int n = clip_near_far(clip, poly);
  if (n < 3) return;
  double wx[8], wy[8], wz[8], iw[8];
  for (int i = 0; i < n; ++i) {
    if (!(poly[i].w > 0.0)) return;
    wx[i] = (poly[i].x / poly[i].w + 1.0) * 0.5 * (double)vw;
    wy[i] = (poly[i].y / poly[i].w + 1.0) * 0.5 * (double)vh;
    wz[i] = (poly[i].z / poly[i].w + 1.0) * 0.5;
    iw[i] = 1.0 / poly[i].w;
    if (!std::isfinite(wx[i]) || !std::isfinite(wy[i]) || !std::isfinite(wz[i])) return;
  }
  for (int f = 1; f + 1 < n; ++f) {
    const int i0 = 0, i1 = f, i2 = f + 1;
    const double x0 = wx[i0], y0 = wy[i0], x1 = wx[i1], y1 = wy[i1], x2 = wx[i2], y2 = wy[i2];
    const double den = (x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0);
    if (den == 0.0 || !std::isfinite(den)) continue;
    const double s = den > 0.0 ? 1.0 : -1.0;
    double lox = std::fmin(x0, std::fmin(x1, x2)), hix = std::fmax(x0, std::fmax(x1, x2));
    double loy = std::fmin(y0, std::fmin(y1, y2)), hiy = std::fmax(y0, std::fmax(y1, y2));
    if (hix < 0.0 || hiy < 0.0 || lox > (double)vw || loy > (double)vh) continue;
    lox = std::fmax(lox, 0.0); loy = std::fmax(loy, 0.0); hix = std::fmin(hix, (double)vw); hiy = std::fmin(hiy, (double)vh);
    int px0 = (int)std::floor(lox - 0.5), px1 = (int)std::ceil(hix - 0.5), py0 = (int)std::floor(loy - 0.5), py1 = (int)std::ceil(hiy - 0.5);
    if (px0 < 0) px0 = 0;
    if (py0 < 0) py0 = 0;
    if (px1 > vw - 1) px1 = vw - 1;
    if (py1 > vh - 1) py1 = vh - 1;
    for (int py = py0; py <= py1; ++py)
      for (int px = px0; px <= px1; ++px) {
        const double cx = (double)px + 0.5, cy = (double)py + 0.5;
        if (!edge_inside(x0, y0, x1, y1, cx, cy, s) || !edge_inside(x1, y1, x2, y2, cx, cy, s) || !edge_inside(x2, y2, x0, y0, cx, cy, s)) continue;
        const double b1 = ((cx - x0) * (y2 - y0) - (x2 - x0) * (cy - y0)) / den;
        const double b2 = ((x1 - x0) * (cy - y0) - (cx - x0) * (y1 - y0)) / den;
        const double b0 = (1.0 - b1) - b2;
        float zw = (float)((b0 * wz[i0] + b1 * wz[i1]) + b2 * wz[i2]);
        if (!(zw > 0.0f)) zw = 0.0f;                                 // the depth range is [0, 1]
        if (zw > 1.0f) zw = 1.0f;
        const double q0 = b0 * iw[i0], q1 = b1 * iw[i1], q2 = b2 * iw[i2];
        const double qs = (q0 + q1) + q2;
        const double p0 = q0 / qs, p1 = q1 / qs, p2 = q2 / qs;
        double B[3];
        for (int k = 0; k < 3; ++k) B[k] = (p0 * poly[i0].b[k] + p1 * poly[i1].b[k]) + p2 * poly[i2].b[k];
        frag(px, py, zw, B);
      }
  }
}

// an attribute at the fragment: ((B0 a0 + B1 a1) + B2 a2) in fp64, rounded to binary32
inline float rinterp(const double* B, float a0, float a1, float a2) { return (float)((B[0] * (double)a0 + B[1] * (double)a1) + B[2] * (double)a2); }

}  // namespace ro

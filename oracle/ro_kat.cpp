// ORACLE (test infrastructure, NOT product code): thread control + scalar known-answer hooks for the math pins.
#include "ro_math.h"
#include "rr_oracle.h"
#include <omp.h>
using namespace ro;
extern "C" {
void ro_set_threads(int n) { omp_set_num_threads(n > 0 ? n : 1); }
int ro_get_max_threads(void) { return omp_get_max_threads(); }
float ro_kat_log2(float x) { return det_log2(x); }
float ro_kat_exp2(float x) { return det_exp2(x); }
float ro_kat_pow(float x, float y) { return gl_pow(x, y); }
void ro_kat_tex3d(const float* vol, int C, int X, int Y, int Z, float s, float t, float r, float* out) {
  if (C == 2) tex3d_linear<2>(vol, X, Y, Z, s, t, r, out, 2);
  else if (C == 3) tex3d_linear<3>(vol, X, Y, Z, s, t, r, out, 3);
  else tex3d_linear<4>(vol, X, Y, Z, s, t, r, out, 4);
}
float ro_kat_tex2d(const float* img, int W, int H, float s, float t, int nearest) {
  return nearest ? tex2d_nearest(img, W, H, 1, 0, s, t) : tex2d_linear(img, W, H, 1, 0, s, t);
}
}

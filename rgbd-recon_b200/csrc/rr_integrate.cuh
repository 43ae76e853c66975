// Shared pieces of the two TSDF integrators (rr_integrate.cu: direct global gathers; rr_integrate_staged.cu: TMA-staged
// operands in shared memory): parameter blocks, the per-voxel arithmetic of glsl/tsdf_integration.vs:23-59 as small
// device functions used by BOTH kernels (one source of truth for the bit-exact results), and the clear stream of the
// fused kernels.
#pragma once

#include "rr_context.h"
#include "rr_math.cuh"

#include <cuda_fp16.h>

namespace rr {

struct IntegrateParams {
  const float4* inv;      // [N][IZ][IY][IX]
  const float4* gather;   // [N][H+1][W+1][2]   (direct kernels only)
  const float2* pairs;    // [N][H+2][pair_pitch] (depth_b.x, quality | silhouette sign), border replicated
  int pair_pitch;
  const float4* ztab;     // [Z]: per fine z the coarse plane pair and weight of the z filter tap: (k0, k1, g, 1-g)
  float* tsdf;
  float* weight;
  const int32_t* ranges;  // [num_bricks][6]
  const uint32_t* occupied;
  const uint32_t* num_occupied;
  int IX, IY, IZ, W, H, X, Y, Z;
  float fW, fH, exmax, eymax;   // (float)W, (float)H, (float)(W-1), (float)(H-1)
  int z_begin, z_end;     // slab
  int z_chunk;
  float limit;
  int wide_loads;         // tunable ldg256: gather texels with one 256-bit load
};

// clear stream of the fused kernels: row masks of the occupied bricks and the two work counters
struct FusedParams {
  IntegrateParams ip;
  const uint32_t* rowmask; const uint8_t* rowany; const int16_t* cand_y; const int16_t* cand_z;
  int mask_words, nby;
  uint32_t* work;              // [0] compute items handed out, [1] fill items handed out
  int max_cols, max_nz, zchunk, n_zchunks;
  int fill_rows; uint32_t fill_items; uint32_t row_begin, row_end;   // fill_rows <= 32
  int fill_warps;
  int chunk;                   // compute items a CTA draws from the global counter at a time
  float fill_value;
};

// x/y part of the trilinear inverse-volume fetch for one coarse plane: lerp(v0, v1, t) = fma(t, v1, (1 - t) * v0),
// x first, then y (the z lerp follows per voxel in tap_coords).
__device__ __forceinline__ float3 plane_reduce(const float4& p00, const float4& p10, const float4& p01, const float4& p11,
                                               float a, float oma, float b, float omb) {
  float3 r;
  r.x = fmaf(b, fmaf(a, p11.x, oma * p01.x), omb * fmaf(a, p10.x, oma * p00.x));
  r.y = fmaf(b, fmaf(a, p11.y, oma * p01.y), omb * fmaf(a, p10.y, oma * p00.y));
  r.z = fmaf(b, fmaf(a, p11.z, oma * p01.z), omb * fmaf(a, p10.z, oma * p00.z));
  return r;
}

// z lerp of the reduced planes -> pos_calib (u, v, d); bilinear footprint at (u, v): weights (wa, wb) and the index of
// its lower-left texel clamped to [-1, W-1] (fmaxf/fminf drop a NaN operand, so NaN -> -1). The footprint index the
// tables are addressed with is that + 1, in [0, W]; callers fold the + 1 into their base addresses where they can.
__device__ __forceinline__ void tap_coords(const float3& A, const float3& B, float g, float omg, float fW, float fH,
                                           float exmax, float eymax, float& wa, float& wb, float& d, int& exm1, int& eym1) {
  const float u = fmaf(g, B.x, omg * A.x), v = fmaf(g, B.y, omg * A.y);
  d = fmaf(g, B.z, omg * A.z);
  const float uu = u * fW - 0.5f, vv = v * fH - 0.5f;
  const float fu = floorf(uu), fv = floorf(vv);
  wa = uu - fu; wb = vv - fv;
  exm1 = (int)fminf(fmaxf(fu, -1.0f), exmax);
  eym1 = (int)fminf(fmaxf(fv, -1.0f), eymax);
}

// tsdf_integration.vs:30-55 for one sensor. d00..d11: depth_b.x at the footprint's four taps; q00..q11: quality at the same
// taps with the silhouette (exactly 0 or 1) in the sign bit.
__device__ __forceinline__ void fuse_tap(float wa, float wb, float d, float d00, float d10, float d01, float d11,
                                         float q00s, float q10s, float q01s, float q11s, float limit, float neg_limit,
                                         float& weighted_tsd, float& total_weight) {
  // silhouette < 1 ? The four taps are exactly 0 or 1. lerp(1,1,t) == 1 and lerp(0,0,t) == 0 exactly for every finite
  // t, so uniform footprints need no arithmetic; NaN weights compare false either way.
  if (weighted_tsd >= limit) {
    const uint32_t bx = __float_as_uint(q00s), by = __float_as_uint(q10s), bz = __float_as_uint(q01s), bw = __float_as_uint(q11s);
    const uint32_t all1 = (bx & by & bz & bw) >> 31, any1 = (bx | by | bz | bw) >> 31;
    bool sil_lt1;
    if (all1) {
      sil_lt1 = false;
    } else if (!any1) {
      sil_lt1 = (wa == wa) && (wb == wb);
    } else {
      const float s00 = (int)bx < 0 ? 1.0f : 0.0f, s10 = (int)by < 0 ? 1.0f : 0.0f;
      const float s01 = (int)bz < 0 ? 1.0f : 0.0f, s11 = (int)bw < 0 ? 1.0f : 0.0f;
      sil_lt1 = lerpf(lerpf(s00, s10, wa), lerpf(s01, s11, wa), wb) < 1.0f;
    }
    if (sil_lt1) { weighted_tsd = neg_limit; return; }
  }
  // NEAREST depth tap = upper tap of the footprint iff the bilinear weight is >= 0.5 (floor(t) == floor(t-0.5)+1);
  // where the subtraction t-0.5 can round (t < 0.5) both taps are the same clamped texel.
  const bool selx = wa >= 0.5f, sely = wb >= 0.5f;
  const float depth = sely ? (selx ? d11 : d01) : (selx ? d10 : d00);
  const float sdist = d - depth;
  if (sdist <= neg_limit) {
    weighted_tsd = neg_limit;
  } else if (sdist >= limit) {
  } else {
    const float w = lerpf(lerpf(fabsf(q00s), fabsf(q10s), wa), lerpf(fabsf(q01s), fabsf(q11s), wa), wb);
    weighted_tsd = (weighted_tsd * total_weight + w * sdist) / (total_weight + w);
    total_weight += w;
  }
}

template <int MODE>
__device__ __forceinline__ void store_voxel(const IntegrateParams& p, unsigned o, float weighted_tsd, float total_weight) {
  if (MODE == 2) {
    // half2 voxel: (tsdf, weight) rounded to nearest-even half, one 4-byte store
    const __half2 h = __floats2half2_rn(weighted_tsd, total_weight);
    reinterpret_cast<uint32_t*>(p.tsdf)[o] = *reinterpret_cast<const uint32_t*>(&h);
  } else {
    p.tsdf[o] = weighted_tsd;
    if (MODE == 1) p.weight[o] = total_weight;
  }
}

// One fill item: rows [row0, row1), at most 32. Lane r classifies row row0 + r (which brick rows cover it, does any of
// them hold an occupied brick); rows without occupied bricks are streamed with 16-byte stores, the others consult the
// row bitmask per 4-voxel group.
template <bool WEIGHT>
__device__ __forceinline__ void fill_rows(const FusedParams& p, uint32_t row0, uint32_t row1, int lane) {
  const int X = p.ip.X, Y = p.ip.Y;
  const float4 v4 = make_float4(p.fill_value, p.fill_value, p.fill_value, p.fill_value);
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool vec = (X & 3) == 0;
  int br0 = -1, br1 = -1, br2 = -1, br3 = -1;
  bool any = false;
  if (row0 + (uint32_t)lane < row1) {
    const uint32_t row = row0 + (uint32_t)lane;
    const int z = (int)(row / (uint32_t)Y), y = (int)(row - (uint32_t)z * (uint32_t)Y);
    const int cy0 = p.cand_y[2 * y], cy1 = p.cand_y[2 * y + 1], cz0 = p.cand_z[2 * z], cz1 = p.cand_z[2 * z + 1];
    br0 = (cy0 >= 0 && cz0 >= 0) ? cz0 * p.nby + cy0 : -1;
    br1 = (cy1 >= 0 && cz0 >= 0) ? cz0 * p.nby + cy1 : -1;
    br2 = (cy0 >= 0 && cz1 >= 0) ? cz1 * p.nby + cy0 : -1;
    br3 = (cy1 >= 0 && cz1 >= 0) ? cz1 * p.nby + cy1 : -1;
    any = (br0 >= 0 && p.rowany[br0]) || (br1 >= 0 && p.rowany[br1]) || (br2 >= 0 && p.rowany[br2]) || (br3 >= 0 && p.rowany[br3]);
  }
  uint32_t anymask = __ballot_sync(0xffffffffu, any);
  const int nrows = (int)(row1 - row0);
  if (nrows < 32) anymask |= ~0u << nrows;          // rows past the item count as "not clean": they end every run
  const int groups = (X + 127) >> 7;                // 128-voxel (32 x float4) groups per row
  for (int r = 0; r < nrows;) {
    if (!((anymask >> r) & 1u)) {
      // a run of rows without occupied bricks is one contiguous range of memory: stream it
      const int run = min(__ffs((int)(anymask >> r)) - 1, nrows - r);     // anymask has a set bit at or above nrows unless nrows == 32
      const int len = (anymask >> r) ? run : nrows - r;
      float* t0 = p.ip.tsdf + (size_t)(row0 + (uint32_t)r) * X;
      float* w0 = WEIGHT ? p.ip.weight + (size_t)(row0 + (uint32_t)r) * X : nullptr;
      const int n = len * X;
      if (vec) {
        float4* t4 = reinterpret_cast<float4*>(t0) + lane;
        float4* w4 = WEIGHT ? reinterpret_cast<float4*>(w0) + lane : nullptr;
        const int n4 = n >> 2;
        int i = lane;
        for (; i + 96 < n4; i += 128, t4 += 128) {
          __stcs(t4, v4); __stcs(t4 + 32, v4); __stcs(t4 + 64, v4); __stcs(t4 + 96, v4);
          if (WEIGHT) { __stcs(w4, z4); __stcs(w4 + 32, z4); __stcs(w4 + 64, z4); __stcs(w4 + 96, z4); w4 += 128; }
        }
        for (; i < n4; i += 32, t4 += 32) {
          __stcs(t4, v4);
          if (WEIGHT) { __stcs(w4, z4); w4 += 32; }
        }
      } else {
        for (int x = lane; x < n; x += 32) { t0[x] = p.fill_value; if (WEIGHT) w0[x] = 0.0f; }
      }
      r += len;
      continue;
    }
    float* trow = p.ip.tsdf + (size_t)(row0 + (uint32_t)r) * X;
    float* wrow = WEIGHT ? p.ip.weight + (size_t)(row0 + (uint32_t)r) * X : nullptr;
    const int b0 = __shfl_sync(0xffffffffu, br0, r), b1 = __shfl_sync(0xffffffffu, br1, r);
    const int b2 = __shfl_sync(0xffffffffu, br2, r), b3 = __shfl_sync(0xffffffffu, br3, r);
    ++r;
    for (int chunk = 0; chunk * 1024 < X; ++chunk) {
      uint32_t comb = 0;
      const int w = chunk * 32 + lane;
      if (w < p.mask_words) {
        if (b0 >= 0) comb |= __ldg(p.rowmask + (size_t)b0 * p.mask_words + w);
        if (b1 >= 0) comb |= __ldg(p.rowmask + (size_t)b1 * p.mask_words + w);
        if (b2 >= 0) comb |= __ldg(p.rowmask + (size_t)b2 * p.mask_words + w);
        if (b3 >= 0) comb |= __ldg(p.rowmask + (size_t)b3 * p.mask_words + w);
      }
      const int xbase = chunk * 1024;
      if (vec) {
        const int jn = min(8, groups - chunk * 8);
        for (int j = 0; j < jn; ++j) {
          const uint32_t word = __shfl_sync(0xffffffffu, comb, j * 4 + (lane >> 3));
          const int x = xbase + (j * 32 + lane) * 4;
          if (x >= X) continue;
          const uint32_t nib = (word >> ((lane & 7) * 4)) & 15u;
          if (nib == 0) {
            __stcs(reinterpret_cast<float4*>(trow + x), v4);
            if (WEIGHT) __stcs(reinterpret_cast<float4*>(wrow + x), z4);
          } else if (nib != 15u) {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (!((nib >> e) & 1u)) { trow[x + e] = p.fill_value; if (WEIGHT) wrow[x + e] = 0.0f; }
          }
        }
      } else {
        for (int i = 0; i < 32; ++i) {
          const uint32_t word = __shfl_sync(0xffffffffu, comb, i);
          const int x = xbase + i * 32 + lane;
          if (x >= X) continue;
          if (!((word >> lane) & 1u)) { trow[x] = p.fill_value; if (WEIGHT) wrow[x] = 0.0f; }
        }
      }
    }
  }
}

// Draw fill items from the global counter until the slab's rows are exhausted (whole warp).
template <bool WEIGHT>
__device__ __forceinline__ void fill_loop(const FusedParams& p, int lane) {
  for (;;) {
    unsigned it = 0;
    if (lane == 0) it = atomicAdd(p.work + 1, 1u);
    it = __shfl_sync(0xffffffffu, it, 0);
    if (it >= p.fill_items) break;
    const uint32_t r0 = p.row_begin + it * (uint32_t)p.fill_rows;
    fill_rows<WEIGHT>(p, r0, min(r0 + (uint32_t)p.fill_rows, p.row_end), lane);
  }
}

// The cleared voxel as the 4 bytes the fill stores write: -limit (R32F), or half2(-limit, 0) for half2 voxels.
float cleared_voxel(int mode, float limit);
// fills the clear-stream half of FusedParams for the current slab (row range, item count, masks)
void setup_fill(const rr_ctx* c, const IntegrateParams& p, int mode, int fill_rows, FusedParams& f);
// k_integrate_bricks over the occupied bricks flagged in the per-brick mask `only` (nullptr: all of them)
int launch_bricks_masked(rr_ctx* c, const IntegrateParams& p, int mode, const uint8_t* only);
// staged (TMA) integrator: returns RR_OK and sets *done = true when it handled the launch
int launch_integrate_staged(rr_ctx* c, const IntegrateParams& p, int mode, bool* done);

}  // namespace rr

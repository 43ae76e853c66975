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
  unsigned plane_elems;   // X * Y
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

// One 32-byte gather texel with a single 256-bit load (LDG.E.256, sm_100+): half the load instructions and L1 requests
// of two LDG.128 on the same sector.
__device__ __forceinline__ void ldg_texel(const float4* p, float4& lo, float4& hi) {
  asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w)
               : "l"(p));
}

// One sensor's lookup for one voxel: bilinear weights, interpolated depth coordinate and the 32-byte gather texel.
struct GatherTap {
  float wa, wb, d;
  float4 lo, hi;
};

// One (x, y) column, z in [zb, ze). All index arithmetic is 32-bit (sizes are validated on the host).
// z (hence the coarse plane pair) is uniform across a warp wherever the callers keep a warp inside one brick / one
// dense tile, so the plane-advance branches below do not diverge.
template <int N, int MODE, bool PAIRS = false>
__device__ __forceinline__ void march_column(const IntegrateParams& p, int x, int y, int zb, int ze) {
  const float stepX = 1.0f / (float)p.X, stepY = 1.0f / (float)p.Y;
  const float px = ((float)x + 0.5f) * stepX, py = ((float)y + 0.5f) * stepY;
  int x0, x1, y0, y1; float a, b;
  lin_coord(px, p.IX, x0, x1, a);
  lin_coord(py, p.IY, y0, y1, b);
  const float oma = 1.0f - a, omb = 1.0f - b;            // lerp(v0, v1, t) = fma(t, v1, (1 - t) * v0)
  const unsigned o00 = y0 * p.IX + x0, o10 = y0 * p.IX + x1, o01 = y1 * p.IX + x0, o11 = y1 * p.IX + x1;
  const unsigned plane_sz = (unsigned)(p.IX * p.IY);
  const unsigned gstride = (unsigned)((p.W + 1) * (p.H + 1) * 2);
  const unsigned grow = (unsigned)(p.W + 1);
  const float limit = p.limit, neg_limit = -p.limit;
  float3 A[N], B[N];
  int ck0 = -1, ck1 = -1;

  auto plane = [&](int s, int k) -> float3 {
    const float4* base = p.inv + (unsigned)(s * p.IZ + k) * plane_sz;
    const float4 p00 = __ldg(base + o00), p10 = __ldg(base + o10), p01 = __ldg(base + o01), p11 = __ldg(base + o11);
    return plane_reduce(p00, p10, p01, p11, a, oma, b, omb);
  };

  unsigned o = (unsigned)((zb * p.Y + y) * p.X + x);
  const unsigned ostep = (unsigned)(p.X * p.Y);
  for (int z = zb; z < ze; ++z, o += ostep) {
    const float4 zt = __ldg(p.ztab + z);
    const int k0 = __float_as_int(zt.x), k1 = __float_as_int(zt.y);
    const float g = zt.z, omg = zt.w;
    if (k0 != ck0) {
      if (k0 == ck1) {
#pragma unroll
        for (int s = 0; s < N; ++s) A[s] = B[s];
      } else {
#pragma unroll
        for (int s = 0; s < N; ++s) A[s] = plane(s, k0);
      }
      ck0 = k0;
    }
    if (k1 != ck1) {
      if (k1 == k0) {
#pragma unroll
        for (int s = 0; s < N; ++s) B[s] = A[s];
      } else {
#pragma unroll
        for (int s = 0; s < N; ++s) B[s] = plane(s, k1);
      }
      ck1 = k1;
    }
    float weighted_tsd = limit, total_weight = 0.0f;

    // z filter tap, bilinear footprint (silhouette, quality) at (u, v) and the gather loads for sensor s
    auto fetch = [&](int s) -> GatherTap {
      GatherTap t;
      int ex, ey;
      tap_coords(A[s], B[s], g, omg, p.fW, p.fH, p.exmax, p.eymax, t.wa, t.wb, t.d, ex, ey);
      ex += 1; ey += 1;
      if (PAIRS) {
        // the pair image the staged integrator tiles (same taps, 8 bytes per pixel): footprint (ex, ey) = pixels ex, ex+1 of rows ey, ey+1
        const float2* q = p.pairs + ((unsigned)(s * (p.H + 2) + ey) * (unsigned)p.pair_pitch + (unsigned)ex);
        const float2 t00 = __ldg(q), t10 = __ldg(q + 1), t01 = __ldg(q + p.pair_pitch), t11 = __ldg(q + p.pair_pitch + 1);
        t.lo = make_float4(t00.x, t10.x, t01.x, t11.x);
        t.hi = make_float4(t00.y, t10.y, t01.y, t11.y);
        return t;
      }
      const float4* g4 = p.gather + ((unsigned)s * gstride + ((unsigned)ey * grow + (unsigned)ex) * 2u);
      if (p.wide_loads) ldg_texel(g4, t.lo, t.hi); else { t.lo = __ldg(g4); t.hi = __ldg(g4 + 1); }
      return t;
    };
    auto fuse = [&](const GatherTap& t) {
      fuse_tap(t.wa, t.wb, t.d, t.lo.x, t.lo.y, t.lo.z, t.lo.w, t.hi.x, t.hi.y, t.hi.z, t.hi.w, limit, neg_limit, weighted_tsd, total_weight);
    };
    // two sensors' gathers are in flight before the first decision chain runs
#pragma unroll
    for (int s = 0; s + 1 < N; s += 2) {
      const GatherTap t0 = fetch(s), t1 = fetch(s + 1);
      fuse(t0);
      fuse(t1);
    }
    if (N & 1) { const GatherTap t = fetch(N - 1); fuse(t); }
    store_voxel<MODE>(p, o, weighted_tsd, total_weight);
  }
}

// The small per-axis tables the clear consults per row (cand_y, cand_z, rowany): global by default, the staged kernel
// keeps a copy in shared memory so that classifying a fill item costs no global round trips.
struct FillTables { const int16_t* cand_y; const int16_t* cand_z; const uint8_t* rowany; };

// Lane r of a fill item classifies row row0 + r: the (at most four) brick rows that cover it and whether any of them holds
// an occupied brick.
__device__ __forceinline__ bool classify_row(const FusedParams& p, const FillTables& ft, uint32_t row0, uint32_t row1, int lane,
                                             int& br0, int& br1, int& br2, int& br3) {
  br0 = br1 = br2 = br3 = -1;
  if (row0 + (uint32_t)lane >= row1) return false;
  const uint32_t row = row0 + (uint32_t)lane;
  const int Y = p.ip.Y;
  const int z = (int)(row / (uint32_t)Y), y = (int)(row - (uint32_t)z * (uint32_t)Y);
  const int cy0 = ft.cand_y[2 * y], cy1 = ft.cand_y[2 * y + 1], cz0 = ft.cand_z[2 * z], cz1 = ft.cand_z[2 * z + 1];
  br0 = (cy0 >= 0 && cz0 >= 0) ? cz0 * p.nby + cy0 : -1;
  br1 = (cy1 >= 0 && cz0 >= 0) ? cz0 * p.nby + cy1 : -1;
  br2 = (cy0 >= 0 && cz1 >= 0) ? cz1 * p.nby + cy0 : -1;
  br3 = (cy1 >= 0 && cz1 >= 0) ? cz1 * p.nby + cy1 : -1;
  return (br0 >= 0 && ft.rowany[br0]) || (br1 >= 0 && ft.rowany[br1]) || (br2 >= 0 && ft.rowany[br2]) || (br3 >= 0 && ft.rowany[br3]);
}

// x mask of a row covered by brick rows b0..b3 (bit set = voxel inside an occupied brick), word `w` of it
__device__ __forceinline__ uint32_t row_mask_word(const FusedParams& p, int b0, int b1, int b2, int b3, int w) {
  uint32_t comb = 0;
  if (b0 >= 0) comb |= __ldg(p.rowmask + (size_t)b0 * p.mask_words + w);
  if (b1 >= 0) comb |= __ldg(p.rowmask + (size_t)b1 * p.mask_words + w);
  if (b2 >= 0) comb |= __ldg(p.rowmask + (size_t)b2 * p.mask_words + w);
  if (b3 >= 0) comb |= __ldg(p.rowmask + (size_t)b3 * p.mask_words + w);
  return comb;
}

// One fill item with per-lane stores (k_integrate_fused): rows [row0, row1), at most 32. Runs of rows without occupied
// bricks are streamed with 16-byte stores, the others consult the row bitmask per 4-voxel group.
template <bool WEIGHT>
__device__ __forceinline__ void fill_rows(const FusedParams& p, const FillTables& ft, uint32_t row0, uint32_t row1, int lane) {
  const int X = p.ip.X;
  const float4 v4 = make_float4(p.fill_value, p.fill_value, p.fill_value, p.fill_value);
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool vec = (X & 3) == 0;
  int br0, br1, br2, br3;
  const bool any = classify_row(p, ft, row0, row1, lane, br0, br1, br2, br3);
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
    const int b0 = __shfl_sync(0xffffffffu, br0, r), b1 = __shfl_sync(0xffffffffu, br1, r);
    const int b2 = __shfl_sync(0xffffffffu, br2, r), b3 = __shfl_sync(0xffffffffu, br3, r);
    float* trow = p.ip.tsdf + (size_t)(row0 + (uint32_t)r) * X;
    float* wrow = WEIGHT ? p.ip.weight + (size_t)(row0 + (uint32_t)r) * X : nullptr;
    ++r;
    for (int chunk = 0; chunk * 1024 < X; ++chunk) {
      const int w = chunk * 32 + lane;
      const uint32_t comb = w < p.mask_words ? row_mask_word(p, b0, b1, b2, b3, w) : 0u;
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

// ---- the clear by TMA bulk stores (staged integrator) ---------------------------------------------------------------
// cp.async.bulk shared -> global: one instruction moves up to a whole buffer of cleared voxels, asynchronously, so the
// issuing warp is not held by the store queue the way per-lane stores hold it (a warp streams only ~4 B/clk of those).
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src_smem, uint32_t bytes, uint64_t policy) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(src_smem), "r"(bytes), "l"(policy) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Source of the bulk stores: a shared-memory buffer holding `bytes` of cleared voxels (and, for a separate weight volume,
// as many zero bytes behind it); policy: their L2 cache hint (evict_first); drop: measurement hook, nothing is stored.
struct FillSource { uint32_t src, bytes; bool drop; uint64_t policy; int depth; bool lsu; };
// Bound on the bulk-store groups a lane leaves pending (tunable stage_fill_depth; < 0: none): the TMA unit serves the
// integrator's tile loads and the clear's stores from one queue, so a deep backlog of stores delays every load behind it.
__device__ __forceinline__ void bulk_throttle(int depth) {
  if (depth < 0) return;
  if (depth == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  else if (depth == 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
  else if (depth == 2) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
  else asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
}

// One fill item, rows [row0, row1), at most 32; requires X % 4 == 0, X <= 1024 and a buffer of at least one row.
// A run of rows without occupied bricks is contiguous, 16-byte aligned memory: one bulk store per buffer-sized piece.
// Consecutive rows that cross occupied bricks and are covered by the same brick rows share one x mask: its runs of
// voxels to clear are found once (lane k keeps run k) and every row of the group then costs one bulk store per run for
// the 16-byte aligned interior plus up to six scalar stores at its ragged ends.
template <bool WEIGHT>
__device__ __forceinline__ void fill_rows_bulk(const FusedParams& p, const FillTables& ft, uint32_t row0, uint32_t row1, int lane, const FillSource& fs) {
  const int X = p.ip.X;
  int br0, br1, br2, br3;
  const bool any = classify_row(p, ft, row0, row1, lane, br0, br1, br2, br3);
  uint32_t anymask = __ballot_sync(0xffffffffu, any);
  const int nrows = (int)(row1 - row0);
  if (nrows < 32) anymask |= ~0u << nrows;          // rows past the item count as "not clean": they end every run
  for (int r = 0; r < nrows;) {
    if (!((anymask >> r) & 1u)) {
      const int run = min(__ffs((int)(anymask >> r)) - 1, nrows - r);
      const int len = (anymask >> r) ? run : nrows - r;
      uint8_t* t0 = reinterpret_cast<uint8_t*>(p.ip.tsdf + (size_t)(row0 + (uint32_t)r) * X);
      uint8_t* w0 = WEIGHT ? reinterpret_cast<uint8_t*>(p.ip.weight + (size_t)(row0 + (uint32_t)r) * X) : nullptr;
      const uint32_t nbytes = (uint32_t)(len * X) * 4u;
      if (fs.lsu) {
        // per-lane streaming stores (experiment: keeps the clear out of the TMA unit's queue)
        const float4 v4 = make_float4(p.fill_value, p.fill_value, p.fill_value, p.fill_value), z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (uint32_t off = (uint32_t)lane * 16u; off < nbytes; off += 512u) {
          __stcs(reinterpret_cast<float4*>(t0 + off), v4);
          if (WEIGHT) __stcs(reinterpret_cast<float4*>(w0 + off), z4);
        }
        r += len;
        continue;
      }
      for (uint32_t off = (uint32_t)lane * fs.bytes; off < nbytes; off += 32u * fs.bytes) {
        if (fs.drop) continue;
        const uint32_t sz = min(fs.bytes, nbytes - off);
        bulk_store(t0 + off, fs.src, sz, fs.policy);
        if (WEIGHT) bulk_store(w0 + off, fs.src + fs.bytes, sz, fs.policy);
      }
      bulk_commit();
      bulk_throttle(fs.depth);
      r += len;
      continue;
    }
    const int b0 = __shfl_sync(0xffffffffu, br0, r), b1 = __shfl_sync(0xffffffffu, br1, r);
    const int b2 = __shfl_sync(0xffffffffu, br2, r), b3 = __shfl_sync(0xffffffffu, br3, r);
    const uint32_t same = __ballot_sync(0xffffffffu, any && br0 == b0 && br1 == b1 && br2 == b2 && br3 == b3) >> r;
    const int glen = min((same == 0xffffffffu) ? 32 : __ffs((int)~same) - 1, nrows - r);
    uint32_t comb = 0xffffffffu;                         // words past the row count as occupied
    if (lane < p.mask_words) {
      comb = row_mask_word(p, b0, b1, b2, b3, lane);
      if (lane * 32 + 32 > X) comb |= ~0u << (X - lane * 32);      // bits past the last voxel of the row
    }
    // transitions: a run to clear starts where a 0 follows a 1 (or the row begins), and ends where a 1 follows a 0
    uint32_t prev = __shfl_up_sync(0xffffffffu, comb >> 31, 1);
    if (lane == 0) prev = 1u;
    const uint32_t t = comb ^ ((comb << 1) | prev);
    uint32_t starts = t & ~comb, ends = t & comb;
    int ra = 0, rb = 0, nruns = 0;
    for (;;) {
      const uint32_t bs = __ballot_sync(0xffffffffu, starts != 0u);
      if (bs == 0u) break;
      const int ls = __ffs((int)bs) - 1;
      const uint32_t sw = __shfl_sync(0xffffffffu, starts, ls);
      const int s_pos = ls * 32 + __ffs((int)sw) - 1;
      if (lane == ls) starts &= starts - 1u;
      const uint32_t be = __ballot_sync(0xffffffffu, ends != 0u);
      int e_pos = X;                                     // a run that reaches a 1024-voxel row's end has no closing transition
      if (be != 0u) {
        const int le = __ffs((int)be) - 1;
        const uint32_t ew = __shfl_sync(0xffffffffu, ends, le);
        e_pos = le * 32 + __ffs((int)ew) - 1;
        if (lane == le) ends &= ends - 1u;
      }
      if (nruns >= 32) {
        // more runs than lanes (a row alternating faster than 32 bricks): the rest with per-lane scalar stores
        for (int g = 0; g < glen; ++g) {
          float* trow = p.ip.tsdf + (size_t)(row0 + (uint32_t)(r + g)) * X;
          float* wrow = WEIGHT ? p.ip.weight + (size_t)(row0 + (uint32_t)(r + g)) * X : nullptr;
          for (int x = s_pos + lane; x < e_pos; x += 32) { trow[x] = p.fill_value; if (WEIGHT) wrow[x] = 0.0f; }
        }
      } else if (lane == nruns) {
        ra = s_pos; rb = e_pos;
      }
      ++nruns;
    }
    const int a4 = min((ra + 3) & ~3, rb), b4 = max(rb & ~3, a4);       // aligned interior [a4, b4) of this lane's run
    for (int g = 0; g < glen; ++g) {
      float* trow = p.ip.tsdf + (size_t)(row0 + (uint32_t)(r + g)) * X;
      float* wrow = WEIGHT ? p.ip.weight + (size_t)(row0 + (uint32_t)(r + g)) * X : nullptr;
      if (lane < nruns) {
        if (b4 > a4 && !fs.drop) {
          bulk_store(trow + a4, fs.src, (uint32_t)(b4 - a4) * 4u, fs.policy);
          if (WEIGHT) bulk_store(wrow + a4, fs.src + fs.bytes, (uint32_t)(b4 - a4) * 4u, fs.policy);
        }
        for (int x = ra; x < a4; ++x) { trow[x] = p.fill_value; if (WEIGHT) wrow[x] = 0.0f; }
        for (int x = b4; x < rb; ++x) { trow[x] = p.fill_value; if (WEIGHT) wrow[x] = 0.0f; }
      }
    }
    bulk_commit();
    bulk_throttle(fs.depth);
    r += glen;
  }
}

// Draw fill items from the global counter until the slab's rows are exhausted (whole warp).
template <bool WEIGHT>
__device__ __forceinline__ void fill_loop(const FusedParams& p, const FillTables& ft, int lane) {
  for (;;) {
    unsigned it = 0;
    if (lane == 0) it = atomicAdd(p.work + 1, 1u);
    it = __shfl_sync(0xffffffffu, it, 0);
    if (it >= p.fill_items) break;
    const uint32_t r0 = p.row_begin + it * (uint32_t)p.fill_rows;
    fill_rows<WEIGHT>(p, ft, r0, min(r0 + (uint32_t)p.fill_rows, p.row_end), lane);
  }
}
template <bool WEIGHT>
__device__ __forceinline__ void fill_loop_bulk(const FusedParams& p, const FillTables& ft, int lane, const FillSource& fs) {
  for (;;) {
    unsigned it = 0;
    if (lane == 0) it = atomicAdd(p.work + 1, 1u);
    it = __shfl_sync(0xffffffffu, it, 0);
    if (it >= p.fill_items) break;
    const uint32_t r0 = p.row_begin + it * (uint32_t)p.fill_rows;
    fill_rows_bulk<WEIGHT>(p, ft, r0, min(r0 + (uint32_t)p.fill_rows, p.row_end), lane, fs);
  }
  bulk_wait_all();      // the buffer must outlive the copies that read it
}

// ---- per-frame verdicts of the staged integrator ---------------------------------------------------------------------
// One warp per work item (brick x y-chunk x z-chunk) of an occupied brick: per sensor, scan the footprint rectangle of the
// item in the pair image (are all silhouette taps 1, what is the range of depth_b.x) and compare with the item's exact range
// of pos_calib.z (k_footprints): sdist = pos_calib.z - depth is monotone in both operands and rounding is monotone, so
//   zlo - dmax >= limit  =>  sdist >= limit for every voxel of the item   (tsdf_integration.vs:45: nothing happens)
//   zhi - dmin <= -limit =>  sdist <= -limit for every voxel              (:41: weighted_tsd = -limit)
// and with every silhouette tap 1 the silhouette test (:32) never fires. Such (item, sensor) pairs need no per-voxel work.
// Verdict = skip mask | front mask << 8. Non-finite values keep a sensor on the voxel-by-voxel path (NaN patterns unchanged).
// The item is appended, with its verdict, to this frame's work list of its cost class = number of sensors left to evaluate
// voxel by voxel (an item with an oversize footprint counts as the most expensive class): the integrator hands out the
// expensive items first, so the last ones to finish are the short ones.
struct ClassifyParams {
  const uint2* fp;          // [items][N]: tile origin tx0 | ty0 << 16, footprint rectangle x offset | oversize << 7 | width << 8 | height << 20
  const float2* zr;         // [items][N]: exact range of pos_calib.z over the item
  const float2* pairs; int pair_pitch, H2;
  uint2* list;              // [N + 1][list_stride] out: (item, verdict) by class
  uint32_t* class_count;    // [N + 1] in/out: entries per class (cleared with the brick counters)
  uint32_t list_stride;
  int per_brick, N;
  float limit;
};

__device__ __forceinline__ void classify_item(const ClassifyParams& q, uint32_t brick, uint32_t sub, int lane) {
  const size_t item = (size_t)brick * q.per_brick + sub;
  const float inf = __int_as_float(0x7f800000);
  uint32_t skip = 0, front = 0, oversize = 0, voxels = 0;
#pragma unroll 1
  for (int s = 0; s < q.N; ++s) {
    const uint2 f = q.fp[item * q.N + s];
    const float2 z = q.zr[item * q.N + s];
    const int rw = (int)((f.y >> 8) & 4095u), rh = (int)(f.y >> 20);
    if (rw * rh == 0) continue;
    voxels = 1;
    if (f.y & 128u) oversize |= 1u << s;
    if (!(z.x <= z.y)) continue;
    const float2* img = q.pairs + ((size_t)s * q.H2 + (f.x >> 16)) * q.pair_pitch + (f.x & 0xffffu) + (f.y & 127u);
    float dlo = inf, dhi = -inf;
    bool ok = true;
    for (int i = lane; i < rw * rh; i += 32) {
      const int ty = i / rw, tx = i - ty * rw;
      const float2 t = __ldg(img + (size_t)ty * q.pair_pitch + tx);
      ok = ok && ((int)__float_as_uint(t.y) < 0) && (fabsf(t.x) < inf);
      dlo = fminf(dlo, t.x); dhi = fmaxf(dhi, t.x);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      dlo = fminf(dlo, __shfl_xor_sync(0xffffffffu, dlo, d)); dhi = fmaxf(dhi, __shfl_xor_sync(0xffffffffu, dhi, d));
    }
    if (!__all_sync(0xffffffffu, ok)) continue;
    if (z.x - dhi >= q.limit) skip |= 1u << s;
    else if (z.y - dlo <= -q.limit) front |= 1u << s;
  }
  if (lane == 0 && voxels) {
    const uint32_t off = skip | front;
    const int cls = (oversize & ~off) ? q.N : q.N - __popc(off);
    const uint32_t idx = atomicAdd(q.class_count + cls, 1u);
    if (idx < q.list_stride) q.list[(size_t)cls * q.list_stride + idx] = make_uint2((uint32_t)item, skip | (front << 8));
  }
}

// The cleared voxel as the 4 bytes the fill stores write: -limit (R32F), or half2(-limit, 0) for half2 voxels.
float cleared_voxel(int mode, float limit);
// fills the clear-stream half of FusedParams for the current slab (row range, item count, masks)
void setup_fill(const rr_ctx* c, const IntegrateParams& p, int mode, int fill_rows, FusedParams& f);
// classification parameters of the current configuration; false when the staged integrator is not selected
bool staged_classify_params(const rr_ctx* c, ClassifyParams& q);
// staged (TMA) integrator: returns RR_OK and sets *done = true when it handled the launch
int launch_integrate_staged(rr_ctx* c, const IntegrateParams& p, int mode, bool* done);

}  // namespace rr

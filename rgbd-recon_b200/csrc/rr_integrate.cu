// TSDF integration for sm_100a: the per-voxel work of glsl/tsdf_integration.vs:23-59, driven like
// ReconIntegration::integrate (framework/reconstruction/recon_integration.cpp:243-270): clear to -limit, then either
// every voxel (dense) or the voxels of every occupied brick.
//
// Kernel shape (DESIGN.md "integrate"): one thread owns an (x, y) column of the volume and marches z. The x/y part of
// the trilinear inverse-calibration lookup (cv_xyz_inv, LINEAR filtering restated in fp32) is reduced once per coarse
// z-plane and kept in registers for all N sensors, so the 8-corner gather of the shader becomes 4 corner loads per
// coarse plane per sensor; consecutive lanes are consecutive x, so the R32F stores are 128-byte coalesced rows.
// Silhouette (bilinear), depth (nearest) and quality (bilinear) come from ONE 32-byte gather texel per voxel-sensor
// (see k_pack_gather). HBM-bound integer/float gather work: no tensor-core path applies.
#include "rr_context.h"
#include "rr_math.cuh"

#include <algorithm>
#include <cstdlib>

namespace rr {

struct IntegrateParams {
  const float4* inv;      // [N][IZ][IY][IX]
  const float4* gather;   // [N][H+1][W+1][2]
  float* tsdf;
  float* weight;
  const int32_t* ranges;  // [num_bricks][6]
  const uint32_t* occupied;
  const uint32_t* num_occupied;
  int IX, IY, IZ, W, H, X, Y, Z;
  int z_begin, z_end;     // slab
  int z_chunk;
  float limit;
};

// One (x, y) column, z in [zb, ze). All index arithmetic is 32-bit (sizes are validated on the host).
template <int N, bool WEIGHT>
__device__ __forceinline__ void march_column(const IntegrateParams& p, int x, int y, int zb, int ze) {
  const float stepX = 1.0f / (float)p.X, stepY = 1.0f / (float)p.Y, stepZ = 1.0f / (float)p.Z;
  const float px = ((float)x + 0.5f) * stepX, py = ((float)y + 0.5f) * stepY;
  int x0, x1, y0, y1; float a, b;
  lin_coord(px, p.IX, x0, x1, a);
  lin_coord(py, p.IY, y0, y1, b);
  const unsigned o00 = y0 * p.IX + x0, o10 = y0 * p.IX + x1, o01 = y1 * p.IX + x0, o11 = y1 * p.IX + x1;
  const unsigned plane_sz = (unsigned)(p.IX * p.IY);
  const unsigned gstride = (unsigned)((p.W + 1) * (p.H + 1) * 2);
  const unsigned grow = (unsigned)(p.W + 1);
  const float limit = p.limit;
  const float fW = (float)p.W, fH = (float)p.H, exmax = (float)(p.W - 1), eymax = (float)(p.H - 1);
  float3 A[N], B[N];
  int ck0 = -1, ck1 = -1;

  auto plane = [&](int s, int k) -> float3 {
    const float4* base = p.inv + (unsigned)(s * p.IZ + k) * plane_sz;
    const float4 p00 = __ldg(base + o00), p10 = __ldg(base + o10), p01 = __ldg(base + o01), p11 = __ldg(base + o11);
    float3 r;
    r.x = lerpf(lerpf(p00.x, p10.x, a), lerpf(p01.x, p11.x, a), b);
    r.y = lerpf(lerpf(p00.y, p10.y, a), lerpf(p01.y, p11.y, a), b);
    r.z = lerpf(lerpf(p00.z, p10.z, a), lerpf(p01.z, p11.z, a), b);
    return r;
  };

  unsigned o = (unsigned)((zb * p.Y + y) * p.X + x);
  const unsigned ostep = (unsigned)(p.X * p.Y);
  for (int z = zb; z < ze; ++z, o += ostep) {
    const float pz = ((float)z + 0.5f) * stepZ;
    int k0, k1; float g;
    lin_coord(pz, p.IZ, k0, k1, g);
    const bool needA = (k0 != ck0), a_from_b = needA && (k0 == ck1);
    const bool needB = (k1 != ck1), b_from_a = needB && (k1 == k0);
    float weighted_tsd = limit, total_weight = 0.0f;
#pragma unroll
    for (int s = 0; s < N; ++s) {
      if (needA) A[s] = a_from_b ? B[s] : plane(s, k0);
      if (needB) B[s] = b_from_a ? A[s] : plane(s, k1);
      const float u = lerpf(A[s].x, B[s].x, g), v = lerpf(A[s].y, B[s].y, g), d = lerpf(A[s].z, B[s].z, g);
      // Bilinear footprint (silhouette, quality) at (u, v): lower-left texel floor(u*W - 0.5), weights (wa, wb).
      const float uu = u * fW - 0.5f, vv = v * fH - 0.5f;
      const float fu = floorf(uu), fv = floorf(vv);
      const float wa = uu - fu, wb = vv - fv;
      // gather-texel index = clamp(footprint, -1, W-1) + 1; fmaxf/fminf drop a NaN operand, so NaN -> entry 0
      const int ex = (int)fminf(fmaxf(fu, -1.0f), exmax) + 1, ey = (int)fminf(fmaxf(fv, -1.0f), eymax) + 1;
      const float4* g4 = p.gather + ((unsigned)s * gstride + ((unsigned)ey * grow + (unsigned)ex) * 2u);
      const float4 lo = __ldg(g4), hi = __ldg(g4 + 1);
      // silhouette < 1 ? The four taps are exactly 0 or 1 (sign bits of hi). lerp(1,1,t) == 1 and lerp(0,0,t) == 0
      // exactly for every finite t, so uniform footprints need no arithmetic; NaN weights compare false either way.
      const uint32_t bx = __float_as_uint(hi.x), by = __float_as_uint(hi.y), bz = __float_as_uint(hi.z), bw = __float_as_uint(hi.w);
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
      if (sil_lt1 && weighted_tsd >= limit) { weighted_tsd = -limit; continue; }
      // NEAREST depth tap = upper tap of the footprint iff the bilinear weight is >= 0.5 (floor(t) == floor(t-0.5)+1);
      // where the subtraction t-0.5 can round (t < 0.5) both taps are the same clamped texel.
      const bool selx = wa >= 0.5f, sely = wb >= 0.5f;
      const float depth = sely ? (selx ? lo.w : lo.z) : (selx ? lo.y : lo.x);
      const float sdist = d - depth;
      if (sdist <= -limit) {
        weighted_tsd = -limit;
      } else if (sdist >= limit) {
      } else {
        const float q00 = fabsf(hi.x), q10 = fabsf(hi.y), q01 = fabsf(hi.z), q11 = fabsf(hi.w);
        const float w = lerpf(lerpf(q00, q10, wa), lerpf(q01, q11, wa), wb);
        weighted_tsd = (weighted_tsd * total_weight + w * sdist) / (total_weight + w);
        total_weight += w;
      }
    }
    ck0 = k0; ck1 = k1;
    p.tsdf[o] = weighted_tsd;
    if (WEIGHT) p.weight[o] = total_weight;
  }
}

template <int N, bool WEIGHT>
__global__ void __launch_bounds__(256) k_integrate_dense(const __grid_constant__ IntegrateParams p) {
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  if (x >= p.X || y >= p.Y) return;
  const int zb = p.z_begin + blockIdx.z * p.z_chunk;
  const int ze = min(zb + p.z_chunk, p.z_end);
  march_column<N, WEIGHT>(p, x, y, zb, ze);
}

// Occupied bricks only (VolumeSampler::sample(indices), volume_sampler.cpp:74-76). Persistent kernel: a fixed grid
// (a multiple of the 148 SMs) strides over work items = (occupied brick, z-chunk, block of BRICK_COLS columns); the
// item count comes from the device-side occupied count, so no host round trip and no empty blocks.
// Bricks may overlap or leave one-voxel gaps (float rounding in divideBox/containedVoxels): overlapping voxels are
// written twice with the same value, gaps keep the cleared -limit.
#define BRICK_MAX_THREADS 320
template <int N, bool WEIGHT, int MINB>
__global__ void __launch_bounds__(BRICK_MAX_THREADS, MINB) k_integrate_bricks(const __grid_constant__ IntegrateParams p, int max_cols, int max_nz, int BRICK_ZCHUNK) {
  const unsigned n_occ = *p.num_occupied;
  const unsigned col_blocks = ((unsigned)max_cols + blockDim.x - 1u) / blockDim.x;
  const unsigned z_blocks = (unsigned)(max_nz + BRICK_ZCHUNK - 1) / BRICK_ZCHUNK;
  const unsigned per_brick = col_blocks * z_blocks;
  const unsigned items = n_occ * per_brick;
  for (unsigned w = blockIdx.x; w < items; w += gridDim.x) {
    const unsigned b = w / per_brick, r = w - b * per_brick;
    const unsigned zc = r / col_blocks, cc = r - zc * col_blocks;
    const int32_t* rg = p.ranges + (size_t)p.occupied[b] * 6;
    const int x0 = rg[0], nx = rg[1] - rg[0], y0 = rg[2], ny = rg[3] - rg[2];
    const int zb = max(rg[4] + (int)zc * BRICK_ZCHUNK, p.z_begin);
    const int ze = min(min(rg[4] + (int)(zc + 1) * BRICK_ZCHUNK, rg[5]), p.z_end);
    const int ci = (int)(cc * blockDim.x + threadIdx.x);
    if (zb >= ze || ci >= nx * ny) continue;
    const int cy = ci / nx, cx = ci - cy * nx;
    march_column<N, WEIGHT>(p, x0 + cx, y0 + cy, zb, ze);
  }
}

// ---- fused clear + integrate (bricks mode) ----------------------------------------------------------------------
// ReconIntegration::integrate clears the whole volume to -limit and then overwrites the voxels of the occupied bricks
// (recon_integration.cpp:248-259). Here ONE persistent kernel writes every voxel of the slab exactly once: the clear is
// an HBM-bound stream of 16-byte stores (4 bytes per voxel), the brick evaluation is issue/latency-bound gather work,
// so warps of both kinds share each SM and overlap. Work is handed out per WARP from two device counters:
//   fill item    = FILL_ROWS consecutive (y, z) voxel rows; voxels inside an occupied brick are skipped (row bitmask
//                  built by k_bricks_masks), everything else gets -limit (and weight 0);
//   compute item = 32 consecutive columns of the flattened (occupied brick, column) space x one z-chunk.
// The first `fill_warps` warps of a CTA start on fill items, the others on compute items; a warp that runs out of its
// own kind helps with the other, so the kernel ends when both counters are exhausted.
struct FusedParams {
  IntegrateParams ip;
  const uint32_t* rowmask; const uint8_t* rowany; const int16_t* cand_y; const int16_t* cand_z;
  int mask_words, nby;
  uint32_t* work;              // [0] compute items handed out, [1] fill items handed out
  int max_cols, max_nz, zchunk, n_zchunks;
  int fill_rows; uint32_t fill_items; uint32_t row_begin, row_end;
  int fill_warps;
  float fill_value;
};

template <bool WEIGHT>
__device__ __forceinline__ void fill_rows(const FusedParams& p, uint32_t row0, uint32_t row1, int lane) {
  const int X = p.ip.X, Y = p.ip.Y;
  const float4 v4 = make_float4(p.fill_value, p.fill_value, p.fill_value, p.fill_value);
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool vec = (X & 3) == 0;
  for (uint32_t row = row0; row < row1; ++row) {
    const int z = (int)(row / (uint32_t)Y), y = (int)(row - (uint32_t)z * (uint32_t)Y);
    const int cy0 = p.cand_y[2 * y], cy1 = p.cand_y[2 * y + 1], cz0 = p.cand_z[2 * z], cz1 = p.cand_z[2 * z + 1];
    // brick rows that contain this voxel row (at most 2 x 2)
    int br[4];
    br[0] = (cy0 >= 0 && cz0 >= 0) ? cz0 * p.nby + cy0 : -1;
    br[1] = (cy1 >= 0 && cz0 >= 0) ? cz0 * p.nby + cy1 : -1;
    br[2] = (cy0 >= 0 && cz1 >= 0) ? cz1 * p.nby + cy0 : -1;
    br[3] = (cy1 >= 0 && cz1 >= 0) ? cz1 * p.nby + cy1 : -1;
    bool any = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) any = any || (br[k] >= 0 && p.rowany[br[k]] != 0);
    float* trow = p.ip.tsdf + (size_t)row * X;
    float* wrow = WEIGHT ? p.ip.weight + (size_t)row * X : nullptr;
    for (int chunk = 0; chunk * 1024 < X; ++chunk) {
      uint32_t comb = 0;
      if (any) {
        const int w = chunk * 32 + lane;
        if (w < p.mask_words) {
#pragma unroll
          for (int k = 0; k < 4; ++k) if (br[k] >= 0) comb |= __ldg(p.rowmask + (size_t)br[k] * p.mask_words + w);
        }
      }
      const int xbase = chunk * 1024;
      if (vec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t word = any ? __shfl_sync(0xffffffffu, comb, j * 4 + (lane >> 3)) : 0u;
          const int x = xbase + (j * 32 + lane) * 4;
          if (x >= X) continue;
          const uint32_t nib = (word >> ((lane & 7) * 4)) & 15u;
          if (nib == 0) {
            __stcs(reinterpret_cast<float4*>(trow + x), v4);
            if (WEIGHT) __stcs(reinterpret_cast<float4*>(wrow + x), z4);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (!((nib >> e) & 1u)) { trow[x + e] = p.fill_value; if (WEIGHT) wrow[x + e] = 0.0f; }
          }
        }
      } else {
        for (int i = 0; i < 32; ++i) {
          const uint32_t word = any ? __shfl_sync(0xffffffffu, comb, i) : 0u;
          const int x = xbase + i * 32 + lane;
          if (x >= X) continue;
          if (!((word >> lane) & 1u)) { trow[x] = p.fill_value; if (WEIGHT) wrow[x] = 0.0f; }
        }
      }
    }
  }
}

#ifndef RR_FUSED_THREADS
#define RR_FUSED_THREADS 512
#endif
template <int N, bool WEIGHT>
__global__ void __launch_bounds__(RR_FUSED_THREADS, 2) k_integrate_fused(const __grid_constant__ FusedParams p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned n_occ = *p.ip.num_occupied;
  const unsigned col_blocks = (n_occ * (unsigned)p.max_cols + 31u) / 32u;
  const unsigned citems = col_blocks * (unsigned)p.n_zchunks;
  bool filling = warp < p.fill_warps;
  for (int phase = 0; phase < 2; ++phase, filling = !filling) {
    if (filling) {
      for (;;) {
        unsigned it = 0;
        if (lane == 0) it = atomicAdd(p.work + 1, 1u);
        it = __shfl_sync(0xffffffffu, it, 0);
        if (it >= p.fill_items) break;
        const uint32_t r0 = p.row_begin + it * (uint32_t)p.fill_rows;
        fill_rows<WEIGHT>(p, r0, min(r0 + (uint32_t)p.fill_rows, p.row_end), lane);
      }
    } else {
      for (;;) {
        unsigned it = 0;
        if (lane == 0) it = atomicAdd(p.work, 1u);
        it = __shfl_sync(0xffffffffu, it, 0);
        if (it >= citems) break;
        const unsigned cb = it / (unsigned)p.n_zchunks, zc = it - cb * (unsigned)p.n_zchunks;
        const unsigned slot = cb * 32u + (unsigned)lane;
        const unsigned b = slot / (unsigned)p.max_cols, col = slot - b * (unsigned)p.max_cols;
        if (b >= n_occ) continue;
        const int32_t* rg = p.ip.ranges + (size_t)p.ip.occupied[b] * 6;
        const int x0 = rg[0], nx = rg[1] - rg[0], y0 = rg[2], ny = rg[3] - rg[2];
        const int zb = max(rg[4] + (int)zc * p.zchunk, p.ip.z_begin);
        const int ze = min(min(rg[4] + (int)(zc + 1) * p.zchunk, rg[5]), p.ip.z_end);
        if (zb >= ze || (int)col >= nx * ny) continue;
        const int cy = (int)col / nx, cx = (int)col - cy * nx;
        march_column<N, WEIGHT>(p.ip, x0 + cx, y0 + cy, zb, ze);
      }
    }
  }
}

// glClearTexImage(-limit) (recon_integration.cpp:250-251): 16-byte streaming stores over the slab.
__global__ void __launch_bounds__(256) k_fill(float* __restrict__ dst, size_t n, float value) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n4 = n / 4;
  float4* d4 = reinterpret_cast<float4*>(dst);
  const float4 v = make_float4(value, value, value, value);
  for (size_t j = i; j < n4; j += stride) __stcs(d4 + j, v);
  for (size_t j = n4 * 4 + i; j < n; j += stride) dst[j] = value;
}

template <int N>
static int launch_n(rr_ctx* c, const IntegrateParams& p, bool bricks, bool weight, bool fused) {
  const dim3 blk(32, 8, 1);
  if (bricks) {
    int max_cols = 0, max_nz = 0;
    for (size_t i = 0; i + 5 < c->h_ranges.size(); i += 6) {
      max_cols = std::max(max_cols, (c->h_ranges[i + 1] - c->h_ranges[i]) * (c->h_ranges[i + 3] - c->h_ranges[i + 2]));
      max_nz = std::max(max_nz, c->h_ranges[i + 5] - c->h_ranges[i + 4]);
    }
    if (max_cols == 0 || max_nz == 0) return RR_OK;
    static const int zchunk_env = getenv("RR_BRICK_ZCHUNK") ? atoi(getenv("RR_BRICK_ZCHUNK")) : 0;
    if (fused) {
      // z-chunks: split a brick's z extent into pieces of ~zchunk voxels of equal size
      const int want = zchunk_env > 0 ? zchunk_env : 13;
      const int n_zchunks = std::max(1, (max_nz + want - 1) / want);
      FusedParams f{};
      f.ip = p;
      f.rowmask = c->d_rowmask; f.rowany = c->d_rowany; f.cand_y = c->d_cand_y; f.cand_z = c->d_cand_z;
      f.mask_words = c->mask_words; f.nby = (int)c->bricks.res[1];
      f.work = c->d_work;
      f.max_cols = max_cols; f.max_nz = max_nz; f.n_zchunks = n_zchunks; f.zchunk = (max_nz + n_zchunks - 1) / n_zchunks;
      static const int fill_rows = getenv("RR_FUSED_FILL_ROWS") ? atoi(getenv("RR_FUSED_FILL_ROWS")) : 32;
      static const int fill_warps = getenv("RR_FUSED_FILL_WARPS") ? atoi(getenv("RR_FUSED_FILL_WARPS")) : 2;
      static const int ctas_per_sm = getenv("RR_FUSED_CTAS") ? atoi(getenv("RR_FUSED_CTAS")) : 2;
      f.fill_rows = fill_rows; f.fill_warps = fill_warps;
      f.row_begin = (uint32_t)p.z_begin * (uint32_t)p.Y; f.row_end = (uint32_t)p.z_end * (uint32_t)p.Y;
      f.fill_items = (f.row_end - f.row_begin + (uint32_t)fill_rows - 1u) / (uint32_t)fill_rows;
      f.fill_value = -p.limit;
      cudaMemsetAsync(c->d_work, 0, 4 * sizeof(uint32_t), c->stream);
      const dim3 grd(148 * ctas_per_sm, 1, 1);
      if (weight) k_integrate_fused<N, true><<<grd, RR_FUSED_THREADS, 0, c->stream>>>(f);
      else k_integrate_fused<N, false><<<grd, RR_FUSED_THREADS, 0, c->stream>>>(f);
      RR_LAUNCH_CHECK(c, "k_integrate_fused");
      return RR_OK;
    }
    // block size: the multiple of 32 (128..320) that wastes the fewest lanes on a brick's column count
    int threads = 256;
    double best = 1e9;
    for (int t = 128; t <= BRICK_MAX_THREADS; t += 32) {
      const double waste = double((max_cols + t - 1) / t * t) / double(max_cols);
      if (waste < best - 1e-9 || (waste < best + 1e-9 && t > threads)) { best = waste; threads = t; }
    }
    const int zchunk = zchunk_env > 0 ? zchunk_env : 9;
    static const int gmult = getenv("RR_BRICK_GRID") ? atoi(getenv("RR_BRICK_GRID")) : 6;
    const dim3 grd(148 * gmult, 1, 1);
    if (weight) k_integrate_bricks<N, true, 2><<<grd, threads, 0, c->stream>>>(p, max_cols, max_nz, zchunk);
    else k_integrate_bricks<N, false, 2><<<grd, threads, 0, c->stream>>>(p, max_cols, max_nz, zchunk);
  } else {
    const int nz = p.z_end - p.z_begin;
    const dim3 grd((p.X + 31) / 32, (p.Y + 7) / 8, (nz + p.z_chunk - 1) / p.z_chunk);
    if (weight) k_integrate_dense<N, true><<<grd, blk, 0, c->stream>>>(p);
    else k_integrate_dense<N, false><<<grd, blk, 0, c->stream>>>(p);
  }
  RR_LAUNCH_CHECK(c, "k_integrate");
  return RR_OK;
}

int launch_integrate(rr_ctx* c) {
  IntegrateParams p{};
  p.inv = c->d_inv; p.gather = c->d_gather; p.tsdf = c->d_tsdf; p.weight = c->d_weight;
  p.ranges = c->d_ranges; p.occupied = c->d_occupied; p.num_occupied = c->d_num_occ;
  p.IX = (int)c->ires[0]; p.IY = (int)c->ires[1]; p.IZ = (int)c->ires[2];
  p.W = c->W; p.H = c->H; p.X = (int)c->res[0]; p.Y = (int)c->res[1]; p.Z = (int)c->res[2];
  p.z_begin = (int)c->slab_z0; p.z_end = (int)c->slab_z1;
  p.z_chunk = 32;
  p.limit = c->cfg.limit;
  const bool weight = c->cfg.store_weight != 0;
  const bool bricks = c->cfg.use_bricks != 0;
  timer_begin(c, "2integrate");
  const size_t plane = (size_t)p.X * p.Y;
  const size_t nslab = plane * (size_t)(p.z_end - p.z_begin);
  // RR_INTEGRATE_FUSED=0 selects the two-kernel path (k_fill, then k_integrate_bricks) for A/B measurements
  const bool fused_env = !(getenv("RR_INTEGRATE_FUSED") && atoi(getenv("RR_INTEGRATE_FUSED")) == 0);
  const bool fused = bricks && fused_env && c->fused_ok;
  if (bricks && !fused && nslab) {
    // dense mode overwrites every voxel, so only the brick path needs the clear
    k_fill<<<148 * 8, 256, 0, c->stream>>>(c->d_tsdf + plane * p.z_begin, nslab, -p.limit);
    RR_LAUNCH_CHECK(c, "k_fill");
    if (weight) {
      k_fill<<<148 * 8, 256, 0, c->stream>>>(c->d_weight + plane * p.z_begin, nslab, 0.0f);
      RR_LAUNCH_CHECK(c, "k_fill");
    }
  }
  int rc = RR_OK;
  if (nslab) {
    switch (c->N) {
      case 1: rc = launch_n<1>(c, p, bricks, weight, fused); break;
      case 2: rc = launch_n<2>(c, p, bricks, weight, fused); break;
      case 3: rc = launch_n<3>(c, p, bricks, weight, fused); break;
      case 4: rc = launch_n<4>(c, p, bricks, weight, fused); break;
      case 5: rc = launch_n<5>(c, p, bricks, weight, fused); break;
      case 6: rc = launch_n<6>(c, p, bricks, weight, fused); break;
      case 7: rc = launch_n<7>(c, p, bricks, weight, fused); break;
      case 8: rc = launch_n<8>(c, p, bricks, weight, fused); break;
      default: return fail(c, RR_ERR_UNSUPPORTED, "integrate: 1..8 sensors supported");
    }
  }
  timer_end(c, "2integrate");
  return rc;
}

}  // namespace rr

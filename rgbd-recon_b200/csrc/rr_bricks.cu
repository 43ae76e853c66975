// Brick occupancy on the device: ReconIntegration::clearOccupiedBricks / updateOccupiedBricks
// (framework/reconstruction/recon_integration.cpp:272-278, 431-446) without the GPU->CPU->GPU round trip.
// The occupied list must equal the CPU loop's order (ascending brick id), so the compaction is an ORDERED
// ballot + prefix scan by a single 1024-thread block (a brick grid is ~10^4 counters: one block is latency-optimal).
#include "rr_context.h"

namespace rr {

__global__ void __launch_bounds__(1024) k_bricks_compact(const uint32_t* __restrict__ counters, uint32_t num_bricks,
                                                         uint32_t min_voxels, uint32_t* __restrict__ occupied,
                                                         uint32_t* __restrict__ num_occupied) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t base;
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  for (uint32_t start = 0; start < num_bricks; start += 1024u) {
    const uint32_t i = start + threadIdx.x;
    const bool occ = (i < num_bricks) && (counters[i] >= min_voxels);
    const unsigned ballot = __ballot_sync(0xffffffffu, occ);
    const uint32_t rank_in_warp = __popc(ballot & ((1u << lane) - 1u));
    if (lane == 0) warp_sums[warp] = __popc(ballot);
    __syncthreads();
    uint32_t warp_off = 0, total = 0;
    for (unsigned w = 0; w < 32; ++w) {
      const uint32_t s = warp_sums[w];
      if (w < warp) warp_off += s;
      total += s;
    }
    if (occ) occupied[base + warp_off + rank_in_warp] = i;
    __syncthreads();
    if (threadIdx.x == 0) base += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *num_occupied = base;
}

// Per-brick and per-brick-row masks derived from the counters (one launch):
//   near_occ[b] = 1 if brick b or any of its 26 neighbours is occupied: the raymarcher may skip the TSDF fetches of
//     samples inside bricks with near_occ == 0 (every trilinear tap there still holds the cleared value -limit);
//   occ_mask[b] = brick b is occupied;
//   rowmask[bz][by][w] = bit x set iff voxel column x lies inside the x range of an occupied brick of row (by, bz);
//   rowany[bz][by]     = the row has an occupied brick  (both read by the fused clear+integrate kernel).
__global__ void __launch_bounds__(256) k_bricks_masks(const uint32_t* __restrict__ counters, uint32_t rx, uint32_t ry, uint32_t rz,
                                                      uint32_t min_voxels, const int32_t* __restrict__ ranges, int mask_words,
                                                      uint8_t* __restrict__ near_occ, uint8_t* __restrict__ occ_mask,
                                                      uint32_t* __restrict__ rowmask, uint8_t* __restrict__ rowany) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rx * ry * rz) {
    const int bx = (int)(i % rx), by = (int)((i / rx) % ry), bz = (int)(i / (rx * ry));
    uint8_t any = 0;
    for (int dz = -1; dz <= 1; ++dz)
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          const int x = bx + dx, y = by + dy, z = bz + dz;
          if (x < 0 || y < 0 || z < 0 || x >= (int)rx || y >= (int)ry || z >= (int)rz) continue;
          if (counters[((size_t)z * ry + y) * rx + x] >= min_voxels) any = 1;
        }
    near_occ[i] = any;
    occ_mask[i] = counters[i] >= min_voxels ? 1 : 0;
  }
  if (rowmask && i < ry * rz * (uint32_t)mask_words) {
    const uint32_t row = i / (uint32_t)mask_words, w = i - row * (uint32_t)mask_words;
    const int w0 = (int)(w * 32u);
    uint32_t m = 0, any = 0;
    for (uint32_t bx = 0; bx < rx; ++bx) {
      if (counters[(size_t)row * rx + bx] < min_voxels) continue;
      any = 1;
      const int lo = max(ranges[bx * 6] - w0, 0), hi = min(ranges[bx * 6 + 1] - w0, 32);   // x range of brick (bx, 0, 0)
      if (lo < hi) m |= (hi - lo == 32) ? 0xFFFFFFFFu : (((1u << (hi - lo)) - 1u) << lo);
    }
    rowmask[i] = m;
    if (w == 0) rowany[row] = (uint8_t)any;
  }
}

int launch_bricks_clear(rr_ctx* c) {
  cudaError_t e = cudaMemsetAsync(c->d_counters, 0, sizeof(uint32_t) * c->bricks.num, c->stream);
  return check(c, e, "bricks clear");
}

int launch_bricks_update(rr_ctx* c) {
  k_bricks_compact<<<1, 1024, 0, c->stream>>>(c->d_counters, c->bricks.num, c->cfg.min_voxels_per_brick, c->d_occupied, c->d_num_occ);
  RR_LAUNCH_CHECK(c, "k_bricks_compact");
  const uint32_t nb = c->bricks.num;
  if (nb == c->bricks.res[0] * c->bricks.res[1] * c->bricks.res[2]) {
    const uint32_t rows_words = c->fused_ok ? c->bricks.res[1] * c->bricks.res[2] * (uint32_t)c->mask_words : 0u;
    const uint32_t threads = nb > rows_words ? nb : rows_words;
    k_bricks_masks<<<(threads + 255) / 256, 256, 0, c->stream>>>(c->d_counters, c->bricks.res[0], c->bricks.res[1], c->bricks.res[2],
                                                             c->cfg.min_voxels_per_brick, c->d_ranges, c->mask_words, c->d_near_occ,
                                                             c->d_occ_mask, c->fused_ok ? c->d_rowmask : nullptr, c->d_rowany);
    RR_LAUNCH_CHECK(c, "k_bricks_masks");
  }
  cudaError_t e = cudaMemcpyAsync(c->h_num_occ, c->d_num_occ, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream);
  return check(c, e, "bricks count copy");
}

}  // namespace rr

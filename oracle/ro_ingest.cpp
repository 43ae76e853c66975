// ORACLE (test infrastructure, NOT product code). Compressed frame ingest of the reference:
//   * colour: S3TC DXT1 blocks (GL_COMPRESSED_RGBA_S3TC_DXT1_EXT, framework/NetKinectArray.cpp:120-131,149-151), decoded
//     by the GL sampler in the reference; restated here from the published BC1 definition and PINNED against the
//     reference's own CPU decoder external/squish (squish::DecompressImage, colourblock.cpp:160-214; used by the reference
//     at NetKinectArray.cpp:635) through oracle/_ref and tests/golden/ref_dxt1.npz. Only .rgb is sampled downstream.
//     DXT5 streams (compress_rgb == 5, KinectCalibrationFile.cpp:336) carry the same colour block behind 8 alpha bytes;
//     pinned the same way (tests/golden/ref_dxt5.npz).
//   * depth: 8-bit GL_LUMINANCE texels (NetKinectArray.cpp:170-172), sampled as normalised fixed point byte/255 and
//     expanded by pre_depth.fs uncompress() (:51-61; restated in ro_preprocess.cpp).
#include "rr_oracle.h"

#include <cstdint>

namespace {
inline void unpack565(const uint8_t* b, uint8_t* rgb, int& value) {
  value = (int)b[0] | ((int)b[1] << 8);
  const int r = (value >> 11) & 0x1f, g = (value >> 5) & 0x3f, bl = value & 0x1f;
  rgb[0] = (uint8_t)((r << 3) | (r >> 2));
  rgb[1] = (uint8_t)((g << 2) | (g >> 4));
  rgb[2] = (uint8_t)((bl << 3) | (bl >> 2));
}
}  // namespace

// DXT1: (W/4)*(H/4) blocks of 8 bytes, row-major over 4x4 tiles. DXT5 (GL_COMPRESSED_RGBA_S3TC_DXT5_EXT,
// NetKinectArray.cpp:125-128,153-156): 16-byte blocks = 8 bytes of alpha (not sampled downstream: skipped) followed by the
// same colour block, always in four-colour mode (squish colourblock.cpp:160-214 with isDxt1 = false).
// out: uint8 [H][W][3]. W, H multiples of 4.
static void decode_bc(const uint8_t* blocks, int W, int H, uint8_t* out, bool dxt5) {
  const int bw = W / 4, bh = H / 4;
  for (int by = 0; by < bh; ++by)
    for (int bx = 0; bx < bw; ++bx) {
      const uint8_t* blk = blocks + ((size_t)by * bw + bx) * (dxt5 ? 16 : 8) + (dxt5 ? 8 : 0);
      uint8_t codes[4][3];
      int a, b;
      unpack565(blk, codes[0], a);
      unpack565(blk + 2, codes[1], b);
      for (int i = 0; i < 3; ++i) {
        const int c = codes[0][i], d = codes[1][i];
        if (!dxt5 && a <= b) { codes[2][i] = (uint8_t)((c + d) / 2); codes[3][i] = 0; }
        else { codes[2][i] = (uint8_t)((2 * c + d) / 3); codes[3][i] = (uint8_t)((c + 2 * d) / 3); }
      }
      for (int py = 0; py < 4; ++py) {
        const uint8_t packed = blk[4 + py];
        for (int px = 0; px < 4; ++px) {
          const int idx = (packed >> (2 * px)) & 3;
          uint8_t* o = out + (((size_t)(by * 4 + py)) * W + (bx * 4 + px)) * 3;
          o[0] = codes[idx][0]; o[1] = codes[idx][1]; o[2] = codes[idx][2];
        }
      }
    }
}

extern "C" void ro_decode_dxt1(const uint8_t* blocks, int W, int H, uint8_t* out) { decode_bc(blocks, W, H, out, false); }
extern "C" void ro_decode_dxt5(const uint8_t* blocks, int W, int H, uint8_t* out) { decode_bc(blocks, W, H, out, true); }

// GL normalised fixed point: float = byte / 255 (GL 4.4 §2.3.4.1, one IEEE division)
extern "C" void ro_depth8_to_float(const uint8_t* in, size_t n, float* out) {
  for (size_t i = 0; i < n; ++i) out[i] = (float)in[i] / 255.0f;
}

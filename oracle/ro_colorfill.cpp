// ORACLE (test infrastructure, NOT product code). The reference ships no tests; this file is pinned against the reference's own
// framebuffer_transfer.fs / tsdf_inpaint.fs / tsdf_colorfill.fs run on the CPU (oracle/glsl_host/, tests/golden/ref_glsl_colorfill.npz).
// Scalar restatement of the colour hole filling that follows the raymarch when m_fill_holes is set (the default,
// framework/reconstruction/recon_integration.cpp:54): ReconIntegration::fillColors (recon_integration.cpp:280-339),
// the ViewLod mip atlas (framework/rendering/view_lod.cpp:24-61), glsl/framebuffer_transfer.fs, glsl/tsdf_inpaint.fs
// and glsl/tsdf_colorfill.fs. The passes are simulated literally, whole framebuffers and all:
//
//   draw():        atlas F (1.5W x H, RGBA32F + depth) cleared to (0,1,0,0) / 1.0, raymarch result in the lod-0 viewport.
//   transfer:      S cleared; S[px,py] = F[ivec2(texcoord * resolution_full)] for the W x H viewport — this squeezes
//                  the WHOLE atlas into the lod-0 viewport (x scaled by 2/3: source column floor(1.5 px + 0.75)).
//   for i = 1..L-1: inpaint(lod = i-1) renders lod i of F from S (4x4 taps around the 2/3-scaled position), then
//                  another transfer refreshes S from F.
//   colorfill:     per window pixel the first lod whose texel has alpha > 0; below lod 0 the colour is a blend of two
//                  bilinear (MIRRORED_REPEAT) fetches from lods level+1 and level+2; depth = raymarch depth; the depth
//                  test GL_LESS against the cleared 1.0 keeps only pixels the raymarch hit.
//
// GL-undefined cases are fixed as NVIDIA behaves: texelFetch outside the texture returns zeros; clamp(x, lo, hi) with
// lo > hi is min(max(x, lo), hi); unset uniform array elements are zero. Varyings: pass_TexCoord = (pixel + 0.5) / size.
#include "ro_math.h"
#include "rr_oracle.h"

#include <algorithm>
#include <cmath>
#include <vector>

using namespace ro;

namespace {

struct Atlas {
  int FW = 0, H = 0;
  std::vector<V4> color;
  std::vector<float> depth;
  void init(int fw, int h) { FW = fw; H = h; color.assign((size_t)fw * h, V4{0.f, 1.f, 0.f, 0.f}); depth.assign((size_t)fw * h, 1.0f); }
  void clear() { std::fill(color.begin(), color.end(), V4{0.f, 1.f, 0.f, 0.f}); std::fill(depth.begin(), depth.end(), 1.0f); }
  bool inside(int x, int y) const { return x >= 0 && y >= 0 && x < FW && y < H; }
  V4 fetch_color(int x, int y) const { return inside(x, y) ? color[(size_t)y * FW + x] : V4{0.f, 0.f, 0.f, 0.f}; }
  float fetch_depth(int x, int y) const { return inside(x, y) ? depth[(size_t)y * FW + x] : 0.0f; }
};

struct Lods {
  int n = 0;
  int off[20][2] = {}, res[20][2] = {};     // uniform uvec2[20]; elements past n stay zero
};

// ViewLod::setResolution (view_lod.cpp:24-52)
Lods make_lods(int W, int H) {
  Lods l;
  l.n = 1 + (int)std::floor(std::log2((float)std::min(W, H)));
  if (l.n > 20) l.n = 20;
  int oy = H;
  for (int i = 0; i < l.n; ++i) {
    l.res[i][0] = (int)std::floor((float)W / std::pow(2.0f, (float)i));
    l.res[i][1] = (int)std::floor((float)H / std::pow(2.0f, (float)i));
    if (i > 0) { oy -= l.res[i][1]; l.off[i][0] = W; l.off[i][1] = oy; }
  }
  return l;
}

// GL 4.4 §8.14.2 MIRRORED_REPEAT
int mirror_wrap(int i, int size) {
  int m = i % (2 * size);
  if (m < 0) m += 2 * size;
  int a = m - size;
  if (a < 0) a = -(1 + a);
  return (size - 1) - a;
}

V4 texture_linear_mirror(const Atlas& t, float s, float r) {
  const float u = s * (float)t.FW - 0.5f, v = r * (float)t.H - 0.5f;
  const float fu = floorf(u), fv = floorf(v);
  const float a = u - fu, b = v - fv;
  const int i0 = mirror_wrap((int)fu, t.FW), i1 = mirror_wrap((int)fu + 1, t.FW);
  const int j0 = mirror_wrap((int)fv, t.H), j1 = mirror_wrap((int)fv + 1, t.H);
  const V4 c00 = t.color[(size_t)j0 * t.FW + i0], c10 = t.color[(size_t)j0 * t.FW + i1];
  const V4 c01 = t.color[(size_t)j1 * t.FW + i0], c11 = t.color[(size_t)j1 * t.FW + i1];
  V4 o;
  o.x = lerpf(lerpf(c00.x, c10.x, a), lerpf(c01.x, c11.x, a), b);
  o.y = lerpf(lerpf(c00.y, c10.y, a), lerpf(c01.y, c11.y, a), b);
  o.z = lerpf(lerpf(c00.z, c10.z, a), lerpf(c01.z, c11.z, a), b);
  o.w = lerpf(lerpf(c00.w, c10.w, a), lerpf(c01.w, c11.w, a), b);
  return o;
}

// framebuffer_transfer.fs over the lod-0 viewport of dst, after ViewLod::enable(0) cleared dst
void transfer(const Atlas& src, Atlas& dst, int W, int H) {
  dst.clear();
  for (int py = 0; py < H; ++py)
    for (int px = 0; px < W; ++px) {
      const float tx = ((float)px + 0.5f) / (float)W, ty = ((float)py + 0.5f) / (float)H;
      const int sx = (int)(tx * (float)src.FW), sy = (int)(ty * (float)src.H);
      dst.color[(size_t)py * dst.FW + px] = src.fetch_color(sx, sy);
      dst.depth[(size_t)py * dst.FW + px] = src.fetch_depth(sx, sy);
    }
}

// tsdf_inpaint.fs with uniform lod = l, rendered into the lod l+1 viewport of dst (no clear), reading src
void inpaint(const Atlas& src, Atlas& dst, const Lods& L, int l) {
  const int ox = L.off[l + 1][0], oy = L.off[l + 1][1], rx = L.res[l + 1][0], ry = L.res[l + 1][1];
  for (int fy = oy; fy < oy + ry; ++fy)
    for (int fx = ox; fx < ox + rx; ++fx) {
      if (!dst.inside(fx, fy)) continue;
      const float tcx = ((float)fx - (float)ox) / (float)rx, tcy = ((float)fy - (float)oy) / (float)ry;
      const int lx = (int)((float)L.off[l][0] + (float)L.res[l][0] * tcx), ly = (int)((float)L.off[l][1] + (float)L.res[l][1] * tcy);
      const int pix = (int)((float)lx * (2.0f / 3.0f)), piy = (int)((float)ly * 1.0f);
      float depth_av = 0.0f;
      int num_samples = 0;
      V4 samples[16];
      for (int x = 0; x < 4; ++x)
        for (int y = 0; y < 4; ++y) {
          const int tx = pix + (int)((float)x - 4.0f * 0.5f + 1.0f), ty = piy + (int)((float)y - 4.0f * 0.5f + 1.0f);
          V4 color = src.fetch_color(tx, ty);
          const float depth = src.fetch_depth(tx, ty);
          if (color.w <= 0.0f) color.x = -1.0f;
          else { depth_av += depth; ++num_samples; }
          samples[x + y * 4] = V4{color.x, color.y, color.z, depth};
        }
      V4 out;
      float out_depth;
      if (num_samples == 0) {
        out_depth = src.fetch_depth(pix, piy);
        out = (out_depth < 1.0f) ? V4{0.f, 0.f, 0.f, -1.f} : V4{0.f, 1.f, 0.f, 0.f};
      } else {
        depth_av /= (float)num_samples;
        float tc[3] = {0.f, 0.f, 0.f}, total_depth = 0.0f, total_weight = 0.0f;
        for (int i = 0; i < 16; ++i)
          if (samples[i].x >= 0.0f && samples[i].w >= depth_av) {
            const float weight = 1.0f;
            tc[0] += samples[i].x * weight; tc[1] += samples[i].y * weight; tc[2] += samples[i].z * weight;
            total_depth += samples[i].w * weight;
            total_weight += weight;
          }
        out = V4{tc[0] / total_weight, tc[1] / total_weight, tc[2] / total_weight, 1.0f};
        out_depth = total_depth / total_weight;
      }
      dst.color[(size_t)fy * dst.FW + fx] = out;
      dst.depth[(size_t)fy * dst.FW + fx] = out_depth;
    }
}

}  // namespace

extern "C" int ro_fill_num_lods(int W, int H) { return make_lods(W, H).n; }

// rgba [H][W][4], depth [H][W] (the raymarch outputs: alpha 1 = blended colour, -1 = fallback colour, 0 = no surface;
// depth 1.0 = no surface). out_rgba [H][W][4]: the colorfill output where depth < 1 (GL_LESS against the cleared depth
// buffer), the input pixel elsewhere. Optional: the final atlas (colour [H][FW][4], depth [H][FW]) for diagnosis.
extern "C" void ro_fill_colors(const float* rgba, const float* depth, int W, int H, float* out_rgba,
                               float* atlas_rgba, float* atlas_depth) {
  const int FW = (int)((float)W * 1.5f);
  const Lods L = make_lods(W, H);
  Atlas F, S;
  F.init(FW, H); S.init(FW, H);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const float* p = rgba + ((size_t)y * W + x) * 4;
      // a pixel the raymarch discarded (alpha 0), or whose fragment failed GL_LESS against the cleared 1.0, keeps the
      // cleared framebuffer value
      const float d = depth[(size_t)y * W + x];
      const bool drawn = (p[3] != 0.0f) && (d < 1.0f);
      F.color[(size_t)y * FW + x] = drawn ? V4{p[0], p[1], p[2], p[3]} : V4{0.f, 1.f, 0.f, 0.f};
      F.depth[(size_t)y * FW + x] = drawn ? d : 1.0f;
    }
  transfer(F, S, W, H);
  for (int i = 1; i < L.n; ++i) {
    inpaint(S, F, L, i - 1);
    transfer(F, S, W, H);
  }
  if (atlas_rgba)
    for (size_t i = 0; i < F.color.size(); ++i) { atlas_rgba[i * 4] = F.color[i].x; atlas_rgba[i * 4 + 1] = F.color[i].y; atlas_rgba[i * 4 + 2] = F.color[i].z; atlas_rgba[i * 4 + 3] = F.color[i].w; }
  if (atlas_depth) std::copy(F.depth.begin(), F.depth.end(), atlas_depth);

  // tsdf_colorfill.fs over the W x H window
  const float inv_fw = 1.0f / (float)FW, inv_h = 1.0f / (float)H;
  for (int py = 0; py < H; ++py)
    for (int px = 0; px < W; ++px) {
      const float tcx = (float)px / (float)L.res[0][0], tcy = (float)py / (float)L.res[0][1];
      const float ptx = ((float)px + 0.5f) / (float)W, pty = ((float)py + 0.5f) / (float)H;
      V4 out{0.f, 0.f, 0.f, 0.f};
      int level = 0;
      for (; level < L.n; ++level) {
        const int cx = (int)((float)L.off[level][0] + (float)L.res[level][0] * tcx);
        const int cy = (int)((float)L.off[level][1] + (float)L.res[level][1] * tcy);
        out = F.fetch_color(cx, cy);
        if (out.w > 0.0f) break;
      }
      if (level > 0) {
        auto lod_pos2 = [&](int lod, float& ox, float& oy) {
          const float o0 = (float)L.off[lod][0], o1 = (float)L.off[lod][1], r0 = (float)L.res[lod][0], r1 = (float)L.res[lod][1];
          ox = gl_clamp(o0 + r0 * ptx, o0 + 0.5f, (float)(L.off[lod][0] + L.res[lod][0]) - 0.5f);
          oy = gl_clamp(o1 + r1 * pty, o1 + 0.5f, (float)(L.off[lod][1] + L.res[lod][1]) - 0.5f);
        };
        float p2x, p2y, p1x, p1y;
        lod_pos2(std::min(level + 2, 19), p2x, p2y);
        lod_pos2(std::min(level + 1, 19), p1x, p1y);
        const V4 c1 = texture_linear_mirror(F, p1x * inv_fw, p1y * inv_h);
        const V4 c2 = texture_linear_mirror(F, p2x * inv_fw, p2y * inv_h);
        const float dx = ptx - floorf(ptx), dy = pty - floorf(pty);
        const float w1 = sqrtf(fmaf(dy, dy, dx * dx));
        const float w2 = 1.0f - w1;
        const float ws = w1 + w2;
        out = V4{(c1.x * w1 + c2.x * w2) / ws, (c1.y * w1 + c2.y * w2) / ws, (c1.z * w1 + c2.z * w2) / ws, (c1.w * w1 + c2.w * w2) / ws};
      }
      const int dx0 = (int)((float)L.off[0][0] + (float)L.res[0][0] * tcx), dy0 = (int)((float)L.off[0][1] + (float)L.res[0][1] * tcy);
      const float frag_depth = F.fetch_depth(dx0, dy0);
      float* o = out_rgba + ((size_t)py * W + px) * 4;
      const float* in = rgba + ((size_t)py * W + px) * 4;
      if (frag_depth < 1.0f) { o[0] = out.x; o[1] = out.y; o[2] = out.z; o[3] = out.w; }
      else { o[0] = in[0]; o[1] = in[1]; o[2] = in[2]; o[3] = in[3]; }
    }
}

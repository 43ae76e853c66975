/* ORACLE (test infrastructure, NOT product code): C entry points of the scalar CPU restatement of
 * steppobeck/rgbd-recon's volumetric-fusion path. Loaded with ctypes by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs ONLY. The reference ships no tests; the
 * restatement is pinned against reference code compiled into oracle/_ref: C++ sources (libref_harness.so, see ro_math.h) and
 * the pre-processing / integration / raymarch / colour-fill SHADERS run on the CPU (libref_glsl.so, oracle/glsl_host/). The
 * space-skipping hull of ro_raymarch.cpp is checked through the shader's skipSpace branch on depth peels produced by the
 * reference's bricks.vs/gs/fs and a rasteriser (glsl_harness.cpp::rg_depth_peels). */
#ifndef RR_ORACLE_H
#define RR_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

void ro_set_threads(int n);
int ro_get_max_threads(void);

/* scalar known-answer hooks for the math pins */
float ro_kat_log2(float x);
float ro_kat_exp2(float x);
float ro_kat_pow(float x, float y);
void ro_kat_tex3d(const float* vol, int C, int X, int Y, int Z, float s, float t, float r, float* out);
float ro_kat_tex2d(const float* img, int W, int H, float s, float t, int nearest);
void ro_kat_rgb_to_lab(const float* rgb, float* lab);

void ro_pre_morph(const float* depth_in, int W, int H, float* depth_out);
void ro_pre_depth(const float* depth_in, int W, int H, const float* cv_xyz, const float* cv_uv, int CX, int CY, int CZ,
                  const uint8_t* color, int CW, int CH, const float* bbox_min, const float* bbox_max,
                  float cv_min_ds, float cv_max_ds, int filter_textures, int compress, float scale, float near_,
                  float scaled_near, float* out_depth, float* out_lab);
void ro_pre_boundary(const float* depth_rg, const float* lab, int W, int H, int refine, float* out_depth_b, float* out_sil);
void ro_pre_normal(const float* depth_b, int W, int H, const float* cv_xyz, int CX, int CY, int CZ,
                   const float* bbox_min, float brick_size, const uint32_t* brick_res, uint32_t num_bricks,
                   uint32_t* bricks, float* out_normal);
void ro_pre_quality(const float* depth_b, const float* normals, int W, int H, const float* cv_xyz, int CX, int CY, int CZ,
                    const float* camera_pos, float* out_quality);

void ro_volume_res(const float* bbox_min, const float* bbox_max, float voxel_size, uint32_t* res_out);
float ro_adjust_brick_size(float voxel_size, float size);
uint32_t ro_divide_box(const float* bbox_min, const float* bbox_max, float brick_size, const uint32_t* res_volume,
                       uint32_t* res_bricks_out, int32_t* ranges);
uint32_t ro_divide_box_args(const float* bbox_min, const float* bbox_max, float brick_size, const uint32_t* res_volume,
                            float* args, int32_t* raw_ranges);
uint32_t ro_occupied_bricks(const uint32_t* counters, uint32_t num_bricks, uint32_t min_voxels, uint32_t* occupied_out);
void ro_integrate(int N, const float* inv, const int32_t* inv_res, const float* sil, const float* depth_b,
                  const float* quality, int W, int H, float limit, const uint32_t* res, int use_bricks,
                  const int32_t* brick_ranges, const uint32_t* occupied, uint32_t num_occupied, float* tsdf, float* weight);

void ro_frustum(const float* cv_xyz, int X, int Y, int Z, float* planes_out, float* campos_out);
int ro_frustum_inside(const float* planes, const float* p);
void ro_calib_invert(const float* cv_xyz, int X, int Y, int Z, const float* bbox_min, const float* bbox_max,
                     const uint32_t* out_res, float* out, uint32_t* neigh_out, int brute);

void ro_raymarch_uniforms(const float* mv, const float* proj, const float* bmin, const float* bmax, int vw, int vh, float* out);
void ro_raymarch_rays(const float* modelview, const float* projection, const float* bbox_min, const float* bbox_max, int vw, int vh,
                      float limit, float* out_points, uint8_t* out_covered);
void ro_raymarch(const float* tsdf, const uint32_t* res, float limit, int N, const float* inv, const int32_t* inv_res,
                 const float* cv_uv, const int32_t* cv_res, const uint8_t* color, int CW, int CH,
                 const float* depth_b, const float* quality, int W, int H, const float* bbox_min, const float* bbox_max,
                 const float* modelview, const float* projection, int vw, int vh, int shade_mode, int skip_space,
                 const uint32_t* occupied, uint32_t n_occ, const uint32_t* brick_res, float brick_size,
                 float* out_rgba, float* out_depth, float* out_samples, float* out_pos);


/* The point-drawing reconstructions: ReconPoints::draw and ReconCalibs::draw, see ro_points.cpp. */
void ro_draw_points(int N, int W, int H, const float* depth_b, const float* normal, const uint8_t* color, int CW, int CH,
                    const float* cv_xyz, const float* cv_uv, const int32_t* cv_res, const float* bmin, const float* bmax,
                    const float* mv, const float* proj, int vw, int vh, int shade_mode, float* out_rgba, float* out_depth);
void ro_draw_calibs(const float* tsdf, const uint32_t* res, const int32_t* inv_res, float limit, const float* bmin, const float* bmax,
                    const float* mv, const float* proj, int vw, int vh, float* out_rgba, float* out_depth);

/* ReconTrigrid::draw, see ro_trigrid.cpp. out_accum / out_depth1 may be null. */
void ro_draw_trigrid(int N, int W, int H, const float* depth_b, const float* quality, const uint8_t* color, int CW, int CH,
                     const float* cv_xyz, const float* cv_uv, const int32_t* cv_res, const float* bmin, const float* bmax,
                     const float* mv, const float* proj, int vw, int vh, int shade_mode, float min_length, float epsilon,
                     float* out_rgba, float* out_depth, float* out_accum, float* out_depth1);

/* Compressed ingest: DXT1 colour blocks -> uint8 [H][W][3]; 8-bit depth -> byte/255. */
void ro_decode_dxt1(const uint8_t* blocks, int W, int H, uint8_t* out_rgb);
void ro_decode_dxt5(const uint8_t* blocks, int W, int H, uint8_t* out_rgb);   /* 16-byte blocks: alpha (skipped) + colour */
void ro_depth8_to_float(const uint8_t* in, size_t n, float* out);

/* Colour hole filling after the raymarch (ReconIntegration::fillColors + ViewLod + framebuffer_transfer / tsdf_inpaint /
 * tsdf_colorfill): rgba [H][W][4] and depth [H][W] in, out_rgba [H][W][4]; atlas_* optional ([H][1.5W][4] / [H][1.5W]). */
int ro_fill_num_lods(int W, int H);
void ro_fill_colors(const float* rgba, const float* depth, int W, int H, float* out_rgba, float* atlas_rgba, float* atlas_depth);

#ifdef __cplusplus
}
#endif
#endif

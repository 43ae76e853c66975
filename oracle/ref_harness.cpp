// ORACLE (test infrastructure, NOT product code): C entry points over REAL reference code, compiled where it lies
// under /root/reference by oracle/Makefile into oracle/_ref/libref_harness.so (git-ignored, never copied into the repo):
//   framework/calibration/frustum.cpp, calibration_inverter.cpp, nearest_neighbour_search.cpp, calibration_volume.hpp,
//   framework/rendering/volume_sampler.cpp, framework/DataTypes.cpp, external/gloost/{Matrix,Point3,Vector3,Ray,
//   BoundingBox,BoundingVolume}.cpp, external/squish/*.cpp (DXT codec) and the header-only glm 0.9.5.3.
// OpenGL / globjects / CGAL / boost are absent from this image: oracle/ref_stubs provides no-op GL and globjects
// declarations and an exact-kNN stand-in for CGAL's Orthogonal_k_neighbor_search (its header states the contract).
// This file only marshals arguments; it contains no algorithmic code of its own except ref_draw_uniforms, which calls
// the same gloost / glm functions in the same order as ReconIntegration::draw (recon_integration.cpp:183-206).
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#define private public   // read Frustum::m_planes for the plane-level comparison (layout is unchanged)
#include "frustum.hpp"
#undef private
#include "calibration_inverter.hpp"
#include "calibration_volume.hpp"
#include "volume_sampler.hpp"
#include <DataTypes.h>
#include <Matrix.h>
#include <BoundingBox.h>
#include <glm/gtc/matrix_inverse.hpp>
#include <glm/gtc/matrix_transform.hpp>
#include <squish.h>

using namespace kinect;

static std::array<glm::fvec3, 8> corners_of(const float* c) {
  std::array<glm::fvec3, 8> a;
  for (int i = 0; i < 8; ++i) a[i] = glm::fvec3{c[i * 3], c[i * 3 + 1], c[i * 3 + 2]};
  return a;
}

extern "C" {

// external/squish (the reference's CPU DXT codec, used at NetKinectArray.cpp:635): rgba uint8 [h][w][4] <-> DXT1 blocks
void ref_squish_compress_dxt1(const unsigned char* rgba, int w, int h, unsigned char* blocks) {
  squish::CompressImage(rgba, w, h, blocks, squish::kDxt1 | squish::kColourRangeFit);
}
void ref_squish_decompress_dxt1(const unsigned char* blocks, int w, int h, unsigned char* rgba) {
  squish::DecompressImage(rgba, w, h, blocks, squish::kDxt1);
}
int ref_squish_storage_dxt1(int w, int h) { return squish::GetStorageRequirements(w, h, squish::kDxt1); }
// the same for DXT5 streams (compress_rgb == 5, NetKinectArray.cpp:125-128)
void ref_squish_compress_dxt5(const unsigned char* rgba, int w, int h, unsigned char* blocks) {
  squish::CompressImage(rgba, w, h, blocks, squish::kDxt5 | squish::kColourRangeFit);
}
void ref_squish_decompress_dxt5(const unsigned char* blocks, int w, int h, unsigned char* rgba) {
  squish::DecompressImage(rgba, w, h, blocks, squish::kDxt5);
}
int ref_squish_storage_dxt5(int w, int h) { return squish::GetStorageRequirements(w, h, squish::kDxt5); }

// kinect::Frustum (frustum.cpp): planes float[6][4], camera position float[3]
void ref_frustum(const float* corners, float* planes_out, float* cam_out) {
  Frustum f{corners_of(corners)};
  for (int i = 0; i < 6; ++i) { planes_out[i * 4] = f.m_planes[i].x; planes_out[i * 4 + 1] = f.m_planes[i].y; planes_out[i * 4 + 2] = f.m_planes[i].z; planes_out[i * 4 + 3] = f.m_planes[i].w; }
  glm::fvec3 c = f.getCameraPos();
  cam_out[0] = c.x; cam_out[1] = c.y; cam_out[2] = c.z;
}

void ref_frustum_inside(const float* corners, const float* points, int n, int* out) {
  Frustum f{corners_of(corners)};
  for (int i = 0; i < n; ++i) out[i] = f.inside(glm::fvec3{points[i * 3], points[i * 3 + 1], points[i * 3 + 2]}) ? 1 : 0;
}

// CalibrationVolume<xyz>::write -> CalibrationInverter(files, bbox) -> calculateInverseVolumes -> writeInverseVolumes
// -> CalibrationVolume<fvec4>::read. dir must end with '/'. Returns 0 on success.
int ref_calib_invert(const char* dir, const float* cv_xyz, unsigned X, unsigned Y, unsigned Z, const float* bmin, const float* bmax,
                     const unsigned* out_res, float* out) {
  std::vector<xyz> vol((size_t)X * Y * Z);
  std::memcpy(vol.data(), cv_xyz, vol.size() * sizeof(xyz));
  const std::string base = std::string(dir) + "sensor0.";
  CalibrationVolume<xyz>{glm::uvec3{X, Y, Z}, glm::fvec2{0.5f, 4.5f}, vol}.write(base + "cv_xyz");
  gloost::BoundingBox bbox;
  bbox.setPMin(gloost::Point3{bmin[0], bmin[1], bmin[2]});
  bbox.setPMax(gloost::Point3{bmax[0], bmax[1], bmax[2]});
  CalibrationInverter inverter{std::vector<std::string>{base + "yml"}, bbox};
  inverter.calculateInverseVolumes(glm::uvec3{out_res[0], out_res[1], out_res[2]});
  inverter.writeInverseVolumes(std::string(dir));
  CalibrationVolume<glm::fvec4> inv{std::string(dir) + "sensor0.cv_xyz_inv"};
  if (inv.res() != glm::uvec3{out_res[0], out_res[1], out_res[2]}) return -1;
  if (inv.depthLimits() != glm::fvec2{0.5f, 4.5f}) return -2;
  std::memcpy(out, inv.volume().data(), inv.volume().size() * sizeof(glm::fvec4));
  return 0;
}

// CalibrationVolume<T> file round trip for T = xyz (3 floats), uv (2 floats), fvec4: writes then reads `path`.
int ref_volume_roundtrip(const char* path, int channels, const unsigned* res, const float* limits, const float* data, float* out,
                         unsigned* res_out, float* limits_out) {
  const size_t n = (size_t)res[0] * res[1] * res[2];
  const glm::uvec3 r{res[0], res[1], res[2]};
  const glm::fvec2 l{limits[0], limits[1]};
  if (channels == 3) {
    std::vector<xyz> v(n); std::memcpy(v.data(), data, n * sizeof(xyz));
    CalibrationVolume<xyz>{r, l, v}.write(path);
    CalibrationVolume<xyz> b{path};
    std::memcpy(out, b.volume().data(), n * sizeof(xyz));
    res_out[0] = b.res().x; res_out[1] = b.res().y; res_out[2] = b.res().z; limits_out[0] = b.depthLimits().x; limits_out[1] = b.depthLimits().y;
  } else if (channels == 2) {
    std::vector<uv> v(n); std::memcpy(v.data(), data, n * sizeof(uv));
    CalibrationVolume<uv>{r, l, v}.write(path);
    CalibrationVolume<uv> b{path};
    std::memcpy(out, b.volume().data(), n * sizeof(uv));
    res_out[0] = b.res().x; res_out[1] = b.res().y; res_out[2] = b.res().z; limits_out[0] = b.depthLimits().x; limits_out[1] = b.depthLimits().y;
  } else if (channels == 4) {
    std::vector<glm::fvec4> v(n); std::memcpy(v.data(), data, n * sizeof(glm::fvec4));
    CalibrationVolume<glm::fvec4>{r, l, v}.write(path);
    CalibrationVolume<glm::fvec4> b{path};
    std::memcpy(out, b.volume().data(), n * sizeof(glm::fvec4));
    res_out[0] = b.res().x; res_out[1] = b.res().y; res_out[2] = b.res().z; limits_out[0] = b.depthLimits().x; limits_out[1] = b.depthLimits().y;
  } else {
    return -1;
  }
  return 0;
}

// VolumeSampler::containedVoxels (volume_sampler.cpp:50-62): writes up to `cap` voxel indices, returns the count.
unsigned ref_contained_voxels(const unsigned* dims, const float* pos, const float* size, unsigned* out, unsigned cap) {
  VolumeSampler s{glm::uvec3{dims[0], dims[1], dims[2]}};
  std::vector<unsigned> idx = s.containedVoxels(glm::fvec3{pos[0], pos[1], pos[2]}, glm::fvec3{size[0], size[1], size[2]});
  for (size_t i = 0; i < idx.size() && i < cap; ++i) out[i] = idx[i];
  return (unsigned)idx.size();
}

// VolumeSampler voxel centres (volume_sampler.cpp:33-48): out float[X*Y*Z][3]
void ref_voxel_positions(const unsigned* dims, float* out) {
  VolumeSampler s{glm::uvec3{dims[0], dims[1], dims[2]}};
  std::memcpy(out, s.voxelPositions().data(), s.voxelPositions().size() * sizeof(glm::fvec3));
}

// kinect::getTrilinear (DataTypes.cpp:115-163), coordinates in voxel units
void ref_get_trilinear(const float* data, unsigned w, unsigned h, unsigned d, float x, float y, float z, float* out) {
  xyz r = getTrilinear(reinterpret_cast<xyz*>(const_cast<float*>(data)), w, h, d, x, y, z);
  out[0] = r.x; out[1] = r.y; out[2] = r.z;
}

float ref_glm_round(float x) { return glm::round(x); }
float ref_glm_distance(const float* a, const float* b) { return glm::distance(glm::fvec3{a[0], a[1], a[2]}, glm::fvec3{b[0], b[1], b[2]}); }

// The uniforms ReconIntegration::draw derives (recon_integration.cpp:183-206), with the reference's own gloost::Matrix
// and glm 0.9.5.3 arithmetic (single precision): out = img_to_eye[16], normal_matrix[16], camera_texturespace[3].
void ref_draw_uniforms(const float* modelview_in, const float* projection_in, const float* bmin, const float* bmax, unsigned vw, unsigned vh, float* out) {
  gloost::Matrix projection_matrix;
  std::memcpy(projection_matrix.data(), projection_in, 16 * sizeof(float));
  gloost::Matrix viewport_translate;
  viewport_translate.setIdentity();
  viewport_translate.setTranslate(1.0, 1.0, 1.0);
  gloost::Matrix viewport_scale;
  viewport_scale.setIdentity();
  viewport_scale.setScale(vw * 0.5, vh * 0.5, 0.5f);
  gloost::Matrix image_to_eye = viewport_scale * viewport_translate * projection_matrix;
  image_to_eye.invert();
  gloost::Matrix modelview;
  std::memcpy(modelview.data(), modelview_in, 16 * sizeof(float));
  glm::fmat4 model_view{modelview};
  // m_mat_vol_to_world exactly as recon_integration.cpp:66-72 builds it
  glm::fvec3 bbox_dimensions{bmax[0] - bmin[0], bmax[1] - bmin[1], bmax[2] - bmin[2]};
  glm::fvec3 bbox_translation{bmin[0], bmin[1], bmin[2]};
  glm::fmat4 vol_to_world = glm::scale(glm::fmat4{1.0f}, bbox_dimensions);
  vol_to_world = glm::translate(glm::fmat4{1.0f}, bbox_translation) * vol_to_world;
  glm::fmat4 normal_matrix = glm::inverseTranspose(model_view * vol_to_world);
  glm::fvec4 camera_world{glm::inverse(model_view) * glm::fvec4{0.0f, 0.0f, 0.0f, 1.0f}};
  glm::vec3 camera_texturespace{glm::inverse(vol_to_world) * camera_world};
  std::memcpy(out, image_to_eye.data(), 16 * sizeof(float));
  std::memcpy(out + 16, &normal_matrix[0][0], 16 * sizeof(float));
  out[32] = camera_texturespace.x; out[33] = camera_texturespace.y; out[34] = camera_texturespace.z;
}

}  // extern "C"

// calib_inverter <file.ks> [-s voxel_size]   — the reference's offline tool (source/calib_inverter.cpp:12-74) on the GPU:
// parse the .ks (kinect files + bbx), volume_res = ceil(bbox / voxel_size) (default 0.007 m), invert every sensor's
// cv_xyz, write <ks dir>/<basename>.cv_xyz_inv in the reference's file format.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iostream>

#include "rr_host.hpp"

int main(int argc, char** argv) {
  float voxel_size = 0.007f;                                   // default_voxel_size, calib_inverter.cpp:10
  std::string ks;
  for (int i = 1; i < argc; ++i) {
    if (!std::strcmp(argv[i], "-s") && i + 1 < argc) voxel_size = (float)std::atof(argv[++i]);
    else ks = argv[i];
  }
  if (ks.empty()) { std::cerr << "usage: calib_inverter <file.ks> [-s voxel_size]" << std::endl; return 1; }
  try {
    kinect::SceneFile sc = kinect::readSceneFile(ks);
    const float dims[3] = {sc.bbox.getPMax()[0] - sc.bbox.getPMin()[0], sc.bbox.getPMax()[1] - sc.bbox.getPMin()[1],
                           sc.bbox.getPMax()[2] - sc.bbox.getPMin()[2]};
    glm::uvec3 volume_res((unsigned)std::ceil(dims[0] / voxel_size), (unsigned)std::ceil(dims[1] / voxel_size), (unsigned)std::ceil(dims[2] / voxel_size));
    std::cout << "using resolution " << volume_res.x << ", " << volume_res.y << ", " << volume_res.z << std::endl;
    kinect::CalibrationFiles cfs(sc.calib_filenames, 16, 16, 16, 16);      // image sizes are irrelevant to the inversion
    kinect::gpu::Context gpu(0, cfs);
    kinect::CalibrationInverter inverter(sc.calib_filenames, sc.bbox);
    inverter.calculateInverseVolumes(volume_res);
    std::cout << "inverted " << sc.calib_filenames.size() << " volume(s) in " << inverter.lastGpuMilliseconds() << " ms (including the download)" << std::endl;
    inverter.writeInverseVolumes(sc.resource_path);
  } catch (std::exception const& e) {
    std::cerr << "calib_inverter: " << e.what() << std::endl;
    return 1;
  }
  return 0;
}

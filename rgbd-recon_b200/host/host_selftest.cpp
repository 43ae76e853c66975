// host_selftest — device-free checks of the host layer's file and wire formats: sensor .yml fields,
// .ks scene files, the server message layout and the feedback struct. Exit code 0 = all passed. Run by tests/test_host_cpp.py.
//   host_selftest <scratch directory>
#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>

#include "rr_host.hpp"

static int g_failed = 0;
#define EXPECT(cond)                                                                        \
  do {                                                                                      \
    if (!(cond)) { std::cerr << "FAILED " << __LINE__ << ": " #cond << std::endl; ++g_failed; } \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 2) { std::cerr << "usage: host_selftest <scratch directory>" << std::endl; return 1; }
  const std::string dir = std::string(argv[1]) + "/";
  using namespace kinect;

  // ---- a .yml as the reference's calibration tools write it (OpenCV-style lists); unknown keys must be skipped
  {
    std::ofstream y(dir + "k0.yml");
    y << "%YAML:1.0\nserial: 012345\nrgb_intrinsics: !!opencv-matrix\n   data: [ 1060., 0., 640., 0., 1060., 540., 0., 0., 1. ]\n"
      << "rgb_size: [ 1280, 1080 ]\ndepth_size: [ 512, 424 ]\nnear_far: [ 0.5, 4.5 ]\ncompress_rgb: [ 1, 0 ]\n"
      << "compress_depth: [ 1, 0 ]\nmin_length: [ 0.02, 0 ]\nT: [ 0.1, 0.2, 0.3 ]\n";
  }
  KinectCalibrationFile k(dir + "k0.yml");
  EXPECT(k.parse());
  EXPECT(k.getWidthC() == 1280 && k.getHeightC() == 1080 && k.getWidth() == 512 && k.getHeight() == 424);
  EXPECT(k.getNear() == 0.5f && k.getFar() == 4.5f);
  EXPECT(k.isCompressedRGB() == 1 && k.isCompressedDepth());
  EXPECT(std::fabs(k.min_length - 0.02f) < 1e-7f);
  KinectCalibrationFile missing(dir + "nope.yml");
  EXPECT(!missing.parse());
  // defaults of the reference's constructor when a key is absent (KinectCalibrationFile.cpp:88-95)
  { std::ofstream y(dir + "k1.yml"); y << "rgb_size: [ 64, 48 ]\ndepth_size: [ 32, 24 ]\n"; }
  KinectCalibrationFile d(dir + "k1.yml");
  EXPECT(d.parse() && d.getNear() == 0.3f && d.getFar() == 7.0f && d.isCompressedRGB() == 1 && !d.isCompressedDepth());

  CalibrationFiles cf({dir + "k0.yml", dir + "k1.yml"});      // sizes and formats of sensor 0 apply to all (calibration_files.cpp:26-33)
  EXPECT(cf.num() == 2 && cf.getWidth() == 512 && cf.getHeightC() == 1080 && cf.isCompressedRGB() == 1 && cf.isCompressedDepth());
  EXPECT(cf.getNear() == 0.5f && cf.getFar() == 4.5f);
  bool threw = false;
  try { CalibrationFiles bad({dir + "nope.yml"}); } catch (std::invalid_argument const&) { threw = true; }
  EXPECT(threw);

  // ---- .ks scene file (kinect_client.cpp:213-234)
  { std::ofstream s(dir + "scene.ks"); s << "serverport 127.0.0.1:7000\nkinect k0.yml\nkinect /abs/k1.yml\nbbx -1 0 -1.5 1 2.2 1.5\n"; }
  SceneFile sc = readSceneFile(dir + "scene.ks");
  EXPECT(sc.calib_filenames.size() == 2 && sc.calib_filenames[0] == dir + "k0.yml" && sc.calib_filenames[1] == "/abs/k1.yml");
  EXPECT(sc.bbox.getPMin()[2] == -1.5f && sc.bbox.getPMax()[1] == 2.2f && sc.resource_path == dir);

  // ---- server message: N x [colour | depth], first 8 bytes double as the frame time (NetKinectArray.cpp:511-538)
  {
    const unsigned N = 3; const std::size_t cs = 24, ds = 16;
    std::vector<uint8_t> msg((cs + ds) * N), col(cs * N), dep(ds * N);
    for (std::size_t i = 0; i < msg.size(); ++i) msg[i] = (uint8_t)(i * 7 + 3);
    const double stamp = 1234.5;
    std::memcpy(msg.data(), &stamp, sizeof(stamp));
    const double t = NetKinectArray::splitMessage(msg.data(), msg.size(), N, cs, ds, col.data(), dep.data());
    EXPECT(t == stamp);
    bool ok = true;
    for (unsigned i = 0; i < N; ++i) {
      ok = ok && !std::memcmp(col.data() + i * cs, msg.data() + i * (cs + ds), cs);
      ok = ok && !std::memcmp(dep.data() + i * ds, msg.data() + i * (cs + ds) + cs, ds);
    }
    EXPECT(ok);
    threw = false;
    try { NetKinectArray::splitMessage(msg.data(), msg.size() - 1, N, cs, ds, col.data(), dep.data()); } catch (std::invalid_argument const&) { threw = true; }
    EXPECT(threw);
  }

  // ---- feedback payload (FeedbackReceiver.h:16-22)
  {
    sys::feedback f{}, g{};
    for (int i = 0; i < 16; ++i) { f.cyclops_mat[i] = (float)i; f.screen_mat[i] = (float)(i * 2); f.model_mat[i] = (float)(i * 3); }
    f.recon_mode = 4; f.stream_slot = 1;
    EXPECT(sys::parseFeedback(&f, sizeof(f), g) && g.recon_mode == 4 && g.stream_slot == 1 && g.model_mat[15] == 45.0f);
    EXPECT(!sys::parseFeedback(&f, sizeof(f) - 4, g));
  }

  // ---- CalibrationVolume<T> file layout round trip (calibration_volume.hpp:13-84)
  {
    std::vector<xyz> data(2 * 3 * 4);
    for (std::size_t i = 0; i < data.size(); ++i) data[i] = xyz{(float)i, (float)i * 0.5f, -(float)i};
    CalibrationVolume<xyz> v(glm::uvec3(2, 3, 4), glm::fvec2(0.5f, 4.5f), data);
    v.write(dir + "v.cv_xyz");
    CalibrationVolume<xyz> r(dir + "v.cv_xyz");
    EXPECT(r.res().x == 2 && r.res().y == 3 && r.res().z == 4 && r.depthLimits().y == 4.5f && r.volume().size() == data.size() && r(1, 2, 3).z == -23.0f);
  }

  if (g_failed) { std::cerr << g_failed << " check(s) failed" << std::endl; return 1; }
  std::cout << "host_selftest ok" << std::endl;
  return 0;
}

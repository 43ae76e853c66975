"""The GL-free C++ host layer (rgbd-recon_b200/host/) and its two programs, driven the way the reference's programs are:
.ks scene file + .cv_xyz/.cv_uv volumes + .stream files in, .cv_xyz_inv / fused volume / image out."""
import os
import subprocess

import numpy as np
import pytest

from conftest import bits_equal, mismatch_report

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "rgbd-recon_b200", "bin")


def write_scene_files(d, scene, frames=None):
    """The on-disk inputs kinect_client / calib_inverter expect."""
    from rrpy import volume_io
    ks = ["serverport 127.0.0.1:7000"]
    for i in range(scene.N):
        open(os.path.join(d, f"sensor{i}.yml"), "w").write("# parsed by the reference's KinectCalibrationFile; unused here\n")
        volume_io.write_volume(os.path.join(d, f"sensor{i}.cv_xyz"), scene.cv_xyz[i])
        volume_io.write_volume(os.path.join(d, f"sensor{i}.cv_uv"), scene.cv_uv[i])
        ks.append(f"kinect sensor{i}.yml")
    ks.append("bbx " + " ".join(repr(float(v)) for v in list(scene.bbox_min) + list(scene.bbox_max)))
    ks_path = os.path.join(d, "scene.ks")
    open(ks_path, "w").write("\n".join(ks) + "\n")
    streams = []
    for i in range(scene.N):
        p = os.path.join(d, f"sensor{i}.stream")
        with open(p, "wb") as f:
            for fr in (frames or [scene]):
                f.write(fr.color[i].tobytes())
                f.write(fr.depth[i].tobytes())
        streams.append(p)
    return ks_path, streams


def test_host_file_and_wire_formats(tmp_path):
    """SURVEY.md 8f-3 without a device: sensor .yml fields, .ks files, server message layout, feedback payload, volume files."""
    r = subprocess.run([os.path.join(BIN, "host_selftest"), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0 and "host_selftest ok" in r.stdout, r.stderr


def test_natural_neighbour_interpolator_properties():
    """kinect::NaturalNeighbourInterpolator without CGAL (host C++). Sibson coordinates are checked through
    what defines them: an affine field is reproduced to float rounding on scattered samples AND on a regular grid (every
    Delaunay cell degenerate), the coordinates reproduce the query position, a cell centre gets its eight corners with weight
    1/8, a sample returns itself, outside the hull there are no coordinates, and the volumes match a Monte Carlo count."""
    r = subprocess.run([os.path.join(BIN, "nni_selftest")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "nni_selftest ok" in r.stdout, r.stderr[-2000:]


def test_programs_are_built_and_report_usage():
    for name in ("calib_inverter", "fusion_playback", "host_selftest"):
        exe = os.path.join(BIN, name)
        assert os.path.exists(exe), "build first: make -C rgbd-recon_b200"
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 1 and "usage" in r.stderr


def test_scene_file_errors_are_reported(tmp_path):
    r = subprocess.run([os.path.join(BIN, "calib_inverter"), str(tmp_path / "missing.ks")], capture_output=True, text=True)
    assert r.returncode == 1 and "cannot open scene file" in r.stderr
    (tmp_path / "empty.ks").write_text("serverport x\n")
    r = subprocess.run([os.path.join(BIN, "calib_inverter"), str(tmp_path / "empty.ks")], capture_output=True, text=True)
    assert r.returncode == 1 and "no 'kinect" in r.stderr


def test_no_gpu_means_loud_failure(tmp_path, small_scene):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    ks, _ = write_scene_files(str(tmp_path), small_scene)
    r = subprocess.run([os.path.join(BIN, "calib_inverter"), ks, "-s", "0.1"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_calib_inverter_program_matches_oracle(tmp_path, small_scene):
    import oracle_py as O
    from rrpy import volume_io
    ks, _ = write_scene_files(str(tmp_path), small_scene)
    r = subprocess.run([os.path.join(BIN, "calib_inverter"), ks, "-s", "0.05"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "using resolution 40, 44, 40" in r.stdout
    for i in range(small_scene.N):
        got, lim = volume_io.read_volume(str(tmp_path / f"sensor{i}.cv_xyz_inv"), 4)
        assert lim == (0.5, 4.5)
        want = O.calib_invert(small_scene.cv_xyz[i], small_scene.bbox_min, small_scene.bbox_max, (40, 44, 40))
        assert bits_equal(got, want).all(), mismatch_report(f"sensor{i}.cv_xyz_inv", got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("devices", [None, "0,0,0", "all"])
def test_fusion_playback_program_matches_oracle(tmp_path, small_scene, devices):
    """The C++ look-alike program against the oracle; with --devices the same run goes through rr_group (z-slabs over several
    devices in one process, here also three slabs sharing GPU 0) and must give the same volume and image bit for bit."""
    import oracle_py as O
    extra = []
    if devices == "all":
        import torch
        if torch.cuda.device_count() < 2:
            pytest.skip("one GPU: the multi-device run is covered by the shared-device case")
        extra = ["--gpus", str(min(4, torch.cuda.device_count()))]
    elif devices:
        extra = ["--devices", devices]
    from rrpy import synth, volume_io
    sc = small_scene
    ks, streams = write_scene_files(str(tmp_path), sc)
    inv = synth.analytic_inverse(sc, (50, 55, 50))
    for i in range(sc.N):
        volume_io.write_volume(str(tmp_path / f"sensor{i}.cv_xyz_inv"), inv[i])
    VW, VH = 320, 180
    mv, pr = synth.look_at((1.6, 1.5, 2.2), (0.0, 1.1, 0.0)), synth.perspective(50.0, VW / VH, 0.1, 10.0)
    np.concatenate([mv, pr]).astype(np.float32).tofile(str(tmp_path / "view.bin"))
    r = subprocess.run([os.path.join(BIN, "fusion_playback"), ks, "--depth", str(sc.W), str(sc.H), "--color", str(sc.CW), str(sc.CH),
                        "--streams", ";".join(streams), "--frames", "3", "--voxel", "0.02", "--view", str(VW), str(VH), "--shade", "1",
                        "--matrices", str(tmp_path / "view.bin"), "--dump-tsdf", str(tmp_path / "tsdf.bin"), "--dump-image", str(tmp_path / "img.bin")] + extra,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "2integrate mean ms" in r.stdout
    assert ("z-slabs over" in r.stdout) == bool(extra)
    grid = O.brick_grid(sc.bbox_min, sc.bbox_max, 0.02, 0.1)
    cams = [O.frustum(sc.cv_xyz[i])[1] for i in range(sc.N)]
    pre = O.preprocess(sc, grid, cams)
    occ = O.occupied_bricks(pre["bricks"], 10)
    want = O.integrate(inv, pre, grid, 0.01, True, occ)
    got = np.fromfile(str(tmp_path / "tsdf.bin"), np.float32).reshape(want.shape)
    assert bits_equal(got, want).all(), mismatch_report("tsdf", got, want)
    img = np.fromfile(str(tmp_path / "img.bin"), np.float32)
    rgba, depth = img[:VW * VH * 4].reshape(VH, VW, 4), img[VW * VH * 4:].reshape(VH, VW)
    rm = O.raymarch(want, 0.01, inv, sc, pre, grid, occ, mv, pr, VW, VH, 1, True)
    assert (rm["depth"] < 1).sum() > 500
    # m_fill_holes is on by default (recon_integration.cpp:54): drawF() = raymarch + fillColors
    filled = O.fill_colors(rm["rgba"], rm["depth"])
    assert bits_equal(depth, rm["depth"]).all() and bits_equal(rgba, filled).all()
    ratio = float(r.stdout.split("occupied ratio")[1].split()[0])
    assert abs(ratio - len(occ) / grid["num_bricks"]) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["points", "trigrid", "calibs"])
def test_fusion_playback_other_reconstructions(tmp_path, small_scene, mode):
    """SURVEY.md 8f-4 through the C++ look-alike classes (ReconPoints, ReconTrigrid, ReconCalibs in host/rr_host.hpp): the
    program draws the last frame set with the named reconstruction, like kinect_client switches its g_recons, and the image
    equals the oracle's bit for bit."""
    import oracle_py as O
    from rrpy import synth, volume_io
    sc = small_scene
    ks, streams = write_scene_files(str(tmp_path), sc)
    for i in range(sc.N):                                     # sizes, formats and min_length arrive the reference's way: the sensors' .yml
        open(os.path.join(str(tmp_path), f"sensor{i}.yml"), "w").write(
            f"serial: 00{i}\nrgb_size: [ {sc.CW}, {sc.CH} ]\ndepth_size: [ {sc.W}, {sc.H} ]\nnear_far: [ 0.5, 4.5 ]\n"
            "compress_rgb: [ 0, 0 ]\ncompress_depth: [ 0, 0 ]\nmin_length: [ 0.06, 0 ]\n")
    inv = synth.analytic_inverse(sc, (40, 44, 40))
    for i in range(sc.N):
        volume_io.write_volume(str(tmp_path / f"sensor{i}.cv_xyz_inv"), inv[i])
    VW, VH = 240, 136
    mv, pr = synth.look_at((1.4, 1.5, 2.0), (0.0, 1.1, 0.0)), synth.perspective(50.0, VW / VH, 0.1, 10.0)
    np.concatenate([mv, pr]).astype(np.float32).tofile(str(tmp_path / "view.bin"))
    r = subprocess.run([os.path.join(BIN, "fusion_playback"), ks,
                        "--streams", ";".join(streams), "--frames", "2", "--voxel", "0.02", "--view", str(VW), str(VH), "--shade", "1",
                        "--matrices", str(tmp_path / "view.bin"), "--recon", mode, "--dump-recon", str(tmp_path / "recon.bin")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert f"recon {mode} covered pixels" in r.stdout
    img = np.fromfile(str(tmp_path / "recon.bin"), np.float32)
    rgba, depth = img[:VW * VH * 4].reshape(VH, VW, 4), img[VW * VH * 4:].reshape(VH, VW)
    grid = O.brick_grid(sc.bbox_min, sc.bbox_max, 0.02, 0.1)
    cams = [O.frustum(sc.cv_xyz[i])[1] for i in range(sc.N)]
    pre = O.preprocess(sc, grid, cams)
    if mode == "points":
        want = O.draw_points(sc, pre, mv, pr, VW, VH, shade_mode=1)
    elif mode == "trigrid":
        want = O.draw_trigrid(sc, pre, mv, pr, VW, VH, shade_mode=1, min_length=0.06)
    else:
        tsdf = O.integrate(inv, pre, grid, 0.01, True, O.occupied_bricks(pre["bricks"], 10))
        want = O.draw_calibs(tsdf, (40, 44, 40), 0.01, sc.bbox_min, sc.bbox_max, mv, pr, VW, VH)
    assert (want[1] < 1.0).sum() > 300
    assert bits_equal(depth, want[1]).all(), mismatch_report(f"{mode} depth", depth, want[1])
    assert bits_equal(rgba, want[0]).all(), mismatch_report(f"{mode} rgba", rgba, want[0])


@pytest.mark.gpu
def test_fusion_playback_from_yml_and_server_messages(tmp_path):
    """The reference's default stream format end to end through the host layer: sizes, DXT1 colour, 8-bit depth and its
    near/far range come from the sensors' .yml (CalibrationFiles), the frames arrive as server messages
    (N x [colour | depth], NetKinectArray.cpp:511-538) through NetKinectArray::pushMessage."""
    import dataclasses
    import oracle_py as O
    from rrpy import synth, volume_io
    sc = synth.make_scene(N=2, W=128, H=106, CW=160, CH=136, cv_res=(32, 32, 64))
    ks, _ = write_scene_files(str(tmp_path), sc)
    for i in range(sc.N):
        open(os.path.join(str(tmp_path), f"sensor{i}.yml"), "w").write(
            f"serial: 00{i}\nrgb_size: [ {sc.CW}, {sc.CH} ]\ndepth_size: [ {sc.W}, {sc.H} ]\nnear_far: [ 0.5, 4.5 ]\n"
            "compress_rgb: [ 1, 0 ]\ncompress_depth: [ 1, 0 ]\n")
    inv = synth.analytic_inverse(sc, (40, 44, 40))
    for i in range(sc.N):
        volume_io.write_volume(str(tmp_path / f"sensor{i}.cv_xyz_inv"), inv[i])
    dxt = [synth.encode_dxt1(sc.color[i]) for i in range(sc.N)]
    d8 = [synth.encode_depth8(sc.depth[i], 0.5, 4.5) for i in range(sc.N)]
    msg = b"".join(dxt[i].tobytes() + d8[i].tobytes() for i in range(sc.N))
    (tmp_path / "messages.bin").write_bytes(msg * 2)
    r = subprocess.run([os.path.join(BIN, "fusion_playback"), ks, "--messages", str(tmp_path / "messages.bin"), "--frames", "3",
                        "--voxel", "0.025", "--view", "64", "36", "--dump-tsdf", str(tmp_path / "tsdf.bin")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert f"depth {sc.W}x{sc.H} 8-bit" in r.stdout and "DXT1" in r.stdout and "near/far 0.5 4.5" in r.stdout
    color = np.stack([O.decode_dxt1(dxt[i], sc.CW, sc.CH) for i in range(sc.N)])
    depth = O.depth8_to_float(np.stack(d8))
    osc = dataclasses.replace(sc, color=color, depth=depth)
    near_far = np.float32([[0.5, 4.5]] * sc.N)
    grid = O.brick_grid(sc.bbox_min, sc.bbox_max, 0.025, 0.1)
    cams = [O.frustum(sc.cv_xyz[i])[1] for i in range(sc.N)]
    pre = O.preprocess(osc, grid, cams, True, False, True, compress=near_far)
    occ = O.occupied_bricks(pre["bricks"], 10)
    want = O.integrate(inv, pre, grid, 0.01, True, occ)
    got = np.fromfile(str(tmp_path / "tsdf.bin"), np.float32).reshape(want.shape)
    assert len(occ) > 20
    assert bits_equal(got, want).all(), mismatch_report("tsdf", got, want)
    stamp = np.frombuffer(msg[:8], np.float64)[0]
    assert f"last frame time {stamp:g}"[:20] in r.stdout or "last frame time" in r.stdout

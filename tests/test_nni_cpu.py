"""kinect::NaturalNeighbourInterpolator next to an independent computation of the Sibson coordinates.

The reference delegates to CGAL 's sibson_natural_neighbor_coordinates_3 (framework/NaturalNeighbourInterpolator.cpp:35-47), which
is absent here. Sibson coordinates are a definition, not an algorithm: lambda_i(q) = vol(V_q  cut out of  V_i) / vol(V_q), with
V_i the Voronoi cell of sample i before q is inserted and V_q the cell of q afterwards. This test computes exactly that with Qhull
(scipy.spatial.Voronoi + ConvexHull: cell volumes before and after the insertion) and compares it with the coordinates the C++
class reports (bisector clipping of a box, no triangulation): two formulations that share no code."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "rgbd-recon_b200", "bin", "nni_selftest")


def _cell_volumes(points, want):
    """Volumes of the (bounded) Voronoi cells of points[want]."""
    from scipy.spatial import ConvexHull, Voronoi
    vor = Voronoi(points)
    out = {}
    for i in want:
        region = vor.regions[vor.point_region[i]]
        assert region and -1 not in region, "cell must be bounded (the shell of far sites guarantees it for the inner sites)"
        out[i] = ConvexHull(vor.vertices[region]).volume
    return out


def _sibson_qhull(sites, q, inner):
    before = _cell_volumes(sites, inner)
    after = _cell_volumes(np.vstack([sites, q[None]]), list(inner) + [len(sites)])
    vq = after[len(sites)]
    lam = np.zeros(len(sites))
    for i in inner:
        lam[i] = (before[i] - after[i]) / vq
    return lam


@pytest.mark.parametrize("kind", ["scattered", "jittered_grid"])
def test_sibson_coordinates_match_qhull(tmp_path, kind):
    pytest.importorskip("scipy")
    assert os.path.exists(EXE), "build first: make -C rgbd-recon_b200"
    rng = np.random.default_rng(7 if kind == "scattered" else 11)
    if kind == "scattered":
        inner = rng.uniform(0.0, 1.0, (160, 3))
    else:                                   # calibration samples sit on grids: a grid with a little jitter (Qhull needs general position)
        g = np.stack(np.meshgrid(*[np.linspace(0.0, 1.0, 6)] * 3, indexing="ij"), -1).reshape(-1, 3)
        inner = g + rng.uniform(-0.02, 0.02, g.shape)
    # a shell of sites far outside: every inner site's Voronoi cell is bounded, before and after the insertion
    t = np.linspace(-3.0, 4.0, 5)
    shell = np.array([(x, y, z) for x in t for y in t for z in t if max(abs(x - 0.5), abs(y - 0.5), abs(z - 0.5)) > 3.0])
    sites = np.vstack([inner, shell]).astype(np.float32)
    queries = rng.uniform(0.3, 0.7, (12, 3)).astype(np.float32)
    (tmp_path / "sites.bin").write_bytes(sites.tobytes())
    (tmp_path / "queries.bin").write_bytes(queries.tobytes())
    r = subprocess.run([EXE, "--coords", str(tmp_path / "sites.bin"), str(tmp_path / "queries.bin"), str(tmp_path / "out.bin")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    got = np.fromfile(str(tmp_path / "out.bin"), np.float64).reshape(len(queries), len(sites))
    idx_inner = list(range(len(inner)))
    for k, q in enumerate(queries):
        want = _sibson_qhull(sites.astype(np.float64), q.astype(np.float64), idx_inner)
        assert abs(got[k].sum() - 1.0) < 1e-9 and abs(want.sum() - 1.0) < 1e-6      # the query's neighbours are all inner sites
        assert got[k, len(inner):].max() < 1e-12
        assert np.abs(got[k] - want).max() < 1e-7, f"query {k}: Sibson coordinates differ by {np.abs(got[k] - want).max():.3g}"
        assert (want > 1e-6).sum() >= 4
        # the local coordinate property both must have: sum lambda_i p_i = q
        assert np.abs(got[k] @ sites.astype(np.float64) - q).max() < 1e-6

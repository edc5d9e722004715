"""The oracle against third-party code that is in this image (scipy, numpy): independent implementations of the
pieces of the path that have one.  The reference itself cannot be built here (ROS + PCL), so these are the
nearest thing to an outside check of the restatement: the levelling convention (src:159-167) against
scipy.spatial.transform, Euclidean clustering (src:269-276) against a KD-tree + connected components from
scipy, the 3DSC support neighbourhood (src:348-350) against scipy's ball query, the elevation formula
(src:147-156) against numpy's evaluation of the literal expression, the crop (src:169-183) against a mask.
Test data hold no pair within 2e-6 of a radius (such points are removed first), so float-vs-double distance
arithmetic cannot decide an outcome."""
import numpy as np
import pytest

scipy_spatial = pytest.importorskip("scipy.spatial")
from scipy.sparse import coo_matrix  # noqa: E402
from scipy.sparse.csgraph import connected_components  # noqa: E402
from scipy.spatial.transform import Rotation  # noqa: E402


def _cloud(seed, n_blobs=60, per=40, spread=0.12, box=20.0):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-box, box, (n_blobs, 3))
    c[:, 2] = rng.uniform(-1.0, 1.0, n_blobs)
    p = c[:, None, :] + spread * rng.normal(size=(n_blobs, per, 3))
    noise = np.concatenate([rng.uniform(-box, box, (400, 2)), rng.uniform(-1, 1, (400, 1))], axis=1)
    xyz = np.concatenate([p.reshape(-1, 3), noise]).astype(np.float32)
    xyz = xyz[rng.permutation(len(xyz))]
    return np.concatenate([xyz, np.zeros((len(xyz), 1), np.float32)], axis=1)


def _drop_pairs_near(cloud, r, band=2e-6):
    """Remove one end of every pair whose distance is within `band` of r (float d^2 and double d differ by ~3e-8
    there): what is left cannot be decided by the arithmetic."""
    while True:
        t = scipy_spatial.cKDTree(cloud[:, :3].astype(np.float64))
        near = t.query_pairs(r + band) - t.query_pairs(r - band)
        if not near:
            return cloud
        cloud = np.delete(cloud, sorted({a for a, _ in near}), axis=0)


def test_levelling_is_Ry_pitch_times_Rx_roll_as_scipy_composes_it(ob):
    rng = np.random.default_rng(3)
    pts = np.zeros((2000, 4), np.float32)
    pts[:, :3] = rng.uniform(-50, 50, (2000, 3))
    for roll, pitch in ((0.03, -0.02), (-0.4, 0.25), (1.2, -0.9), (0.0, 0.0)):
        got = ob.rotate_cloud(pts, roll, pitch)
        # Eigen: AngleAxisf(pitch, UnitY) * AngleAxisf(roll, UnitX) applied to a column vector = Ry(Rx v)
        R = Rotation.from_euler("y", pitch) * Rotation.from_euler("x", roll)
        want = R.apply(pts[:, :3].astype(np.float64))
        assert np.abs(got[:, :3] - want).max() < 2e-5 * 50   # float matrix entries and float accumulation
        assert np.array_equal(got[:, 3], pts[:, 3])


@pytest.mark.parametrize("seed,tol,lo,hi", [(1, 0.3, 5, 500), (2, 0.25, 1, 10000), (5, 0.45, 10, 60)])
def test_euclidean_clusters_are_the_components_of_scipys_radius_graph(ob, seed, tol, lo, hi):
    cloud = _drop_pairs_near(_cloud(seed), tol)
    xyz = cloud[:, :3].astype(np.float64)
    pairs = np.array(sorted(scipy_spatial.cKDTree(xyz).query_pairs(tol)), np.int64).reshape(-1, 2)
    n = len(xyz)
    g = coo_matrix((np.ones(len(pairs)), (pairs[:, 0], pairs[:, 1])), shape=(n, n))
    _, lab = connected_components(g, directed=False)
    want = {}
    for i, l in enumerate(lab):
        want.setdefault(l, []).append(i)
    want = {frozenset(v) for v in want.values() if lo <= len(v) <= hi}
    for mode in (0, 1):  # brute force and the oracle's own KD-tree
        got = ob.extract_clusters(cloud, tol, lo, hi, mode=mode)
        assert {frozenset(int(i) for i in c) for c in got} == want
        sizes = [len(c) for c in got]
        assert sizes == sorted(sizes, reverse=True)        # src:275-276 hands them out largest first


def test_3dsc_support_neighbourhood_is_scipys_ball_query(ob):
    P = ob.node_default()
    cloud = _cloud(7, n_blobs=80, per=60, spread=0.5)
    xyz = cloud[:, :3].astype(np.float64)
    rng = np.random.default_rng(11)
    kp = cloud[rng.choice(len(cloud), 25, replace=False)].copy()
    kp[:, :3] += rng.normal(scale=0.05, size=(25, 3)).astype(np.float32)
    R = P.descriptor_radius
    t = scipy_spatial.cKDTree(xyz)
    inner = t.query_ball_point(kp[:, :3].astype(np.float64), R - 1e-5, return_length=True)
    outer = t.query_ball_point(kp[:, :3].astype(np.float64), R + 1e-5, return_length=True)
    assert np.array_equal(inner, outer)                    # nothing on the boundary
    for mode in (0, 1):
        d, margin, nn = ob.estimate_descriptors(P, cloud, kp, mode=mode)
        assert np.array_equal(nn, inner)
        assert np.isfinite(d[nn > 0]).all() and np.isnan(d[nn == 0]).all()


def test_elevation_equals_numpys_evaluation_of_the_literal_formula(ob):
    rng = np.random.default_rng(4)
    pts = np.zeros((200000, 4), np.float32)
    pts[:, :3] = rng.uniform(-80, 80, (200000, 3))
    pts[:1000, 2] = 0.0
    pts[1000:2000, 1] = 0.0
    got = ob.get_elevation_angles(pts)[:, 3]
    x, y, z = (pts[:, k].astype(np.float64) for k in range(3))
    az = np.arctan2(y, x)
    want = (np.arctan2(z, np.cos(az) * x + np.sin(az) * y) * 180.0 / 3.14159265358979323846).astype(np.float32)
    # numpy calls the same libm; a last-ulp difference of a vectorised routine may flip a float in 1e-7 of the cases
    assert (got.view(np.uint32) != want.view(np.uint32)).mean() < 1e-5
    assert np.abs(got - want).max() < 1e-5


def test_crop_is_three_inclusive_interval_masks_in_order(ob):
    P = ob.node_default()
    rng = np.random.default_rng(8)
    pts = np.zeros((50000, 4), np.float32)
    pts[:, :3] = rng.uniform(-40, 40, (50000, 3))
    pts[:, 3] = np.arange(50000, dtype=np.float32)
    pts[::97, 0] = np.float32(P.x_max)                     # on a limit: kept (PassThrough is inclusive)
    pts[::89, 1] = np.nan
    got = ob.filter_cloud(P, pts)
    x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
    with np.errstate(invalid="ignore"):
        m = ((z >= np.float32(P.z_min)) & (z <= np.float32(P.z_max)) & (y >= np.float32(P.y_min)) & (y <= np.float32(P.y_max)) &
             (x >= np.float32(P.x_min)) & (x <= np.float32(P.x_max)))
    assert np.array_equal(got.view(np.uint32), pts[m].view(np.uint32))

"""Self-consistency of the oracle — the reference has no tests of its own."""
import numpy as np
import pytest


def _pole(x, y, r, n_az=7, rings=(5, 6, 7, 8, 9), jitter=0.0, seed=0):
    """points on the sensor-facing side of a vertical cylinder, one arc per ring; intensity = the
    ring's elevation (what getElevationAngles would have written)."""
    rng = np.random.default_rng(seed)
    d = np.hypot(x, y)
    pts = []
    for ring in rings:
        el = (ring - 7) * 2 - 1
        z = d * np.tan(np.deg2rad(el))
        for a in np.linspace(-1.0, 1.0, n_az):
            px = x - r * np.cos(a) + jitter * rng.normal()
            py = y + r * np.sin(a) + jitter * rng.normal()
            pts.append((px, py, z, el))
    return np.array(pts, np.float32)


def test_single_pole_gives_one_keypoint_on_its_axis(ob):
    P = ob.node_default()
    cloud = _pole(10.0, 2.0, 0.08)
    # interleave rings so ring selection has to be stable
    cloud = cloud.reshape(5, 7, 4).transpose(1, 0, 2).reshape(-1, 4).copy()
    kp, kc, kf = ob.estimate_keypoints(P, cloud)
    assert len(kf) == 5           # one centroid per ring
    assert len(kp) == 1           # stacked vertically -> one keypoint
    assert abs(kp[0, 0] - (10.0 - 0.08 * np.mean(np.cos(np.linspace(-1, 1, 7))))) < 1e-3
    assert abs(kp[0, 1] - 2.0) < 2e-3
    assert kp[0, 3] == np.float32(-5.0)  # intensity of indices[0]: the lowest ring's first point
    assert len(kc) == 35


def test_trunk_is_rejected_by_the_diameter_gate(ob):
    P = ob.node_default()
    cloud = _pole(10.0, 0.0, 0.25, n_az=9)   # xy bbox diagonal > 2*0.15
    kp, kc, kf = ob.estimate_keypoints(P, cloud)
    assert len(kf) == 0 and len(kp) == 0 and len(kc) == 0


def test_min_and_max_cluster_size_gate_whole_components(ob):
    line = np.zeros((60, 4), np.float32)
    line[:, 0] = 5.0 + 0.3 * np.arange(60)      # one chain of 60 points, spacing 0.3 < 0.65
    assert len(ob.extract_clusters(line, 0.65, 5, 50)) == 0      # oversize: dropped whole, not split
    assert len(ob.extract_clusters(line[:50], 0.65, 5, 50)) == 1
    assert len(ob.extract_clusters(line[:4], 0.65, 5, 50)) == 0


def test_strict_threshold(ob):
    r2f = ob.radius_sq_float(0.65)
    d = np.float32(np.sqrt(np.float64(r2f)))
    # find the largest float spacing whose float square is still < r2f
    while np.float32(d * d) >= r2f:
        d = np.nextafter(d, np.float32(0))
    a = np.zeros((2, 4), np.float32)
    a[1, 0] = d
    assert len(ob.extract_clusters(a, 0.65, 2, 10)) == 1
    a[1, 0] = np.nextafter(d, np.float32(1))
    if np.float32(a[1, 0] * a[1, 0]) >= r2f:
        assert len(ob.extract_clusters(a, 0.65, 2, 10)) == 0


def test_brute_force_and_kdtree_agree(ob, synth):
    pts, offs, rp = synth.generate(2, 3)
    P = ob.node_default()
    for s in range(3):
        sc = pts[offs[s]:offs[s + 1]]
        a = ob.process_scan(P, sc, rp[s, 0], rp[s, 1], mode=0)
        b = ob.process_scan(P, sc, rp[s, 0], rp[s, 1], mode=1)
        for k in ("keypoints", "keypoint_cloud", "cloud", "cloud_full", "descriptors"):
            assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), k


def test_cluster_sets_invariant_under_permutation(ob, synth):
    pts, offs, rp = synth.generate(2, 1)
    P = ob.node_default()
    r = ob.process_scan(P, pts, rp[0, 0], rp[0, 1], mode=1)
    ring = ob.select_ring(r["cloud"], 6)
    assert len(ring) > 10
    base = ob.extract_clusters(ring, 0.65, 1, 100000, mode=1)
    rng = np.random.default_rng(3)
    perm = rng.permutation(len(ring))
    other = ob.extract_clusters(ring[perm], 0.65, 1, 100000, mode=1)
    as_sets = lambda cl, m: sorted(tuple(sorted(int(m[i]) for i in c)) for c in cl)
    assert as_sets(base, np.arange(len(ring))) == as_sets(other, perm)


def test_ring_windows_share_their_end_points(ob):
    c = np.zeros((3, 4), np.float32)
    c[:, 0] = [5, 6, 7]
    c[:, 3] = [-14.0, -13.0, 16.5]
    assert len(ob.select_ring(c, 0)) == 1 and len(ob.select_ring(c, 1)) == 2
    assert sum(len(ob.select_ring(c, i)) for i in range(16)) == 3  # -14 twice, -13 once, 16.5 never


def test_bbox_sentinels_reproduce_the_reference_quirk(ob):
    # minx starts at +1000 / maxx at -1000 (src:289-290): beyond +-1000 m the diameter is wrong
    P = ob.node_default()
    P.x_max = 5000.0
    c = np.zeros((5, 4), np.float32)
    c[:, 0] = 2000.0 + 0.01 * np.arange(5)
    cen, cc = ob.get_cylinder_segments(P, c)
    assert len(cen) == 0  # |1000 - (-1000)| style diameter fails the gate although the cluster is tiny
    c[:, 0] -= 1990.0
    cen, cc = ob.get_cylinder_segments(P, c)
    assert len(cen) == 1


def test_descriptor_sum_identity(ob, synth):
    pts, offs, rp = synth.generate(1, 1)
    P = ob.launch_playback()
    r = ob.process_scan(P, pts, rp[0, 0], rp[0, 1], mode=0)
    assert len(r["keypoints"]) > 0
    radii, theta, phi, lut = ob.sc3d_tables(P.descriptor_radius)
    d = r["descriptors"]
    assert np.all(np.isfinite(d)) and np.all(d >= 0)
    # every contribution is lut[bin]/rho with rho >= 1: a bin never exceeds n_neighbors * lut[bin]
    assert np.all(d <= r["n_neighbors"][:, None] * lut[None, :] * (1 + 1e-5))


def test_empty_inputs_return_empty(ob):
    P = ob.node_default()
    e = np.zeros((0, 4), np.float32)
    r = ob.process_scan(P, e, 0.0, 0.0)
    assert len(r["keypoints"]) == 0 and len(r["cloud"]) == 0
    assert ob.extract_clusters(e, 0.65, 1, 10) == []
    kp, kc, kf = ob.estimate_keypoints(P, e)
    assert len(kp) == 0


def test_zero_neighbour_keypoint_is_nan_and_draws_nothing(ob):
    P = ob.node_default()
    cloud = np.array([[10, 0, 0, 0], [10.2, 0, 0.1, 0]], np.float32)
    kps = np.array([[50, 0, 0, 0], [10.1, 0, 0, 0]], np.float32)
    d, m, nn = ob.estimate_descriptors(P, cloud, kps)
    assert np.all(np.isnan(d[0])) and nn[0] == 0
    d2, _, _ = ob.estimate_descriptors(P, cloud, kps[1:])
    assert np.array_equal(d[1].view(np.uint32), d2[0].view(np.uint32))  # same RNG draws (first triple)


def test_elevation_precedes_rotation(ob):
    # intensity is the SENSOR-frame elevation (src:87 before src:92)
    p = np.array([[10, 0, 10 * np.tan(np.deg2rad(5.0)), 0]], np.float32)
    el = ob.get_elevation_angles(p)
    assert abs(el[0, 3] - 5.0) < 1e-5
    rot = ob.rotate_cloud(el, 0.3, -0.2)
    assert rot[0, 3] == el[0, 3]
    assert not np.allclose(rot[0, :3], el[0, :3])


def test_boundary_report_is_the_same_from_both_search_back_ends(ob, synth):
    """The north star's "within 1e-6 m of a tolerance boundary" report: brute force sees every pair, the
    KD-tree prunes with a margin wider than the band, so both count the same pairs; a pair built to sit
    4e-7 m inside the cluster tolerance is reported, one 1 mm inside is not."""
    P = ob.node_default()
    pts, offs, rp = synth.generate(2, 6, scan_index_base=8200)
    for eps in (1e-6, 1e-4):
        a = ob.process_batch_boundary(P, pts, offs, rp, eps, mode=0, n_threads=4)
        b = ob.process_batch_boundary(P, pts, offs, rp, eps, mode=1, n_threads=4)
        assert a.shape == (6, 4) and np.array_equal(a, b)
    assert a.sum() > 0
    P.cluster_min_count = 1
    t = np.tan(np.deg2rad(-1.0))
    for gap, expect in ((0.6499996, True), (0.6489996, False)):
        xs = [2.0 + 0.01 * k for k in range(5)] + [2.04 + gap / np.sqrt(1 + t * t)]
        sc = np.array([(x, 0.5, x * t, 0.0) for x in xs], np.float32)
        for mode in (0, 1):
            r = ob.process_batch_boundary(P, sc, np.array([0, len(sc)], np.int64), np.zeros((1, 2)), 1e-6, mode=mode)
            assert (r[0, 0] >= 1) == expect, (gap, mode, r)


def test_oracle_exports_the_product_symbols(ob, synth):
    """the oracle's shared object exports the same create / process / destroy / last_error /
    version symbols as the product library, with the same structs — a harness written against
    include/fe_b200.h can be pointed at either .so."""
    import ctypes as C
    import os
    from feature_extraction_b200 import _native as N   # struct definitions only; the product .so is not called here
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    L = C.CDLL(os.path.join(root, "oracle", "libfe_oracle.so"))
    L.fe_version.restype = C.c_char_p
    L.fe_last_error.restype = C.c_char_p
    L.fe_create.argtypes = [C.c_int, C.POINTER(N.Params), C.POINTER(N.Limits), C.POINTER(C.c_void_p)]
    L.fe_process_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(N.BatchResult)]
    L.fe_destroy.argtypes = [C.c_void_p]
    assert b"oracle" in L.fe_version() and L.fe_device_count() == 0
    P = N.Params()
    L.fe_params_node_default(C.byref(P))
    ctx = C.c_void_p()
    assert L.fe_create(0, C.byref(P), None, C.byref(ctx)) == 0
    pts, offs, rp = synth.generate(2, 4, scan_index_base=31)
    rpf = np.ascontiguousarray(rp, np.float64).reshape(-1)
    res = N.BatchResult()
    assert L.fe_process_batch(ctx, pts.ctypes.data_as(C.c_void_p), offs.ctypes.data_as(C.c_void_p), rpf.ctypes.data_as(C.c_void_p), 4, C.byref(res)) == 0
    K = int(res.n_keypoints)
    ko = np.ctypeslib.as_array(res.keypoint_offsets, shape=(5,)).copy()
    kp = np.ctypeslib.as_array(C.cast(res.keypoints, C.POINTER(C.c_float)), shape=(K, 4)).copy()
    d = np.ctypeslib.as_array(C.cast(res.descriptors, C.POINTER(C.c_float)), shape=(K, 1980)).copy()
    L.fe_destroy(ctx)
    ko_o, kp_o, d_o, _ = ob.process_batch(ob.node_default(), pts, offs, rp, mode=1)
    assert np.array_equal(ko, ko_o) and np.array_equal(kp.view(np.uint32), kp_o.view(np.uint32))
    assert np.array_equal(d.view(np.uint32), d_o.view(np.uint32))

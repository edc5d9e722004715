"""Edge cases the reference handles by early return (src:209-210, 234-235, 263-264, 278-279,
331-332) and size-independent properties at BASELINE.json's full batch size (needs a B200)."""
import numpy as np
import pytest

from util import bits_equal, check_descriptors, to_fe_params

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def node(ob):
    from feature_extraction_b200 import FeatureExtractionNode
    n = FeatureExtractionNode(to_fe_params(ob.node_default()), max_points=8 << 20, max_scans=512, max_keypoints=1 << 15)
    yield n
    n.close()


def test_empty_batch_and_empty_scans(ob, synth, node):
    e = np.zeros((0, 4), np.float32)
    ko, kp, d = node.processBatch(e, np.zeros(1, np.int64), np.zeros((0, 2)))
    assert list(ko) == [0] and len(kp) == 0
    ko, kp, d = node.processBatch(e, np.zeros(4, np.int64), np.zeros((3, 2)))
    assert list(ko) == [0, 0, 0, 0] and len(kp) == 0 and d.shape == (0, 1980)
    # ragged: empty scans between real ones
    pts, offs, rp = synth.generate(2, 3)
    offs2 = np.array([0, 0, offs[1], offs[1], offs[2], offs[3], offs[3]], np.int64)
    rp2 = np.array([[0, 0], rp[0], [0, 0], rp[1], rp[2], [0, 0]])
    ko, kp, d = node.processBatch(pts, offs2, rp2)
    ko_o, kp_o, d_o, m = ob.process_batch(ob.node_default(), pts, offs2, rp2, mode=0)
    assert np.array_equal(ko, ko_o) and bits_equal(kp, kp_o)
    assert check_descriptors(d, d_o)[1] == 0
    for fn in (node.getElevationAngles, node.rotateCloud, node.filterCloud):
        assert len(fn(e)) == 0
    assert node.extractClusters(e, 0.65, 1, 10) == []
    assert all(len(x) == 0 for x in node.getCylinderSegments(e))
    assert all(len(x) == 0 for x in node.estimateKeypoints(e))
    assert node.estimateDescriptors(e, e).shape == (0, 1980)


def test_non_finite_points_and_out_of_crop(ob, synth, node):
    pts, offs, rp = synth.generate(2, 1, scan_index_base=7)
    pts = pts.copy()
    pts[5] = [np.nan, 1, 1, 0]
    pts[100] = [np.inf, 0, 0, 0]
    pts[200] = [1, -np.inf, 0, 0]
    pts[300] = [0, 0, 0, 0]           # atan2(0,0)
    pts[400] = [1e6, 1e6, 1e6, 0]
    P = ob.node_default()
    r = ob.process_scan(P, pts, rp[0, 0], rp[0, 1], mode=0)
    ko, kp, d = node.processBatch(pts, offs, rp)
    assert bits_equal(kp, r["keypoints"])
    assert check_descriptors(d, r["descriptors"])[1] == 0
    el_g = node.getElevationAngles(pts)
    el_o = ob.get_elevation_angles(pts)
    fin = np.isfinite(el_o[:, 3])
    assert bits_equal(el_g[fin], el_o[fin]) and np.array_equal(np.isnan(el_g[:, 3]), np.isnan(el_o[:, 3]))
    assert bits_equal(node.filterCloud(r["cloud_full"]), r["cloud"])
    # nothing inside the crop box -> no keypoints, not an error
    far = pts.copy()
    far[:, 0] -= 500.0
    ko, kp, d = node.processBatch(far, offs, np.zeros((1, 2)))
    assert ko[-1] == 0


def test_point_on_a_ring_boundary_belongs_to_both_rings(ob, node):
    # elevation exactly -14.0 lies in ring 0 ([-16,-14]) and ring 1 ([-14,-12]) (src:200-202)
    P = ob.node_default()
    c = []
    for k in range(6):
        c.append((10.0 + 0.02 * k, 1.0, -2.0, -15.0))
    for k in range(6):
        c.append((10.0 + 0.02 * k, 1.0, -1.6, -13.0))
    c.append((10.05, 1.0, -1.8, -14.0))
    c = np.array(c, np.float32)
    kp_o, kc_o, kf_o = ob.estimate_keypoints(P, c)
    kp_g, kc_g = node.estimateKeypoints(c)
    assert len(kf_o) == 2 and len(kc_o) == 14  # the boundary point is dumped twice
    assert bits_equal(kp_g, kp_o) and bits_equal(kc_g, kc_o)


def test_boundary_points_overflow_the_whole_scan_gather_and_fall_back_to_ring_groups(ob, node):
    # 1200 crop survivors fit the fast ring-clustering capacity (1408 entries), but 720 of them sit on
    # ring boundaries and count twice: the uncounted whole-scan gather must give way to ring groups
    rng = np.random.default_rng(5)
    c = []
    for k in range(240):
        cx, cy = rng.uniform(5, 70), rng.uniform(-28, 28)
        elev = [-14.0, -12.0, -15.0, -10.0, -13.0][k % 5]  # 3 of 5 clusters are on a boundary
        for j in range(5):
            c.append((cx + 0.03 * j, cy + rng.uniform(-0.02, 0.02), rng.uniform(-1.0, 2.0) * 0.01, elev))
    c = np.array(c, np.float32)
    P = ob.node_default()
    kp_o, kc_o, kf_o = ob.estimate_keypoints(P, c)
    kp_g, kc_g = node.estimateKeypoints(c)
    assert len(kf_o) > 200
    assert bits_equal(kp_g, kp_o) and bits_equal(kc_g, kc_o)


def test_zero_neighbour_keypoint_gives_nan_descriptor(ob, node):
    P = ob.node_default()
    cloud = np.array([[10, 0, 0, 0], [10.2, 0, 0.1, 0], [10.1, 0.3, 0.0, 0]], np.float32)
    kps = np.array([[50, 0, 0, 0], [10.1, 0, 0, 0], [np.nan, 0, 0, 0], [10.0, 0.1, 0.05, 0]], np.float32)
    d_o, m, nn = ob.estimate_descriptors(P, cloud, kps)
    d_g = node.estimateDescriptors(cloud, kps)
    assert np.all(np.isnan(d_g[0])) and np.all(np.isnan(d_g[2]))
    assert check_descriptors(d_g, d_o)[1] == 0
    # neighbour straight above the keypoint: NaN azimuth lands in bin l=0 on both sides
    cloud2 = np.array([[10, 0, 1.0, 0], [10.3, 0.1, 0.0, 0]], np.float32)
    kp2 = np.array([[10, 0, 0, 0]], np.float32)
    d_o, m, nn = ob.estimate_descriptors(P, cloud2, kp2)
    assert bits_equal(node.estimateDescriptors(cloud2, kp2), d_o)


def test_capacity_overflow_is_an_error_not_truncation(ob, synth):
    from feature_extraction_b200 import FeatureExtractionNode, FeatureExtractionError, _native
    pts, offs, rp = synth.generate(2, 4)
    small = FeatureExtractionNode(to_fe_params(ob.node_default()), max_points=20000, max_scans=4, max_keypoints=1)
    with pytest.raises(FeatureExtractionError) as e:
        small.processBatch(pts, offs, rp)
    assert e.value.status == _native.FE_ERR_CAPACITY
    small.close()
    tiny = FeatureExtractionNode(to_fe_params(ob.node_default()), max_points=4096, max_scans=4, max_keypoints=64)
    with pytest.raises(FeatureExtractionError) as e:
        tiny.processBatch(pts, offs, rp)
    assert e.value.status == _native.FE_ERR_CAPACITY
    tiny.close()


def test_idempotence_and_identity_properties(ob, synth, node):
    pts, offs, rp = synth.generate(2, 1, scan_index_base=3)
    node.roll, node.pitch = 0.0, 0.0
    el = node.getElevationAngles(pts)
    assert bits_equal(node.rotateCloud(el)[:, :3] + 0.0, el[:, :3] + 0.0)  # identity rotation (up to -0)
    f1 = node.filterCloud(el)
    assert bits_equal(node.filterCloud(f1), f1)                             # crop is idempotent
    assert bits_equal(node.getElevationAngles(el), el)                      # el depends on xyz only


def test_full_batch_size_properties(ob, synth):
    """BASELINE.json config 2 at full size (10k scans): sub-batching must not change any result,
    results are ordered by scan, and a random sample agrees with the oracle."""
    from feature_extraction_b200 import FeatureExtractionNode
    B = 10000
    pts, offs, rp = synth.generate(2, B, scan_index_base=10_000)
    P = ob.node_default()
    a = FeatureExtractionNode(to_fe_params(P), max_points=48 << 20, max_scans=3000, max_keypoints=1 << 17)
    ko_a, kp_a, d_a = a.processBatch(pts, offs, rp)
    a.close()
    b = FeatureExtractionNode(to_fe_params(P), max_points=6 << 20, max_scans=257, max_keypoints=1 << 15)
    ko_b, kp_b, d_b = b.processBatch(pts, offs, rp)
    assert np.array_equal(ko_a, ko_b) and bits_equal(kp_a, kp_b)
    from util import rel_err
    assert rel_err(d_a, d_b).max() < 1e-5           # only the float atomics' summation order differs
    assert np.all(np.diff(ko_a) >= 0) and ko_a[-1] == len(kp_a) > B // 2
    rng = np.random.default_rng(0)
    sample = np.sort(rng.choice(B, 48, replace=False))
    for s in sample:
        r = ob.process_scan(P, pts[offs[s]:offs[s + 1]], rp[s, 0], rp[s, 1], mode=1)
        assert bits_equal(kp_a[ko_a[s]:ko_a[s + 1]], r["keypoints"]), s
        assert check_descriptors(d_a[ko_a[s]:ko_a[s + 1]], r["descriptors"])[1] == 0
    # a scan's result does not depend on its neighbours in the batch: reversed batch order
    sel = sample[::-1]
    p2 = np.concatenate([pts[offs[s]:offs[s + 1]] for s in sel])
    o2 = np.concatenate([[0], np.cumsum([offs[s + 1] - offs[s] for s in sel])]).astype(np.int64)
    ko_c, kp_c, d_c = b.processBatch(p2, o2, rp[sel])
    for i, s in enumerate(sel):
        assert bits_equal(kp_c[ko_c[i]:ko_c[i + 1]], kp_a[ko_a[s]:ko_a[s + 1]])
    b.close()


def test_elevation_fast_path_rounds_like_the_double_evaluation(ob, node):
    """Row A on 3M points: sensor-like geometry plus arbitrary directions and magnitudes.  The fast
    path accepts a value only when its float rounding is not in doubt, so the result must equal the
    oracle's double evaluation bit for bit (a 1-ulp difference is possible only through libm's own
    double rounding, ~1e-8 per point: at most a couple of points here)."""
    rng = np.random.default_rng(42)
    n = 1_000_000
    az = rng.uniform(-np.pi, np.pi, n)
    el = np.deg2rad(rng.choice(np.arange(-15, 16, 2), n) + rng.normal(0, 0.05, n))
    r = rng.uniform(0.5, 100.0, n)
    a = np.stack([r * np.cos(el) * np.cos(az), r * np.cos(el) * np.sin(az), r * np.sin(el), np.zeros(n)], 1)
    b = rng.normal(0, 1, (n, 4)) * np.exp(rng.uniform(-20, 20, (n, 1)))       # any direction, any scale
    c = rng.normal(0, 30, (n, 4))
    c[:1000, 0] = 0.0                                                          # on the axes
    c[1000:2000, 1] = 0.0
    c[2000:3000, 2] = 0.0
    c[3000:3100, :2] = 0.0
    pts = np.concatenate([a, b, c]).astype(np.float32)
    pts[:, 3] = 0
    g = node.getElevationAngles(pts)
    o = ob.get_elevation_angles(pts)
    diff = np.flatnonzero(g[:, 3].view(np.uint32) != o[:, 3].view(np.uint32))
    assert len(diff) <= 3, (len(diff), g[diff[:5]], o[diff[:5]])
    if len(diff):
        assert np.all(np.abs(g[diff, 3] - o[diff, 3]) <= np.spacing(np.abs(o[diff, 3])))
    assert bits_equal(g[:, :3], pts[:, :3])


def test_oversized_rings_take_the_global_memory_path(ob, node):
    """A ring with more cropped returns than any shared-memory instantiation holds (6144) is not an
    error: the clustering falls through to the instantiation whose arrays live in global memory."""
    rng = np.random.default_rng(5)
    P = ob.node_default()
    # ~9000 returns in ONE ring (elevation -1 deg): 900 small clumps of 10 points on a grid in the crop box
    pts = []
    for c in range(900):
        cx, cy = 2.0 + 1.6 * (c % 45), -28.0 + 2.8 * (c // 45)
        for k in range(10):
            pts.append((cx + 0.02 * k + 0.003 * rng.normal(), cy + 0.003 * rng.normal(), 0.1 * rng.normal(), -1.0))
    pts = np.array(pts, np.float32)
    pts = pts[rng.permutation(len(pts))]
    kp_o, kc_o, kf_o = ob.estimate_keypoints(P, pts, mode=1)
    kp_g, kc_g = node.estimateKeypoints(pts)
    assert len(kf_o) > 500
    assert bits_equal(kp_g, kp_o) and bits_equal(kc_g, kc_o)
    co = ob.extract_clusters(pts, 0.65, 5, 50, mode=1)
    cg = node.extractClusters(pts, 0.65, 5, 50)
    assert len(co) == len(cg) and all(np.array_equal(a, b) for a, b in zip(co, cg))


def test_descriptors_off_and_parameter_changes_on_a_live_context(ob, synth):
    """estimate_descriptors=false (src:112) skips the 3DSC stage; fe_set_params re-derives thresholds,
    grids and tables on an existing context (the reference reads its members on every callback)."""
    from feature_extraction_b200 import FeatureExtractionNode
    pts, offs, rp = synth.generate(2, 6, scan_index_base=40)
    P = ob.node_default()
    nd = FeatureExtractionNode(to_fe_params(P), max_points=1 << 20, max_scans=16, max_keypoints=4096)
    ko1, kp1, d1 = nd.processBatch(pts, offs, rp)
    P.estimate_descriptors = 0
    nd.set_params(to_fe_params(P))
    ko, kp, d = nd.processBatch(pts, offs, rp)
    assert d is None and np.array_equal(ko, ko1) and bits_equal(kp, kp1)
    # launch preset on the same context, then back: results follow the parameters both times
    L = ob.launch_playback()
    nd.set_params(to_fe_params(L))
    ko, kp, d = nd.processBatch(pts, offs, rp)
    ko_o, kp_o, d_o, m = ob.process_batch(L, pts, offs, rp, mode=1, n_threads=4)
    assert np.array_equal(ko, ko_o) and bits_equal(kp, kp_o) and check_descriptors(d, d_o)[1] == 0
    P.estimate_descriptors = 1
    P.descriptor_radius = 1.5
    nd.set_params(to_fe_params(P))
    ko, kp, d = nd.processBatch(pts, offs, rp)
    ko_o, kp_o, d_o, m = ob.process_batch(P, pts, offs, rp, mode=1, n_threads=4)
    assert np.array_equal(ko, ko_o) and bits_equal(kp, kp_o) and check_descriptors(d, d_o)[1] == 0
    nd.close()


def test_non_finite_imu_state_yields_no_keypoints(ob, synth, node):
    """roll/pitch are uninitialised in the reference until the first IMU message: a NaN
    rotation makes every point non-finite, which the crop drops — empty result, no error."""
    pts, offs, rp = synth.generate(2, 2, scan_index_base=77)
    bad = np.array([[np.nan, 0.0], [0.0, np.inf]])
    ko, kp, d = node.processBatch(pts, offs, bad)
    ko_o, kp_o, d_o, _ = ob.process_batch(ob.node_default(), pts, offs, bad, mode=1)
    assert ko[-1] == 0 and ko_o[-1] == 0


def test_random_parameter_sets_match_the_oracle(ob, synth):
    """Derived constants (grid sizes and caps, key widths, trusted-cell flag, surface box, shape-context
    tables) must hold for arbitrary ROS parameter values, not just the two presets."""
    from feature_extraction_b200 import FeatureExtractionNode
    rng = np.random.default_rng(2024)
    scans = {2: synth.generate(2, 2, scan_index_base=600), 3: synth.generate(3, 1, scan_index_base=601)}
    nd = None
    for trial in range(14):
        P = ob.node_default()
        P.x_min = float(rng.uniform(-60, 5)); P.x_max = float(P.x_min + rng.uniform(8, 120))
        P.y_min = float(rng.uniform(-60, -2)); P.y_max = float(rng.uniform(2, 60))
        P.z_min = float(rng.uniform(-2.2, -0.5)); P.z_max = float(rng.uniform(0.5, 8))
        P.cluster_tolerance = float(rng.choice([0.05, 0.2, 0.65, 1.0, 2.0]))
        P.cluster_min_count = int(rng.integers(1, 8)); P.cluster_max_count = int(rng.choice([10, 50, 400, 2000]))
        P.cluster_radius_threshold = float(rng.uniform(0.05, 0.6))
        P.number_detection_channels = int(rng.integers(1, 5))
        P.descriptor_radius = float(rng.choice([0.3, 1.0, 2.5, 4.0, 6.0]))
        cfg = 3 if trial % 4 == 3 else 2
        pts, offs, rp = scans[cfg]
        if nd is None:
            nd = FeatureExtractionNode(to_fe_params(P), max_points=1 << 21, max_scans=8, max_keypoints=1 << 14)
        else:
            nd.set_params(to_fe_params(P))
        ko, kp, d = nd.processBatch(pts, offs, rp)
        ko_o, kp_o, d_o, m = ob.process_batch(P, pts, offs, rp, mode=1, n_threads=8)
        tag = (trial, P.cluster_tolerance, P.cluster_radius_threshold, P.descriptor_radius, int(ko_o[-1]))
        assert np.array_equal(ko, ko_o), tag
        assert bits_equal(kp, kp_o), tag
        assert check_descriptors(d, d_o)[1] == 0, tag
    nd.close()


def test_very_large_scan_takes_the_deferred_paths(ob, synth):
    """One scan with > 65,535 surface points (beyond the 16-bit cell counters of the counting-sort K4a)
    and a symmetric crop: the deferred instantiations (radix K4a, large K2) must give the same result."""
    from feature_extraction_b200 import FeatureExtractionNode
    pts, offs, rp = synth.generate(3, 1, scan_index_base=77, azimuth_steps=14400)
    assert len(pts) > 100000
    P = ob.node_default()
    P.x_min = -75.0
    nd = FeatureExtractionNode(to_fe_params(P), max_points=1 << 20, max_scans=4, max_keypoints=1 << 14)
    ko, kp, d = nd.processBatch(pts, offs, rp)
    # device-side counters of this call come from the host path's slot 0
    r = ob.process_scan(P, pts, rp[0, 0], rp[0, 1], mode=1)
    assert len(r["cloud_full"]) == len(pts)
    assert bits_equal(kp, r["keypoints"])
    assert check_descriptors(d, r["descriptors"])[1] == 0
    nd.close()


def test_bbox_sentinels_and_extreme_size_gates_follow_the_reference(ob, node):
    """src:289-290 initialises the cluster box with +-1000: a cluster beyond 1000 m keeps the sentinel as
    one of its bounds and fails the diameter gate; a cluster straddling nothing special passes.  Also the
    size gate at its extremes (min 1: single returns are clusters; max smaller than a wall run)."""
    from feature_extraction_b200 import FeatureExtractionNode
    P = ob.node_default()
    P.x_min, P.x_max, P.y_min, P.y_max = -2000.0, 2000.0, -2000.0, 2000.0
    P.cluster_min_count, P.cluster_max_count = 1, 6
    rng = np.random.default_rng(21)
    pts = []
    # every cluster sits a little below the sensor plane, so its elevation is in ring 7's window [-2, 0] at any range
    for cx, cy in ((1500.0, 3.0), (-1500.0, -7.0), (30.0, 1200.0), (12.0, -1100.0), (999.95, 0.0), (1000.02, 40.0), (20.0, 4.0), (25.0, -3.0)):
        for k in range(4):
            pts.append((cx + 0.03 * k, cy + 0.01 * rng.normal(), -0.3 - 0.01 * k, 0.0))
    for k in range(9):                      # a run of 9 returns: above max_count, dropped whole
        pts.append((40.0 + 0.2 * k, 10.0, -0.4, 0.0))
    pts.append((55.0, -20.0, -0.5, 0.0))    # a single return: a cluster of its own (min 1)
    pts = np.array(pts, np.float32)
    offs = np.array([0, len(pts)], np.int64)
    rp = np.zeros((1, 2))
    nd = FeatureExtractionNode(to_fe_params(P), max_points=1 << 16, max_scans=4, max_keypoints=1024)
    nd.enableCloudOutputs(True)
    ko, kp, d = nd.processBatch(pts, offs, rp)
    co, cloud, kco, kcloud = nd.cloudOutputs(1)
    nd.close()
    r = ob.process_scan(P, pts, 0.0, 0.0, mode=0)
    assert bits_equal(kp, r["keypoints"]) and bits_equal(kcloud, r["keypoint_cloud"]) and bits_equal(cloud, r["cloud"])
    assert check_descriptors(d, r["descriptors"])[1] == 0
    xs, ys = r["keypoints"][:, 0], r["keypoints"][:, 1]
    assert len(r["cloud"]) == len(pts)
    assert not np.any(np.abs(xs) > 1400.0) and not np.any(np.abs(ys) > 1000.0)   # far clusters keep a sentinel bound: gate fails
    assert np.any(np.abs(xs - 1000.065) < 0.01)      # ... except just beyond 1000 m, where min x = 1000 still gives a small box
    assert np.any(np.abs(xs - 55.0) < 1e-3)          # the single return made it through
    assert not np.any(np.abs(xs - 40.8) < 1.0)       # the oversize run did not

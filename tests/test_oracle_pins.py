"""Known-answer pins of the oracle (SURVEY.md §8c [PROBE] constants).

The reference ships no tests or golden vectors and PCL/Eigen/FLANN/Boost are absent, so these
constants — measured from libstdc++/glibc during the survey — are the only hard pins there are.
"""
import numpy as np


def test_mt19937_seed_12345_stream(ob):
    raw, f = ob.mt19937_draws(12345, 9)
    assert list(raw) == [3992670690, 3823185381, 1358822685, 561383553, 789925284, 170765737,
                         878579710, 3549516158, 2438360421]
    want = np.array([0.929616094, 0.890154719, 0.316375554, 0.130707294, 0.183918819, 0.0397594981,
                     0.20456028, 0.826436102, 0.567725003], np.float32)
    assert np.array_equal(f, want)


def test_radius_thresholds_as_flann_sees_them(ob):
    assert float.hex(float(ob.radius_sq_float(0.65))) == "0x1.b0a3d40000000p-2"
    assert ob.radius_sq_float(0.65) == np.float32(0.422499955)
    assert ob.radius_sq_float(1.0) == np.float32(1.0)
    assert float.hex(float(ob.radius_sq_float(0.15))) == "0x1.70a3d80000000p-6"
    assert float.hex(float(ob.radius_sq_float(0.2))) == "0x1.47ae160000000p-5"
    assert ob.radius_sq_float(2.5, narrow_first=False) == np.float32(6.25)
    assert ob.radius_sq_float(2.5 / 5.0, narrow_first=False) == np.float32(0.25)
    assert np.float32(0.65) == np.float32(0.649999976)


def test_shape_context_edges_R_2_5(ob):
    radii, theta, phi, lut = ob.sc3d_tables(2.5)
    want = np.array([0.25, 0.291478604, 0.339839101, 0.396223307, 0.461962461, 0.53860867, 0.62797159,
                     0.732161164, 0.853637278, 0.995267987, 1.16039729, 1.35292387, 1.57739341,
                     1.83910573, 2.14423966, 2.5], np.float32)
    assert np.array_equal(radii, want)
    assert theta[0] == 0 and theta[11] == np.float32(180.0)
    assert phi[0] == 0 and phi[12] == np.float32(360.0)
    assert np.all(np.isfinite(lut)) and np.all(lut > 0)
    # the table does not depend on the azimuth bin
    assert np.array_equal(lut[:165], lut[165:330])


def test_std_sort_reverse_is_stable_size_desc_up_to_16(ob):
    rng = np.random.default_rng(1)
    for _ in range(200):
        n = int(rng.integers(1, 17))
        sizes = rng.integers(1, 6, n).astype(np.int32)
        order = ob.std_sort_cluster_order(sizes)
        stable = np.argsort(-sizes, kind="stable")
        assert np.array_equal(order, stable)


def test_std_sort_reverse_is_not_stable_beyond_16(ob):
    rng = np.random.default_rng(2)
    diff = 0
    for _ in range(200):
        n = int(rng.integers(17, 120))
        sizes = rng.integers(1, 6, n).astype(np.int32)
        order = ob.std_sort_cluster_order(sizes)
        assert np.array_equal(np.sort(order), np.arange(n))
        assert np.all(np.diff(sizes[order]) <= 0)
        diff += int(not np.array_equal(order, np.argsort(-sizes, kind="stable")))
    assert diff > 150  # 0/200 match at n >= 17


def test_param_presets(ob):
    d = ob.node_default()
    assert (d.x_min, d.x_max, d.y_min, d.y_max, d.z_min, d.z_max) == (0.0, 75.0, -30.0, 30.0, -1.5, 5.0)
    assert (d.cluster_tolerance, d.cluster_min_count, d.cluster_max_count) == (0.65, 5, 50)
    assert (d.cluster_radius_threshold, d.number_detection_channels, d.descriptor_radius) == (0.15, 1, 2.5)
    l = ob.launch_playback()
    assert (l.x_max, l.y_min, l.y_max, l.z_max) == (100.0, -50.0, 50.0, 4.0)
    assert (l.cluster_tolerance, l.cluster_min_count, l.cluster_max_count) == (1.0, 1, 1000)
    assert (l.cluster_radius_threshold, l.number_detection_channels) == (0.2, 2)

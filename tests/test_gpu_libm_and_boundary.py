"""Round-2 parity items (needs a B200):
  * the device's atan2f / acosf are the host libm's, bit for bit (row L of SURVEY.md §8a: the 3DSC bin of
    a neighbour hangs off the last bit of those two calls, reference src:353 -> PCL 3dsc.hpp);
  * the tolerance-boundary report of BASELINE.json's north star ("points lying within 1e-6 m of a
    tolerance boundary reported separately"): the device's counts equal the oracle's;
  * error paths flagged by the round-1 review: capacity checks before any write, a context stays usable
    after a failed call, shape validation in the Python mirror."""
import platform

import numpy as np
import pytest

from util import bits_equal, check_descriptors, to_fe_params

pytestmark = pytest.mark.gpu


def _params(ob, cfg):
    P = ob.launch_playback() if cfg == 1 else ob.node_default()
    if cfg == 4:
        P.descriptor_radius = 5.0
    return P


@pytest.fixture(scope="module")
def node(ob):
    from feature_extraction_b200 import FeatureExtractionNode
    n = FeatureExtractionNode(to_fe_params(ob.node_default()), max_points=8 << 20, max_scans=512, max_keypoints=1 << 15)
    yield n
    n.close()


def _fdlibm_host():
    name, ver = platform.libc_ver()
    try:
        return name == "glibc" and tuple(int(x) for x in ver.split(".")[:2]) < (2, 41)
    except ValueError:
        return False


@pytest.mark.skipif(not _fdlibm_host(), reason="host libm is not the fdlibm-based glibc (<= 2.40)")
def test_device_atan2f_and_acosf_are_the_host_libm_bit_for_bit(ob, node):
    rng = np.random.default_rng(7)
    n = 1 << 23
    # acosf: uniform in [-1, 1], every float within 4096 ulps of the cosine of a 3DSC elevation edge, and
    # raw bit patterns (|x| > 1, NaN, denormals)
    edges = np.cos(np.arange(12) * np.pi / 11).astype(np.float32)
    near = (edges.view(np.uint32)[:, None] + np.arange(-4096, 4097, dtype=np.int64)[None, :]).astype(np.uint32).reshape(-1).view(np.float32)
    a = np.concatenate([rng.uniform(-1, 1, n).astype(np.float32), near, rng.integers(0, 1 << 32, n // 4, dtype=np.uint64).astype(np.uint32).view(np.float32),
                        np.array([0.0, -0.0, 1.0, -1.0, 0.5, -0.5, np.nan, np.inf, 1e-30, 2.0 ** -26, 2.0 ** -27], np.float32)])
    got, want = node.debugLibm(1, a), ob.libm_f32(1, a)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    assert np.array_equal(got[ok].view(np.uint32), want[ok].view(np.uint32))
    # atan2f as 3DSC calls it — (norm of a cross product, dot product) of two unit vectors — plus every
    # float within 4096 ulps of an azimuth edge (k*30 deg), arbitrary magnitudes and raw bit patterns
    ang = rng.uniform(-np.pi, np.pi, n)
    y = np.abs(np.sin(ang)).astype(np.float32)
    x = np.cos(ang).astype(np.float32)
    ea = np.arange(13) * np.pi / 6
    ex = (np.cos(ea).astype(np.float32).view(np.uint32)[:, None] + np.arange(-4096, 4097, dtype=np.int64)[None, :]).astype(np.uint32).view(np.float32)
    ey = np.repeat(np.abs(np.sin(ea)).astype(np.float32)[:, None], ex.shape[1], 1)
    rb = rng.integers(0, 1 << 32, (2, n // 4), dtype=np.uint64).astype(np.uint32).view(np.float32)
    yy = np.concatenate([y, ey.reshape(-1), rng.uniform(0, 1, n).astype(np.float32), rb[0], np.array([0.0, 0.0, 1.0, 0.0, -0.0], np.float32)])
    xx = np.concatenate([x, ex.reshape(-1), rng.uniform(-1, 1, n).astype(np.float32), rb[1], np.array([1.0, -1.0, 0.0, 0.0, -1.0], np.float32)])
    got, want = node.debugLibm(0, yy, xx), ob.libm_f32(0, yy, xx)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    assert np.array_equal(got[ok].view(np.uint32), want[ok].view(np.uint32))
    # atanf over raw bit patterns
    t = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32).view(np.float32)
    got, want = node.debugLibm(2, t), ob.libm_f32(2, t)
    ok = ~np.isnan(want)
    assert np.array_equal(got[ok].view(np.uint32), want[ok].view(np.uint32))


def test_correctly_rounded_mode_is_a_parameter_and_differs_only_on_bin_edges(ob, synth):
    """fe_set_angle_libm(FE_LIBM_CORRECTLY_ROUNDED) — for a host with glibc >= 2.41 — moves at most a
    handful of contributions by one bin; the default mode is the one that equals the oracle exactly."""
    from feature_extraction_b200 import FeatureExtractionNode
    P = ob.node_default()
    pts, offs, rp = synth.generate(2, 24, scan_index_base=8100)
    nd = FeatureExtractionNode(to_fe_params(P), max_points=1 << 21, max_scans=32, max_keypoints=4096)
    ko, kp, d = nd.processBatch(pts, offs, rp)
    nd.setAngleLibm(True)
    ko2, kp2, d2 = nd.processBatch(pts, offs, rp)
    nd.setAngleLibm(False)
    ko3, kp3, d3 = nd.processBatch(pts, offs, rp)
    nd.close()
    assert bits_equal(kp, kp2) and bits_equal(kp, kp3) and bits_equal(d, d3)
    differing_rows = int((d.view(np.uint32) != d2.view(np.uint32)).any(axis=1).sum())
    assert differing_rows <= max(2, len(d) // 100)


@pytest.mark.parametrize("cfg,nscans", [(1, 4), (2, 32), (3, 2), (4, 3)])
def test_boundary_report_equals_the_oracles(ob, synth, cfg, nscans):
    """Pairs within 1e-6 m of each radius predicate (ring clustering, cross-ring merge, 3DSC support,
    3DSC density), per scan: device audit kernels vs the oracle's own searches.  The report changes no
    result."""
    from feature_extraction_b200 import FeatureExtractionNode
    P = _params(ob, cfg)
    pts, offs, rp = synth.generate(cfg, nscans, scan_index_base=8200)
    nd = FeatureExtractionNode(to_fe_params(P), max_points=1 << 22, max_scans=64, max_keypoints=1 << 14)
    ko0, kp0, d0 = nd.processBatch(pts, offs, rp)
    for eps in (1e-6, 1e-4):      # the north star's 1e-6 m, and a wider band so that the counts are not all zero
        nd.enableBoundaryReport(eps)
        ko, kp, d = nd.processBatch(pts, offs, rp)
        rep = nd.boundaryReport()
        want = ob.process_batch_boundary(P, pts, offs, rp, eps_m=eps, mode=1, n_threads=8)
        assert rep.shape == (nscans, 4)
        assert np.array_equal(rep, want), (eps, rep.sum(0), want.sum(0))
        assert np.array_equal(ko, ko0) and bits_equal(kp, kp0) and bits_equal(d, d0)
        if eps == 1e-4 and cfg != 1:
            assert rep.sum() > 0
    # sub-batching does not change the report; the device-resident entry point fills it too
    import torch
    small = FeatureExtractionNode(to_fe_params(P), max_points=1 << 22, max_scans=max(1, nscans // 3), max_keypoints=1 << 14)
    small.enableBoundaryReport(1e-4)
    small.processBatch(pts, offs, rp)
    assert np.array_equal(small.boundaryReport(), want)
    small.close()
    dev = torch.from_numpy(pts).cuda()
    torch.cuda.synchronize()
    nd.processBatchDevice(dev.data_ptr(), offs, rp)
    assert np.array_equal(nd.boundaryReport(), want)
    nd.enableBoundaryReport(0.0)
    nd.close()


def test_boundary_report_flags_a_pair_sitting_on_the_cluster_tolerance(ob):
    """Two returns of one ring 0.65 m apart to within 4e-7 m: whether they link is decided inside the
    1e-6 m band, so the scan is reported; moved 1e-3 m closer it is not."""
    from feature_extraction_b200 import FeatureExtractionNode
    P = ob.node_default()
    P.cluster_min_count = 1
    t = np.tan(np.deg2rad(-1.0))
    base = [(2.0 + 0.01 * k, 0.5, (2.0 + 0.01 * k) * t, 0.0) for k in range(5)]

    def scan(dx):
        pts = np.array(base + [(2.04 + dx, 0.5, (2.04 + dx) * t, 0.0)], np.float32)
        return pts, np.array([0, len(pts)], np.int64), np.zeros((1, 2))

    nd = FeatureExtractionNode(to_fe_params(P), max_points=1 << 16, max_scans=4, max_keypoints=256)
    nd.enableBoundaryReport(1e-6)
    seen = []
    for dx in (0.6499996 / np.sqrt(1 + t * t), 0.6489996 / np.sqrt(1 + t * t)):
        pts, offs, rp = scan(dx)
        nd.processBatch(pts, offs, rp)
        rep = nd.boundaryReport()
        want = ob.process_batch_boundary(P, pts, offs, rp, eps_m=1e-6, mode=0)
        assert np.array_equal(rep, want)
        seen.append(int(rep[0, 0]))
    nd.close()
    assert seen[0] >= 1 and seen[1] == 0, seen


def test_capacity_is_checked_before_anything_is_written(ob, synth):
    """Round-1 review: a device call with more scans than max_scans_per_call overran the pinned staging
    arrays before it was refused.  It must come back as FE_ERR_CAPACITY with the context still usable."""
    import torch
    from feature_extraction_b200 import FeatureExtractionNode, FeatureExtractionError, _native
    P = ob.node_default()
    pts, offs, rp = synth.generate(2, 40, scan_index_base=8300)
    nd = FeatureExtractionNode(to_fe_params(P), max_points=1 << 21, max_scans=8, max_keypoints=4096)
    dev = torch.from_numpy(pts).cuda()
    torch.cuda.synchronize()
    with pytest.raises(FeatureExtractionError) as e:
        nd.processBatchDevice(dev.data_ptr(), offs, rp)
    assert e.value.status == _native.FE_ERR_CAPACITY
    ko, K, p_kp, p_d = nd.processBatchDevice(dev.data_ptr(), offs[:9], rp[:8])
    ko_o, kp_o, _, _ = ob.process_batch(P, pts, offs[:9], rp[:8], mode=1, n_threads=4)
    assert np.array_equal(ko, ko_o) and bits_equal(nd.download(p_kp, (K, 4)), kp_o)
    # fe_extract_clusters with more points than the context stages
    tiny = FeatureExtractionNode(to_fe_params(P), max_points=4096, max_scans=4, max_keypoints=64)
    with pytest.raises(FeatureExtractionError) as e:
        tiny.extractClusters(np.zeros((5000, 4), np.float32), 0.65, 1, 10)
    assert e.value.status == _native.FE_ERR_CAPACITY
    tiny.close()
    nd.close()


def test_context_is_reusable_after_a_failed_host_call(ob, synth):
    """Round-1 review: a capacity error in the middle of a double-buffered host call left a slot busy and
    the next call finalised the stale sub-batch into its own results."""
    from feature_extraction_b200 import FeatureExtractionNode, FeatureExtractionError, _native
    P = ob.node_default()
    pts, offs, rp = synth.generate(2, 40, scan_index_base=8400)
    # 4 sub-batches of 10 scans on two slots; one keypoint per sub-batch at most -> the keypoint pool overflows
    nd = FeatureExtractionNode(to_fe_params(P), max_points=1 << 19, max_scans=10, max_keypoints=1)
    for _ in range(2):
        with pytest.raises(FeatureExtractionError) as e:
            nd.processBatch(pts, offs, rp)
        assert e.value.status == _native.FE_ERR_CAPACITY
    # a scan without keypoints fits: the same context answers correctly, twice
    far = pts[offs[0]:offs[3]].copy()
    far[:, 0] -= 500.0
    for _ in range(2):
        ko, kp, d = nd.processBatch(far, offs[:4], rp[:3])
        assert list(ko) == [0, 0, 0, 0] and len(kp) == 0
    nd.close()
    ok = FeatureExtractionNode(to_fe_params(P), max_points=1 << 19, max_scans=10, max_keypoints=4096)
    ko, kp, d = ok.processBatch(pts, offs, rp)
    ko_o, kp_o, _, _ = ob.process_batch(P, pts, offs, rp, mode=1, n_threads=8)
    assert np.array_equal(ko, ko_o) and bits_equal(kp, kp_o)
    ok.close()


def test_python_mirror_validates_shapes(ob, synth, node):
    pts, offs, rp = synth.generate(2, 3, scan_index_base=8500)
    with pytest.raises(ValueError):
        node.processBatch(pts, offs, rp[:2])                 # short roll_pitch
    with pytest.raises(ValueError):
        node.processBatch(pts[:100], offs, rp)               # offsets past the end of the points
    bad = offs.copy()
    bad[0] = -5
    with pytest.raises(ValueError):
        node.processBatch(pts, bad, rp)
    with pytest.raises(ValueError):
        node.processBatchLayout(np.zeros((10, 12), np.uint8), 12, 0, 4, 8, offs, rp)

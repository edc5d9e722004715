"""ctypes binding of the CPU oracle (oracle/libfe_oracle.so).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs — never from the feature_extraction_b200 package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

DESC_LEN = 1980
POINT = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("intensity", "<f4")])


class Params(C.Structure):
    """fe_params_t of include/fe_b200.h (reference src:9-34)."""
    _fields_ = [
        ("x_min", C.c_double), ("x_max", C.c_double),
        ("y_min", C.c_double), ("y_max", C.c_double),
        ("z_min", C.c_double), ("z_max", C.c_double),
        ("cluster_tolerance", C.c_double),
        ("cluster_min_count", C.c_int32), ("cluster_max_count", C.c_int32),
        ("cluster_radius_threshold", C.c_double),
        ("number_detection_channels", C.c_int32),
        ("estimate_descriptors", C.c_int32),
        ("descriptor_radius", C.c_double),
    ]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libfe_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.feo_version.restype = C.c_char_p
        _LIB.feo_radius_sq_float.restype = C.c_float
        _LIB.feo_radius_sq_float.argtypes = [C.c_double, C.c_int32]
    return _LIB


def _pts(a):
    a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 4)
    return a


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def node_default():
    p = Params()
    lib().feo_params_node_default(C.byref(p))
    return p


def launch_playback():
    p = Params()
    lib().feo_params_launch_playback(C.byref(p))
    return p


def get_elevation_angles(cloud):
    c = _pts(cloud).copy()
    lib().feo_get_elevation_angles(_p(c), C.c_int64(len(c)))
    return c


def rotation_matrix(roll, pitch):
    m = np.zeros(9, np.float32)
    lib().feo_rotation_matrix(C.c_double(roll), C.c_double(pitch), _p(m))
    return m.reshape(3, 3)


def rotate_cloud(cloud, roll, pitch):
    c = _pts(cloud).copy()
    lib().feo_rotate_cloud(_p(c), C.c_int64(len(c)), C.c_double(roll), C.c_double(pitch))
    return c


def filter_cloud(params, cloud):
    c = _pts(cloud)
    out = np.zeros_like(c)
    n = C.c_int64(0)
    st = lib().feo_filter_cloud(C.byref(params), _p(c), C.c_int64(len(c)), _p(out), C.c_int64(len(c)), C.byref(n))
    assert st == 0
    return out[: n.value].copy()


def select_ring(cloud, ring):
    c = _pts(cloud)
    out = np.zeros_like(c)
    n = C.c_int64(0)
    st = lib().feo_select_ring(_p(c), C.c_int64(len(c)), C.c_int(ring), _p(out), C.c_int64(len(c)), C.byref(n))
    assert st == 0
    return out[: n.value].copy()


def extract_clusters(cloud, tolerance, min_size, max_size, mode=0):
    """-> list of int32 index arrays in PCL's output order."""
    c = _pts(cloud)
    n = len(c)
    offs = np.zeros(n + 2, np.int32)
    idx = np.zeros(max(n, 1), np.int32)
    nc = C.c_int32(0)
    st = lib().feo_extract_clusters(_p(c), C.c_int64(n), C.c_double(tolerance), C.c_int32(min_size),
                                    C.c_int32(max_size), C.c_int32(mode), _p(offs), C.c_int32(n + 1),
                                    _p(idx), C.c_int64(max(n, 1)), C.byref(nc))
    assert st == 0
    return [idx[offs[i]: offs[i + 1]].copy() for i in range(nc.value)]


def get_cylinder_segments(params, ring_cloud, mode=0):
    c = _pts(ring_cloud)
    n = len(c)
    cen = np.zeros((max(n, 1), 4), np.float32)
    cc = np.zeros((max(n, 1), 4), np.float32)
    n1, n2 = C.c_int64(0), C.c_int64(0)
    st = lib().feo_get_cylinder_segments(C.byref(params), _p(c), C.c_int64(n), C.c_int32(mode),
                                         _p(cen), C.c_int64(len(cen)), C.byref(n1),
                                         _p(cc), C.c_int64(len(cc)), C.byref(n2))
    assert st == 0
    return cen[: n1.value].copy(), cc[: n2.value].copy()


def estimate_keypoints(params, cloud, mode=0):
    """-> keypoints, keypoint_cloud, keypoints_full"""
    c = _pts(cloud)
    n = len(c)
    cap = max(2 * n, 1)
    kp = np.zeros((cap, 4), np.float32)
    kc = np.zeros((cap, 4), np.float32)
    kf = np.zeros((cap, 4), np.float32)
    n1, n2, n3 = C.c_int64(0), C.c_int64(0), C.c_int64(0)
    st = lib().feo_estimate_keypoints(C.byref(params), _p(c), C.c_int64(n), C.c_int32(mode),
                                      _p(kp), C.c_int64(cap), C.byref(n1), _p(kc), C.c_int64(cap), C.byref(n2),
                                      _p(kf), C.c_int64(cap), C.byref(n3))
    assert st == 0
    return kp[: n1.value].copy(), kc[: n2.value].copy(), kf[: n3.value].copy()


def estimate_descriptors(params, cloud_full, keypoints, mode=0):
    """-> descriptors (k,1980), edge_margin (k,), n_neighbors (k,)"""
    c = _pts(cloud_full)
    kp = _pts(keypoints)
    k = len(kp)
    d = np.zeros((max(k, 1), DESC_LEN), np.float32)
    m = np.zeros(max(k, 1), np.float32)
    nn = np.zeros(max(k, 1), np.int32)
    st = lib().feo_estimate_descriptors(C.byref(params), _p(c), C.c_int64(len(c)), _p(kp), C.c_int64(k),
                                        C.c_int32(mode), _p(d), _p(m), _p(nn))
    assert st == 0
    return d[:k].copy(), m[:k].copy(), nn[:k].copy()


def sc3d_tables(search_radius):
    radii = np.zeros(16, np.float32)
    theta = np.zeros(12, np.float32)
    phi = np.zeros(13, np.float32)
    lut = np.zeros(DESC_LEN, np.float32)
    lib().feo_sc3d_tables(C.c_double(search_radius), _p(radii), _p(theta), _p(phi), _p(lut))
    return radii, theta, phi, lut


def radius_sq_float(tol, narrow_first=True):
    return np.float32(lib().feo_radius_sq_float(C.c_double(tol), C.c_int32(1 if narrow_first else 0)))


def mt19937_draws(seed, n):
    raw = np.zeros(n, np.uint32)
    f = np.zeros(n, np.float32)
    lib().feo_mt19937_draws(C.c_uint32(seed), C.c_int32(n), _p(raw), _p(f))
    return raw, f


def std_sort_cluster_order(sizes):
    s = np.ascontiguousarray(sizes, np.int32)
    out = np.zeros(len(s), np.int32)
    lib().feo_std_sort_cluster_order(_p(s), C.c_int32(len(s)), _p(out))
    return out


def imu_to_roll_pitch(quat_xyzw, cloud_leveling=True):
    q = np.ascontiguousarray(quat_xyzw, np.float64)
    r, p = C.c_double(0), C.c_double(0)
    lib().feo_imu_to_roll_pitch(_p(q), C.c_int32(1 if cloud_leveling else 0), C.byref(r), C.byref(p))
    return float(r.value), float(p.value)


def process_scan(params, points, roll, pitch, mode=0, want_clouds=False):
    """cloudCallback (src:83-117) for one scan.  -> dict"""
    c = _pts(points)
    n = len(c)
    cap = max(2 * n, 16)
    kp = np.zeros((cap, 4), np.float32)
    nk = C.c_int64(0)
    # first pass: keypoints only (count), then descriptors with the right size
    kc = np.zeros((cap, 4), np.float32)
    cl = np.zeros((max(n, 1), 4), np.float32)
    cf = np.zeros((max(n, 1), 4), np.float32)
    n_kc, n_cl = C.c_int64(0), C.c_int64(0)
    p2 = Params.from_buffer_copy(params)
    p2.estimate_descriptors = 0
    st = lib().feo_process_scan(C.byref(p2), _p(c), C.c_int64(n), C.c_double(roll), C.c_double(pitch), C.c_int32(mode),
                                _p(kp), C.c_int64(cap), C.byref(nk), None, None,
                                _p(kc), C.c_int64(cap), C.byref(n_kc), _p(cl), C.c_int64(len(cl)), C.byref(n_cl), _p(cf))
    assert st == 0, st
    k = nk.value
    res = {"keypoints": kp[:k].copy(), "keypoint_cloud": kc[: n_kc.value].copy(),
           "cloud": cl[: n_cl.value].copy(), "cloud_full": cf[:n].copy()}
    if params.estimate_descriptors:
        d, m, nn = estimate_descriptors(params, res["cloud_full"], res["keypoints"], mode)
        res["descriptors"], res["edge_margin"], res["n_neighbors"] = d, m, nn
    return res


def process_batch(params, points, scan_offsets, roll_pitch, mode=1, n_threads=1, want_desc=True, want_margin=False):
    """-> keypoint_offsets, keypoints, descriptors (or None), edge_margin (or None)"""
    c = _pts(points)
    offs = np.ascontiguousarray(scan_offsets, np.int64)
    rp = np.ascontiguousarray(roll_pitch, np.float64).reshape(-1)
    B = len(offs) - 1
    ko = np.zeros(B + 1, np.int64)
    tot = C.c_int64(0)
    p2 = Params.from_buffer_copy(params)
    if not want_desc:
        p2.estimate_descriptors = 0
    cap = max(64 * B, 1024)
    while True:
        kp = np.zeros((cap, 4), np.float32)
        d = np.zeros((cap, DESC_LEN), np.float32) if (want_desc and params.estimate_descriptors) else None
        m = np.zeros(cap, np.float32) if (want_margin and d is not None) else None
        st = lib().feo_process_batch(C.byref(p2), _p(c), _p(offs), _p(rp), C.c_int32(B), C.c_int32(mode),
                                     C.c_int32(n_threads), _p(ko), _p(kp), _p(d) if d is not None else None,
                                     _p(m) if m is not None else None, C.c_int64(cap), C.byref(tot))
        if st == 3:
            cap = int(tot.value) + 16
            continue
        assert st == 0, st
        break
    k = tot.value
    return ko, kp[:k].copy(), (d[:k].copy() if d is not None else None), (m[:k].copy() if m is not None else None)


def process_batch_boundary(params, points, scan_offsets, roll_pitch, eps_m=1e-6, mode=1, n_threads=1):
    """Tolerance-boundary report of the oracle: (B, 4) int64 pair counts within eps_m of the radius of
    [ring clustering, cross-ring merge, 3DSC support, 3DSC density] (see feo_process_batch_boundary)."""
    c = _pts(points)
    offs = np.ascontiguousarray(scan_offsets, np.int64)
    rp = np.ascontiguousarray(roll_pitch, np.float64).reshape(-1)
    B = len(offs) - 1
    out = np.zeros((max(B, 1), 4), np.int64)
    st = lib().feo_process_batch_boundary(C.byref(params), _p(c), _p(offs), _p(rp), C.c_int32(B), C.c_int32(mode),
                                          C.c_int32(n_threads), C.c_double(eps_m), _p(out))
    assert st == 0, st
    return out[:B]


def libm_f32(op, a, b=None):
    """The host libm (glibc) element-wise: op 0 atan2f(a, b), 1 acosf(a), 2 atanf(a)."""
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b if b is not None else a, np.float32)
    out = np.empty_like(a)
    lib().feo_libm_f32(C.c_int32(op), _p(a), _p(b), _p(out), C.c_int64(a.size))
    return out

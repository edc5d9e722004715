"""Regenerates tests/golden/*.npz from the CPU oracle.

The reference ships no golden vectors and cannot run here (ROS + PCL), so these fixtures are
produced by the oracle itself (brute-force mode): they pin the oracle against regressions and
platform drift (libm), and give the GPU tests inputs/outputs that do not depend on the generator.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from feature_extraction_b200 import synth  # noqa: E402
from oracle import oracle_binding as ob  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def one(name, cfg, scan_index, params):
    pts, offs, rp = synth.generate(cfg, 1, scan_index_base=scan_index, n_threads=1)
    r = ob.process_scan(params, pts, rp[0, 0], rp[0, 1], mode=0)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), points=pts, roll_pitch=rp[0],
        params=np.array([getattr(params, f) for f, _ in params._fields_], np.float64),
        keypoints=r["keypoints"], descriptors=r["descriptors"], edge_margin=r["edge_margin"],
        keypoint_cloud=r["keypoint_cloud"], cloud=r["cloud"],
        elevation=r["cloud_full"][:, 3].copy())
    print(name, "N", len(pts), "crop", len(r["cloud"]), "keypoints", len(r["keypoints"]))


if __name__ == "__main__":
    one("config1_scan0_launch_playback", 1, 0, ob.launch_playback())
    one("config2_scan5_node_default", 2, 5, ob.node_default())
    one("config3_scan2_dense_urban", 3, 2, ob.node_default())
    p4 = ob.node_default()
    p4.descriptor_radius = 5.0
    one("config4_scan1_descriptor_radius_5", 4, 1, p4)

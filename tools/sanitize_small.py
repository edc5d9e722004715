"""A small pass of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize_small.py
Covers: fused batch (configs 1-4, few scans), unordered input (grid fallback of K2), boundary report, cloud outputs,
record output, stage entry points, single-scan graph replay, and one call of more than 16 scans (the throughput
instantiations of K2 and K4d; calls of at most 16 scans run the latency ones)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from feature_extraction_b200 import FeatureExtractionNode, launch_playback, node_default, synth  # noqa: E402

rng = np.random.default_rng(0)
for cfg, n in ((1, 2), (2, 3), (3, 1), (4, 1)):
    P = launch_playback() if cfg == 1 else node_default()
    if cfg == 4:
        P.descriptor_radius = 5.0
    pts, offs, rp = synth.generate(cfg, n, scan_index_base=77)
    nd = FeatureExtractionNode(P, max_points=1 << 19, max_scans=4, max_keypoints=4096)
    nd.enableCloudOutputs(True)
    nd.enableBoundaryReport(1e-6)
    ko, kp, d = nd.processBatch(pts, offs, rp)
    nd.cloudOutputs(n)
    nd.boundaryReport()
    nd.enableBoundaryReport(0.0)
    nd.enableRecordOutput(True)
    nd.processBatch(pts, offs, rp)
    nd.enableRecordOutput(False)
    # unordered input: every entry is a run of its own, dense rings take the grid fallback
    sh = pts[: offs[1]][rng.permutation(int(offs[1]))]
    nd.processBatch(sh, offs[:2], rp[:1])
    # single-scan calls: eager, capture, replay
    for _ in range(3):
        nd.processBatch(pts[: offs[1]], offs[:2], rp[:1])
    # stage entry points
    el = nd.getElevationAngles(pts[: offs[1]])
    rc = nd.filterCloud(nd.rotateCloud(el))
    nd.estimateKeypoints(rc)
    if len(kp):
        nd.estimateDescriptors(el, kp[: min(len(kp), 4)])
    nd.extractClusters(rc[:2000], 0.65, 1, 1000)
    print("config", cfg, "ok:", len(kp), "keypoints")
    nd.close()

# more than 16 scans in one call: the batch instantiations (k_ring_runs<4>, k_desc_hist_warp)
P = node_default()
pts, offs, rp = synth.generate(2, 20, scan_index_base=177)
nd = FeatureExtractionNode(P, max_points=1 << 19, max_scans=32, max_keypoints=4096)
ko, kp, d = nd.processBatch(pts, offs, rp)
print("batch of 20 ok:", len(kp), "keypoints")
nd.close()

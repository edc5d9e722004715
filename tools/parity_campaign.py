"""Large GPU-vs-oracle parity sweep (run under gpurun); writes a markdown summary to gpurun_out/.

No whitelist: a descriptor row either meets BASELINE.json's 1e-5 relative bar or is counted as failing.
Beside it the north star's tolerance-boundary report: pairs within 1e-6 m of each radius predicate
(ring clustering, cross-ring merge, 3DSC support, 3DSC density), counted by the device's audit kernels
and by the oracle's own searches — the two must agree.

  python tools/parity_campaign.py [SHIFT]     other scan indices = other seeded scenes
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from feature_extraction_b200 import FeatureExtractionNode, synth  # noqa: E402
from oracle import oracle_binding as ob  # noqa: E402
from util import to_fe_params  # noqa: E402

SHIFT = int(sys.argv[1]) if len(sys.argv) > 1 else 0
EPS = 1e-6
cores = len(os.sched_getaffinity(0))
rows, brows = [], []
for cfg, nsc, base in ((1, 400, 5000), (2, 4000, 50000), (3, 160, 7000), (4, 240, 9000)):
    P = ob.launch_playback() if cfg == 1 else ob.node_default()
    if cfg == 4:
        P.descriptor_radius = 5.0
    pts, offs, rp = synth.generate(cfg, nsc, scan_index_base=base + SHIFT)
    nd = FeatureExtractionNode(to_fe_params(P), max_points=int(offs[-1]) + 4096, max_scans=nsc, max_keypoints=max(4096, nsc * 64))
    t = time.time()
    ko, kp, d = nd.processBatch(pts, offs, rp)
    tg = time.time() - t
    nd.enableBoundaryReport(EPS)
    nd.processBatch(pts, offs, rp)
    rep = nd.boundaryReport()
    nd.enableBoundaryReport(0.0)
    stats_unordered = 0
    t = time.time()
    ko_o, kp_o, d_o, _ = ob.process_batch(P, pts, offs, rp, mode=1, n_threads=cores)
    tc = time.time() - t
    rep_o = ob.process_batch_boundary(P, pts, offs, rp, eps_m=EPS, mode=1, n_threads=cores)
    same_off = bool(np.array_equal(ko, ko_o))
    kp_bits = bool(same_off and np.array_equal(kp.view(np.uint32), kp_o.view(np.uint32)))
    row_exact = (d.view(np.uint32) == d_o.view(np.uint32)).all(axis=1) if same_off else np.zeros(0, bool)
    with np.errstate(invalid="ignore", divide="ignore"):
        rel = np.abs(d.astype(np.float64) - d_o) / np.maximum(np.abs(d_o), 1e-300)
    rel = np.where((d == d_o) | (np.isnan(d) & np.isnan(d_o)), 0, rel).max(axis=1) if same_off else np.zeros(0)
    rows.append((cfg, nsc, int(offs[-1]), len(kp_o), same_off, kp_bits, int(row_exact.sum()), int((~row_exact).sum()),
                 float(rel.max()) if len(rel) else 0.0, int((rel > 1e-5).sum()), tg, tc))
    brows.append((cfg, nsc, bool(np.array_equal(rep, rep_o)), int((rep.sum(1) > 0).sum())) + tuple(int(v) for v in rep.sum(0)) +
                 tuple(int(v) for v in rep_o.sum(0)))
    nd.close()
out = ["# Parity campaign (GPU C-ABI vs CPU oracle, KD-tree mode), round 2, scan-index shift %d" % SHIFT, "",
       "No whitelist.  Bar: keypoints bit-equal, descriptor values within 1e-5 relative.", "",
       "| config | scans | points | keypoints | counts equal | keypoints bit-equal | descriptor rows bit-equal | rows not bit-equal | max rel err | rows > 1e-5 | GPU s (pageable host) | oracle s (all cores) |",
       "|---|---|---|---|---|---|---|---|---|---|---|---|"]
for r in rows:
    out.append("| %d | %d | %d | %d | %s | %s | %d | %d | %.3g | %d | %.3f | %.2f |" % r)
out += ["", "## Tolerance-boundary report (pairs within %g m of a radius predicate)" % EPS, "",
        "| config | scans | device == oracle | scans with any boundary pair | device: ring clustering | cross-ring merge | 3DSC support | 3DSC density | oracle: ring clustering | cross-ring merge | 3DSC support | 3DSC density |",
        "|---|---|---|---|---|---|---|---|---|---|---|---|"]
for r in brows:
    out.append("| %d | %d | %s | %d | %d | %d | %d | %d | %d | %d | %d | %d |" % r)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "parity_campaign_r2%s.md" % ("_%d" % SHIFT if SHIFT else "")), "w").write("\n".join(out) + "\n")
print("\n".join(out))

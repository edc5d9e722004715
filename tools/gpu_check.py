"""Quick GPU-vs-oracle check used during development (run under gpurun)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle_binding as ob
from feature_extraction_b200 import node as fn, synth

def to_fe(P):
    q = fn.node_default()
    for f, _ in q._fields_:
        setattr(q, f, getattr(P, f))
    return q

def cmp(name, a, b, exact=True):
    a = np.asarray(a); b = np.asarray(b)
    if a.shape != b.shape:
        print("  %-28s SHAPE MISMATCH gpu %s oracle %s" % (name, a.shape, b.shape)); return False
    if a.size == 0:
        print("  %-28s ok (empty)" % name); return True
    eq = np.array_equal(a.view(np.uint32), b.view(np.uint32)) if a.dtype == np.float32 else np.array_equal(a, b)
    if eq:
        print("  %-28s bit-exact %s" % (name, a.shape)); return True
    with np.errstate(invalid='ignore', divide='ignore'):
        d = np.abs(a - b); rel = d / np.maximum(np.abs(b), 1e-30)
    nbad = int((~((a == b) | (np.isnan(a) & np.isnan(b)))).sum())
    print("  %-28s DIFF: %d elems differ, max abs %.3g, max rel %.3g" % (name, nbad, np.nanmax(d), np.nanmax(np.where(d > 0, rel, 0))))
    return False

for cfg in (1, 2, 3, 4):
    nsc = 3
    pts, offs, rp = synth.generate(cfg, nsc)
    P = ob.launch_playback() if cfg == 1 else ob.node_default()
    if cfg == 4: P.descriptor_radius = 5.0
    nd = fn.FeatureExtractionNode(to_fe(P))
    print("== config", cfg)
    for s in range(nsc):
        sc = pts[offs[s]:offs[s+1]]
        r = ob.process_scan(P, sc, rp[s, 0], rp[s, 1], mode=0)
        print(" scan", s, "N", len(sc), "crop", len(r['cloud']), "kp", len(r['keypoints']))
        el_o = ob.get_elevation_angles(sc); el_g = nd.getElevationAngles(sc)
        cmp("getElevationAngles", el_g, el_o)
        nd.roll, nd.pitch = rp[s]
        cmp("rotateCloud", nd.rotateCloud(el_o), ob.rotate_cloud(el_o, rp[s, 0], rp[s, 1]))
        cmp("filterCloud", nd.filterCloud(r['cloud_full']), r['cloud'])
        kp_o, kc_o, kf_o = ob.estimate_keypoints(P, r['cloud'])
        ring = ob.select_ring(r['cloud'], 5)
        if len(ring) and len(ring) <= 2900:
            co = ob.extract_clusters(ring, P.cluster_tolerance, P.cluster_min_count, P.cluster_max_count)
            cg = nd.extractClusters(ring, P.cluster_tolerance, P.cluster_min_count, P.cluster_max_count)
            same = len(co) == len(cg) and all(np.array_equal(x, y) for x, y in zip(co, cg))
            print("  %-28s %s (%d clusters, %d pts)" % ("extractClusters ring5", "identical" if same else "MISMATCH %d vs %d" % (len(cg), len(co)), len(co), len(ring)))
            cen_o, cc_o = ob.get_cylinder_segments(P, ring)
            cen_g, cc_g = nd.getCylinderSegments(ring)
            cmp("getCylinderSegments cen", cen_g, cen_o); cmp("getCylinderSegments cloud", cc_g, cc_o)
        kp_g, kc_g = nd.estimateKeypoints(r['cloud'])
        cmp("estimateKeypoints kp", kp_g, kp_o); cmp("estimateKeypoints cloud", kc_g, kc_o)
        if len(kp_o):
            d_g = nd.estimateDescriptors(r['cloud_full'], kp_o)
            ok = cmp("estimateDescriptors", d_g, r['descriptors'])
            if not ok:
                with np.errstate(invalid='ignore'):
                    rel = np.abs(d_g - r['descriptors']) / np.maximum(np.abs(r['descriptors']), 1e-30)
                    rel = np.where(np.abs(d_g - r['descriptors']) > 0, rel, 0)
                print("    per-kp max rel:", np.nanmax(rel, axis=1)[:12], "margin", r['edge_margin'][:12])
    # fused batch
    nd.enableCloudOutputs(True)
    nd.enableStageTiming(True)
    ko, kp, d = nd.processBatch(pts, offs, rp)
    ko_o, kp_o, d_o, m_o = ob.process_batch(P, pts, offs, rp, mode=0, n_threads=8, want_margin=True)
    cmp("BATCH keypoint_offsets", ko, ko_o); cmp("BATCH keypoints", kp, kp_o)
    if d is not None and d.shape == d_o.shape and len(d):
        with np.errstate(invalid='ignore'):
            rel = np.abs(d - d_o) / np.maximum(np.abs(d_o), 1e-30); rel = np.where(np.abs(d - d_o) > 0, rel, 0)
        print("  BATCH descriptors max rel %.3g; kp over 1e-5: %d of %d; min margin %.3g" % (np.nanmax(rel), int((np.nanmax(rel, axis=1) > 1e-5).sum()), len(d), m_o.min()))
    print("  stage times:", nd.stageTimes(), "launches", nd.last_launches)
    nd.close()

# throughput smoke: config 2, 2000 scans
pts, offs, rp = synth.generate(2, 2000)
nd = fn.FeatureExtractionNode(to_fe(ob.node_default()), max_points=40 << 20, max_scans=2048, max_keypoints=1 << 17)
nd.enableStageTiming(True)
for it in range(3):
    t = time.time(); ko, kp, d = nd.processBatch(pts, offs, rp, copy=False); dt = time.time() - t
    print("config2 x2000: %.1f ms -> %.0f scans/s, K=%d" % (dt * 1e3, 2000 / dt, len(kp)), nd.stageTimes())

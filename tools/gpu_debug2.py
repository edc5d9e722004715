import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
from oracle import oracle_binding as ob
from feature_extraction_b200 import FeatureExtractionNode, synth
from util import to_fe_params, bits_equal, rel_err
P = ob.node_default()
nd = FeatureExtractionNode(to_fe_params(P), max_points=8 << 20, max_scans=512, max_keypoints=1 << 15)
for sidx in (137, 111):
    pts, offs, rp = synth.generate(2, 1, scan_index_base=sidx)
    r = ob.process_scan(P, pts, rp[0, 0], rp[0, 1], mode=0)
    kp_o, kc_o, kf_o = ob.estimate_keypoints(P, r['cloud'])
    kp_g, kc_g = nd.estimateKeypoints(r['cloud'])
    print("scan", sidx, "estimateKeypoints equal:", bits_equal(kp_g, kp_o), "kc equal:", bits_equal(kc_g, kc_o), len(kp_o), len(kf_o))
    kf_g = nd.debugKeypointsFull(r['cloud'])
    print("  GPU kf:", kf_g.shape, "oracle kf:", kf_o.shape, "equal", bits_equal(kf_g, kf_o))
    if kf_g.shape == kf_o.shape:
        bad = np.where((kf_g.view(np.uint32) != kf_o.view(np.uint32)).any(axis=1))[0]
        for b in bad: print("   row", b, "gpu", kf_g[b], "oracle", kf_o[b])
    else:
        print(kf_g)
    allc = []
    for ring in range(16):
        rc = ob.select_ring(r['cloud'], ring)
        cen_o, cc_o = ob.get_cylinder_segments(P, rc)
        cen_g, cc_g = nd.getCylinderSegments(rc)
        if not bits_equal(cen_g, cen_o):
            print("  ring", ring, "centroids differ", len(cen_g), len(cen_o))
        allc.append(cen_o)
    kf = np.concatenate(allc)
    print("  kf equal to oracle's:", bits_equal(kf, kf_o))
    # merge stage alone: cluster kf with pseudo z
    kfz = kf.copy()
    kfz[:, 2] = (kf[:, 3].astype(np.float64) * 0.75 * P.cluster_radius_threshold / 2).astype(np.float32)
    co = ob.extract_clusters(kfz, P.cluster_radius_threshold, P.number_detection_channels, 16)
    cg = nd.extractClusters(kfz, P.cluster_radius_threshold, P.number_detection_channels, 16)
    print("  merge clusters equal:", len(co)==len(cg) and all(np.array_equal(a,b) for a,b in zip(co,cg)))


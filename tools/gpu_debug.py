import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
from oracle import oracle_binding as ob
from feature_extraction_b200 import FeatureExtractionNode, synth
from util import to_fe_params, bits_equal, rel_err
for cfg, nsc, base in ((2, 48, 100), (4, 4, 100)):
    P = ob.node_default()
    if cfg == 4: P.descriptor_radius = 5.0
    nd = FeatureExtractionNode(to_fe_params(P), max_points=8 << 20, max_scans=512, max_keypoints=1 << 15)
    pts, offs, rp = synth.generate(cfg, nsc, scan_index_base=base)
    ko, kp, d = nd.processBatch(pts, offs, rp)
    ko_o, kp_o, d_o, m_o = ob.process_batch(P, pts, offs, rp, mode=0, n_threads=8, want_margin=True)
    print("cfg", cfg, "offsets equal", np.array_equal(ko, ko_o), "K", len(kp), len(kp_o))
    if np.array_equal(ko, ko_o):
        bad = np.where((kp.view(np.uint32) != kp_o.view(np.uint32)).any(axis=1))[0]
        print(" kp rows differing:", bad[:20], "of", len(kp))
        for b in bad[:6]:
            s = np.searchsorted(ko, b, side='right') - 1
            print("  scan", s, "local", b - ko[s], "gpu", kp[b], "oracle", kp_o[b], "diff", kp[b] - kp_o[b])
            # check whether it's a permutation within the scan
            a = kp[ko[s]:ko[s+1]]; o = kp_o[ko[s]:ko[s+1]]
            sa = a[np.lexsort(a.T)]; so = o[np.lexsort(o.T)]
            print("   same set within scan:", bits_equal(sa, so))
        e = rel_err(d, d_o).max(axis=1)
        print(" desc: max rel", e.max(), "n>1e-5", (e > 1e-5).sum(), "margins of those", m_o[e > 1e-5][:10])
    else:
        dd = np.where(np.diff(ko) != np.diff(ko_o))[0]
        print(" scans with different counts:", dd[:20], np.diff(ko)[dd[:20]], np.diff(ko_o)[dd[:20]])
    nd.close()

"""How much would two halves of a device-resident batch gain from running concurrently on two streams?
Two contexts on the same GPU, each given one half from its own host thread, against one context with the whole
batch.  (Probe only: wall-clock around synchronous calls, several repetitions, best of.)
    python tools/concurrency_probe.py [scans] [ways]"""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from feature_extraction_b200 import FeatureExtractionNode, node_default, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
W = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = int(sys.argv[3]) if len(sys.argv) > 3 else 2
P = node_default()
pts, offs, rp = synth.generate(cfg, B, scan_index_base=0)
d = torch.from_numpy(pts).cuda()
cuts = [B * i // W for i in range(W + 1)]


def make(n0, n1):
    npt = int(offs[n1] - offs[n0])
    return FeatureExtractionNode(P, max_points=npt + 4096, max_scans=n1 - n0, max_keypoints=max(4096, (n1 - n0) * 64))


whole = make(0, B)
parts = [make(cuts[i], cuts[i + 1]) for i in range(W)]


def run_whole():
    whole.processBatchDevice(d.data_ptr(), offs, rp)


def run_part(i):
    a, b = cuts[i], cuts[i + 1]
    parts[i].processBatchDevice(d.data_ptr() + int(offs[a]) * 16, offs[a:b + 1] - offs[a], rp[a:b])


def timed(fn, reps=8):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t)
    return best * 1e3


def run_parts():
    th = [threading.Thread(target=run_part, args=(i,)) for i in range(W)]
    for t in th:
        t.start()
    for t in th:
        t.join()


def run_serial():
    for i in range(W):
        run_part(i)


for _ in range(3):
    run_whole(); run_parts()
print("config %d, %d scans: whole %.3f ms | %d parts one after the other %.3f ms | %d parts concurrently %.3f ms"
      % (cfg, B, timed(run_whole), W, timed(run_serial), W, timed(run_parts)))

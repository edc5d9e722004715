"""N>1 path on CPU: world_size-2 gloo.  The compute inside each rank is the oracle here (no GPU in
this container); what is under test is the shard split and the host-side CSR gather."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_partition_the_batch():
    from feature_extraction_b200.sharding import shard_range
    for n in (0, 1, 7, 8, 100, 100001):
        for w in (1, 2, 3, 4, 8):
            r = [shard_range(n, g, w) for g in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from feature_extraction_b200 import synth
    from feature_extraction_b200.sharding import gather_results, shard_inputs
    from oracle import oracle_binding as ob
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    pts, offs, rp = synth.generate(2, 7, n_threads=1)
    P = ob.node_default()
    p, o, r = shard_inputs(pts, offs, rp, rank, world)
    ko, kp, d, _ = ob.process_batch(P, p, o, r, mode=1, n_threads=1)
    res = gather_results(ko, kp, d)
    dist.barrier()
    # the shared-memory form of the same gather (what bench.py times for config 5), twice: the second call
    # reuses the segment, and a call without descriptors must work too
    os.environ["MASTER_PORT"] = str(port)
    from feature_extraction_b200.sharding import SharedGather
    sg = SharedGather(tag="fe_gather_test")
    res2 = sg.gather(ko, kp, d)
    if rank == 0:
        res2 = tuple(np.array(x) for x in res2)
    res3 = sg.gather(ko, kp, d)
    if rank == 0:
        res3 = tuple(np.array(x) for x in res3)
    res4 = sg.gather(ko, kp, None)
    if rank == 0:
        res4 = (np.array(res4[0]), np.array(res4[1]), res4[2])
    sg.close()
    if rank == 0:
        ko_a, kp_a, d_a, _ = ob.process_batch(P, pts, offs, rp, mode=1, n_threads=1)
        ok = True
        for r3 in (res, res2, res3):
            ok = ok and (np.array_equal(r3[0], ko_a) and np.array_equal(r3[1].view(np.uint32), kp_a.view(np.uint32))
                         and np.array_equal(r3[2].view(np.uint32), d_a.view(np.uint32)))
        ok = ok and np.array_equal(res4[0], ko_a) and np.array_equal(res4[1].view(np.uint32), kp_a.view(np.uint32)) and res4[2] is None
        q.put(bool(ok))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_shard_and_gather_matches_unsharded():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True

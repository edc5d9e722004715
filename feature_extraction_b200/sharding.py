"""Scan-parallel sharding across the GPUs of one box.

Scans are independent (the reference reads no cross-scan state, src:72-145), so GPU g of G takes
the contiguous scan range [g*B/G, (g+1)*B/G) and results are concatenated on the host in scan
order.  There is no data-path collective; torch.distributed is used only to collect the per-rank
CSR results (and by bench.py for the barrier / max-over-ranks timing).
"""
import numpy as np


def shard_range(n_scans, rank, world_size):
    """Contiguous scan range [lo, hi) of `rank`."""
    lo = (n_scans * rank) // world_size
    hi = (n_scans * (rank + 1)) // world_size
    return lo, hi


def shard_inputs(points, scan_offsets, roll_pitch, rank, world_size):
    """The slice of a CSR batch that `rank` processes (offsets rebased to 0)."""
    offs = np.asarray(scan_offsets, np.int64)
    lo, hi = shard_range(len(offs) - 1, rank, world_size)
    p0, p1 = int(offs[lo]), int(offs[hi])
    return points[p0:p1], offs[lo:hi + 1] - p0, np.asarray(roll_pitch).reshape(-1, 2)[lo:hi]


def concat_csr(parts):
    """parts: list of (keypoint_offsets, keypoints, descriptors|None) in rank order -> one CSR."""
    offs = [np.zeros(1, np.int64)]
    run = 0
    for ko, _, _ in parts:
        ko = np.asarray(ko, np.int64)
        offs.append(ko[1:] + run)
        run += int(ko[-1])
    kp = np.concatenate([p[1].reshape(-1, 4) for p in parts], axis=0)
    descs = [p[2] for p in parts]
    d = None if any(x is None for x in descs) else np.concatenate(descs, axis=0)
    return np.concatenate(offs), kp, d


def gather_results(keypoint_offsets, keypoints, descriptors, dst=0):
    """Host-side gather of the per-rank results on rank `dst` (None elsewhere)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return keypoint_offsets, keypoints, descriptors
    world = dist.get_world_size()
    mine = (np.asarray(keypoint_offsets), np.asarray(keypoints), None if descriptors is None else np.asarray(descriptors))
    out = [None] * world if dist.get_rank() == dst else None
    dist.gather_object(mine, out, dst=dst)
    if dist.get_rank() != dst:
        return None
    return concat_csr(out)

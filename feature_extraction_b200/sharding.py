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


def _parallel_copy(dst, src, threads=4, min_rows=4096):
    """dst[:] = src with a few threads (numpy releases the GIL inside large copies): the descriptor block of a
    100k-scan sweep is gigabytes, one thread moves it at a fraction of the host's memory bandwidth."""
    n = len(src)
    if n < min_rows or threads <= 1:
        dst[:] = src
        return
    from concurrent.futures import ThreadPoolExecutor
    cuts = [n * i // threads for i in range(threads + 1)]
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(lambda ab: np.copyto(dst[ab[0]:ab[1]], src[ab[0]:ab[1]]), zip(cuts[:-1], cuts[1:])))


class SharedGather:
    """Host-side gather of per-rank CSR results through one POSIX shared-memory segment (ranks of ONE box).

    What crosses torch.distributed is one small all_gather of (scans, keypoints) per rank; every rank then
    copies its keypoints and descriptors straight into its slice of the segment, in parallel with the
    others, and rank `dst` reads the concatenated arrays in scan order.  No data-path collective, no
    pickling of the 7.9 KB-per-keypoint descriptors (compare gather_results)."""

    def __init__(self, tag="fe_gather"):
        import os
        import torch.distributed as dist
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        port = os.environ.get("MASTER_PORT", "0")
        self.name = "%s_%s" % (tag, port)
        self.path = None
        self.cap = 0
        self.mm = None

    @staticmethod
    def _pick_dir(nbytes):
        """A directory every rank of the box sees with room for the segment: /dev/shm when it is large enough
        (a container's default is 64 MB — writing past a tmpfs's size is a SIGBUS, not an exception), else /tmp."""
        import os
        for d in ("/dev/shm", "/tmp"):
            try:
                st = os.statvfs(d)
                if st.f_bavail * st.f_frsize > nbytes * 1.2 + (64 << 20):
                    return d
            except OSError:
                pass
        return None

    def _all_counts(self, n_scans, n_kp):
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return np.array([[n_scans, n_kp]], np.int64)
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        mine = torch.tensor([n_scans, n_kp], dtype=torch.int64, device=dev)
        out = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(out, mine)
        return np.stack([t.cpu().numpy() for t in out])

    def _map(self, nbytes):
        import os
        import torch.distributed as dist
        if nbytes <= self.cap and self.mm is not None:
            return
        cap = max(int(nbytes * 1.25) + 4096, 1 << 20)
        self.mm = None
        if self.path:
            if self.rank == 0:
                try:
                    os.unlink(self.path)
                except OSError:
                    pass
            self.path = None
        d = self._pick_dir(cap)   # the same answer on every rank of the box
        if d is None:
            raise MemoryError("SharedGather: no shared directory with %d bytes free" % cap)
        self.path = os.path.join(d, self.name)
        if self.rank == 0:
            with open(self.path, "wb") as f:
                f.truncate(cap)
        if self.world > 1:
            dist.barrier()
        self.mm = np.memmap(self.path, dtype=np.uint8, mode="r+", shape=(cap,))
        self.cap = cap

    def gather(self, keypoint_offsets, keypoints, descriptors, dst=0):
        """-> (keypoint_offsets, keypoints, descriptors) of the whole job on rank `dst`, None elsewhere.
        The arrays alias the shared segment (valid until the next gather)."""
        import torch.distributed as dist
        ko = np.asarray(keypoint_offsets, np.int64)
        kp = np.ascontiguousarray(keypoints, np.float32).reshape(-1, 4)
        d = None if descriptors is None else np.ascontiguousarray(descriptors, np.float32)
        if self.world == 1:
            return ko, kp, d
        counts = self._all_counts(len(ko) - 1, len(kp))
        S, K = int(counts[:, 0].sum()), int(counts[:, 1].sum())
        dl = 0 if d is None else (d.shape[1] if d.ndim == 2 else 0)
        # layout: offsets int64[S+1] | keypoints float32[K,4] | descriptors float32[K,dl]
        o_kp = (S + 1) * 8
        o_d = o_kp + K * 16
        self._map(o_d + K * dl * 4)
        s0, k0 = int(counts[: self.rank, 0].sum()), int(counts[: self.rank, 1].sum())
        offs = self.mm[: (S + 1) * 8].view(np.int64)
        offs[s0 + 1: s0 + len(ko)] = ko[1:] + k0
        if self.rank == 0:
            offs[0] = 0
        self.mm[o_kp + k0 * 16: o_kp + (k0 + len(kp)) * 16].view(np.float32).reshape(-1, 4)[:] = kp
        if dl:
            dst_d = self.mm[o_d + k0 * dl * 4: o_d + (k0 + len(kp)) * dl * 4].view(np.float32).reshape(-1, dl)
            _parallel_copy(dst_d, d)
        dist.barrier()
        if self.rank != dst:
            return None
        return (offs, self.mm[o_kp: o_kp + K * 16].view(np.float32).reshape(-1, 4),
                self.mm[o_d: o_d + K * dl * 4].view(np.float32).reshape(-1, dl) if dl else None)

    def close(self):
        import os
        import torch.distributed as dist
        self.mm = None
        if self.world > 1 and dist.is_initialized():
            dist.barrier()
        if self.rank == 0 and self.path:
            try:
                os.unlink(self.path)
            except OSError:
                pass

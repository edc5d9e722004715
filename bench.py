#!/usr/bin/env python
"""bench.py — scans/sec of the per-scan keypoint pipeline (BASELINE.json metric) on N B200s.

A "step" is one pass of the hot path (reference cloudCallback, src:83-117) over one batch of
synthetic VLP-16 scans: BASELINE.json configs[1] — 10k scans, node_default parameters — per GPU.

  value : whole-job scans/s with the points already resident in HBM (fe_process_batch_device),
          timed with CUDA events on the library's stream, max over ranks.
  e2e   : the same metric through the host-buffer C-ABI call (fe_process_batch): pinned host points
          in, H2D + kernels + D2H of keypoints/descriptors inside the timed region.
  roofline     : the dominant kernel's algorithmic bytes / its CUDA-event time vs measured HBM peak.
                 The timed pipeline runs the surface-grid kernel on a side stream next to the
                 clustering kernels; per-kernel times therefore come from the same number of steps
                 repeated right after the timed region with the stages serialised
                 (fe_enable_stage_timing), each bracketed by CUDA events on the launching stream.
  cpu_baseline : the CPU oracle (a port of the reference's PCL path, KD-tree mode) on this box's
                 host cores, on a bounded sample of the same workload.

`--impl reference` times that CPU path alone with all host threads (the reference itself needs
ROS + PCL and cannot be built in this image).

Multi-GPU: scans are independent, so every rank processes its own batch (weak scaling) with no
data-path collective; torch.distributed is only the barrier and the max-over-ranks of the time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIG_DESC = {
    1: "config1: single synthetic VLP-16 scan, launch_playback params",
    2: "config2: batch of 10k synthetic VLP-16 scans (16 rings x 1800 azimuth steps), node_default params",
    3: "config3: dense urban scans at 4x azimuth density, node_default params",
    4: "config4: descriptor-heavy (descriptor_radius 5.0, hundreds of poles), node_default params",
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--scans", type=int, default=0, help="scans per GPU per step (default: the config's batch)")
    ap.add_argument("--cpu-sample", type=int, default=1536, help="scans of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def default_scans(cfg):
    return {1: 1, 2: 10000, 3: 2500, 4: 1000}[cfg]


def oracle_params(ob, cfg):
    P = ob.launch_playback() if cfg == 1 else ob.node_default()
    if cfg == 4:
        P.descriptor_radius = 5.0
    return P


def product_params(cfg):
    from feature_extraction_b200 import launch_playback, node_default
    P = launch_playback() if cfg == 1 else node_default()
    if cfg == 4:
        P.descriptor_radius = 5.0
    return P


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_indices):
        """One poller for ALL GPUs of the job (rank 0 starts it; ranks > 0 pass None): a poller per rank
        multiplies the driver queries without adding information."""
        self.gpus = gpu_indices
        self.rows = []
        self.proc = None

    def start(self):
        if self.gpus is None:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", ",".join(str(g) for g in self.gpus), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100" if len(self.gpus) == 1 else "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.gpus is None:
            return None
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "gpus": list(self.gpus)}


def host_threads():
    return len(os.sched_getaffinity(0))


def run_reference(args, rank, world):
    """The reference's CPU path (oracle port, KD-tree mode), all host threads, rank 0 only."""
    if rank != 0:
        return
    from feature_extraction_b200 import synth
    from oracle import oracle_binding as ob
    cfg = args.config
    P = oracle_params(ob, cfg)
    cores = host_threads()
    n = max(8, min(args.cpu_sample, args.scans or default_scans(cfg)))
    pts, offs, rp = synth.generate(cfg, n)
    for _ in range(max(args.warmup, 1)):
        ob.process_batch(P, pts[: offs[min(n, 64)]], offs[: min(n, 64) + 1], rp[: min(n, 64)], mode=1, n_threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ob.process_batch(P, pts, offs, rp, mode=1, n_threads=cores)
    dt = (time.perf_counter() - t0) / args.steps
    v = n / dt
    sample = "%d scans of the workload per step, KD-tree oracle, %d threads scan-parallel" % (n, cores)
    out = {
        "metric": "scans/sec", "value": v, "unit": "scans/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "impl": "reference",
        "config": {"workload": CONFIG_DESC[cfg], "scans_per_step": n, "points_per_scan_mean": float(offs[-1]) / n},
        "cpu_baseline": {"value": v, "unit": "scans/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "mpoints_per_s": v * float(offs[-1]) / n / 1e6,
    }
    emit(out)


def stage_bytes(st, desc_len=1980):
    """Algorithmic (compulsory) bytes of every timed stage for one launch — DESIGN.md §Kernels."""
    N, Ns, Nc, Kf, K, M = (st[k] for k in ("points", "surface_points", "crop_points", "ring_clusters", "keypoints", "neighbours"))
    return {
        "K1 level+crop+ring": 16 * N + 16 * Ns + 20 * Nc,
        "K2 ring clusters": 20 * Nc + 16 * Kf,
        "K3 merge keypoints": 16 * Kf + 16 * K,
        "keypoint CSR": 32 * K,
        "K4a surface grid": 16 * Ns + 16 * Ns,
        "K4b mark neighbours": 16 * M + 4 * M,
        "K4c density": 20 * M,
        "K4d shape context": 20 * M + 4 * desc_len * K,
    }


_RESULT_FD = None


def quiet_stdout():
    """stdout carries exactly one JSON line: whatever libraries print there at C level (NCCL's version
    banner, ...) is sent to stderr for the duration of the run; the result goes to the original fd."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def emit(out):
    line = (json.dumps(out) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_RESULT_FD, line)


def main():
    args = parse()
    quiet_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from feature_extraction_b200 import FeatureExtractionNode, PinnedBuffer, synth

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    cfg = args.config
    B = args.scans or default_scans(cfg)
    A = synth.default_azimuth_steps(cfg)

    # ---- synthetic input, straight into pinned host memory (each rank its own scans) ----
    cap = 16 * A * B
    pin = PinnedBuffer((cap, 4), np.float32)
    pts, offs, rp = synth.generate(cfg, B, scan_index_base=rank * B, out=pin.array)
    npts = int(offs[-1])

    P = product_params(cfg)
    est_kp = max(4096, B * (64 if cfg in (3, 4) else 16))
    dev_node = FeatureExtractionNode(P, device=local_rank, max_points=npts + 4096, max_scans=B, max_keypoints=est_kp,
                                     max_ring_clusters=max(1 << 20, B * 512))
    d_pts = torch.empty((max(npts, 1), 4), dtype=torch.float32, device="cuda")
    d_pts[:npts].copy_(torch.from_numpy(pts), non_blocking=False)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: device-resident hot path ----
    # clocks / throttle reasons are sampled from before the warm-up to the end of the e2e phase (the
    # device-resident timed region alone is shorter than one nvidia-smi sampling period)
    phys = [g.strip() for g in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if g.strip()]
    job_gpus = [phys[i] if i < len(phys) else i for i in range(world)]  # nvidia-smi wants physical indices / UUIDs
    sampler = ClockSampler(job_gpus if rank == 0 else None)
    sampler.start()
    for _ in range(args.warmup):
        dev_node.processBatchDevice(d_pts.data_ptr(), offs, rp)
    barrier()
    launches = 0
    dev_node.timerBegin()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ko, K, p_kp, p_d = dev_node.processBatchDevice(d_pts.data_ptr(), offs, rp)
        launches += dev_node.last_launches
    ev_ms = dev_node.timerEnd()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    ms_step = max_over_ranks(ev_ms / args.steps)
    value = world * B / (ms_step * 1e-3)
    stats = dev_node.batchStats()
    # Per-kernel durations for the roofline: the timed pipeline above runs K4a on a side stream next to
    # K2/K3, so an event pair there would time two kernels at once.  The same steps are repeated right
    # here with the stages serialised (fe_enable_stage_timing) and every stage bracketed by CUDA events
    # on the launching stream; `serial_ms_per_step` is the step time of that mode.
    stage_acc = {}
    dev_node.enableStageTiming(True)
    dev_node.processBatchDevice(d_pts.data_ptr(), offs, rp)
    dev_node.timerBegin()
    for _ in range(args.steps):
        dev_node.processBatchDevice(d_pts.data_ptr(), offs, rp)
        for nm, ms in dev_node.stageTimes():
            stage_acc[nm] = stage_acc.get(nm, 0.0) + ms
    serial_ms_step = dev_node.timerEnd() / args.steps
    dev_node.enableStageTiming(False)
    stage_ms = {k: v / args.steps for k, v in stage_acc.items()}

    # ---- roofline of the dominant kernel ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    sb = stage_bytes(stats)
    kernels = {}
    for nm, ms in stage_ms.items():
        gbs = sb.get(nm, 0) / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        kernels[nm] = {"ms": ms, "algorithmic_bytes": sb.get(nm, 0), "gbs": gbs, "frac": gbs / peak}
    dom = max(stage_ms, key=stage_ms.get) if stage_ms else None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if dom and os.path.exists(tpath):
        rec = json.load(open(tpath)).get(dom)
        if rec and rec.get("scans"):
            traffic = rec["dram_bytes_per_launch"] * (B / rec["scans"])
    roofline = None
    if dom:
        roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["gbs"], "peak": peak, "unit": "GB/s",
                    "frac": kernels[dom]["frac"], "traffic": traffic, "peak_source": peak_src,
                    "share_of_step": stage_ms[dom] / max(sum(stage_ms.values()), 1e-9)}
        # K2/K3/K4c/K4d keep a scan (or a keypoint) in shared memory: their time is issue / shared-memory
        # bound and their DRAM traffic equals their (small) algorithmic bytes, so an HBM fraction says
        # little about them.  The kernels that do stream HBM are reported beside the dominant one.
        streaming = [k for k in ("K1 level+crop+ring", "K4a surface grid") if k in kernels]
        if streaming:
            best = max(streaming, key=lambda k: stage_ms[k])
            roofline["largest_hbm_streaming_kernel"] = {"kernel": best, "achieved": kernels[best]["gbs"], "frac": kernels[best]["frac"],
                                                        "share_of_step": stage_ms[best] / max(sum(stage_ms.values()), 1e-9)}

    # ---- e2e: host buffers through fe_process_batch (H2D + kernels + D2H inside the timed region) ----
    dev_node.close()
    del d_pts
    torch.cuda.empty_cache()
    sub_scans = max(64, min(1024, B))
    sub_pts = int(min(npts + 4096, (npts / max(B, 1)) * sub_scans * 1.5 + 16 * A * 4))
    host_node = FeatureExtractionNode(P, device=local_rank, max_points=sub_pts, max_scans=sub_scans,
                                      max_keypoints=max(4096, sub_scans * (64 if cfg in (3, 4) else 16)))
    for _ in range(args.warmup):
        ko, kp, d = host_node.processBatch(pts, offs, rp, copy=False)
    barrier()
    t0 = time.perf_counter()
    e2e_launches = 0
    for _ in range(args.steps):
        ko, kp, d = host_node.processBatch(pts, offs, rp, copy=False)
        e2e_launches += host_node.last_launches
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)
    if world > 1:
        dist.barrier()
    e2e_value = world * B / (e2e_ms * 1e-3)
    # the same call with packed 12-byte xyz records (the path never reads the input intensity, src:154)
    pin12 = PinnedBuffer((max(npts, 1), 3), np.float32)
    pin12.array[:npts] = pts[:, :3]
    for _ in range(2):
        host_node.processBatchLayout(pin12.array[:npts], 12, 0, 4, 8, offs, rp, copy=False)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ko12, kp12, d12 = host_node.processBatchLayout(pin12.array[:npts], 12, 0, 4, 8, offs, rp, copy=False)
    torch.cuda.synchronize()
    e2e12_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)
    same12 = bool(np.array_equal(ko12, ko) and np.array_equal(kp12.view(np.uint32), kp.view(np.uint32)))
    ko, kp, d = host_node.processBatch(pts, offs, rp, copy=False)  # leave the float4 results in place for the checks below
    pin12.free()
    clocks = sampler.stop()
    h2d = npts * 16 + (B + 1) * 12 + B * 36
    d2h = int(len(kp)) * 16 + (0 if d is None else int(len(kp)) * 1980 * 4) + (B + 1) * 4 + 32

    out = {
        "metric": "scans/sec", "value": value, "unit": "scans/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": CONFIG_DESC[cfg], "scans_per_gpu": B, "points_per_scan_mean": npts / max(B, 1),
                   "input_bytes_per_gpu": npts * 16, "l2": "inputs larger than L2, no flush needed" if npts * 16 > 256e6 else "inputs smaller than L2",
                   "parallelism": "scan-parallel x%d, no collective" % world},
        "mpoints_per_s": value * npts / max(B, 1) / 1e6,
        "wall_ms_per_step": wall_ms / args.steps,
        "e2e": {"value": e2e_value, "unit": "scans/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "mpoints_per_s": e2e_value * npts / max(B, 1) / 1e6, "sub_batch_scans": sub_scans},
        "e2e_packed_xyz": {"value": world * B / (e2e12_ms * 1e-3), "unit": "scans/s", "ms_per_step": e2e12_ms,
                           "h2d_bytes_per_step": npts * 12 + (B + 1) * 12 + B * 36, "same_keypoints_as_float4": same12,
                           "note": "fe_process_batch_layout with 12-byte xyz records"},
        "gpu_launches": launches,
        "gpu_launches_e2e": e2e_launches,
        "clocks": clocks,
        "roofline": roofline,
        "kernels": kernels,
        "kernels_note": "per-stage CUDA-event times of %d extra steps run right after the timed region with the stages serialised "
                        "(fe_enable_stage_timing; %.3f ms per step in that mode); the timed pipeline overlaps K4a with K2/K3 on two streams" % (args.steps, serial_ms_step),
        "serial_ms_per_step": serial_ms_step,
        "work": stats,
    }

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only) ----
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle_binding as ob
        OP = oracle_params(ob, cfg)
        n = max(1, min(args.cpu_sample, B))
        o2 = offs[: n + 1]
        t0 = time.perf_counter()
        ko_o, kp_o, d_o, _ = ob.process_batch(OP, pts, o2, rp[:n], mode=1, n_threads=1)
        dt1 = time.perf_counter() - t0
        cores = host_threads()
        t0 = time.perf_counter()
        ob.process_batch(OP, pts, o2, rp[:n], mode=1, n_threads=cores)
        dtn = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": n / dt1, "unit": "scans/s", "cores": 1, "kind": "port",
                               "sample": "first %d scans of the workload, KD-tree oracle, 1 thread (how the reference runs: ros::spin)" % n,
                               "all_cores": {"value": n / dtn, "cores": cores}}
        # the sample doubles as an in-bench parity check of the timed path
        same = bool(np.array_equal(ko[: n + 1], ko_o) and np.array_equal(kp[: ko[n]].view(np.uint32), kp_o.view(np.uint32)))
        out["parity_vs_oracle_on_sample"] = same
        if same and d is not None and d_o is not None:
            dg = d[: ko[n]]
            with np.errstate(invalid="ignore", divide="ignore"):
                rel = np.abs(dg.astype(np.float64) - d_o) / np.maximum(np.abs(d_o), 1e-300)
            rel = np.where((dg == d_o) | (np.isnan(dg) & np.isnan(d_o)), 0.0, rel)
            rel = np.where(np.isnan(rel), np.inf, rel)
            rows = rel.max(axis=1) if len(rel) else np.zeros(0)
            out["descriptor_parity_on_sample"] = {
                "keypoints": int(len(dg)), "rows_bit_identical": int((dg.view(np.uint32) == d_o.view(np.uint32)).all(axis=1).sum()),
                "rows_beyond_1e-5": int((rows > 1e-5).sum()), "max_rel_err": float(rows.max()) if len(rows) else 0.0}
    if rank == 0:
        emit(out)
    host_node.close()
    pin.free()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

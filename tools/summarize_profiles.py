#!/usr/bin/env python
"""Turn what tools/final_profile.sh left in gpurun_out/<tag>/ into the tracked files under profiles/:
    <tag>_bench_config{1,2,3,4}.json, <tag>_bench_reference_arm.json   (the un-profiled bench lines)
    <tag>_launches_config2_2000scans.csv + <tag>_launch_shares.md      (ncu launch list vs CUDA-event shares)
    <tag>_ncu_full_raw_10000scans.csv + <tag>_ncu_summary.md            (ncu --set full extracts)
    dram_traffic.json                                                   (per-stage DRAM bytes, read by bench.py)
usage: python tools/summarize_profiles.py <tag>
"""
import csv
import json
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGE_OF = [("k_level_crop_ring", "K1 level+crop+ring"), ("k_ring_runs", "K2 ring clusters"), ("k_cluster_rings", "K2 ring clusters"),
            ("k_merge_keypoints", "K3 merge keypoints"), ("k_kp_rank", "K4b mark neighbours"), ("k_kp_", "keypoint CSR"),
            ("k_surface_grid", "K4a surface grid"), ("k_desc_mark", "K4b mark neighbours"),
            ("k_density", "K4c density"), ("k_desc_hist", "K4d shape context")]


def stage_of(kernel):
    for pat, st in STAGE_OF:
        if pat in kernel:
            return st
    return None


def short(kernel):
    k = re.sub(r"^void\s+", "", kernel)
    k = re.sub(r"\(.*$", "", k)
    return k.replace("fe::", "")


def read_launches(path):
    lines = [l for l in open(path, errors="replace") if l.startswith('"')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    iK, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    to_us = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
    return [(r[iK], float(r[iV].replace(",", "")) * to_us.get(r[iU], 1.0)) for r in rows[1:] if len(r) > iV and r[iV] not in ("", "Metric Value")]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2_final"
    src = os.path.join(ROOT, "gpurun_out", tag)
    dst = os.path.join(ROOT, "profiles")
    for name in ("bench_default", "bench_config1", "bench_config2", "bench_config3", "bench_config4", "bench_reference_arm"):
        p = os.path.join(src, name + ".json")
        if os.path.exists(p) and os.path.getsize(p) > 0:
            shutil.copy(p, os.path.join(dst, "%s_%s.json" % (tag, name)))
    bench = json.load(open(os.path.join(src, "bench_config2.json")))

    # ---- launch list ----
    lp = os.path.join(src, "launches_config2_2000scans.csv")
    shutil.copy(lp, os.path.join(dst, tag + "_launches_config2_2000scans.csv"))
    L = read_launches(lp)
    # every step starts with k_level_crop_ring; warm-up 1 + 2 steps: the third sequence is the last
    # device-resident step (the e2e phase with its 1024-scan sub-batches follows it)
    # a step = a run of K1 launches (one per scan range) followed by the other kernels
    starts = [i for i, (k, _) in enumerate(L) if "k_level_crop_ring" in k and (i == 0 or "k_level_crop_ring" not in L[i - 1][0])]
    seq = L[starts[2]:starts[3]] if len(starts) > 3 else L[starts[-1]:]
    tot = sum(v for _, v in seq)
    md = ["# %s — ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`), config 2, 2000 scans per step" % tag, "",
          "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ... python bench.py --scans 2000 "
          "--steps 2 --warmup 1 --no-cpu-baseline` (raw: `%s_launches_config2_2000scans.csv`; script `tools/final_profile.sh`)." % tag,
          "Per-launch times under ncu are cold-cache and serialised: compare SHARES with the CUDA-event stage times of "
          "`%s_bench_config2.json`, not absolutes." % tag, "", "One device-resident step (third launch sequence):", "",
          "| kernel | µs | share | stage |", "|---|---|---|---|"]
    by_stage = {}
    for k, v in seq:
        st = stage_of(k) or "-"
        by_stage[st] = by_stage.get(st, 0.0) + v
        md.append("| `%s` | %.1f | %.1f %% | %s |" % (short(k), v, 100 * v / tot, st))
    md.append("| total | %.1f | | |" % tot)
    md += ["", "Stage shares, ncu launch list vs CUDA events of the 10k-scan bench (`%s_bench_config2.json`):" % tag, "",
           "| stage | ncu share | CUDA-event ms | CUDA-event share | algorithmic GB/s | frac of measured HBM peak (%.1f GB/s) |" % bench["roofline"]["peak"],
           "|---|---|---|---|---|---|"]
    ktot = sum(v["ms"] for v in bench["kernels"].values())
    for st, v in bench["kernels"].items():
        md.append("| %s | %.1f %% | %.3f | %.1f %% | %.0f | %.3f |" % (st, 100 * by_stage.get(st, 0.0) / tot, v["ms"], 100 * v["ms"] / ktot, v["gbs"], v["frac"]))
    open(os.path.join(dst, tag + "_launch_shares.md"), "w").write("\n".join(md) + "\n")

    # ---- full-set capture ----
    rp = os.path.join(src, "ncu_full_raw_10000scans.csv")
    shutil.copy(rp, os.path.join(dst, tag + "_ncu_full_raw_10000scans.csv"))
    rows = list(csv.reader(open(rp, errors="replace")))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def g(r, name, scale=1.0):
        v = r[col[name]].replace(",", "")
        u = units[col[name]]
        try:
            x = float(v)
        except ValueError:
            return float("nan")
        if u == "Kbyte": x *= 1e3
        if u == "Mbyte": x *= 1e6
        if u == "Gbyte": x *= 1e9
        if u == "ns": x *= 1e-3
        if u == "ms": x *= 1e3
        return x * scale

    md = ["# %s — `ncu --set full --clock-control none` extracts, config 2, 10000 scans per launch" % tag, "",
          "Command: `ncu --set full --clock-control none --import-source on -k regex:\"k_ring_runs|k_level_crop|k_surface_grid_cells|"
          "k_density|k_desc_hist|k_desc_mark|k_merge\" -c 20 -o ... python bench.py --config 2 --scans 10000 --steps 1 --warmup 0 --no-cpu-baseline` "
          "(raw page: `%s_ncu_full_raw_10000scans.csv`; script `tools/final_profile.sh`).  Cold caches, serialised launches." % tag, "",
          "| kernel | time µs | DRAM read MB | DRAM write MB | regs | grid×block | warps active % | DRAM % | issue active % | thr/inst | L2 hit % |",
          "|---|---|---|---|---|---|---|---|---|---|---|"]
    traffic = {}
    seen_other = False
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        k = r[col["Kernel Name"]]
        if "k_level_crop_ring" in k:
            if seen_other:
                break  # the second step (e2e phase) starts here; a device step launches K1 in up to four scan ranges
        else:
            seen_other = True
        rd, wr = g(r, "dram__bytes_read.sum"), g(r, "dram__bytes_write.sum")
        grid = r[col["Grid Size"]].strip("()").split(",")[0].strip()
        blk = r[col["Block Size"]].strip("()").split(",")[0].strip()
        md.append("| `%s` | %.1f | %.1f | %.1f | %s | %s×%s | %.1f | %.1f | %.1f | %.2f | %.1f |" % (
            short(k), g(r, "gpu__time_duration.sum"), rd / 1e6, wr / 1e6, r[col["launch__registers_per_thread"]], grid, blk,
            g(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            g(r, "dram__bytes_read.sum.pct_of_peak_sustained_elapsed") + g(r, "dram__bytes_write.sum.pct_of_peak_sustained_elapsed"),
            g(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            g(r, "smsp__thread_inst_executed_per_inst_executed.ratio"), g(r, "lts__t_sector_hit_rate.pct")))
        st = stage_of(k)
        if st:
            t = traffic.setdefault(st, {"scans": 10000, "dram_bytes_per_launch": 0.0, "source": "profiles/%s_ncu_full_raw_10000scans.csv" % tag})
            t["dram_bytes_per_launch"] += rd + wr
    md += ["", "DRAM bytes per stage (read + write, all instantiations of the stage summed) against the algorithmic bytes of the bench, both per 10000 scans:", "",
           "| stage | DRAM MB (ncu) | algorithmic MB | ratio |", "|---|---|---|---|"]
    B = bench["config"]["scans_per_gpu"]
    for st, t in traffic.items():
        alg = bench["kernels"].get(st, {}).get("algorithmic_bytes", 0) * 10000.0 / B
        md.append("| %s | %.1f | %.1f | %.2f |" % (st, t["dram_bytes_per_launch"] / 1e6, alg / 1e6, t["dram_bytes_per_launch"] / alg if alg else float("nan")))
    open(os.path.join(dst, tag + "_ncu_summary.md"), "w").write("\n".join(md) + "\n")
    json.dump(traffic, open(os.path.join(dst, "dram_traffic.json"), "w"), indent=1)
    print("wrote profiles/%s_*" % tag)


if __name__ == "__main__":
    main()

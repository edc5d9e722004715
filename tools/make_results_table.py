#!/usr/bin/env python
"""Render the results table of BASELINE.md §4 from the bench lines under profiles/:
    python tools/make_results_table.py r2_final  > /tmp/table.md
(config 1-4 lines: profiles/<tag>_bench_config{1,2,3,4}.json; sub-records and config 5: <tag>_bench_default.json)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else "r2_final"


def load(name):
    p = os.path.join(ROOT, "profiles", "%s_%s.json" % (tag, name))
    return json.load(open(p)) if os.path.exists(p) and os.path.getsize(p) else None


def k(v):
    return "%.2f M" % (v / 1e6) if v >= 1e6 else ("%.1f k" % (v / 1e3) if v >= 1e4 else "%.0f" % v)


default = load("bench_default")
rows = []
params = {1: "`launch_playback`", 2: "`node_default`", 3: "`node_default`", 4: "`node_default` + R=5"}
names = {1: "1 single scan, few poles", 2: "2 batch 10k scans", 3: "3 dense urban, 4× azimuth", 4: "4 descriptor-heavy (R=5.0)"}
for c in (1, 2, 3, 4):
    d = load("bench_config%d" % c)
    if d is None:
        continue
    sub = default if (c == 2 and default) else (default or {}).get("configs", {}).get(str(c), {})
    cpu = (d.get("cpu_baseline") or (sub or {}).get("cpu_baseline") or {})
    cpu1 = cpu.get("value")
    cpun = (cpu.get("all_cores") or {})
    ker = sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms"])
    kstr = " · ".join("%s %.2f" % (n.split()[0], v["ms"]) for n, v in ker[:6] if v["ms"] > 0.005)
    B, npt = d["config"]["scans_per_gpu"], d["config"]["points_per_scan_mean"]
    e2e, pk = d["e2e"]["value"], (d.get("e2e_packed_xyz") or {}).get("value")
    rows.append("| %s | %s × %.1f k | %s | %s | %s | **%s** (%.0f) — %.3f ms per batch | %s%s | %s / %s | %s |" % (
        names[c], "{:,}".format(B), npt / 1e3, params[c], ("%.1f" % cpu1) if cpu1 else "—",
        ("%s @%d" % (k(cpun["value"]), cpun["cores"])) if cpun else "—", k(d["value"]), d["mpoints_per_s"], d["ms_per_step"],
        k(e2e), (" (%s packed)" % k(pk)) if pk else "",
        ("%s×" % "{:,.0f}".format(d["value"] / cpu1)) if cpu1 else "—", ("%s×" % "{:,.0f}".format(e2e / cpu1)) if cpu1 else "—", kstr))
lines = []
_print = print


def print(x):  # noqa: A001 - collect the table so that --write can splice it into BASELINE.md
    lines.append(x)
    _print(x)


print("| config (`BASELINE.json:configs`) | scans × mean pts | params | CPU (i) 1 thread, scans/s | CPU (ii) all cores | 1×B200 device scans/s (Mpts/s) | 1×B200 e2e scans/s | device / e2e speed-up vs (i) | kernels, ms per batch (serialised) |")
print("|---|---|---|---|---|---|---|---|---|")
for r in rows:
    print(r)
if default and "config5" in default and "value" in default["config5"]:
    c5 = default["config5"]
    print("| 5 sweep, %s scans, 1 GPU (strong-scaling base) | %s × 14.0 k | `node_default` | — | — | **%s** (%.0f) | %s (%s packed); host gather %.1f ms of %.0f ms; %.2f of the bare H2D copy (%.1f GB/s) | — | `profiles/%s_bench_default.json:config5` |" % (
        "{:,}".format(c5["total_scans"]), "{:,}".format(c5["total_scans"]), k(c5["value"]), c5["mpoints_per_s"], k(c5["e2e"]["value"]),
        k(c5["e2e_packed_xyz"]["value"]), c5["e2e"]["host_gather_ms_per_step"], c5["e2e"]["ms_per_step"], c5["h2d_probe"]["e2e_over_probe"],
        c5["h2d_probe"]["gbs"], tag))

if "--write" in sys.argv:
    p = os.path.join(ROOT, "BASELINE.md")
    txt = open(p).read()
    a, b = txt.index("<!-- results-table:begin -->"), txt.index("<!-- results-table:end -->")
    open(p, "w").write(txt[:a] + "<!-- results-table:begin -->\n" + "\n".join(lines) + "\n" + txt[b:])

#!/usr/bin/env python
"""Source-level hot spots of one kernel from an ncu report (compiled with -lineinfo, captured with
--import-source on):  python tools/ncu_hot_lines.py <report.ncu-rep> [top] [kernel-name regex | launch index in the report]
Prints, per CUDA source line, its share of executed warp instructions and of stall samples and the mean
active lanes per instruction."""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    flt = []
    if len(sys.argv) > 3:
        flt = ["--launch-skip", sys.argv[3], "-c", "1"] if sys.argv[3].isdigit() else ["-k", "regex:" + sys.argv[3], "-c", "1"]
    txt = subprocess.run(["ncu", "-i", rep] + flt + ["--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    cur, hdr, data = None, None, []
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif len(r) > 8 and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[2] == "-":  # a source line (SASS rows carry an address)
            iS = hdr.index("# Samples") if "# Samples" in hdr else hdr.index("Warp Stall Sampling (All Samples)")
            iI, iT = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
            try:
                data.append((int(r[iI]), int(r[iS]), int(r[iT]), cur, r[0], r[1]))
            except ValueError:
                pass
    tot = sum(d[0] for d in data) or 1
    ts = sum(d[1] for d in data) or 1
    print("total warp instructions %d, samples %d" % (tot, ts))
    for d in sorted(data, key=lambda x: -x[1])[:top]:
        print("%5.1f%% inst %5.1f%% smp  lanes %4.1f  %s:%s  %s" % (100 * d[0] / tot, 100 * d[1] / ts, d[2] / max(d[0], 1), d[3], d[4], d[5].strip()[:100]))


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Per-CUDA-source-line sample / instruction shares from an .ncu-rep (needs -lineinfo and --import-source on).
usage: scripts/ncu_lines.py report.ncu-rep [kernel_regex] [min_pct]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    kern = sys.argv[2] if len(sys.argv) > 2 else None
    min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.7
    cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
    if kern:
        cmd += ["--kernel-name", "regex:" + kern]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, lines, seen_kernel = None, [], 0
    for r in rows:
        if r and r[0] == "Kernel Name":
            seen_kernel += 1
            if seen_kernel > 1:
                break  # first launch only
        if "# Samples" in r:
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or not r[0]:
            continue  # SASS rows have an empty line number
        try:
            lines.append((int(r[0]), r[1], int(r[hdr.index("# Samples")]), int(r[hdr.index("Instructions Executed")])))
        except ValueError:
            pass
    ts = sum(l[2] for l in lines) or 1
    te = sum(l[3] for l in lines) or 1
    print(f"samples {ts}  warp instructions {te}")
    for ln, src, smp, ex in lines:
        if smp * 100 / ts >= min_pct or ex * 100 / te >= min_pct:
            print(f"{ln:5d} {smp*100/ts:6.2f}% smp {ex*100/te:6.2f}% ins  {src.strip()[:130]}")


if __name__ == "__main__":
    main()

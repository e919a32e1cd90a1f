#!/usr/bin/env python3
"""Where a kernel's warp instructions go: `ncu -i rep --page source --csv --kernel-name regex:K > f.csv`, then
scripts/ncu_regions.py f.csv [units]  -- runs of SASS instructions with the same execution count, with their share;
`units` (e.g. the number of 128-byte lines or sequences the launch processed) turns totals into instructions per unit."""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    units = float(sys.argv[2]) if len(sys.argv) > 2 else None
    hdr = next(r for r in rows if r and r[0] == "Address")
    ai, si, ii = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed")
    seen, data = set(), []
    for r in rows:
        if not r or not r[ai].startswith("0x") or r[ai] in seen:
            continue
        seen.add(r[ai])
        data.append((int(r[ai], 16), r[si].strip(), int(r[ii])))
    base = data[0][0]
    tot = sum(d[2] for d in data)
    print("total warp instructions %.3f G" % (tot / 1e9), ("= %.1f per unit" % (tot / units)) if units else "")
    segs, prev, start, acc, n = [], None, 0, 0, 0
    for a, s, c in data:
        if prev is None or abs(c - prev) > 0.15 * max(c, prev, 1):
            if n:
                segs.append((start, a - base, acc, n, prev))
            start, acc, n = a - base, 0, 0
        acc += c
        n += 1
        prev = c
    segs.append((start, data[-1][0] - base, acc, n, prev))
    for st, en, acc, n, c in segs:
        if acc > 0.004 * tot:
            print("0x%05x-0x%05x n=%4d execs/instr=%8.2fM total=%7.1fM (%4.1f%%)%s" % (st, en, n, c / 1e6, acc / 1e6, 100 * acc / tot, (" per unit %.1f" % (acc / units)) if units else ""))


if __name__ == "__main__":
    main()

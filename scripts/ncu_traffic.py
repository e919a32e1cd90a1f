#!/usr/bin/env python3
"""Writes profiles/<tag>_ncu_traffic.json from an `ncu --set full` report of the default bench command: per kernel the DRAM
bytes of one launch (dram__bytes_read.sum, dram__bytes_write.sum), its duration and warp instructions, stamped with the
commit.  bench.py reads the newest such file for `roofline.traffic` instead of a typed-in constant.
usage: scripts/ncu_traffic.py report.ncu-rep tag workload frames"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGES = {
    "k_huffman_literals": ["k_build_huf_tables", "k_decode_literals"],
    "k_sequences": ["k_build_seq_tables", "k_decode_sequences", "k_decode_sequences_multi"],
    "k_scan_blocks": ["k_scan_blocks"],
    "k_execute": ["k_place_zero", "k_resolve", "k_frame_verdict", "k_execute_bodies", "k_place", "k_execute2", "k_execute", "k_execute_pair", "k_execute_pair2", "k_execute_team", "k_long_hist",
                  "k_long_blockscan", "k_long_compose", "k_long_emit", "k_long_jump", "k_long_verdict"],
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def main():
    rep, tag, workload, frames = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]

    def val(r, key):
        i = hdr.index(key)
        return float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)

    kernels = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").split("<")[0].strip()  # "void k_execute2<0>(...)" -> k_execute2
        if name in kernels:  # the first launch of every kernel
            continue
        kernels[name] = {"dram_read": val(r, "dram__bytes_read.sum"), "dram_write": val(r, "dram__bytes_write.sum"),
                         "ms": val(r, "gpu__time_duration.sum"), "warp_instructions": val(r, "smsp__inst_executed.sum")}
    commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    doc = {"commit": commit, "workload": workload, "frames": frames, "report": os.path.basename(rep), "stages": STAGES, "kernels": kernels,
           "note": "one launch per kernel under ncu --set full --clock-control none: cold caches, kernels serialised"}
    path = os.path.join(ROOT, "profiles", f"{tag}_ncu_traffic.json")
    with open(path, "w") as f:
        json.dump(doc, f, indent=1)
    print(path)


if __name__ == "__main__":
    main()

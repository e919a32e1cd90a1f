#!/usr/bin/env python3
"""Regenerates tests/golden/ from the read-only reference checkout (/root/reference).

Run in the build container only (the GPU box has no /root/reference).  What it writes:

  decodecorpus/zNNNNNN.zst   the 100 golden inputs of the reference (decodecorpus_files/, DATA not source)
  manifest.json              per file: compressed size+sha256, original size+sha256 (originals are NOT copied)
  ll_table.json              the 64-entry predefined LL decode table KAT, parsed out of fse/fse_test.go:8-41
                             as {baseline, additional_bits, number_of_bits, symbol}

The bit-reader and ring-buffer KATs (reversebitstream_test.go, bitstream_test.go, ringbuffer_test.go)
are tiny and are restated by hand in tests/test_oracle_kats.py with their file:line.
"""
import hashlib
import json
import os
import re
import shutil
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def sha(b: bytes) -> str:
    return hashlib.sha256(b).hexdigest()


def main() -> None:
    src = os.path.join(REF, "decodecorpus_files")
    dst = os.path.join(HERE, "decodecorpus")
    os.makedirs(dst, exist_ok=True)
    entries = []
    fingerprint = hashlib.sha256()
    for name in sorted(os.listdir(src)):
        data = open(os.path.join(src, name), "rb").read()
        fingerprint.update(name.encode() + bytes.fromhex(sha(data)))
        if not name.endswith(".zst"):
            continue
        orig = open(os.path.join(src, name[:-4]), "rb").read()
        shutil.copyfile(os.path.join(src, name), os.path.join(dst, name))
        entries.append(
            {
                "name": name,
                "compressed_size": len(data),
                "compressed_sha256": sha(data),
                "original_size": len(orig),
                "original_sha256": sha(orig),
            }
        )
    manifest = {
        "source": "KillingSpark/sparkzstd decodecorpus_files/ (Readme.md:4,75)",
        "files": entries,
        "total_compressed": sum(e["compressed_size"] for e in entries),
        "total_original": sum(e["original_size"] for e in entries),
        "fingerprint_sha256_sorted_name_sha": fingerprint.hexdigest(),
    }
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)

    text = open(os.path.join(REF, "fse", "fse_test.go")).read()
    body = text[text.index("expectedLLDecodingTable") : text.index("func TestBuilding")]
    cells = re.findall(r"\{(\d+),\s*(\d+),\s*(\d+),\s*(\d+)\}", body)
    assert len(cells) == 64, len(cells)
    table = [
        {"baseline": int(a), "additional_bits": int(b), "number_of_bits": int(c), "symbol": int(d)}
        for a, b, c, d in cells
    ]
    with open(os.path.join(HERE, "ll_table.json"), "w") as f:
        json.dump({"source": "fse/fse_test.go:8-41 (FSETableEntry{Baseline, NumberOfAdditionalBits, NumberOfBits, Symbol})", "table": table}, f, indent=1)
    print(f"{len(entries)} corpus files, {manifest['total_compressed']} -> {manifest['total_original']} bytes")
    print("fingerprint", manifest["fingerprint_sha256_sorted_name_sha"])


if __name__ == "__main__":
    main()

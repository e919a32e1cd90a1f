"""Command-line twin of the reference's test CLI (cmd/sparkzstd/main.go:113-195).

    python -m sparkzstd_b200.cli [--verify-checksum] X.zst [Y.zst ...]

For every argument X.zst the frame is decoded on the GPU through the FrameReader API; when a file X
exists next to it the output is compared byte for byte (CompareWithFile, main.go:46-111).  Prints the
per-file verdict, the summary and the average speed in MB/s like the reference does (main.go:143-191).
Exit status 1 when any file failed or differed.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="sparkzstd_b200.cli", description=__doc__.splitlines()[0])
    ap.add_argument("files", nargs="+", help="X.zst files; X is used as the expected output when present")
    ap.add_argument("--verify-checksum", action="store_true", help="also verify the content checksum (not done by the reference)")
    ap.add_argument("-o", "--output-dir", help="write decoded files here")
    ap.add_argument("--dict", dest="dictionary", help="a zstd dictionary file (raw content or formatted) the files were compressed with "
                    "(szb_decode_batch_dict; the reference has no dictionary support)")
    ap.add_argument("--stream", action="store_true", help="treat every file as a zstd stream: concatenated frames, skippable frames "
                    "(szb_decode_stream; the reference decodes exactly one frame per file)")
    args = ap.parse_args(argv)

    from . import decompression as D

    ctx = D.default_context()
    comp = D.NewFrameReader(None, ctx)  # one reusable reader, Reset per file: main.go:126,59
    total_bytes = 0
    total_s = 0.0
    ok = differ = failed = 0
    for path in args.files:
        t0 = time.perf_counter()
        try:
            if args.dictionary:
                with open(args.dictionary, "rb") as df, open(path, "rb") as f:
                    out = ctx.decode_batch_dict([f.read()], df.read())[0]
            elif args.stream:
                with open(path, "rb") as f:
                    out = ctx.decode_stream(f.read(), verify_checksum=args.verify_checksum)
            else:
                with open(path, "rb") as f:
                    comp.Reset(f)
                    chunks = []
                    while True:
                        c = comp.Read(1 << 20)
                        if not c:
                            break
                        chunks.append(c)
                out = b"".join(chunks)
                if args.verify_checksum:
                    data = np.fromfile(path, dtype=np.uint8)
                    dst = np.empty(max(len(out), 1), dtype=np.uint8)
                    _, _, st = ctx.decode_batch_into(np.concatenate([data, np.zeros(8, np.uint8)]), np.array([0], np.uint64),
                                                     np.array([len(data)], np.uint64), dst, verify_checksum=True)
                    if st[0] != 0:
                        raise D.error_for(int(st[0]))
        except Exception as e:  # the reference prints the error and goes on to the next file
            failed += 1
            print(f"{path}: ERROR {e}")
            continue
        dt = time.perf_counter() - t0
        total_s += dt
        total_bytes += len(out)
        verdict = "decoded"
        original = path[:-4] if path.endswith(".zst") else None
        if original and os.path.exists(original):
            with open(original, "rb") as f:
                want = f.read()
            if want == out:
                verdict = "identical to original"
                ok += 1
            else:
                n = next((i for i, (x, y) in enumerate(zip(want, out)) if x != y), min(len(want), len(out)))
                verdict = f"DIFFERS from original at byte {n} (sizes {len(out)} vs {len(want)})"
                differ += 1
        else:
            ok += 1
        if args.output_dir:
            os.makedirs(args.output_dir, exist_ok=True)
            name = os.path.basename(original or path + ".out")
            with open(os.path.join(args.output_dir, name), "wb") as f:
                f.write(out)
        print(f"{path}: {len(out)} bytes, {verdict}")
    print(f"Files: {len(args.files)}  ok: {ok}  different: {differ}  errors: {failed}")
    if total_s > 0:
        print(f"Average detected Speed: {total_bytes / total_s / 1e6:.1f} MB/s (includes file I/O, like the reference)")
    return 1 if (differ or failed) else 0


if __name__ == "__main__":
    sys.exit(main())

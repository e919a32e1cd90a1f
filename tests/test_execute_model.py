"""A CPU model of the stage-4 algorithm of k_execute (kernels.cuh), run against the oracle's per-block trace.

This is not the oracle and not the product: it restates, in a few lines of Python, the two ideas the CUDA kernel rests on, so
that they are checked without a GPU:
  * every sequence becomes up to two SEGMENTS (literal run, match) whose source is `word + position`, found per output byte
    as "the last segment that starts at or before the byte" (the kernel does that with a popcount over a bitmap);
  * the output is produced in address order, one 128-byte line per step: a source below the line is read from the output
    produced so far; a source inside the line is another byte of the step -- bytes of earlier 32-byte chunks are final when
    a chunk is resolved, bytes of the own chunk are found by following the chain of in-chunk sources (pointer jumping).
"""
import bisect

import pytest

from oracle import pyszo


def model_execute(trace) -> bytes:
    out = bytearray()
    for blk in trace.blocks:
        if blk.type != 2:  # Raw / RLE bodies are written by k_execute_bodies; the model takes them from the trace sizes
            raise AssertionError("model_execute expects frames of compressed blocks only")
        # ---- producer: segments (start position, kind, source) ----
        starts, segs = [], []
        pos, lit = len(out), 0
        for (ll, ml, _), off in zip(blk.sequences, blk.real_offsets):
            if ll:
                starts.append(pos)
                segs.append(("lit", lit - pos))  # literal index = word + position
                pos += ll
                lit += ll
            starts.append(pos)
            segs.append(("match", -off))  # source position = word + position
            pos += ml
        if lit < len(blk.literals):
            starts.append(pos)
            segs.append(("lit", lit - pos))
            pos += len(blk.literals) - lit
        end = pos
        # ---- consumer: 128-byte lines in address order ----
        p = len(out)
        out.extend(b"\0" * (end - p))
        line = p & ~127
        while line < end:
            lo, hi = max(line, p), min(line + 128, end)
            for chunk in range(line, line + 128, 32):
                c_lo, c_hi = max(chunk, lo), min(chunk + 32, hi)
                if c_lo >= c_hi:
                    continue
                val, parent = {}, {}
                for q in range(c_lo, c_hi):
                    kind, word = segs[bisect.bisect_right(starts, q) - 1]
                    if kind == "lit":
                        val[q] = blk.literals[word + q]
                    else:
                        src = word + q
                        assert 0 <= src < q
                        if src >= chunk and src >= lo:
                            parent[q] = src  # another byte of this very chunk: not produced yet
                        else:
                            val[q] = out[src]  # below the chunk: earlier chunk of the line, earlier line, earlier block
                rounds = 0
                while parent:  # pointer jumping: a chain of in-chunk sources halves every round
                    nxt = {}
                    for q, s in parent.items():
                        if s in val:
                            val[q] = val[s]
                        else:
                            nxt[q] = parent[s]
                    for q in parent:
                        if q in val and q in nxt:
                            del nxt[q]
                    parent = {q: s for q, s in nxt.items() if q not in val}
                    rounds += 1
                    assert rounds <= 6  # log2(32) + 1
                for q in range(c_lo, c_hi):
                    out[q] = val[q]
            line += 128
    return bytes(out)


@pytest.mark.parametrize("pick", [0, 1, 2])
def test_model_matches_oracle_on_corpus_frames(corpus, pick):
    done = 0
    for name, data, size, _ in corpus[pick::3]:
        if not (0 < size <= 60_000):
            continue
        want, tr = pyszo.decode_frame(data, True)
        if any(b.type != 2 for b in tr.blocks):
            continue
        assert model_execute(tr) == want, name
        done += 1
        if done == 4:
            break
    assert done >= 1

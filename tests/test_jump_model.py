"""A CPU model of the block-parallel stage 4 for long frames (k_long_hist / k_long_compose / k_long_emit / k_long_jump in
kernels.cuh), run against the oracle's per-block trace.

Not the oracle and not the product: it restates the three ideas those kernels rest on, so that they are checked without a GPU:
  * the repeat-offset history a block starts with (sequence_execution.go:65-114) is obtained without executing the blocks
    before it: every block is walked once with a SYMBOLIC history -- an entry is a constant or "entry i of the history the
    block started with, minus k" -- which gives the block's transfer function; the functions are composed in block order;
  * every output byte gets a DISTANCE: 0 for a literal byte (and the bytes of Raw/RLE blocks), else how far back its source
    byte lies; inside an overlapping match the distance is a multiple of the offset, so that it points in front of the match;
  * distances are resolved by pointer jumping, in any order and in place: d[p] += d[p - d[p]] until p - d[p] is a literal
    byte.  Any value d[q] ever holds points at an ancestor of q, which is why no ordering between the bytes is needed.
"""
import random

import pytest

from oracle import pyszo

SYM = "sym"


def dec(v, k=1):
    if isinstance(v, tuple):
        return (SYM, v[1], v[2] + k)
    return (v - k) & 0xFFFFFFFF


def next_offset(h, ofv, ll_zero):
    """sequence_execution.go:65-114 over values that may be symbolic."""
    if ofv > 3:
        off = ofv - 3
        return off, [off, h[0], h[1]]
    idx = ofv - 1 + (1 if ll_zero else 0)
    if idx == 0:
        return h[0], h
    if idx == 1:
        return h[1], [h[1], h[0], h[2]]
    off = h[2] if idx == 2 else dec(h[0])
    return off, [off, h[0], h[1]]


def transfer(blk):
    h = [(SYM, 0, 0), (SYM, 1, 0), (SYM, 2, 0)]
    for ll, _, ofv in blk.sequences:
        _, h = next_offset(h, ofv, ll == 0)
    return h


def apply(t, h):
    return [dec(h[e[1]], e[2]) if isinstance(e, tuple) else e for e in t]


def model_decode(trace, rng) -> bytes:
    # history at the start of every block, by composition
    h = [1, 4, 8]
    hist_in = []
    for blk in trace.blocks:
        hist_in.append(list(h))
        if blk.type == 2 and blk.sequences:
            h = apply(transfer(blk), h)
            assert tuple(h) == tuple(blk.hist_after)
    # emit: literal bytes and distances, block by block (any order)
    total = sum(b.out_len for b in trace.blocks)
    out = bytearray(total)
    d = [None] * total
    order = list(range(len(trace.blocks)))
    rng.shuffle(order)
    for bi in order:
        blk = trace.blocks[bi]
        pos = blk.out_off
        if blk.type != 2 or not blk.sequences:
            body = blk.literals if blk.type == 2 else None
            for k in range(blk.out_len):
                d[pos + k] = 0
            if body is not None:
                out[pos:pos + blk.out_len] = body
            else:
                out[pos:pos + blk.out_len] = b"\xAA" * blk.out_len  # placeholder, patched by the caller (k_execute_bodies)
            continue
        h = hist_in[bi]
        lit = 0
        for (ll, ml, ofv), want_off in zip(blk.sequences, blk.real_offsets):
            off, h = next_offset(h, ofv, ll == 0)
            assert off == want_off
            for k in range(ll):
                out[pos + k] = blk.literals[lit + k]
                d[pos + k] = 0
            pos += ll
            lit += ll
            for m in range(ml):
                d[pos + m] = off if m < off else off * (m // off + 1)
            pos += ml
        rest = len(blk.literals) - lit
        for k in range(rest):
            out[pos + k] = blk.literals[lit + k]
            d[pos + k] = 0
        assert pos + rest == blk.out_off + blk.out_len
    assert all(x is not None for x in d)
    return out, d


def jump(out, d, rng):
    idx = [p for p in range(len(d)) if d[p]]
    rng.shuffle(idx)
    for p in idx:  # any order; writes to d are visible to later bytes (path compression), as on the device
        while True:
            e = d[p - d[p]]
            if e == 0:
                break
            d[p] += e
        out[p] = out[p - d[p]]


@pytest.mark.parametrize("pick", [0, 1, 2])
def test_jump_model_matches_oracle_on_corpus_frames(corpus, pick):
    rng = random.Random(1234 + pick)
    done = 0
    for name, data, size, _ in corpus[pick::3]:
        if not (0 < size <= 80_000):
            continue
        want, tr = pyszo.decode_frame(data, True)
        out, d = model_decode(tr, rng)
        for blk in tr.blocks:  # Raw / RLE bodies come from k_execute_bodies
            if blk.type != 2:
                out[blk.out_off:blk.out_off + blk.out_len] = want[blk.out_off:blk.out_off + blk.out_len]
        jump(out, d, rng)
        assert bytes(out) == want, name
        done += 1
        if done == 6:
            break
    assert done >= 1


def test_symbolic_history_composes():
    # h0 - 1 chains and permutations through two blocks
    class B:
        pass

    b1, b2 = B(), B()
    b1.sequences = [(0, 3, 3), (0, 3, 3), (5, 3, 2), (0, 3, 1)]  # ll == 0 shifts the repeat codes
    b2.sequences = [(0, 3, 3), (1, 3, 3), (1, 3, 100), (0, 3, 2)]
    for h0 in ([1, 4, 8], [9, 5, 7], [1000, 3, 2]):
        h = list(h0)
        for b in (b1, b2):
            for ll, _, ofv in b.sequences:
                _, h = next_offset(h, ofv, ll == 0)
        assert apply(transfer(b2), apply(transfer(b1), list(h0))) == h

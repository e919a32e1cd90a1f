"""Hand-assembled frames (tests/crafted_frames.py): RLE literals next to sequences and sequences longer than the
execute stage's segment ring -- shapes libzstd level 3 never produces.  The CPU half pins the crafted
frames against the oracle; the GPU half compares the CUDA path with the independently computed output."""
import numpy as np
import pytest

from oracle import pyszo
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import crafted_frames as crafted  # noqa: E402


def test_crafted_tables_match_the_golden_ll_table():
    assert crafted.check_against_golden_ll()


@pytest.mark.parametrize("name", sorted(crafted.cases()))
def test_oracle_decodes_crafted_frame(name):
    frame, expected = crafted.cases()[name]
    got, tr = pyszo.decode_frame(frame, True)
    assert got == expected
    blk = tr.blocks[1]
    assert blk.nseq == 1 and blk.modes == (0, 0, 0)


def test_oracle_decodes_oversize_block():
    frame, expected = crafted.oversize_block_case()
    assert len(expected) == 1024 + 20 + 131074
    assert pyszo.decode_frame(frame) == expected


@pytest.mark.gpu
def test_gpu_decodes_crafted_frames_in_one_batch():
    from sparkzstd_b200.decompression import Context

    ctx = Context(0)
    try:
        cs = crafted.cases()
        names = sorted(cs)
        outs = ctx.decode_batch([cs[n][0] for n in names])
        bad = [n for n, o in zip(names, outs) if o != cs[n][1]]
        assert not bad, bad
        # the same frames many times over, interleaved: every warp of a CTA on a different shape
        many = [cs[names[(i * 5) % len(names)]] for i in range(256)]
        outs = ctx.decode_batch([f for f, _ in many])
        assert all(o == e for o, (_, e) in zip(outs, many))
        f, e = crafted.oversize_block_case()
        assert ctx.decode_batch([f, cs[names[0]][0], f]) == [e, cs[names[0]][1], e]
    finally:
        ctx.close()

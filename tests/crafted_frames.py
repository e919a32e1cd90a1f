"""Hand-assembled zstd frames for paths libzstd level 3 never emits (RLE literals next to sequences,
sequences longer than the execute stage's segment ring).  A frame here is a Raw block of seed bytes followed
by ONE compressed block with ONE sequence coded with the predefined FSE tables (RFC 8878 3.1.1.3.2.2), so
that the encoder side stays trivial: a single sequence needs initial states only, no state updates.

The expected output is computed here, independently of any decoder."""
import json
import os

import numpy as np

# RFC 8878 default distributions (the reference holds the same constants in fse/predefined.go:22-78)
LL_NORM = [4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1]
ML_NORM = [1, 4, 3, 2, 2, 2, 2, 2, 2] + [1] * 37 + [-1] * 7
OF_NORM = [1, 1, 1, 1, 1, 1, 2, 2, 2] + [1] * 15 + [-1] * 5
LL_BASE = list(range(16)) + [16, 18, 20, 22, 24, 28, 32, 40, 48, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536]
LL_BITS = [0] * 16 + [1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16]
ML_BASE = list(range(3, 35)) + [35, 37, 39, 41, 43, 47, 51, 59, 67, 83, 99, 131, 259, 515, 1027, 2051, 4099, 8195, 16387, 32771, 65539]
ML_BITS = [0] * 32 + [1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16]


def spread(norm, al):
    """Symbol of every cell of the FSE decoding table (the placement half of fse.go:136-230)."""
    size = 1 << al
    cells = [None] * size
    high = size - 1
    for s, p in enumerate(norm):
        if p == -1:
            cells[high] = s
            high -= 1
    pos = 0
    step = (size >> 1) + (size >> 3) + 3
    for s, p in enumerate(norm):
        for _ in range(max(p, 0)):
            cells[pos] = s
            pos = (pos + step) & (size - 1)
            while pos > high:
                pos = (pos + step) & (size - 1)
    assert pos == 0 and None not in cells
    return cells


LL_CELLS, ML_CELLS, OF_CELLS = spread(LL_NORM, 6), spread(ML_NORM, 6), spread(OF_NORM, 5)


def check_against_golden_ll():
    """The LL placement must agree with the reference's golden table (fse/fse_test.go:8-41)."""
    with open(os.path.join(os.path.dirname(__file__), "golden", "ll_table.json")) as fh:
        gold = json.load(fh)["table"]
    return all(LL_BASE[LL_CELLS[i]] == gold[i]["symbol"] and LL_BITS[LL_CELLS[i]] == gold[i]["additional_bits"] for i in range(64))


def _code(value, base, bits):
    for c in range(len(base) - 1, -1, -1):
        if base[c] <= value:
            assert value - base[c] < (1 << bits[c]) or bits[c] == 0 and value == base[c], (value, c)
            return c, value - base[c], bits[c]
    raise ValueError(value)


def one_sequence_bitstream(ll, ml, offset):
    """The backward bitstream of a single sequence with predefined tables: the decoder reads LL, OF, ML states,
    then OF, ML, LL extra bits (sequences.go:145-176)."""
    llc, llx, lln = _code(ll, LL_BASE, LL_BITS)
    mlc, mlx, mln = _code(ml, ML_BASE, ML_BITS)
    ofv = offset + 3
    ofc = ofv.bit_length() - 1
    ofx = ofv - (1 << ofc)
    acc = 1  # the padding marker bit
    for val, n in ((LL_CELLS.index(llc), 6), (OF_CELLS.index(ofc), 5), (ML_CELLS.index(mlc), 6), (ofx, ofc), (mlx, mln), (llx, lln)):
        acc = (acc << n) | val
    return acc.to_bytes((acc.bit_length() + 7) // 8, "little")


def _block_header(last, btype, size):
    return ((size << 3) | (btype << 1) | last).to_bytes(3, "little")


def _raw_or_rle_literals_header(ltype, regen):
    assert regen < (1 << 20)
    if regen < 32:
        return bytes([(regen << 3) | ltype])
    if regen < 4096:
        return bytes([((regen & 15) << 4) | (1 << 2) | ltype, regen >> 4])
    return bytes([((regen & 15) << 4) | (3 << 2) | ltype, (regen >> 4) & 255, regen >> 12])


def frame(seed_bytes: bytes, lit_kind: str, literals: bytes, ll: int, ml: int, offset: int, oversize: bool = False):
    """-> (frame bytes, expected output).  lit_kind 'rle': every literal is literals[0] and len(literals) is the
    regenerated size; 'raw': literals as given.  One sequence (ll, ml, offset), the rest are trailing literals.
    oversize: allow a block that regenerates more than Block_Maximum_Size (the reference does not check it)."""
    regen = len(literals)
    assert ll <= regen and offset >= 4  # offsets 1..3 would be repeat codes
    if lit_kind == "rle":
        assert len(set(literals)) == 1
        lit_section = _raw_or_rle_literals_header(1, regen) + literals[:1]
    else:
        lit_section = _raw_or_rle_literals_header(0, regen) + literals
    seq_section = bytes([1, 0]) + one_sequence_bitstream(ll, ml, offset)  # one sequence, three predefined tables
    body = lit_section + seq_section
    out = bytearray(seed_bytes)
    out += literals[:ll]
    assert offset <= len(out)
    for _ in range(ml):
        out.append(out[-offset])
    out += literals[ll:]
    total = len(out)
    assert regen + ml <= 128 * 1024 or oversize
    hdr = bytes([0x28, 0xB5, 0x2F, 0xFD, 0xA0]) + total.to_bytes(4, "little")  # single segment, 4-byte content size
    blocks = _block_header(0, 0, len(seed_bytes)) + seed_bytes + _block_header(1, 2, len(body)) + body
    return hdr + blocks, bytes(out)


def cases(seed: int = 7):
    """name -> (frame, expected)"""
    rng = np.random.default_rng(seed)
    rnd = lambda n: rng.integers(0, 256, n, dtype=np.uint8).tobytes()  # noqa: E731
    seed_bytes = rnd(1000)
    out = {}
    # RLE literals: runs that fit a fill-table row, runs that do not, a trailing run of either kind
    out["rle_lit_short_runs"] = frame(seed_bytes, "rle", bytes([0xAB]) * 150, 100, 40, 700)
    out["rle_lit_long_run"] = frame(seed_bytes, "rle", bytes([0x5C]) * 3000, 2000, 100, 2300)
    out["rle_lit_long_tail"] = frame(seed_bytes, "rle", bytes([0x11]) * 5000, 200, 9, 1100)
    out["rle_lit_zero_ll"] = frame(seed_bytes, "rle", bytes([0xEE]) * 256, 0, 300, 512)
    # sequences longer than the segment ring: literal run, match with a far source, overlapping matches
    out["long_match_far"] = frame(seed_bytes, "raw", rnd(10), 10, 60000, 1010)
    out["long_match_period_5"] = frame(seed_bytes, "raw", rnd(20), 20, 40000, 5)
    out["long_match_period_40"] = frame(seed_bytes, "raw", rnd(20), 7, 5000, 40)
    out["long_literal_run"] = frame(seed_bytes, "raw", rnd(70000), 65000, 33, 64999)
    # just around the ring span and the 128-byte line
    out["match_3839"] = frame(seed_bytes, "raw", rnd(4), 1, 3839, 300)
    out["match_127_off_4"] = frame(seed_bytes, "raw", rnd(300), 200, 127, 4)
    out["match_128_off_129"] = frame(seed_bytes, "raw", rnd(300), 131, 128, 129)
    return out


def oversize_block_case(seed: int = 11):
    """(frame, expected): the compressed block regenerates 20 + 131 074 bytes, more than Block_Maximum_Size (128 KiB).
    The reference decodes it (no check); the block-parallel long-frame path sizes its scratch for 128 KiB per block and
    must hand the frame to k_execute_pair."""
    rng = np.random.default_rng(seed)
    seed_bytes = rng.integers(0, 256, 1024, dtype=np.uint8).tobytes()
    return frame(seed_bytes, "raw", rng.integers(0, 256, 20, dtype=np.uint8).tobytes(), 7, 131074, 900, oversize=True)


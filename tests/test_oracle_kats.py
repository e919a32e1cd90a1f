"""Known-answer tests that pin the ORACLE's building blocks to the reference's own unit
tests.  Each test names the reference test it restates."""
import ctypes as C
import json
import os

import pytest

from oracle import pyszo
from oracle.pyszo import _FBits, _FseTable, _HufTable, _RBits, _Sequence

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rb(data: bytes):
    L = pyszo.lib()
    buf = (C.c_uint8 * max(1, len(data))).from_buffer_copy(data.ljust(1, b"\0"))
    r = _RBits()
    L.szo_rbits_init(C.byref(r), buf, len(data))
    r._keep = buf
    return L, r


def _fb(data: bytes):
    L = pyszo.lib()
    buf = (C.c_uint8 * max(1, len(data))).from_buffer_copy(data.ljust(1, b"\0"))
    r = _FBits()
    L.szo_fbits_init(C.byref(r), buf, len(data))
    r._keep = buf
    return L, r


RAMP = bytes(range(256))


def test_reverse_bitstream_ramp():
    """bitstream/reversebitstream_test.go:7-170 TestReverseBitStream"""
    L, r = _rb(RAMP)
    for i in range(256):
        assert L.szo_rbits_read(C.byref(r), 8) == RAMP[255 - i]
    L, r = _rb(RAMP)
    for i in range(512):
        x = L.szo_rbits_read(C.byref(r), 4)
        b = RAMP[256 - (i // 2) - 1]
        assert x == ((b >> 4) if i % 2 == 0 else (b & 0xF))
    L, r = _rb(RAMP)  # 8 x 5 bits then a full byte
    for i in range(256 // 9):
        for _ in range(8):
            L.szo_rbits_read(C.byref(r), 5)
        idx = (i + 1) * 6 - 1
        assert L.szo_rbits_read(C.byref(r), 8) == RAMP[256 - idx - 1]
    for pattern in ([3] * 8, [6, 3, 3, 3, 3, 6], [7, 7, 7, 3]):
        L, r = _rb(RAMP)
        for i in range(256 // 4):
            for n in pattern:
                L.szo_rbits_read(C.byref(r), n)
            idx = (i + 1) * 4 - 1
            assert L.szo_rbits_read(C.byref(r), 8) == RAMP[256 - idx - 1]


def test_reverse_bitstream_edges():
    """bitstream/reversebitstream_test.go:172-229 TestEdges"""
    L, r = _rb(bytes([64, 58, 169, 224]))
    reads = [3, 4, 4, 1, 3, 5, 1, 3, 3, 0, 4]
    want = [7, 0, 5, 0, 4, 19, 1, 2, 2, 0, 0]
    got = [L.szo_rbits_read(C.byref(r), n) for n in reads]
    assert got == want


def test_reverse_bitstream_overread():
    """reads past the start return zero-filled low bits and keep decrementing"""
    L, r = _rb(bytes([0b00000101]))
    assert L.szo_rbits_read(C.byref(r), 5) == 0
    assert L.szo_rbits_read(C.byref(r), 5) == 20
    assert L.szo_rbits_bits_still_in_stream(C.byref(r)) == -3
    assert L.szo_rbits_read(C.byref(r), 7) == 0
    assert L.szo_rbits_bits_still_in_stream(C.byref(r)) == -10
    # 64-bit reads across 9 bytes
    data = bytes(range(1, 13))
    L, r = _rb(data)
    L.szo_rbits_read(C.byref(r), 3)
    v = int.from_bytes(data, "little")
    assert L.szo_rbits_read(C.byref(r), 64) == (v >> (96 - 3 - 64)) & ((1 << 64) - 1)


def test_reverse_bitstream_matches_bigint_model():
    import random

    rng = random.Random(7)
    for _ in range(200):
        n = rng.randrange(0, 40)
        data = bytes(rng.randrange(256) for _ in range(n))
        v = int.from_bytes(data, "little")
        L, r = _rb(data)
        off = n * 8 - 1
        for _ in range(60):
            k = rng.randrange(0, 65)
            got = L.szo_rbits_read(C.byref(r), k)
            if k == 0:
                want = 0
            elif off <= -1:
                want = 0
                off -= k
            else:
                lo = off - k + 1
                want = (v >> lo) & ((1 << k) - 1) if lo >= 0 else ((v & ((1 << (off + 1)) - 1)) << (-lo))
                off -= k
            assert got == want
            assert L.szo_rbits_bits_still_in_stream(C.byref(r)) == off


def test_forward_bitstream_ramp():
    """bitstream/bitstream_test.go:9-149 (same patterns, LSB first)"""
    L, b = _fb(RAMP)
    v = C.c_uint64()
    for i in range(256):
        assert L.szo_fbits_read(C.byref(b), 8, C.byref(v)) == 0 and v.value == RAMP[i]
    L, b = _fb(RAMP)
    for i in range(512):
        L.szo_fbits_read(C.byref(b), 4, C.byref(v))
        assert v.value == ((RAMP[i // 2] & 0xF) if i % 2 == 0 else (RAMP[i // 2] >> 4))
    for pattern in ([3] * 8, [6, 3, 3, 3, 3, 6]):
        L, b = _fb(RAMP)
        for i in range(256 // 4):
            for n in pattern:
                L.szo_fbits_read(C.byref(b), n, C.byref(v))
            L.szo_fbits_read(C.byref(b), 8, C.byref(v))
            assert v.value == RAMP[(i + 1) * 4 - 1]


def test_forward_bitstream_unwind_and_eof():
    L, b = _fb(bytes([0b10110101, 0xFF]))
    v = C.c_uint64()
    L.szo_fbits_read(C.byref(b), 3, C.byref(v))
    assert v.value == 0b101
    assert L.szo_fbits_unwind_bit(C.byref(b)) == 0
    L.szo_fbits_read(C.byref(b), 2, C.byref(v))
    assert v.value == 0b01  # bits 2,3 of the byte: bit2=1, bit3=0 -> 0b01
    L.szo_fbits_read(C.byref(b), 5, C.byref(v))  # bits 4..8 -> crosses into the next byte
    assert v.value == ((0b10110101 >> 4) | (1 << 4)) & 0x1F
    assert L.szo_fbits_unwind_bit(C.byref(b)) == 0  # un-reads the byte just fetched (bitstream.go:27-35)
    assert b.pos == 1 and b.offset == 8
    assert L.szo_fbits_read(C.byref(b), 16, C.byref(v)) == -32  # only 8 bits left -> EOF


def test_predefined_ll_table_matches_reference_golden():
    """fse/fse_test.go:8-41 expectedLLDecodingTable (the test function itself no longer compiles
    in the reference, the vector is valid: SURVEY.md section 4)"""
    L = pyszo.lib()
    with open(os.path.join(GOLDEN, "ll_table.json")) as f:
        want = json.load(f)["table"]
    t = _FseTable()
    assert L.szo_fse_build_ll_table(C.byref(t)) == 0
    assert t.table_size == 64 and t.accuracy_log == 6
    for i, w in enumerate(want):
        e = t.table[i]
        assert (e.baseline, e.additional_bits, e.number_of_bits, e.symbol) == (
            w["baseline"],
            w["additional_bits"],
            w["number_of_bits"],
            w["symbol"],
        ), i
    L.szo_fse_table_free(C.byref(t))


def test_predefined_ml_of_tables_are_well_formed():
    L = pyszo.lib()
    for fn, size in (("szo_fse_build_ml_table", 64), ("szo_fse_build_of_table", 32)):
        t = _FseTable()
        assert getattr(L, fn)(C.byref(t)) == 0
        assert t.table_size == size
        # every state's [baseline, baseline + 2^nb) range stays inside the table
        for i in range(size):
            e = t.table[i]
            assert e.baseline + (1 << e.number_of_bits) <= size
        L.szo_fse_table_free(C.byref(t))


def test_match_copy_semantics():
    """decompression/ringbuffer_test.go:85-154 TestRepeat, restated on a flat buffer:
    Repeat(n, after) == RepeatBeforeIndex(n, after+n) appends out[len-after-n : len-after]."""
    L = pyszo.lib()
    buf = (C.c_uint8 * 256)()
    pos = C.c_size_t(0)

    def push(s: bytes):
        for i, ch in enumerate(s):
            buf[pos.value + i] = ch
        pos.value += len(s)

    def repeat(n, after):
        assert L.szo_match_copy(buf, C.byref(pos), 256, n, after + n) == 0

    push(b"Teststring")
    repeat(4, 0)  # RepeatLast(4)
    assert bytes(buf[: pos.value]) == b"Teststringring"
    repeat(8, 0)
    assert bytes(buf[: pos.value])[-10:] == b"ngringring"
    push(b"1234567890")
    repeat(5, 3)
    assert bytes(buf[: pos.value])[-10:] == b"6789034567"
    repeat(3, 7)
    assert bytes(buf[: pos.value])[-10:] == b"9034567678"
    # overlapping match (offset < length): periodic extension, ringbuffer.go:247-274
    pos.value = 0
    push(b"ab")
    assert L.szo_match_copy(buf, C.byref(pos), 256, 7, 2) == 0
    assert bytes(buf[: pos.value]) == b"ababababa"
    assert L.szo_match_copy(buf, C.byref(pos), 256, 3, 1) == 0
    assert bytes(buf[: pos.value]) == b"ababababaaaa"
    # reaching before the start of the frame is an error (ringbuffer.go:203-214)
    assert L.szo_match_copy(buf, C.byref(pos), 256, 3, 100) == -30


def test_random_overlap_property():
    """decompression/ringbuffer_test.go:223-317 TestRandomRepeates: after[start+j] == after[start-oldest+j]"""
    import random

    L = pyszo.lib()
    rng = random.Random(3)
    buf = (C.c_uint8 * 4096)()
    pos = C.c_size_t(0)
    for i in range(64):
        buf[i] = rng.randrange(256)
    pos.value = 64
    for _ in range(300):
        n = rng.randrange(1, 12)
        oldest = rng.randrange(1, min(pos.value, 100) + 1)
        start = pos.value
        assert L.szo_match_copy(buf, C.byref(pos), 4096, n, oldest) == 0
        for j in range(n):
            assert buf[start + j] == buf[start - oldest + j]


def test_next_offset_table():
    """decompression/sequence_execution.go:65-114, the table in SURVEY.md A.9"""
    L = pyszo.lib()

    def step(h, ov, ll):
        hist = (C.c_int64 * 3)(*h)
        seq = _Sequence(3, ll, ov)
        off = L.szo_next_offset(C.byref(hist), C.byref(seq))
        return off, list(hist)

    h = [10, 20, 30]
    assert step(h, 1, 5) == (10, [10, 20, 30])
    assert step(h, 2, 5) == (20, [20, 10, 30])
    assert step(h, 3, 5) == (30, [30, 10, 20])
    assert step(h, 1, 0) == (20, [20, 10, 30])
    assert step(h, 2, 0) == (30, [30, 10, 20])
    assert step(h, 3, 0) == (9, [9, 10, 20])
    assert step(h, 7, 0) == (4, [4, 10, 20])
    assert step(h, 7, 9) == (4, [4, 10, 20])


def test_z000085_stage_trace(corpus):
    """Worked stage-level trace of decodecorpus_files/z000085.zst."""
    data = dict((n, d) for n, d, _, _ in corpus)["z000085.zst"]
    out, tr = pyszo.decode_frame(data, want_trace=True)
    assert len(out) == 179 and tr.window_size == 3840 and not tr.has_fcs
    b1, b2, b3, b4 = tr.blocks
    assert (b1.type, b1.block_size) == (0, 0)
    assert (b2.type, b2.block_size, b2.lit_type, b2.lit_streams, b2.lit_regen, b2.huf_max_bits) == (2, 48, 2, 1, 76, 4)
    assert b2.literals.hex().startswith("f3f3f3f3f3f3f3f3a4f3f3f3a4a4f3a4")
    assert b2.modes == (2, 2, 1)
    assert b2.sequences == [(1, 3, 4), (5, 3, 12), (9, 3, 5), (15, 3, 32), (0, 3, 31), (46, 3, 45)]
    assert b2.real_offsets == [1, 9, 2, 29, 28, 42]
    assert b2.hist_after == (42, 28, 29) and b2.out_len == 94
    assert (b3.lit_streams, b3.lit_regen, b3.huf_max_bits, b3.modes) == (4, 61, 8, (2, 2, 0))
    assert b3.sequences == [(39, 3, 72), (1, 6, 115), (5, 3, 1)]
    assert b3.real_offsets == [69, 112, 112]
    assert b3.hist_after == (112, 69, 42) and b3.out_len == 73
    assert (b4.type, b4.last, b4.block_size) == (0, 1, 12)
    assert out[-12:].hex() == "1b7199f47b95e2aa585c6728"


def test_huffman_direct_weights_build_and_decode():
    """structure/huffman.go:40-264 on a hand-made tree: weights [2,1] + implied last = 1 ->
    maxBits 2; code lengths {0:1, 1:2, 2:2}; longest codes at the lowest table indices."""
    L = pyszo.lib()
    desc = bytes([127 + 2, 0x21])
    w = (C.c_uint8 * 4096)()
    nw, used = C.c_int(), C.c_int()
    buf = (C.c_uint8 * len(desc)).from_buffer_copy(desc)
    assert L.szo_huf_decode_tree_desc(buf, len(desc), w, C.byref(nw), C.byref(used)) == 0
    assert (nw.value, used.value, list(w[:2])) == (2, 2, [2, 1])
    t = _HufTable()
    assert L.szo_huf_build(w, 2, C.byref(t)) == 0
    assert t.max_bits == 2 and [t.symbols[i] for i in range(4)] == [1, 2, 0, 0]
    assert [t.number_of_bits[i] for i in range(4)] == [2, 2, 1, 1]
    # stream read MSB-first from the last byte: padding '1' marker then codes 1(0) 00(1) 01(2) 1(0)
    # bits after marker: 1 00 01 1 -> last byte 0b1_1_00_01_1 = marker(1) + 6 bits + ... build explicitly
    bits = "1" + "1" + "00" + "01" + "1"  # marker, then symbols 0,1,2,0
    bits = bits.rjust(8, "0")
    stream = bytes([int(bits, 2)])
    sb = (C.c_uint8 * 1).from_buffer_copy(stream)
    out = (C.c_uint8 * 16)()
    n = C.c_int()
    assert L.szo_huf_decode_stream(C.byref(t), sb, 1, out, 16, C.byref(n)) == 0
    assert list(out[: n.value]) == [0, 1, 2, 0]
    L.szo_huf_table_free(C.byref(t))

"""CPU-side checks of the product's host logic: the C ABI library loads and exports every symbol
include/szb200.h declares, the header walker agrees with the oracle's per-block trace, and the
engine refuses to run without a CUDA device (no CPU fallback).  No GPU compute here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import pyszo
from tools import corpus as cg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import sparkzstd_b200

    L = sparkzstd_b200.load()
    hdr = open(os.path.join(ROOT, "include", "szb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(szb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    from sparkzstd_b200._lib import SYMBOLS

    assert declared == {s[0] for s in SYMBOLS}
    assert b"sm_100a" in L.szb_version()
    assert L.szb_strerror(-1) == b"Magicnum is not correct"  # framedecompressor.go:128


def test_struct_layouts_match_the_header():
    from sparkzstd_b200._lib import BlockDesc, FrameDesc

    assert C.sizeof(FrameDesc) == 64
    assert C.sizeof(BlockDesc) == 80
    # szb_abi_layout(): what the library was compiled with -- the ctypes mirror must agree field by field ...
    import sparkzstd_b200

    L = sparkzstd_b200.load()
    buf = (C.c_uint32 * 64)()
    n = L.szb_abi_layout(buf, 64)
    got = list(buf[:n])
    ff = [f[0] for f in FrameDesc._fields_]
    bf = [f[0] for f in BlockDesc._fields_ if not f[0].startswith("_pad")]
    want = [C.sizeof(FrameDesc)] + [getattr(FrameDesc, f).offset for f in ff] + [C.sizeof(BlockDesc)] + [getattr(BlockDesc, f).offset for f in bf]
    assert n == len(want) == 38 and got == want
    # ... and so must the Go mirror (go/szb200/szb200.go: plain structs with exported fields, same order, same widths; Go
    # aligns every field naturally, as the C compiler does): recompute its layout from the declaration
    import os
    import re

    go = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "go", "szb200", "szb200.go")).read()
    width = {"uint64": 8, "uint32": 4, "int32": 4, "uint8": 1}

    def go_layout(name):
        body = re.search(r"type %s struct \{(.*?)\n\}" % name, go, re.S).group(1)
        off, offs, align = 0, [], 1
        for line in body.splitlines():
            m = re.match(r"\s*(\w+)\s+(uint64|uint32|int32|uint8)\b", line)
            if not m:
                continue
            w = width[m.group(2)]
            off = (off + w - 1) // w * w
            offs.append((m.group(1), off))
            off += w
            align = max(align, w)
        return (off + align - 1) // align * align, offs

    fsize, foffs = go_layout("FrameDesc")
    bsize, boffs = go_layout("BlockDesc")
    boffs = [o for o in boffs if not o[0].startswith("Pad")]
    assert "Flags" in [n_ for n_, _ in boffs]
    assert [fsize] + [o for _, o in foffs] + [bsize] + [o for _, o in boffs] == got
    # same field order by name (snake_case <-> CamelCase)
    camel = lambda s: "".join(p.upper() if p in ("ll", "of", "ml", "id") else p.capitalize() for p in s.split("_"))
    fix = {"Nblocks": "NBlocks", "Nseq": "NSeq"}
    assert [fix.get(camel(f), camel(f)) for f in ff] == [n_ for n_, _ in foffs]
    assert [fix.get(camel(f), camel(f)) for f in bf] == [n_ for n_, _ in boffs]


def test_no_cpu_fallback_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from sparkzstd_b200 import SzbError
    from sparkzstd_b200.decompression import Context

    with pytest.raises(SzbError) as e:
        Context(0)
    assert e.value.code == -67


def _walk(data: bytes):
    from sparkzstd_b200.decompression import Walk

    src = np.frombuffer(data + b"\0" * 8, dtype=np.uint8)
    with Walk(src, [0], [len(data)]) as w:
        return w.frames()[0], w.blocks()


def test_walker_matches_oracle_trace_on_the_corpus(corpus):
    for name, data, size, _ in corpus:
        fr, blocks = _walk(data)
        _, tr = pyszo.decode_frame(data, want_trace=True)
        assert fr.status == 0, name
        assert fr.src_len == tr.bytes_consumed, name
        assert fr.window_size == tr.window_size and bool(fr.has_content_size) == tr.has_fcs, name
        if tr.has_fcs:
            assert fr.content_size == tr.frame_content_size == size
        assert fr.has_checksum == 1  # every corpus frame carries one; the reference leaves it unread
        assert len(blocks) == len(tr.blocks), name
        for d, b in zip(blocks, tr.blocks):
            assert (d.type, d.last, d.block_size) == (b.type, b.last, b.block_size), name
            if b.type == 2:
                assert (d.lit_type, d.lit_streams, d.lit_regen, d.nseq) == (b.lit_type, b.lit_streams, b.lit_regen, b.nseq), name
                if b.nseq:
                    assert (d.seq_modes >> 6, (d.seq_modes >> 4) & 3, (d.seq_modes >> 2) & 3) == b.modes, name


def test_walker_origin_resolution(corpus):
    """Treeless literals and Repeat-mode tables point at an earlier block of the same frame whose own
    mode is not Repeat (carry rules framedecompressor.go:283-294)."""
    seen = {"treeless": 0, "ll3": 0, "of3": 0, "ml3": 0}
    for name, data, _, _ in corpus:
        fr, blocks = _walk(data)
        for i, d in enumerate(blocks):
            if d.type != 2:
                continue
            if d.lit_type == 3:
                o = blocks[d.huf_origin]
                assert d.huf_origin < i and o.type == 2 and o.lit_type == 2
                seen["treeless"] += 1
            elif d.lit_type == 2:
                assert d.huf_origin == i
            if d.nseq:
                for key, org, sh in (("ll3", d.ll_origin, 6), ("of3", d.of_origin, 4), ("ml3", d.ml_origin, 2)):
                    if (d.seq_modes >> sh) & 3 == 3:
                        o = blocks[org]
                        assert org < i and o.nseq > 0 and (o.seq_modes >> sh) & 3 != 3
                        seen[key] += 1
                    else:
                        assert org == i
    assert seen == {"treeless": 520, "ll3": 83, "of3": 208, "ml3": 147}


def test_walker_error_codes():
    fr, blocks = _walk(b"\x28\xb5\x2f\xfe\x00\x00\x00\x00")
    assert fr.status == -1 and not blocks
    fr, _ = _walk(b"\x28\xb5")
    assert fr.status == -32
    # reserved block type 3 (block.go:45-47)
    fr, _ = _walk(bytes.fromhex("28b52ffd0400") + bytes([0b111, 0, 0]))
    assert fr.status == -7
    # block size > 128 KiB (block.go:50-52)
    fr, _ = _walk(bytes.fromhex("28b52ffd0400") + bytes([0b001 | (1 << 3), 0, 0x20]))
    assert fr.status == -8
    # treeless literals with no earlier table (literals.go:247-252)
    fr, _ = _walk(bytes.fromhex("28b52ffd0400") + bytes([0b101 | (4 << 3), 0, 0]) + bytes([0x03, 0x00, 0x00, 0x00]))
    assert fr.status == -14


def test_walker_multi_frame_discovery(corpus):
    """frame_off == NULL: concatenated frames, checksum stepped over, skippable frames skipped."""
    from sparkzstd_b200.decompression import Walk

    skippable = bytes.fromhex("502a4d18") + (5).to_bytes(4, "little") + b"hello"
    blob = corpus[0][1] + skippable + corpus[1][1] + corpus[2][1]
    src = np.frombuffer(blob + b"\0" * 8, dtype=np.uint8)[: len(blob)]
    with Walk(src) as w:
        fr = w.frames()
        assert w.nframes == 3 and all(f.status == 0 for f in fr)
        assert fr[1].src_off == len(corpus[0][1]) + len(skippable)
        assert fr[2].src_off + fr[2].src_len + 4 == len(blob)


def test_synthetic_corpora_have_the_expected_shapes():
    """BASELINE.md section 2 / SURVEY Appendix C: the generator reproduces the named shapes."""
    c = cg.config2_text_frames(16)
    for i in range(4):
        _, tr = pyszo.decode_frame(c.frame(i), want_trace=True)
        assert tr.single_segment and tr.has_fcs and len(tr.blocks) == 1
        b = tr.blocks[0]
        assert 5000 < b.nseq < 7000 and 11000 < b.lit_regen < 16000 and b.lit_streams == 4 and b.modes == (2, 2, 2)
    assert 20000 < c.compressed_bytes / c.nframes < 26000
    c4 = cg.config4_literal_heavy(1, 1 << 20)
    _, tr = pyszo.decode_frame(c4.frame(0), want_trace=True)
    assert all(b.nseq == 0 and b.lit_regen == 131072 and b.lit_streams == 4 for b in tr.blocks)
    c3 = cg.config3_single_frame(4 << 20, window_log=20)
    _, tr = pyszo.decode_frame(c3.frame(0), want_trace=True)
    assert not tr.has_fcs and not tr.single_segment and tr.window_size == 1 << 20
    assert sum(1 for b in tr.blocks if b.type == 2 and b.lit_type == 3) >= len(tr.blocks) // 2


def _xxh64(data: bytes) -> int:
    M = (1 << 64) - 1
    P1, P2, P3, P4, P5 = 0x9E3779B185EBCA87, 0xC2B2AE3D27D4EB4F, 0x165667B19E3779F9, 0x85EBCA77C2B2AE63, 0x27D4EB2F165667C5
    rotl = lambda x, r: ((x << r) | (x >> (64 - r))) & M
    rnd = lambda acc, v: (rotl((acc + v * P2) & M, 31) * P1) & M
    n, p = len(data), 0
    if n >= 32:
        v = [(P1 + P2) & M, P2, 0, (-P1) & M]
        while p + 32 <= n:
            for i in range(4):
                v[i] = rnd(v[i], int.from_bytes(data[p + 8 * i : p + 8 * i + 8], "little"))
            p += 32
        h = (rotl(v[0], 1) + rotl(v[1], 7) + rotl(v[2], 12) + rotl(v[3], 18)) & M
        for x in v:
            h = ((h ^ rnd(0, x)) * P1 + P4) & M
    else:
        h = P5
    h = (h + n) & M
    while p + 8 <= n:
        h = (rotl(h ^ rnd(0, int.from_bytes(data[p : p + 8], "little")), 27) * P1 + P4) & M
        p += 8
    if p + 4 <= n:
        h = (rotl(h ^ (int.from_bytes(data[p : p + 4], "little") * P1 & M), 23) * P2 + P3) & M
        p += 4
    while p < n:
        h = (rotl(h ^ (data[p] * P5 & M), 11) * P1) & M
        p += 1
    h ^= h >> 33
    h = (h * P2) & M
    h ^= h >> 29
    h = (h * P3) & M
    h ^= h >> 32
    return h


def test_walker_reports_the_content_checksum(corpus):
    """The 4 bytes after the last block are XXH64(content) & 0xFFFFFFFF (zstd format); the walker hands
    them to the optional GPU verification (the reference never reads them)."""
    assert _xxh64(b"") == 0xEF46DB3751D8E999
    checked = 0
    for name, data, size, _ in corpus:
        if size > 20000:
            continue
        fr, _ = _walk(data)
        assert fr.checksum_valid == 1
        assert fr.checksum == _xxh64(pyszo.decode_frame(data)) & 0xFFFFFFFF, name
        checked += 1
    assert checked > 40


def test_parallel_walk_makes_the_same_tables(monkeypatch):
    """Batches of >= 4096 frames with given extents are walked on several threads (SZB_WALK_THREADS caps them); the merged
    tables -- first blocks, table origins, scratch offsets, a dictionary's row as origin -- equal the single-threaded ones."""
    import sparkzstd_b200
    from sparkzstd_b200._lib import BlockDesc, FrameDesc

    L = sparkzstd_b200.load()
    L.szb_walk_literal_bytes.restype = C.c_uint64
    L.szb_walk_sequences.restype = C.c_uint64
    from tools import corpus as cg

    text = cg.config2_text_frames(3000, 6000)
    mixed = cg.config5_mixed(96 << 20)  # multi-block frames with Treeless literals and Repeat modes, the golden frames
    # 3000 + ~900 frames are not 4096 yet: the same frames twice
    src = np.concatenate([text.src[: text.compressed_bytes], mixed.src])
    off = np.concatenate([text.frame_off, mixed.frame_off + np.uint64(text.compressed_bytes)] * 2)
    ln = np.concatenate([text.frame_len, mixed.frame_len] * 2)
    assert len(off) >= 4096

    def walk(threads, dict_blk=None):
        monkeypatch.setenv("SZB_WALK_THREADS", str(threads))
        h = C.c_void_p()
        if dict_blk is None:
            rc = L.szb_walk_create(src.ctypes.data, src.nbytes, off.ctypes.data, ln.ctypes.data, len(off), C.byref(h))
        else:
            rc = L.szb_walk_create_dict(src.ctypes.data, src.nbytes, off.ctypes.data, ln.ctypes.data, len(off), C.byref(dict_blk), 0, C.byref(h))
        assert rc == 0
        nf, nb = L.szb_walk_nframes(h), L.szb_walk_nblocks(h)
        out = (C.string_at(L.szb_walk_frames(h), nf * C.sizeof(FrameDesc)), C.string_at(L.szb_walk_blocks(h), nb * C.sizeof(BlockDesc)),
               L.szb_walk_literal_bytes(h), L.szb_walk_sequences(h))
        L.szb_walk_destroy(h)
        return out

    assert walk(1) == walk(5) == walk(8)
    blk = BlockDesc()
    blk.type, blk.lit_type, blk.nseq, blk.block_size, blk.lit_streams, blk.lit_hdr_bytes, blk.seq_hdr_bytes, blk.seq_modes = 2, 2, 1, 10, 1, 3, 2, 0xA8
    assert walk(1, blk) == walk(7, blk)

"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every check goes through the C ABI
(libszb200.so) and compares with the ORACLE on the same inputs, or with the golden corpus."""
import hashlib
import io

import numpy as np
import pytest

from oracle import pyszo
from tools import corpus as cg

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["exec2", "legacy", "place"])
def stage4_path(request, monkeypatch):
    """Every test runs three times: stage 4 of the frames one warp executes by k_execute2 (the default), by k_execute, and by
    k_resolve + k_place (SZB_EXEC is read when a batch's tables are built)."""
    monkeypatch.setenv("SZB_EXEC", request.param)
    return request.param


@pytest.fixture(scope="module")
def ctx():
    from sparkzstd_b200.decompression import Context

    c = Context(0)
    yield c
    c.close()


def _batch_arrays(frames):
    src = np.frombuffer(b"".join(frames) + b"\0" * 16, dtype=np.uint8)
    lens = np.array([len(f) for f in frames], dtype=np.uint64)
    offs = (np.cumsum(lens) - lens).astype(np.uint64)
    return src, offs, lens


# ---- config 1: the reference's golden corpus ------------------------------------------------------
def test_decodecorpus_batch_byte_exact(ctx, corpus):
    """All 100 decodecorpus frames in ONE batch launch, byte-exact vs the originals' sha256."""
    frames = [d for _, d, _, _ in corpus]
    outs = ctx.decode_batch(frames)
    bad = [n for (n, _, size, sha), o in zip(corpus, outs) if len(o) != size or hashlib.sha256(o).hexdigest() != sha]
    assert not bad, bad


def test_decodecorpus_frame_reader_api(ctx, corpus):
    """NewFrameReader(r) / Reset(r) / Read, as cmd/sparkzstd/main.go:59-68,126 uses it."""
    from sparkzstd_b200.decompression import NewFrameReader

    comp = NewFrameReader(None, ctx)
    for name, data, size, sha in corpus[::7]:
        comp.Reset(io.BytesIO(data))
        h = hashlib.sha256()
        total = 0
        while True:
            chunk = comp.Read(4096)
            if not chunk:
                break
            h.update(chunk)
            total += len(chunk)
        assert total == size and h.hexdigest() == sha, name


def test_decodecorpus_frame_decompressor_api(ctx, corpus):
    """NewFrameDecompressor(src, dst).Decompress(), as cmd/sparkzstd/main.go:22-40 uses it."""
    from sparkzstd_b200.decompression import NewFrameDecompressor

    for name, data, size, sha in corpus[3::9]:
        target = io.BytesIO()
        fd = NewFrameDecompressor(io.BytesIO(data), target, ctx)
        fd.Decompress()
        out = target.getvalue()
        assert len(out) == size and hashlib.sha256(out).hexdigest() == sha, name
        assert fd.BlockCounter == len(pyszo.decode_frame(data, want_trace=True)[1].blocks)


def test_stage_level_parity_with_oracle_trace(ctx, corpus):
    """Literal buffers (stage 2), sequence triples (stage 3) and per-block regenerated sizes of every
    compressed block of the corpus equal the oracle's per-block trace."""
    from sparkzstd_b200.decompression import Batch

    frames = [d for _, d, _, _ in corpus]
    src, offs, lens = _batch_arrays(frames)
    b = Batch(ctx, src, offs, lens)
    try:
        d_src = b.upload(src)
        b.decode_entropy(d_src)
        total, foff, flen = b.sizes()
        sizes, status = b.read_block_results()
        assert not status.any()
        assert total == sum(s for _, _, s, _ in corpus)
        bi = 0
        checked_lit = checked_seq = 0
        for fi, (name, data, size, _) in enumerate(corpus):
            _, tr = pyszo.decode_frame(data, want_trace=True)
            assert int(flen[fi]) == size, name
            for blk in tr.blocks:
                assert int(sizes[bi]) == blk.out_len, (name, bi)
                if blk.type == 2 and fi % 3 == 0:
                    if blk.lit_type >= 2:
                        assert b.read_literals(bi, blk.lit_regen) == blk.literals, (name, bi)
                        checked_lit += 1
                    if blk.nseq:
                        ll, ml, of = b.read_sequences(bi, blk.nseq)
                        want = np.array(blk.sequences, dtype=np.int64)
                        assert (ll == want[:, 0]).all() and (ml == want[:, 1]).all() and (of == want[:, 2]).all(), (name, bi)
                        checked_seq += 1
                bi += 1
        assert bi == b.nblocks and checked_lit > 100 and checked_seq > 100
    finally:
        b.close()


# ---- synthetic corpora of the BASELINE.json shapes, small enough for the oracle ----------------------
def _check_corpus_vs_oracle(ctx, c):
    frames = [c.frame(i) for i in range(c.nframes)]
    outs = ctx.decode_batch(frames)
    for i, (f, o) in enumerate(zip(frames, outs)):
        want = pyszo.decode_frame(f)
        assert o == want, (c.name, i, len(o), len(want))
        if int(c.raw_hash[i]):
            assert cg.hash_bytes(np.frombuffer(o, dtype=np.uint8)) == int(c.raw_hash[i])


def test_config2_text_frames_small(ctx):
    _check_corpus_vs_oracle(ctx, cg.config2_text_frames(96))


def test_config2_with_checksums(ctx):
    _check_corpus_vs_oracle(ctx, cg.config2_text_frames(8, checksum=1))


def test_config4_literal_heavy_small(ctx):
    _check_corpus_vs_oracle(ctx, cg.config4_literal_heavy(4, 1 << 20))


def test_config3_single_frame_small(ctx):
    """One streamed multi-block frame: no content size, Treeless literals, cross-block matches."""
    c = cg.config3_single_frame(12 << 20, window_log=20)
    _check_corpus_vs_oracle(ctx, c)


def test_config5_mixed_small(ctx):
    c = cg.config5_mixed(24 << 20)
    _check_corpus_vs_oracle(ctx, c)


def test_ragged_frame_sizes(ctx):
    sizes = [0, 1, 2, 3, 7, 63, 64, 65, 255, 256, 1000, 4095, 4096, 4097, 65535, 65537, 131071, 131072, 131073, 300001]
    kinds = [cg.KIND_TEXT] * len(sizes)
    c = cg.generate("ragged", np.array(kinds, np.int32), 77 + np.arange(len(sizes), dtype=np.uint64), np.array(sizes, np.uint64))
    _check_corpus_vs_oracle(ctx, c)
    c = cg.generate("ragged_const", np.full(len(sizes), cg.KIND_CONSTANT, np.int32), 99 + np.arange(len(sizes), dtype=np.uint64),
                    np.array(sizes, np.uint64))
    _check_corpus_vs_oracle(ctx, c)


# ---- size-independent properties at a larger size -----------------------------------------------------
def test_config2_medium_hash_of_hashes(ctx):
    """4096 frames (256 MiB): per-frame hash of the GPU output equals the generator's hash of the original."""
    c = cg.config2_text_frames(4096)
    dst = np.empty(c.decompressed_bytes + 64, dtype=np.uint8)
    out_off, out_len, status = ctx.decode_batch_into(c.src, c.frame_off, c.frame_len, dst)
    assert not status.any()
    assert (out_len == c.raw_size).all()
    got = cg.hash_frames(dst, out_off, out_len)
    assert (got == c.raw_hash).all()
    # idempotence: decoding the same batch again gives the same bytes
    dst2 = np.empty_like(dst)
    ctx.decode_batch_into(c.src, c.frame_off, c.frame_len, dst2)
    assert (cg.hash_frames(dst2, out_off, out_len) == c.raw_hash).all()


def test_large_host_batch_takes_the_pipelined_path(ctx):
    """> 192 MB of compressed frames with content sizes: szb_decode_batch overlaps H2D / kernels / D2H in
    chunks of consecutive frames; the result must be indistinguishable from the plain path."""
    c = cg.config2_text_frames(10000)
    assert c.compressed_bytes > (192 << 20)
    dst = np.empty(c.decompressed_bytes + 64, dtype=np.uint8)
    out_off, out_len, status = ctx.decode_batch_into(c.src, c.frame_off, c.frame_len, dst)
    assert not status.any() and (out_len == c.raw_size).all()
    assert (out_off == np.cumsum(c.raw_size) - c.raw_size).all()  # frames back to back, in frame order
    assert (cg.hash_frames(dst, out_off, out_len) == c.raw_hash).all()


def test_pipelined_path_with_a_frame_that_lies_about_its_size(ctx):
    """One frame of a > 192 MB batch declares 5000 bytes, regenerates 1000 and then fails (reserved block type): the chunked
    path places frames by their declared sizes, so it must notice that the failed frame shifted its neighbours and leave the
    batch to the plain path.  Every other frame must come back whole."""
    c = cg.config2_text_frames(10000)
    assert c.compressed_bytes > (192 << 20)
    liar = bytes([0x28, 0xB5, 0x2F, 0xFD, 0xA0]) + (5000).to_bytes(4, "little") + bytes([0x40, 0x1F, 0x00]) + bytes(range(250)) * 4 + bytes([0x07, 0x00, 0x00])
    k = 5000
    src = c.src.copy()
    off, length = int(c.frame_off[k]), int(c.frame_len[k])
    assert len(liar) <= length
    src[off : off + len(liar)] = np.frombuffer(liar, dtype=np.uint8)
    frame_len = c.frame_len.copy()
    frame_len[k] = len(liar)
    with pytest.raises(pyszo.OracleError) as oracle_says:
        pyszo.decode_frame(liar)
    dst = np.empty(c.decompressed_bytes + 64, dtype=np.uint8)
    out_off, out_len, status = ctx.decode_batch_into(src, c.frame_off, frame_len, dst)
    assert status[k] == oracle_says.value.code and out_len[k] == 0
    ok = np.ones(c.nframes, dtype=bool)
    ok[k] = False
    assert not status[ok].any() and (out_len[ok] == c.raw_size[ok]).all()
    got = cg.hash_frames(dst, out_off, out_len)
    assert (got[ok] == c.raw_hash[ok]).all()


def test_streamed_reader_hands_out_pieces_and_reads_in_pieces(ctx):
    """szb_decompress_reader behind FrameReader.Read / FrameDecompressor.Decompress: a 96 MiB frame read from a
    source that returns short counts; the output arrives in pieces (the first Read returns before everything was copied back)
    and equals the generator's original."""
    import hashlib as H

    from sparkzstd_b200 import decompression as D

    c = cg.config3_single_frame(96 << 20, 20)
    frame = c.frame(0)

    class Dribble(io.RawIOBase):  # a reader that never returns more than 1 MiB + 17 bytes
        def __init__(self, data):
            self.data, self.pos, self.calls = data, 0, 0

        def read(self, n=-1):
            self.calls += 1
            n = len(self.data) - self.pos if n is None or n < 0 else n
            n = min(n, (1 << 20) + 17)
            out = self.data[self.pos : self.pos + n]
            self.pos += len(out)
            return out

    src = Dribble(frame)
    r = D.NewFrameReader(src, ctx)
    assert src.pos <= 18  # only the frame header was taken eagerly (framereader.go:22-31)
    h, total, reads = H.sha256(), 0, 0
    first = r.Read(1 << 16)
    assert 0 < len(first) <= (1 << 16)
    h.update(first)
    total += len(first)
    while True:
        chunk = r.Read(8 << 20)
        if not chunk:
            break
        h.update(chunk)
        total += len(chunk)
        reads += 1
    assert total == int(c.raw_size[0]) and reads >= 3 and src.calls > 20
    target = io.BytesIO()
    D.NewFrameDecompressor(Dribble(frame), target, ctx).Decompress()
    assert H.sha256(target.getvalue()).digest() == h.digest()
    got = cg.hash_frames(np.frombuffer(target.getvalue(), dtype=np.uint8), np.array([0], np.uint64), np.array([total], np.uint64))
    assert got[0] == c.raw_hash[0]
    # errors come through the same way: a truncated frame, a failing source, a failing target
    with pytest.raises(D.ErrUnexpectedEOF):
        D.NewFrameDecompressor(io.BytesIO(frame[: len(frame) // 2]), io.BytesIO(), ctx).Decompress()

    class Boom(io.RawIOBase):
        def read(self, n=-1):
            raise OSError("source went away")

        def write(self, b):
            raise OSError("disk full")

    with pytest.raises(OSError, match="source went away"):
        D.NewFrameDecompressor(Boom(), io.BytesIO(), ctx).Decompress()
    with pytest.raises(OSError, match="disk full"):
        D.NewFrameDecompressor(io.BytesIO(frame), Boom(), ctx).Decompress()
    # a reader that is dropped half way must not wedge the context
    r = D.NewFrameReader(io.BytesIO(frame), ctx)
    assert len(r.Read(1000)) > 0
    r.close()
    assert ctx.decode_batch([frame[: len(frame)]])[0][:16] == target.getvalue()[:16]


def test_pageable_host_buffers_go_through_the_pinned_rings(ctx):
    """szb_decode_batch with ordinary (pageable) numpy buffers and > 192 MB of input: the chunked path stages both directions
    through the context's pinned slots; same bytes as with the caller's memory pinned."""
    c = cg.config2_text_frames(10000)
    src = np.array(c.src, copy=True)  # plain malloc'd memory
    dst = np.empty(c.decompressed_bytes + 64, dtype=np.uint8)
    out_off, out_len, status = ctx.decode_batch_into(src, c.frame_off, c.frame_len, dst)
    assert not status.any() and (out_len == c.raw_size).all()
    assert (cg.hash_frames(dst, out_off, out_len) == c.raw_hash).all()


def test_dictionaries_formatted_and_raw(ctx):
    """SURVEY 8f-4 (not in the reference, which ignores Dictionary_ID): frames libzstd compressed with a dictionary, decoded in one
    batch with szb_decode_batch_dict, equal the originals and the oracle's dictionary extension; matches reach into the
    dictionary's content, first blocks use its Huffman / FSE tables and its repeat offsets."""
    from sparkzstd_b200 import decompression as D

    msgs = cg.dictionary_messages(400)
    formatted = cg.train_dictionary(msgs[:300], 8 << 10)
    raw = b"".join(msgs[:40])
    bigger = [bytes(cg.fill(cg.KIND_TEXT, 4242 + i, 30000 + 977 * i, 0)) for i in range(3)]
    originals = msgs[300:400] + bigger + [b"", b"x"]
    for d in (formatted, raw):
        frames = [cg.compress_with_dict(m, d) for m in originals]
        outs = ctx.decode_batch_dict(frames, d)
        assert outs == originals
        assert outs == [pyszo.decode_frame(f, dictionary=d) for f in frames]
    # a frame that names another dictionary id, and a frame decoded without its dictionary
    frames = [cg.compress_with_dict(m, formatted) for m in originals[:8]]
    other = bytearray(formatted)
    other[4] ^= 0x55  # another Dictionary_ID
    with pytest.raises(D.SzbError) as e:
        ctx.decode_batch_dict(frames, bytes(other))
    assert e.value.code == -70
    # plain frames decode with a dictionary given, too (they name no dictionary and never reach in front of themselves)
    plain = [c for c in (cg.config2_text_frames(3).frame(i) for i in range(3))]
    assert ctx.decode_batch_dict(plain, raw) == [pyszo.decode_frame(f) for f in plain]


# ---- edge cases and error behaviour --------------------------------------------------------------------
def test_empty_batch_and_empty_frames(ctx, corpus):
    assert ctx.decode_batch([]) == []
    empties = [d for _, d, s, _ in corpus if s == 0]
    assert len(empties) == 3
    assert ctx.decode_batch(empties) == [b"", b"", b""]


def test_wrong_magic_and_truncation(ctx, corpus):
    from sparkzstd_b200 import decompression as D

    with pytest.raises(D.ErrWrongMagicnumber):
        D.NewFrameReader(io.BytesIO(b"\x28\xb5\x2f\xfe" + b"\0" * 16), ctx)
    name, data, size, _ = corpus[0]
    src, offs, lens = _batch_arrays([data[: len(data) // 2], data, b"\x00\x01\x02\x03\x04"])
    dst = np.empty(size * 2 + 64, dtype=np.uint8)
    out_off, out_len, status = ctx.decode_batch_into(src, offs, lens, dst)
    assert status[0] == -32 and status[1] == 0 and status[2] == -1
    assert int(out_len[1]) == size and int(out_len[0]) == 0
    o = int(out_off[1])
    assert hashlib.sha256(dst[o : o + size].tobytes()).hexdigest() == corpus[0][3]


def test_destination_too_small(ctx, corpus):
    name, data, size, _ = corpus[0]
    src, offs, lens = _batch_arrays([data])
    dst = np.empty(size - 1, dtype=np.uint8)
    _, out_len, status = ctx.decode_batch_into(src, offs, lens, dst)
    assert status[0] == -64 and int(out_len[0]) == 0


# Error identity: a frame the oracle rejects must end with the oracle's code.  The engine is deliberately
# STRICTER in two places; nothing else may differ:
#   -19  a single-stream Huffman block whose stream decodes fewer symbols than its regenerated size: the reference leaves
#        the rest of its (reused) literal buffer as it was and goes on; the engine reports ErrDidntUseAllBitsToDecodeHuffman
#   -35  beyond the zstd format limits the engine enforces (accuracy logs, symbol counts, Huffman depth, offset codes)
ENGINE_STRICTER = (-19, -35)
# and where both fail, the engine may name an EARLIER or equivalent cause for the same corrupt bytes:
#   the reference panics (-33: index out of range, negative slice) where the engine's bounds check names the field
EQUIVALENT_CODES = {(-33, -32), (-33, -35), (-33, -10), (-33, -15), (-33, -19), (-32, -33), (-15, -33), (-19, -15), (-15, -19), (-20, -33), (-33, -20), (-2, -32), (-32, -2)}


def _mutants(corpus, rng, n, flips=3):
    out = []
    small = [(nm, d, sz) for nm, d, sz, _ in corpus if len(d) >= 40 and sz <= 300_000]
    for k in range(n):
        nm, data, size = small[k % len(small)]
        buf = bytearray(data)
        for _ in range(1 + k % flips):
            p = int(rng.integers(6, len(buf) - 4))
            buf[p] ^= 1 << int(rng.integers(0, 8))
        out.append((nm, bytes(buf)))
    return out


def test_corrupted_payloads_end_with_the_oracles_verdict(ctx, corpus, stage4_path):
    """Bit flips in the payload: the engine must return (never hang or fault).  Whenever the oracle still decodes the frame
    the GPU output is identical (or the engine is stricter in one of the two documented ways); whenever the oracle rejects it
    the engine rejects it too, and with the oracle's own code (error identity)."""
    rng = np.random.default_rng(5)
    muts = _mutants(corpus, rng, 600)
    frames = [f for _, f in muts]
    src, offs, lens = _batch_arrays(frames)
    dst = np.empty(96 << 20, dtype=np.uint8)
    out_off, out_len, status = ctx.decode_batch_into(src, offs, lens, dst)
    same_bytes = same_code = stricter = 0
    other, unexplained = [], []
    for i, (nm, f) in enumerate(muts):
        want, code = None, 0
        try:
            want = pyszo.decode_frame(f)
        except pyszo.OracleError as e:
            code = e.code
        st = int(status[i])
        if want is not None:
            if st == 0:
                o, l = int(out_off[i]), int(out_len[i])
                assert dst[o : o + l].tobytes() == want, (nm, i)
                same_bytes += 1
            else:
                assert st in ENGINE_STRICTER, (nm, i, st)
                stricter += 1
        else:
            assert st != 0, (nm, i, "the engine decoded a frame the oracle rejects", code)
            if st == code:
                same_code += 1
            elif st in ENGINE_STRICTER or (code, st) in EQUIVALENT_CODES:
                other.append((code, st))
            else:
                unexplained.append((nm, i, "oracle", code, "engine", st))
    assert not unexplained, unexplained
    assert same_bytes >= 100 and same_code >= 100, (same_bytes, same_code, stricter, other)
    assert len(other) <= same_code // 4, other  # equivalents stay the exception


def test_content_checksums_verify_on_the_gpu(ctx, corpus):
    """every decodecorpus frame carries a content checksum the reference never reads; the
    engine can verify it (XXH64 low 32 bits) and must flag a frame whose checksum bytes were tampered with."""
    frames = [d for _, d, _, _ in corpus]
    bad_idx = 5
    tampered = bytearray(frames[bad_idx])
    tampered[-1] ^= 0x40
    frames[bad_idx] = bytes(tampered)
    src, offs, lens = _batch_arrays(frames)
    dst = np.empty(sum(s for _, _, s, _ in corpus) + 64, dtype=np.uint8)
    out_off, out_len, status = ctx.decode_batch_into(src, offs, lens, dst, verify_checksum=True)
    assert status[bad_idx] == -68 and int(out_len[bad_idx]) == 0
    assert not np.delete(status, bad_idx).any()
    # without the flag the same tampered frame decodes fine, like in the reference
    _, _, status2 = ctx.decode_batch_into(src, offs, lens, dst)
    assert not status2.any()
    # and the synthetic frames with checksums
    c = cg.config2_text_frames(8, checksum=1)
    d2 = np.empty(c.decompressed_bytes + 64, dtype=np.uint8)
    _, _, st3 = ctx.decode_batch_into(c.src, c.frame_off, c.frame_len, d2, verify_checksum=True)
    assert not st3.any()


def test_concatenated_and_skippable_stream_through_the_decode_entry(ctx, corpus, tmp_path):
    """szb_decode_stream (not a reference behaviour: the reference decodes one frame per reader and knows no
    skippable frames): frames back to back with skippable frames in between and content checksums after each -- the stream's
    content is the concatenation of what the oracle decodes frame by frame."""
    picks = [corpus[i] for i in (0, 3, 7, 11, 19, 23)]
    skip = lambda k, body: (0x184D2A50 + k).to_bytes(4, "little") + len(body).to_bytes(4, "little") + body
    blob = skip(0, b"") + picks[0][1] + skip(5, b"metadata") + picks[1][1] + picks[2][1] + skip(15, bytes(300)) + picks[3][1] + picks[4][1] + picks[5][1]
    want = b"".join(pyszo.decode_frame(d) for _, d, _, _ in picks)
    assert ctx.decode_stream(blob) == want
    assert ctx.decode_stream(blob, verify_checksum=True) == want  # every golden frame carries its XXH64
    assert ctx.decode_stream(b"") == b"" and ctx.decode_stream(skip(1, b"only a skippable frame")) == b""
    # a flipped content byte of the third frame: the checksum of that frame no longer matches
    bad = bytearray(blob)
    at = len(skip(0, b"")) + len(picks[0][1]) + len(skip(5, b"metadata")) + len(picks[1][1]) + len(picks[2][1]) - 5
    bad[at] ^= 0xFF
    with pytest.raises(Exception):
        ctx.decode_stream(bytes(bad), verify_checksum=True)
    # trailing garbage is no frame: the walk ends with its error
    with pytest.raises(Exception):
        ctx.decode_stream(blob + b"\x01\x02\x03\x04\x05\x06")
    # the CLI's --stream mode writes the stream's content
    from sparkzstd_b200 import cli

    f = tmp_path / "many.zst"
    f.write_bytes(blob)
    (tmp_path / "many").write_bytes(want)
    assert cli.main(["--stream", str(f)]) == 0


def test_cli_compares_with_originals(ctx, corpus, tmp_path):
    """cmd/sparkzstd equivalent: decode X.zst, compare with X (main.go:46-111)."""
    from sparkzstd_b200 import cli

    paths = []
    for name, data, size, _ in corpus[10:14]:
        p = tmp_path / name
        p.write_bytes(data)
        (tmp_path / name[:-4]).write_bytes(pyszo.decode_frame(data))
        paths.append(str(p))
    assert cli.main(paths + ["--verify-checksum"]) == 0
    (tmp_path / corpus[10][0][:-4]).write_bytes(b"not the original")
    assert cli.main(paths) == 1


def test_descriptor_level_entry(ctx, corpus):
    """szb_decode_blocks: host-built descriptor tables + device pointers (what the Go walker feeds)."""
    import ctypes as C

    import torch

    from sparkzstd_b200.decompression import Walk

    frames = [d for _, d, _, _ in corpus[:20]]
    src, offs, lens = _batch_arrays(frames)
    total = sum(s for _, _, s, _ in corpus[:20])
    with Walk(src, offs, lens) as w:
        d_src = torch.from_numpy(src.copy()).cuda()
        d_dst = torch.zeros(total + 64, dtype=torch.uint8, device="cuda")
        n = w.nframes
        out_off = np.zeros(n, np.uint64)
        out_len = np.zeros(n, np.uint64)
        status = np.zeros(n, np.int32)
        rc = ctx._L.szb_decode_blocks(ctx._h, C.c_void_p(d_src.data_ptr()), src.nbytes, w.frames_ptr(), n, w.blocks_ptr(), w.nblocks,
                                      C.c_void_p(d_dst.data_ptr()), total + 64, out_off.ctypes.data, out_len.ctypes.data,
                                      status.ctypes.data)
        assert rc == 0 and not status.any()
        host = d_dst.cpu().numpy()
        for i, (name, _, size, sha) in enumerate(corpus[:20]):
            o = int(out_off[i])
            assert int(out_len[i]) == size and hashlib.sha256(host[o : o + size].tobytes()).hexdigest() == sha, name


def test_kernels_were_launched(ctx):
    assert ctx.launch_count() > 0
    t = ctx.last_timing()
    assert t["total"] > 0


@pytest.mark.parametrize("mode", ["jump", "pair", "pair1", "team"])
def test_long_frame_paths_on_every_frame(corpus, tmp_path, mode):
    """The stage-4 paths for LONG frames (normally frames with >= 65 536 sequences) forced onto every frame: the
    block-parallel kernels of execute_long.cuh ("jump"), k_execute_pair2 ("pair"), k_execute_pair ("pair1": SZB_PAIR2=0) and
    k_execute_team ("team": SZB_PAIR2=2).  Corpus, crafted frames, a text batch,
    a streamed multi-block frame and the mixed corpus: same bytes; bit-flipped payloads: same statuses and bytes as the
    one-warp-per-frame path.  Run in a subprocess so that a hang cannot take the test session with it."""
    import os
    import subprocess
    import sys
    import textwrap

    code = textwrap.dedent(
        """
        import hashlib, os, sys
        import numpy as np
        sys.path.insert(0, "tests")
        import crafted_frames
        from tools import corpus as cg
        from oracle import pyszo
        from sparkzstd_b200.decompression import Context
        ctx = Context(0)
        gold = cg.golden_frames()
        outs = ctx.decode_batch([d for _, d, _, _ in gold])
        assert all(hashlib.sha256(o).hexdigest() == sha for o, (_, _, _, sha) in zip(outs, gold))
        cs = crafted_frames.cases()
        names = sorted(cs)
        outs = ctx.decode_batch([cs[n][0] for n in names])
        assert all(o == cs[n][1] for o, n in zip(outs, names))
        # a block that regenerates more than 128 KiB: beyond the scratch bound of the block-parallel path, which must
        # leave that frame (and only that one) to k_execute_pair
        f, e = crafted_frames.oversize_block_case()
        assert ctx.decode_batch([f, cs[names[0]][0], f]) == [e, cs[names[0]][1], e]
        t = cg.config2_text_frames(700)
        outs = ctx.decode_batch([t.frame(i) for i in range(t.nframes)])
        assert all(cg.hash_bytes(np.frombuffer(o, dtype="uint8")) == int(t.raw_hash[i]) for i, o in enumerate(outs))
        for c in (cg.config3_single_frame(12 << 20, window_log=20), cg.config5_mixed(24 << 20)):
            frames = [c.frame(i) for i in range(c.nframes)]
            outs = ctx.decode_batch(frames)
            assert all(o == pyszo.decode_frame(f) for f, o in zip(frames, outs)), c.name
        # bit flips: statuses and bytes must not depend on the path
        rng = np.random.default_rng(5)
        frames = []
        for name, data, size, _ in gold[:80]:
            if len(data) < 40:
                continue
            buf = bytearray(data)
            for _ in range(2):
                p = int(rng.integers(12, len(buf) - 4))
                buf[p] ^= 1 << int(rng.integers(0, 8))
            frames.append(bytes(buf))
        src = np.frombuffer(b"".join(frames) + bytes(64), dtype=np.uint8)
        lens = np.array([len(f) for f in frames], dtype=np.uint64)
        offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64)
        res = []
        for thr in ("1", "4000000000"):
            os.environ["SZB_LONG_SEQS"] = thr
            dst = np.zeros(64 << 20, dtype=np.uint8)
            out_off, out_len, status = ctx.decode_batch_into(src, offs, lens, dst)
            res.append((np.array(out_off), np.array(out_len), np.array(status), dst))
        a, b = res
        assert (a[2] == b[2]).all(), (a[2], b[2])
        assert (a[1] == b[1]).all()
        for i in range(len(frames)):
            if a[2][i] == 0:
                o, l = int(a[0][i]), int(a[1][i])
                o2 = int(b[0][i])
                assert (a[3][o:o + l] == b[3][o2:o2 + l]).all(), i
        print("long ok", ctx.launch_count(), int((a[2] != 0).sum()), "of", len(frames), "fail")
        """
    )
    env = dict(os.environ, SZB_LONG_SEQS="1", SZB_LONG_MODE="jump" if mode == "jump" else "pair",
               SZB_PAIR2={"pair1": "0", "team": "2"}.get(mode, "1"))
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=int(os.environ.get("SZB_TEST_TIMEOUT", 300)),
                         cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert res.returncode == 0 and "long ok" in res.stdout, res.stdout + res.stderr


def test_sequence_kernel_walks_several_groups_per_cta():
    """k_decode_sequences with a capped grid (SZB_SEQ_CTAS_PER_SM=1: 148 CTAs, each walking the groups blockIdx.x, + gridDim.x, ...
    through one mbarrier whose phase flips per group): same bytes as the oracle / the generator's hashes.  In a subprocess: the
    switch is read once per process."""
    import os
    import subprocess
    import sys
    import textwrap

    code = textwrap.dedent(
        """
        import hashlib, sys
        import numpy as np
        from tools import corpus as cg
        from oracle import pyszo
        from sparkzstd_b200.decompression import Context
        ctx = Context(0)
        gold = cg.golden_frames()
        outs = ctx.decode_batch([d for _, d, _, _ in gold])
        assert all(hashlib.sha256(o).hexdigest() == sha for o, (_, _, _, sha) in zip(outs, gold))
        t = cg.config2_text_frames(7000)   # 334 groups of 21 blocks on 148 CTAs: two to three groups per CTA
        for _ in range(2):
            outs = ctx.decode_batch([t.frame(i) for i in range(t.nframes)])
            assert all(cg.hash_bytes(np.frombuffer(o, dtype="uint8")) == int(t.raw_hash[i]) for i, o in enumerate(outs))
        c = cg.config5_mixed(24 << 20)
        frames = [c.frame(i) for i in range(c.nframes)]
        outs = ctx.decode_batch(frames)
        assert all(o == pyszo.decode_frame(f) for f, o in zip(frames, outs))
        print("capped ok", ctx.launch_count())
        """
    )
    env = dict(os.environ, SZB_SEQ_CTAS_PER_SM="1")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=int(os.environ.get("SZB_TEST_TIMEOUT", 300)),
                         cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert res.returncode == 0 and "capped ok" in res.stdout, res.stdout + res.stderr

"""Pins the ORACLE (oracle/szo.c) against the reference's golden corpus
(decodecorpus_files/, Readme.md:4,75 = BASELINE.json configs[0])."""
import hashlib

import pytest

from oracle import pyszo


def test_manifest_fingerprint(manifest):
    # sha256 over sorted name||sha256(file) of all 200 reference files, recorded in SURVEY.md section 8c
    assert manifest["fingerprint_sha256_sorted_name_sha"] == "e55ad86227510adcab1b82a7ec99de350494c85b98345c85175e1b6ab3dbebdd"
    assert len(manifest["files"]) == 100
    assert manifest["total_compressed"] == 5063476
    assert manifest["total_original"] == 11537163


def test_oracle_decodes_all_100_files_byte_exact(corpus):
    bad = []
    for name, data, size, sha in corpus:
        out, tr = pyszo.decode_frame(data, want_trace=True)
        if len(out) != size or hashlib.sha256(out).hexdigest() != sha:
            bad.append(name)
        # the reference never reads the 4-byte content checksum
        assert tr.bytes_consumed == len(data) - 4, name
    assert not bad


def test_block_trace_is_consistent(corpus):
    """Stage-level outputs add up: per block regenerated size = literals + sum(match lengths)."""
    for name, data, size, _ in corpus[:40]:
        out, tr = pyszo.decode_frame(data, want_trace=True)
        pos = 0
        for b in tr.blocks:
            assert b.out_off == pos
            if b.type == 2:
                assert b.out_len == b.lit_regen + sum(s[1] for s in b.sequences), name
                assert sum(s[0] for s in b.sequences) <= b.lit_regen or b.lit_type == 1
            else:
                assert b.out_len == b.block_size
            pos += b.out_len
        assert pos == size


def test_tiny_known_answer_frames():
    # SURVEY.md section 8c, parsed by hand from the corpus files
    assert pyszo.decode_frame(bytes.fromhex("28b52ffd042a01000099e9d851")) == b""
    assert pyszo.decode_frame(bytes.fromhex("28b52ffd044d1900000c0c0cd46aefda")) == b"\x0c\x0c\x0c"
    assert pyszo.decode_frame(bytes.fromhex("28b52ffd84280100000008000" "0a1010000af6d0ee6")) == b"\xa1"
    # last block is RLE with size 0: still consumes its payload byte
    assert pyszo.decode_frame(bytes.fromhex("28b52ffd04530a000001000000000000030000" "2a30e7211b")) == b"\x01"


def test_wrong_magic():
    with pytest.raises(pyszo.OracleError) as e:
        pyszo.decode_frame(b"\x28\xb5\x2f\xfe\x00\x00\x00\x00")
    assert e.value.code == -1


def test_truncated_inputs_error_cleanly(corpus):
    name, data, size, _ = corpus[0]
    for cut in (3, 5, 9, 100, len(data) // 2, len(data) - 5):
        with pytest.raises(pyszo.OracleError):
            pyszo.decode_frame(data[:cut])

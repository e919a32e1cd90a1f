"""Dictionaries.  The reference has no dictionary support (Readme.md:59-61; frame.go:38-47 parses the
Dictionary_ID and ignores it), so there is no golden vector of the reference's for this row.  The oracle's extension restates
RFC 8878 section 5 and is pinned here against a KNOWN ANSWER: frames libzstd 1.5.5 compressed with a dictionary must decode
to the original bytes (and libzstd's own dictionary decoder agrees)."""
import numpy as np
import pytest

from oracle import pyszo
from tools import corpus as cg


@pytest.fixture(scope="module")
def dict_cases():
    msgs = cg.dictionary_messages(400)
    formatted = cg.train_dictionary(msgs[:300], 8 << 10)
    assert formatted[:4] == bytes.fromhex("37a430ec")
    raw = b"".join(msgs[:40])  # raw content: no magic, any bytes
    assert raw[:4] != formatted[:4]
    text = cg.config2_text_frames(2, 20000)
    bigger = [bytes(cg.fill(cg.KIND_TEXT, 4242 + i, 30000 + 977 * i, 0)) for i in range(3)]
    cases = []
    for name, d in (("formatted", formatted), ("raw", raw)):
        for i, m in enumerate(msgs[300:340] + bigger + [b"", b"x"]):
            cases.append((f"{name}-{i}", d, m, cg.compress_with_dict(m, d)))
    del text
    return cases


def test_oracle_dictionary_extension_gives_back_the_originals(dict_cases):
    used_dict_tables = 0
    for name, d, original, frame in dict_cases:
        assert cg.zstd_decompress_with_dict(frame, d, len(original) + 16) == original, name  # the inputs are what they claim
        out, tr = pyszo.decode_frame(frame, want_trace=True, dictionary=d)
        assert out == original, name
        first = next((b for b in tr.blocks if b.type == 2), None)
        if first is not None and (first.lit_type == 3 or 3 in first.modes):
            used_dict_tables += 1  # Treeless literals / Repeat modes in a frame's FIRST compressed block: the dictionary's tables
    assert used_dict_tables >= 10
    # without the dictionary these frames do not decode to the original (matches reach in front of the frame, tables are missing)
    wrong = 0
    for name, d, original, frame in dict_cases[:30]:
        try:
            wrong += pyszo.decode_frame(frame) != original
        except pyszo.OracleError:
            wrong += 1
    assert wrong >= 25

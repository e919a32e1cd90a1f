"""Device code of the CUDA path compiled for the host and driven by the product's header walker, checked against the
golden corpus and the oracle: the serial device functions of stages 1-3 (bits.cuh, fse.cuh, huffman.cuh, sequences.cuh:
the code one lane runs) and the kernels of stage 4 (execute.cuh, execute_long.cuh) on an emulated CTA (warpsim.h).
tests/host_sim/ is test infrastructure only."""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np
import pytest

from oracle import pyszo
from tools import corpus as cg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIM = os.path.join(ROOT, "tests", "host_sim")


def _build_sim(lib_name, defines=()):
    lib = os.path.join(SIM, lib_name)
    srcs = [os.path.join(SIM, "hostsim.cpp"), os.path.join(ROOT, "sparkzstd_b200", "csrc", "walker.cpp")]
    deps = srcs + [os.path.join(SIM, "warpsim.h")] + [os.path.join(ROOT, "sparkzstd_b200", "csrc", f) for f in ("bits.cuh", "fse.cuh", "huffman.cuh", "sequences.cuh", "batch.cuh", "execute.cuh", "execute_long.cuh", "place.cuh", "exec2.cuh")]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", *[f"-D{d}" for d in defines], "-o", lib, *srcs], check=True)
    L = C.CDLL(lib)
    L.hostsim_decode_frame.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    return L


@pytest.fixture(scope="module")
def sim():
    return _build_sim("libhostsim.so")


@pytest.fixture(scope="module")
def sim_refill():
    """The same with the experimental k_long_jump of execute_long.cuh (SZB_JUMP_REFILL: lanes refill finished walks)."""
    return _build_sim("libhostsim_refill.so", ("SZB_JUMP_REFILL=1", "SZB_JUMP_CHAINS=4"))


@pytest.fixture(scope="module")
def sim_team4():
    """The same with four consumer warps per frame in k_execute_team (the default build has two)."""
    return _build_sim("libhostsim_team4.so", ("SZB_X2_TEAM=4",))


@pytest.fixture(scope="module")
def sim_team1():
    return _build_sim("libhostsim_team1.so", ("SZB_X2_TEAM=1",))


def _decode(L, data: bytes, cap: int):
    out = np.empty(cap + 16, dtype=np.uint8)
    n = C.c_size_t()
    rc = L.hostsim_decode_frame(data, len(data), out.ctypes.data, cap + 16, C.byref(n))
    return rc, out[: n.value].tobytes()


def test_serial_device_code_decodes_the_golden_corpus(sim, corpus):
    for name, data, size, sha in corpus:
        rc, out = _decode(sim, data, size)
        assert rc == 0 and len(out) == size and hashlib.sha256(out).hexdigest() == sha, name


def test_serial_device_code_matches_oracle_on_synthetic_shapes(sim):
    for c in (cg.config2_text_frames(6), cg.config4_literal_heavy(1, 1 << 19), cg.config3_single_frame(3 << 20, 20), cg.config5_mixed(4 << 20, with_golden=False)):
        for i in range(min(c.nframes, 12)):
            f = c.frame(i)
            want = pyszo.decode_frame(f)
            rc, out = _decode(sim, f, len(want))
            assert rc == 0 and out == want, (c.name, i)


def test_serial_device_code_error_paths(sim, corpus):
    name, data, size, _ = corpus[0]
    rc, _ = _decode(sim, data[: len(data) // 3], size)
    assert rc == -32
    rc, _ = _decode(sim, b"\0\0\0\0" + data[4:], size)
    assert rc == -1


# ---- the block-parallel stage 4 of long frames (execute_long.cuh) on an emulated warp (warpsim.h) ----
K_EXECUTE, K_EXECUTE_PAIR, K_LONG = 0, 1, 2


def _stage4(L, data: bytes, cap: int, path: int, order: int, two: int, checksum: int = 0):
    L.hostsim_stage4.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_int, C.c_int, C.c_int, C.c_int]
    out = np.empty(2 * cap + 256, dtype=np.uint8)
    n = C.c_size_t()
    rc = L.hostsim_stage4(data, len(data), out.ctypes.data, 2 * cap + 16, C.byref(n), path, order, two, checksum)
    return rc, out[: n.value].tobytes()


def _decode_long(L, data: bytes, cap: int, order: int, two: int):
    return _stage4(L, data, cap, K_LONG, order, two)


def test_long_frame_kernels_decode_golden_frames(sim, corpus):
    done = 0
    for k, (name, data, size, sha) in enumerate(corpus):
        if size > 40_000:
            continue
        order = k % 4 if k % 4 < 2 else 1000 + k
        rc, out = _decode_long(sim, data, size, order, k % 2)
        assert rc == 0 and len(out) == size and hashlib.sha256(out).hexdigest() == sha, (name, order)
        done += 1
    assert done >= 20


def test_long_frame_kernels_match_oracle_on_synthetic_shapes(sim):
    for c in (cg.config2_text_frames(2), cg.config3_single_frame(1 << 19, 20), cg.config5_mixed(1 << 20, with_golden=False)):
        for i in range(min(c.nframes, 6)):
            f = c.frame(i)
            want = pyszo.decode_frame(f)
            if len(want) > 600_000:
                continue
            rc, out = _decode_long(sim, f, len(want), 1 if i % 2 else 77 + i, 0)
            assert rc == 0 and out == want, (c.name, i)


def test_long_frame_kernels_on_crafted_frames(sim):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import crafted_frames as crafted

    for k, (name, (frame, expected)) in enumerate(sorted(crafted.cases().items())):
        rc, out = _decode_long(sim, frame, len(expected), k % 3, k % 2)
        assert rc == 0 and out == expected, name
    # a match that reaches in front of the frame: the first failing round decides (ringbuffer.go:203-214)
    seed = bytes(range(200)) * 5
    frame, _ = crafted.frame(seed, "raw", bytes(50), 10, 50, 900)
    cut = frame[:9] + crafted._block_header(0, 0, 10) + seed[:10] + frame[9 + 3 + len(seed):]
    with pytest.raises(pyszo.OracleError) as e:
        pyszo.decode_frame(cut)
    assert e.value.code == -30
    rc, _ = _decode_long(sim, cut, 4096, 0, 0)
    assert rc == -30
    # a block that regenerates more than 128 KiB is beyond the path's scratch bound: every kernel must leave the frame alone
    big, expected = crafted.oversize_block_case()
    rc, out = _decode_long(sim, big, len(expected), 0, 1)
    assert rc == 1 and out == expected  # 1: left to k_execute_pair, which decoded it


# ---- k_resolve + k_place (place.cuh): one lane per frame resolves, one warp per frame places ----
K_PLACE, K_PLACE_FALLBACK = 3, 4


def _stage4_at(L, data: bytes, cap: int, path: int, two: int, checksum: int = 0, mis: int = 0):
    """`mis`: the output starts that many bytes into a 128-byte aligned allocation (lines are lines of memory)."""
    L.hostsim_stage4_at.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.POINTER(C.c_int)]
    raw = np.empty(2 * cap + 512, dtype=np.uint8)
    base = (-raw.ctypes.data) % 128 + mis
    out = raw[base:]
    n = C.c_size_t()
    err = C.c_int(1)  # stays 1 when the frame fails before stage 4
    rc = L.hostsim_stage4_at(data, len(data), out.ctypes.data, 2 * cap + 16, C.byref(n), path, 0, two, checksum, C.byref(err))
    _stage4_at.exec_status = err.value
    return rc, out[: n.value].tobytes()


def test_place_kernels_decode_golden_frames(sim, corpus):
    done = 0
    for k, (name, data, size, sha) in enumerate(corpus):
        if size > 60_000:
            continue
        rc, out = _stage4_at(sim, data, size, K_PLACE, k % 2, 1, (k * 37) % 128)
        assert rc == 0 and len(out) == size and hashlib.sha256(out).hexdigest() == sha, name
        done += 1
    assert done >= 20
    for k, (name, data, size, sha) in enumerate(corpus[:30]):  # bitmaps too small: every frame is left to k_execute
        if not (0 < size <= 30_000):
            continue
        rc, out = _stage4_at(sim, data, size, K_PLACE_FALLBACK, k % 2, 0, (k * 11) % 128)
        assert rc == 0 and hashlib.sha256(out).hexdigest() == sha, name


def test_place_kernels_on_crafted_and_synthetic_frames(sim):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import crafted_frames as crafted

    for k, (name, (frame, expected)) in enumerate(sorted(crafted.cases().items())):
        for mis in (0, 1 + (k * 29) % 127):
            rc, out = _stage4_at(sim, frame, len(expected), K_PLACE, k % 2, 0, mis)
            assert rc == 0 and out == expected, (name, mis)
    big, expected = crafted.oversize_block_case()  # more output than the host's bound: k_execute's
    rc, out = _stage4_at(sim, big, len(expected), K_PLACE, 0, 0, 5)
    assert rc == 0 and out == expected
    for c in (cg.config2_text_frames(3), cg.config3_single_frame(1 << 19, 20), cg.config5_mixed(1 << 20, with_golden=False)):
        for i in range(min(c.nframes, 4)):
            f = c.frame(i)
            want = pyszo.decode_frame(f)
            if len(want) > 600_000:
                continue
            rc, out = _stage4_at(sim, f, len(want), K_PLACE, i % 2, 0, (i * 53) % 128)
            assert rc == 0 and out == want, (c.name, i)


def test_place_kernels_report_the_oracles_errors(sim, corpus):
    """k_resolve walks a frame in the reference's order, so a corrupted frame ends with the reference's error (the oracle's code),
    not just with some error: bit flips in small golden frames; whatever still decodes must decode to the oracle's bytes."""
    rng = np.random.default_rng(5)
    small = [(n, d, s) for n, d, s, _ in corpus if 200 <= s <= 12_000 and len(d) >= 60]
    same_bytes = same_code = 0
    for k in range(400):
        name, data, size = small[k % len(small)]
        buf = bytearray(data)
        for _ in range(1 + k % 3):
            p = int(rng.integers(10, len(buf) - 4))
            buf[p] ^= 1 << int(rng.integers(0, 8))
        frame = bytes(buf)
        want, code = None, 0
        try:
            want = pyszo.decode_frame(frame)
        except pyszo.OracleError as e:
            code = e.code
        cap = max(size, len(want) if want is not None else 0, 1) * 4 + 4096
        rc, out = _stage4_at(sim, frame, cap, K_PLACE, k % 2, 0, (k * 7) % 128)
        if want is not None and rc == 0:
            assert out == want, (name, k)
            same_bytes += 1
        elif _stage4_at.exec_status < 0:  # found while executing (stages 1-3 were fine)
            assert rc == _stage4_at.exec_status == code, (name, k, rc, code)
            same_code += 1
    assert same_bytes >= 60 and same_code >= 3


# ---- k_execute2 (exec2.cuh): 32-bit positions and entries, lines of memory, four 32-byte rows per line ----
K_EXECUTE2 = 5
K_EXECUTE_PAIR2 = 6
K_EXECUTE_TEAM = 7


@pytest.mark.parametrize("x2path", [K_EXECUTE2, K_EXECUTE_PAIR2, K_EXECUTE_TEAM])
def test_execute2_decodes_golden_frames(sim, corpus, x2path):
    done = 0
    for k, (name, data, size, sha) in enumerate(corpus):
        if size > 60_000:
            continue
        rc, out = _stage4_at(sim, data, size, x2path, k % 2, 1, (k * 37) % 128)
        assert rc == 0 and len(out) == size and hashlib.sha256(out).hexdigest() == sha, name
        done += 1
    assert done >= 20


@pytest.mark.parametrize("x2path", [K_EXECUTE2, K_EXECUTE_PAIR2, K_EXECUTE_TEAM])
def test_execute2_on_crafted_and_synthetic_frames(sim, x2path):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import crafted_frames as crafted

    for k, (name, (frame, expected)) in enumerate(sorted(crafted.cases().items())):
        for mis in (0, 1 + (k * 29) % 127):
            rc, out = _stage4_at(sim, frame, len(expected), x2path, k % 2, 0, mis)
            assert rc == 0 and out == expected, (name, mis)
    big, expected = crafted.oversize_block_case()
    rc, out = _stage4_at(sim, big, len(expected), x2path, 0, 0, 5)
    assert rc == 0 and out == expected
    for c in (cg.config2_text_frames(3), cg.config3_single_frame(1 << 19, 20), cg.config5_mixed(1 << 20, with_golden=False)):
        for i in range(min(c.nframes, 4)):
            f = c.frame(i)
            want = pyszo.decode_frame(f)
            if len(want) > 600_000:
                continue
            rc, out = _stage4_at(sim, f, len(want), x2path, i % 2, 0, (i * 53) % 128)
            assert rc == 0 and out == want, (c.name, i)


def test_execute2_agrees_with_execute_on_corrupted_frames(sim, corpus):
    """Bit flips in small golden frames: k_execute2 must end like k_execute (status and bytes), and with the oracle's bytes
    whenever the oracle still decodes the frame."""
    rng = np.random.default_rng(23)
    small = [(n, d, s) for n, d, s, _ in corpus if 200 <= s <= 12_000 and len(d) >= 60]
    ok = errs = 0
    for k in range(300):
        name, data, size = small[k % len(small)]
        buf = bytearray(data)
        for _ in range(1 + k % 3):
            p = int(rng.integers(10, len(buf) - 4))
            buf[p] ^= 1 << int(rng.integers(0, 8))
        frame = bytes(buf)
        try:
            want = pyszo.decode_frame(frame)
        except pyszo.OracleError:
            want = None
        cap = max(size, len(want) if want is not None else 0, 1) * 4 + 4096
        a = _stage4_at(sim, frame, cap, K_EXECUTE, k % 2, 0, (k * 7) % 128)
        b = _stage4_at(sim, frame, cap, K_EXECUTE2, k % 2, 0, (k * 7) % 128)
        assert a[0] == b[0], (name, k, a[0], b[0])
        # the two-warp kernels: k_execute_pair2 must end like k_execute_pair
        p1 = _stage4_at(sim, frame, cap, K_EXECUTE_PAIR, k % 2, 0, (k * 7) % 128)
        p2 = _stage4_at(sim, frame, cap, K_EXECUTE_PAIR2, k % 2, 0, (k * 7) % 128)
        assert p1[0] == p2[0] and (p1[0] != 0 or p1[1] == p2[1]), (name, k, p1[0], p2[0])
        p3 = _stage4_at(sim, frame, cap, K_EXECUTE_TEAM, k % 2, 0, (k * 7) % 128)
        assert p1[0] == p3[0] and (p1[0] != 0 or p1[1] == p3[1]), (name, k, p1[0], p3[0])
        if a[0] == 0:
            assert a[1] == b[1] == p2[1], (name, k)
            if want is not None:
                assert b[1] == want, (name, k)
                ok += 1
        else:
            errs += 1
    assert ok >= 60 and errs >= 20


@pytest.mark.parametrize("nblocks", [1, 31, 32, 33, 64, 100, 1000])
def test_long_frame_history_scan(sim, nblocks):
    sim.hostsim_compose_selftest.argtypes = [C.c_uint32, C.c_uint32]
    assert sim.hostsim_compose_selftest(nblocks, 1234 + nblocks) == 0


# ---- k_execute and k_execute_pair themselves (execute.cuh), with k_scan_blocks / k_frame_verdict / k_execute_bodies ----
@pytest.mark.parametrize("path", [K_EXECUTE, K_EXECUTE_PAIR])
def test_execute_kernels_decode_golden_frames(sim, corpus, path):
    done = 0
    for k, (name, data, size, sha) in enumerate(corpus):
        if size > 60_000:
            continue
        rc, out = _stage4(sim, data, size, path, 0, k % 2, 1)  # the golden frames carry content checksums: verified too
        assert rc == 0 and len(out) == size and hashlib.sha256(out).hexdigest() == sha, (name, path)
        done += 1
    assert done >= 20


def test_checksum_kernel_notices_a_flipped_byte(sim, corpus):
    seen = 0
    for name, data, size, _ in corpus[:40]:
        if not (0 < size <= 60_000):
            continue
        rc, _ = _stage4(sim, data, size, K_EXECUTE, 0, 0, 2)
        assert rc in (2, 3), name  # 3: mismatch reported; 2: the frame has no checksum
        seen += rc == 3
    assert seen >= 5


@pytest.mark.parametrize("path", [K_EXECUTE, K_EXECUTE_PAIR])
def test_execute_kernels_on_crafted_and_synthetic_frames(sim, path):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import crafted_frames as crafted

    for k, (name, (frame, expected)) in enumerate(sorted(crafted.cases().items())):
        rc, out = _stage4(sim, frame, len(expected), path, 0, k % 2)
        assert rc == 0 and out == expected, name
    big, expected = crafted.oversize_block_case()
    rc, out = _stage4(sim, big, len(expected), path, 0, 0)
    assert rc == 0 and out == expected
    for c in (cg.config2_text_frames(2), cg.config3_single_frame(1 << 19, 20)):
        for i in range(min(c.nframes, 2)):
            f = c.frame(i)
            want = pyszo.decode_frame(f)
            rc, out = _stage4(sim, f, len(want), path, 0, 0)
            assert rc == 0 and out == want, (c.name, i)


def test_stage4_paths_agree_on_corrupted_frames(sim, corpus):
    """Bit flips in small golden frames through all three stage-4 paths on the emulated CTA: the paths must agree with each
    other (status and bytes), and with the oracle whenever the oracle still decodes the frame."""
    rng = np.random.default_rng(11)
    small = [(n, d, s) for n, d, s, _ in corpus if 200 <= s <= 12_000 and len(d) >= 60]
    assert len(small) >= 8
    ran = agree_ok = stage4_errors = 0
    for k in range(240):
        name, data, size = small[k % len(small)]
        buf = bytearray(data)
        for _ in range(1 + k % 3):
            p = int(rng.integers(10, len(buf) - 4))
            buf[p] ^= 1 << int(rng.integers(0, 8))
        frame = bytes(buf)
        try:
            want = pyszo.decode_frame(frame)
        except pyszo.OracleError:
            want = None
        cap = max(size, len(want) if want is not None else 0, 1) * 4 + 4096
        res = [_stage4(sim, frame, cap, path, k % 2, 0) for path in (K_EXECUTE, K_EXECUTE_PAIR, K_LONG)]
        rcs = [r for r, _ in res]
        assert rcs[0] == rcs[1] == rcs[2], (name, k, rcs)
        if rcs[0] == 0:
            assert res[0][1] == res[1][1] == res[2][1], (name, k)
            if want is not None:
                assert res[0][1] == want, (name, k)
                agree_ok += 1
        elif rcs[0] in (-28, -30, -33) and want is None:
            stage4_errors += 1  # found by stage 4 itself: literals ran dry, match before the frame, reference panic
        ran += 1
    assert ran == 240 and agree_ok >= 40 and stage4_errors >= 2


def test_long_frame_kernels_refill_variant(sim_refill, corpus):
    """The build switch SZB_JUMP_REFILL (a k_long_jump whose lanes hand a finished walk's slot to their next byte; not
    measured yet, off by default) must decode the same bytes."""
    done = 0
    for k, (name, data, size, sha) in enumerate(corpus):
        if size > 40_000:
            continue
        rc, out = _stage4(sim_refill, data, size, K_LONG, [0, 1, 1000 + k][k % 3], k % 2, 1)
        assert rc == 0 and len(out) == size and hashlib.sha256(out).hexdigest() == sha, name
        done += 1
    assert done >= 20
    f = cg.config3_single_frame(1 << 19, 20).frame(0)
    want = pyszo.decode_frame(f)
    rc, out = _stage4(sim_refill, f, len(want), K_LONG, 1, 0)
    assert rc == 0 and out == want


@pytest.mark.parametrize("slice_seqs", [32, 96, 1024])
def test_long_frame_kernels_with_sliced_blocks(sim, corpus, slice_seqs, monkeypatch):
    """SZB_LONG_SLICE: k_long_hist / k_long_emit run one warp per slice of a block (k_long_blockscan gives every slice the
    history and the positions it starts with).  Same bytes, same statuses."""
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import crafted_frames as crafted

    monkeypatch.setenv("SZB_LONG_SLICE", str(slice_seqs))
    done = 0
    for k, (name, data, size, sha) in enumerate(corpus):
        if size > (30_000 if slice_seqs < 1024 else 120_000):
            continue
        rc, out = _stage4(sim, data, size, K_LONG, [0, 1, 500 + k][k % 3], k % 2, 1)
        assert rc == 0 and len(out) == size and hashlib.sha256(out).hexdigest() == sha, (name, slice_seqs)
        done += 1
    assert done >= 20
    for k, (name, (frame, expected)) in enumerate(sorted(crafted.cases().items())):
        rc, out = _stage4(sim, frame, len(expected), K_LONG, k % 3, 0)
        assert rc == 0 and out == expected, name
    f = cg.config3_single_frame(1 << 19, 20).frame(0)  # several blocks of ~12 000 sequences, matches across blocks
    want = pyszo.decode_frame(f)
    rc, out = _stage4(sim, f, len(want), K_LONG, 1, 0)
    assert rc == 0 and out == want
    if slice_seqs != 32:
        return
    # errors: the first failing round decides, whichever slice finds it
    rng = np.random.default_rng(11)
    small = [(n, d, s) for n, d, s, _ in corpus if 200 <= s <= 12_000 and len(d) >= 60]
    errors = 0
    for k in range(240):
        name, data, size = small[k % len(small)]
        buf = bytearray(data)
        for _ in range(1 + k % 3):
            p = int(rng.integers(10, len(buf) - 4))
            buf[p] ^= 1 << int(rng.integers(0, 8))
        frame = bytes(buf)
        cap = size * 4 + 4096
        a = _stage4(sim, frame, cap, K_EXECUTE, 0, 0)
        b = _stage4(sim, frame, cap, K_LONG, k % 2, 0)
        assert a == b, (name, k, a[0], b[0])
        errors += a[0] in (-28, -30, -33)
    assert errors >= 1


@pytest.mark.parametrize("team", [1, 4])
def test_execute_team_with_one_and_four_consumer_warps(sim_team1, sim_team4, corpus, team):
    L = sim_team4 if team == 4 else sim_team1
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import crafted_frames as crafted

    done = 0
    for k, (name, data, size, sha) in enumerate(corpus):
        if size > 40_000:
            continue
        rc, out = _stage4_at(L, data, size, K_EXECUTE_TEAM, k % 2, 1, (k * 37) % 128)
        assert rc == 0 and len(out) == size and hashlib.sha256(out).hexdigest() == sha, name
        done += 1
    assert done >= 20
    for k, (name, (frame, expected)) in enumerate(sorted(crafted.cases().items())):
        rc, out = _stage4_at(L, frame, len(expected), K_EXECUTE_TEAM, k % 2, 0, 1 + (k * 29) % 127)
        assert rc == 0 and out == expected, name
    c = cg.config2_text_frames(2)
    for i in range(c.nframes):
        f = c.frame(i)
        want = pyszo.decode_frame(f)
        rc, out = _stage4_at(L, f, len(want), K_EXECUTE_TEAM, 0, 0, 77)
        assert rc == 0 and out == want

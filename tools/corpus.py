"""Seeded synthetic corpora for BASELINE.json configs 2-5 (bench.py and tests use the same recipes).

Generation + compression happen in tools/corpusgen.c (gcc, pthreads) with the system libzstd 1.5.5
loaded through dlopen -- only as the producer of .zst inputs, never on the decode path.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
_SRC = os.path.join(HERE, "corpusgen.c")
_LIB = os.path.join(HERE, "_build", "libcorpusgen.so")

BASE_SEED = 20261017
KIND_TEXT, KIND_SKEWED, KIND_RANDOM, KIND_CONSTANT, KIND_LONGRANGE, KIND_MIXED = range(6)

_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_LIB), exist_ok=True)
        subprocess.run(["gcc", "-O2", "-std=gnu11", "-fPIC", "-shared", "-pthread", "-o", _LIB, _SRC, "-ldl"], check=True)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.cg_zstd_available.restype = C.c_int
        L.cg_compress_bound.restype = C.c_size_t
        L.cg_compress_bound.argtypes = [C.c_size_t]
        L.cg_fill.argtypes = [C.c_int, C.c_uint64, C.c_void_p, C.c_size_t, C.c_size_t]
        L.cg_hash.restype = C.c_uint64
        L.cg_hash.argtypes = [C.c_void_p, C.c_size_t]
        L.cg_hash_frames.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_int]
        L.cg_compress.restype = C.c_size_t
        L.cg_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_int]
        L.cg_compress_stream.restype = C.c_size_t
        L.cg_compress_stream.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int]
        L.cg_zstd_decompress.restype = C.c_size_t
        L.cg_zstd_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.cg_compress_dict.restype = C.c_size_t
        L.cg_compress_dict.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
        L.cg_zstd_decompress_dict.restype = C.c_size_t
        L.cg_zstd_decompress_dict.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.cg_train_dictionary.restype = C.c_size_t
        L.cg_train_dictionary.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_uint]
        L.cg_zstd_decompress_batch_mt.restype = C.c_uint32
        L.cg_zstd_decompress_batch_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int]
        L.cg_generate_frames.restype = C.c_uint64
        L.cg_generate_frames.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def zstd_available() -> int:
    return lib().cg_zstd_available()


def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


@dataclass
class Corpus:
    """Compressed frames in one host arena."""

    name: str
    src: np.ndarray          # uint8 arena holding every compressed frame
    frame_off: np.ndarray    # uint64
    frame_len: np.ndarray    # uint64
    raw_size: np.ndarray     # uint64 decompressed size per frame
    raw_hash: np.ndarray     # uint64 cg_hash of each original
    meta: dict

    @property
    def nframes(self) -> int:
        return len(self.frame_off)

    @property
    def compressed_bytes(self) -> int:
        return int(self.frame_len.sum())

    @property
    def decompressed_bytes(self) -> int:
        return int(self.raw_size.sum())

    def frame(self, i: int) -> bytes:
        o, l = int(self.frame_off[i]), int(self.frame_len[i])
        return self.src[o : o + l].tobytes()

    def subset(self, idx) -> "Corpus":
        idx = np.asarray(idx)
        return Corpus(self.name, self.src, self.frame_off[idx], self.frame_len[idx], self.raw_size[idx], self.raw_hash[idx], dict(self.meta))


def fill(kind: int, seed: int, n: int, lr_window: int = 0) -> np.ndarray:
    out = np.empty(max(n, 1), dtype=np.uint8)
    lib().cg_fill(kind, seed, out.ctypes.data, n, lr_window)
    return out[:n]


def hash_bytes(a: np.ndarray) -> int:
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return int(lib().cg_hash(a.ctypes.data, a.nbytes))


def hash_frames(base: np.ndarray, off: np.ndarray, length: np.ndarray, nthreads: Optional[int] = None) -> np.ndarray:
    off = np.ascontiguousarray(off, dtype=np.uint64)
    length = np.ascontiguousarray(length, dtype=np.uint64)
    out = np.zeros(len(off), dtype=np.uint64)
    lib().cg_hash_frames(base.ctypes.data, off.ctypes.data, length.ctypes.data, len(off), out.ctypes.data, nthreads or host_threads())
    return out


def zstd_decode_batch_mt(c: "Corpus", nframes: int, nthreads: int) -> int:
    """libzstd (dlopen) over the first nframes frames, outputs discarded; returns how many frames failed."""
    off = np.ascontiguousarray(c.frame_off[:nframes], dtype=np.uint64)
    ln = np.ascontiguousarray(c.frame_len[:nframes], dtype=np.uint64)
    raw = np.ascontiguousarray(c.raw_size[:nframes], dtype=np.uint64)
    return int(lib().cg_zstd_decompress_batch_mt(c.src.ctypes.data, off.ctypes.data, ln.ctypes.data, raw.ctypes.data, nframes, nthreads))


def generate(name: str, kinds, seeds, sizes, level: int = 3, checksum: int = 0, nthreads: Optional[int] = None,
             arena: Optional[np.ndarray] = None) -> Corpus:
    kinds = np.ascontiguousarray(kinds, dtype=np.int32)
    seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
    sizes = np.ascontiguousarray(sizes, dtype=np.uint64)
    n = len(kinds)
    L = lib()
    if not L.cg_zstd_available():
        raise RuntimeError("libzstd.so.1 not loadable: cannot generate the synthetic corpus")
    cap = int(sizes.sum() + (sizes // 16384).sum() + 256 * n + (1 << 16))
    if arena is None or arena.nbytes < cap:
        arena = np.empty(cap, dtype=np.uint8)
    off = np.zeros(n, dtype=np.uint64)
    ln = np.zeros(n, dtype=np.uint64)
    hs = np.zeros(n, dtype=np.uint64)
    used = L.cg_generate_frames(kinds.ctypes.data, seeds.ctypes.data, sizes.ctypes.data, n, level, checksum,
                                nthreads or host_threads(), arena.ctypes.data, arena.nbytes, off.ctypes.data,
                                ln.ctypes.data, hs.ctypes.data)
    if used == 0 and n:
        raise RuntimeError("corpus generation failed")
    return Corpus(name, arena[: int(used) + 16], off, ln, sizes.copy(), hs, {"level": level, "checksum": checksum})


# ---- BASELINE.json configs ---------------------------------------------------------------------
def config2_text_frames(nframes: int = 65536, frame_size: int = 65536, base_seed: int = BASE_SEED, **kw) -> Corpus:
    """configs[1]: independent 64 KiB synthetic-text frames, zstd level 3, one frame each."""
    c = generate("text64k", np.zeros(nframes, np.int32), base_seed + np.arange(nframes, dtype=np.uint64),
                 np.full(nframes, frame_size, np.uint64), **kw)
    c.meta.update(workload=f"{nframes} x {frame_size} B synthetic text frames, zstd L3", base_seed=base_seed)
    return c


def config4_literal_heavy(nframes: int = 4096, frame_size: int = 1 << 20, base_seed: int = BASE_SEED + 10_000_000, **kw) -> Corpus:
    """configs[3]: skewed i.i.d. bytes -> 128 KiB blocks of 4-stream Huffman literals, ~0 sequences."""
    c = generate("literal_heavy", np.full(nframes, KIND_SKEWED, np.int32), base_seed + np.arange(nframes, dtype=np.uint64),
                 np.full(nframes, frame_size, np.uint64), **kw)
    c.meta.update(workload=f"{nframes} x {frame_size} B skewed-literal frames, zstd L3", base_seed=base_seed)
    return c


def config3_single_frame(size: int = 4 << 30, window_log: int = 23, seed: int = BASE_SEED + 20_000_000, level: int = 3) -> Corpus:
    """configs[2]: ONE streamed frame (window descriptor, no content size) with long-range matches."""
    L = lib()
    raw = fill(KIND_LONGRANGE, seed, size, 1 << window_log)
    cap = int(L.cg_compress_bound(size))
    dst = np.empty(cap, dtype=np.uint8)
    n = L.cg_compress_stream(raw.ctypes.data, size, dst.ctypes.data, cap, level, window_log, 0)
    if n == 0:
        raise RuntimeError("streaming compression failed")
    h = np.array([hash_bytes(raw)], dtype=np.uint64)
    return Corpus("single_frame", dst[: n + 16], np.array([0], np.uint64), np.array([n], np.uint64), np.array([size], np.uint64), h,
                  {"workload": f"one {size} B frame, windowLog {window_log}, zstd L{level} streamed", "seed": seed})


def golden_frames():
    """The 100 decodecorpus frames (tests/golden): the only source of Repeat / RLE FSE modes and RLE literals."""
    import json

    g = os.path.join(ROOT, "tests", "golden")
    with open(os.path.join(g, "manifest.json")) as f:
        man = json.load(f)
    out = []
    for e in man["files"]:
        with open(os.path.join(g, "decodecorpus", e["name"]), "rb") as f:
            out.append((e["name"], f.read(), e["original_size"], e["original_sha256"]))
    return out


def config5_plan(total_bytes: int = 16 << 30, seed: int = BASE_SEED + 30_000_000):
    """The frame list of configs[4] without generating a byte: (kinds, seeds, sizes) of the synthetic frames.  Every rank of a
    frame-sharded run computes the same plan and generates only the frames of its shard."""
    rng = np.random.default_rng(seed)
    kinds, sizes = [], []
    acc = 0
    menu = [
        (KIND_TEXT, 65536, 40),        # custom FSE tables, 1 block
        (KIND_RANDOM, 300_000, 6),     # Raw blocks
        (KIND_CONSTANT, 300_000, 6),   # RLE blocks
        (KIND_TEXT, 200, 8),           # tiny: predefined tables
        (KIND_TEXT, 2 << 20, 25),      # multi-block: Treeless literals
        (KIND_MIXED, 690_000, 10),     # Raw + Compressed in one frame
        (KIND_SKEWED, 512 << 10, 5),   # literal heavy
    ]
    w = np.array([m[2] for m in menu], dtype=np.float64)
    w /= w.sum()
    while acc < total_bytes:
        k = int(rng.choice(len(menu), p=w))
        kinds.append(menu[k][0])
        sizes.append(menu[k][1])
        acc += menu[k][1]
    n = len(kinds)
    return np.array(kinds, np.int32), seed + np.arange(n, dtype=np.uint64), np.array(sizes, np.uint64)


def config5_mixed(total_bytes: int = 16 << 30, seed: int = BASE_SEED + 30_000_000, with_golden: bool = True, shard=None, **kw) -> Corpus:
    """configs[4]: seeded mix of frame kinds (Raw / RLE / Compressed blocks, predefined + custom tables,
    Treeless literals in multi-block frames) plus the replicated decodecorpus frames.
    shard = (rank, world): only the frames szb_shard_frames gives this rank (by decompressed size) are generated; the golden
    replicas are dealt out round robin.  meta["shard"] records the split."""
    kinds, seeds, sizes = config5_plan(total_bytes, seed)
    mine = None
    if shard is not None:
        from sparkzstd_b200.sharding import shard_frames

        rank, world = shard
        parts = shard_frames(sizes, world)
        mine = parts[rank]
        loads = [int(sizes[p].sum()) for p in parts]
        kinds, seeds, sizes = kinds[mine], seeds[mine], sizes[mine]
    c = generate("mixed", kinds, seeds, sizes, **kw)
    if shard is not None:
        c.meta["shard"] = {"rank": int(shard[0]), "world": int(shard[1]), "frames": int(len(mine)), "shard_bytes": loads}
    if with_golden:
        gold = golden_frames()
        reps = max(1, min(64, (total_bytes // 100) // max(1, sum(g[2] for g in gold))))
        my_reps = reps if shard is None else len(range(shard[0], reps, shard[1]))
        extra = b"".join(g[1] for g in gold)
        base = len(c.src)
        arena = np.concatenate([c.src, np.frombuffer(extra * my_reps, dtype=np.uint8), np.zeros(16, np.uint8)])
        offs, lens, raws, hashes = [], [], [], []
        p = base
        one = []
        for name, data, osz, _ in gold:
            one.append((len(data), osz))
        # hashes of the golden originals are not stored (only sha256 in the manifest): use 0 = "check via oracle"
        for _ in range(my_reps):
            for ln_, osz in one:
                offs.append(p)
                lens.append(ln_)
                raws.append(osz)
                hashes.append(0)
                p += ln_
        c = Corpus("mixed", arena, np.concatenate([c.frame_off, np.array(offs, np.uint64)]),
                   np.concatenate([c.frame_len, np.array(lens, np.uint64)]),
                   np.concatenate([c.raw_size, np.array(raws, np.uint64)]),
                   np.concatenate([c.raw_hash, np.array(hashes, np.uint64)]), dict(c.meta))
        c.meta["golden_reps"] = int(my_reps)
    c.meta.update(workload=f"mixed corpus ~{total_bytes} B, {c.nframes} frames", seed=seed)
    return c


# ---- dictionaries ---------------------------------------------------------------
def train_dictionary(samples, size: int = 16 << 10) -> bytes:
    """A formatted zstd dictionary (ZDICT_trainFromBuffer) from a list of byte strings."""
    L = lib()
    blob = np.frombuffer(b"".join(samples), dtype=np.uint8)
    sizes = np.array([len(x) for x in samples], dtype=np.uint64)  # size_t
    out = np.empty(size, dtype=np.uint8)
    n = L.cg_train_dictionary(out.ctypes.data, size, blob.ctypes.data, sizes.ctypes.data, len(samples))
    if n == 0:
        raise RuntimeError("ZDICT_trainFromBuffer failed (too few / too uniform samples?)")
    return out[: int(n)].tobytes()


def compress_with_dict(data: bytes, dictionary: bytes, level: int = 3) -> bytes:
    L = lib()
    src = np.frombuffer(data, dtype=np.uint8) if data else np.zeros(1, np.uint8)
    d = np.frombuffer(dictionary, dtype=np.uint8)
    cap = int(L.cg_compress_bound(len(data))) + 64
    dst = np.empty(cap, dtype=np.uint8)
    n = L.cg_compress_dict(src.ctypes.data, len(data), d.ctypes.data, len(dictionary), dst.ctypes.data, cap, level)
    if n == 0:
        raise RuntimeError("ZSTD_compress_usingDict failed")
    return dst[: int(n)].tobytes()


def zstd_decompress_with_dict(frame: bytes, dictionary: bytes, cap: int) -> bytes:
    L = lib()
    src = np.frombuffer(frame, dtype=np.uint8)
    d = np.frombuffer(dictionary, dtype=np.uint8)
    dst = np.empty(max(cap, 1), dtype=np.uint8)
    n = L.cg_zstd_decompress_dict(src.ctypes.data, len(frame), d.ctypes.data, len(dictionary), dst.ctypes.data, cap)
    if n == (1 << 64) - 1:
        raise RuntimeError("ZSTD_decompress_usingDict failed")
    return dst[: int(n)].tobytes()


def dictionary_messages(n: int = 200, seed: int = 77):
    """Small JSON-like records with a shared vocabulary: the classic dictionary workload (many small frames whose matches
    point into the dictionary).  Deterministic."""
    rng = np.random.default_rng(seed)
    keys = ["user_id", "session", "timestamp", "event_type", "payload", "region", "device", "latency_ms", "status", "trace"]
    vals = ["click", "view", "purchase", "eu-west-1", "us-east-2", "ap-south-1", "android", "ios", "desktop", "ok", "retry", "timeout"]
    out = []
    for i in range(n):
        parts = []
        for k in rng.permutation(len(keys))[: int(rng.integers(4, len(keys)))]:
            v = vals[int(rng.integers(len(vals)))] if rng.random() < 0.6 else str(int(rng.integers(1 << 30)))
            parts.append('"%s": "%s"' % (keys[int(k)], v))
        out.append(("{" + ", ".join(parts) + ', "seq": %d}' % i).encode() * int(rng.integers(1, 4)))
    return out


def algorithmic_bytes(c: Corpus, match_bytes: int) -> int:
    """C + D + M."""
    return c.compressed_bytes + c.decompressed_bytes + int(match_bytes)

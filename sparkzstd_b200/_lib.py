"""ctypes loader for libszb200.so (the C ABI declared in include/szb200.h).

There is no fallback: if the library is missing, or no CUDA device is usable when a context
is created, the caller gets an exception -- never a CPU decode.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))


class FrameDesc(C.Structure):
    _fields_ = [
        ("src_off", C.c_uint64),
        ("src_len", C.c_uint64),
        ("window_size", C.c_uint64),
        ("content_size", C.c_uint64),
        ("dictionary_id", C.c_uint64),
        ("first_block", C.c_uint32),
        ("nblocks", C.c_uint32),
        ("checksum", C.c_uint32),
        ("status", C.c_int32),
        ("descriptor", C.c_uint8),
        ("single_segment", C.c_uint8),
        ("has_checksum", C.c_uint8),
        ("has_content_size", C.c_uint8),
        ("checksum_valid", C.c_uint32),
    ]


class BlockDesc(C.Structure):
    _fields_ = [
        ("src_off", C.c_uint64),
        ("lit_buf_off", C.c_uint64),
        ("seq_buf_off", C.c_uint64),
        ("block_size", C.c_uint32),
        ("frame", C.c_uint32),
        ("lit_regen", C.c_uint32),
        ("lit_comp", C.c_uint32),
        ("nseq", C.c_uint32),
        ("seq_off", C.c_uint32),
        ("huf_origin", C.c_uint32),
        ("ll_origin", C.c_uint32),
        ("of_origin", C.c_uint32),
        ("ml_origin", C.c_uint32),
        ("type", C.c_uint8),
        ("last", C.c_uint8),
        ("lit_type", C.c_uint8),
        ("lit_streams", C.c_uint8),
        ("lit_hdr_bytes", C.c_uint8),
        ("seq_hdr_bytes", C.c_uint8),
        ("seq_modes", C.c_uint8),
        ("flags", C.c_uint8),
        ("hdr_status", C.c_int32),
    ]


# every symbol include/szb200.h declares: (name, restype, argtypes)
_P = C.c_void_p
SYMBOLS = [
    ("szb_strerror", C.c_char_p, [C.c_int]),
    ("szb_walk_create", C.c_int, [_P, C.c_size_t, _P, _P, C.c_uint32, C.POINTER(_P)]),
    ("szb_walk_create_dict", C.c_int, [_P, C.c_size_t, _P, _P, C.c_uint32, _P, C.c_uint32, C.POINTER(_P)]),
    ("szb_walk_destroy", None, [_P]),
    ("szb_walk_nframes", C.c_uint32, [_P]),
    ("szb_walk_nblocks", C.c_uint32, [_P]),
    ("szb_walk_frames", C.POINTER(FrameDesc), [_P]),
    ("szb_walk_blocks", C.POINTER(BlockDesc), [_P]),
    ("szb_walk_literal_bytes", C.c_uint64, [_P]),
    ("szb_walk_sequences", C.c_uint64, [_P]),
    ("szb_walk_known_output_size", C.c_uint64, [_P]),
    ("szb_ctx_create", C.c_int, [C.c_int, _P, C.POINTER(_P)]),
    ("szb_ctx_destroy", None, [_P]),
    ("szb_abi_layout", C.c_uint32, [_P, C.c_uint32]),
    ("szb_shard_frames", C.c_int, [_P, C.c_uint32, C.c_uint32, _P, _P]),
    ("szb_ctx_stream", _P, [_P]),
    ("szb_ctx_last_error", C.c_char_p, [_P]),
    ("szb_decode_batch", C.c_int, [_P, _P, C.c_size_t, _P, _P, C.c_uint32, _P, C.c_size_t, _P, _P, _P, C.c_uint32]),
    ("szb_decode_stream", C.c_int, [_P, _P, C.c_size_t, _P, C.c_size_t, _P, _P, _P, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.c_uint32]),
    ("szb_ctx_last_frame", C.c_int, [_P, _P]),
    ("szb_decompress_reader", C.c_int, [_P, _P, _P, _P, _P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_uint32]),
    ("szb_dict_create", C.c_int, [_P, _P, C.c_size_t, C.POINTER(_P)]),
    ("szb_dict_destroy", None, [_P]),
    ("szb_dict_id", C.c_uint32, [_P]),
    ("szb_decode_batch_dict", C.c_int, [_P, _P, _P, C.c_size_t, _P, _P, C.c_uint32, _P, C.c_size_t, _P, _P, _P, C.c_uint32]),
    ("szb_decode_blocks", C.c_int, [_P, _P, C.c_size_t, _P, C.c_uint32, _P, C.c_uint32, _P, C.c_size_t, _P, _P, _P]),
    ("szb_batch_create", C.c_int, [_P, _P, C.c_size_t, _P, _P, C.c_uint32, C.POINTER(_P)]),
    ("szb_batch_create_from_tables", C.c_int, [_P, C.c_size_t, _P, C.c_uint32, _P, C.c_uint32, C.POINTER(_P)]),
    ("szb_batch_destroy", None, [_P]),
    ("szb_batch_nframes", C.c_uint32, [_P]),
    ("szb_batch_nblocks", C.c_uint32, [_P]),
    ("szb_batch_decode_entropy", C.c_int, [_P, _P]),
    ("szb_batch_sizes", C.c_int, [_P, C.POINTER(C.c_uint64), _P, _P]),
    ("szb_batch_execute", C.c_int, [_P, _P, _P, C.c_size_t]),
    ("szb_batch_run", C.c_int, [_P, _P, _P, C.c_size_t]),
    ("szb_batch_verify_checksums", C.c_int, [_P, _P]),
    ("szb_batch_finish", C.c_int, [_P, _P]),
    ("szb_batch_read_literals", C.c_int, [_P, C.c_uint32, _P, C.c_size_t]),
    ("szb_batch_read_sequences", C.c_int, [_P, C.c_uint32, _P, _P, _P, C.c_size_t]),
    ("szb_batch_read_block_results", C.c_int, [_P, _P, _P, C.c_size_t]),
    ("szb_last_timing", C.c_int, [_P, C.POINTER(C.c_float), C.c_int]),
    ("szb_launch_count", C.c_uint64, [_P]),
    ("szb_decompress_frame", C.c_int, [_P, _P, C.c_size_t, C.POINTER(_P), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    ("szb_free", None, [_P]),
    ("szb_version", C.c_char_p, []),
]

_lib = None


def lib_path() -> str:
    return os.environ.get("SZB200_LIB", os.path.join(HERE, "libszb200.so"))


def load():
    """Loads libszb200.so (building it first when sources are newer and nvcc exists)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if "SZB200_LIB" not in os.environ:
        from . import build as _build

        try:
            if _build.is_stale(path):
                _build.build(out=path)
        except RuntimeError:
            if not os.path.exists(path):
                raise
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `python -m sparkzstd_b200.build` (needs nvcc). There is no CPU fallback.")
    L = C.CDLL(path)
    for name, res, args in SYMBOLS:
        fn = getattr(L, name)  # AttributeError here = the .so does not match include/szb200.h
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


class SzbError(Exception):
    def __init__(self, code: int, detail: str = ""):
        self.code = code
        msg = load().szb_strerror(code).decode()
        super().__init__(f"[{code}] {msg}" + (f" ({detail})" if detail else ""))

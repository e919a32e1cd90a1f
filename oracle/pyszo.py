"""ctypes binding of the ORACLE (oracle/szo.c) -- test infrastructure, NOT product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product package (sparkzstd_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libszo.so")


def build(force: bool = False) -> str:
    """Compile oracle/szo.c with gcc (make).  Returns the library path."""
    src = os.path.join(_HERE, "szo.c")
    hdr = os.path.join(_HERE, "szo.h")
    stale = (
        force
        or not os.path.exists(_LIB_PATH)
        or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr))
    )
    if stale:
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return _LIB_PATH


class _Sequence(C.Structure):
    _fields_ = [("match_length", C.c_int64), ("literal_length", C.c_int64), ("offset", C.c_int64)]


class _BlockTrace(C.Structure):
    _fields_ = [
        ("type", C.c_int),
        ("last", C.c_int),
        ("block_size", C.c_uint32),
        ("out_off", C.c_uint64),
        ("out_len", C.c_uint64),
        ("lit_type", C.c_int),
        ("lit_streams", C.c_int),
        ("lit_regen", C.c_uint32),
        ("lit_off", C.c_uint64),
        ("nseq", C.c_uint32),
        ("seq_off", C.c_uint64),
        ("ll_mode", C.c_int),
        ("of_mode", C.c_int),
        ("ml_mode", C.c_int),
        ("huf_max_bits", C.c_int),
        ("hist_after", C.c_int64 * 3),
    ]


class _Trace(C.Structure):
    _fields_ = [
        ("blocks", C.POINTER(_BlockTrace)),
        ("nblocks", C.c_size_t),
        ("blocks_cap", C.c_size_t),
        ("literals", C.POINTER(C.c_uint8)),
        ("nliterals", C.c_size_t),
        ("literals_cap", C.c_size_t),
        ("sequences", C.POINTER(_Sequence)),
        ("real_offsets", C.POINTER(C.c_int64)),
        ("nsequences", C.c_size_t),
        ("sequences_cap", C.c_size_t),
        ("window_size", C.c_uint64),
        ("frame_content_size", C.c_uint64),
        ("has_fcs", C.c_int),
        ("single_segment", C.c_int),
        ("bytes_consumed", C.c_size_t),
    ]


class _FseEntry(C.Structure):
    _fields_ = [
        ("baseline", C.c_uint16),
        ("additional_bits", C.c_uint8),
        ("number_of_bits", C.c_uint8),
        ("symbol", C.c_int32),
    ]


class _FseTable(C.Structure):
    _fields_ = [
        ("accuracy_log", C.c_int),
        ("nvalues", C.c_int),
        ("values", C.c_int64 * 512),
        ("table", C.POINTER(_FseEntry)),
        ("table_size", C.c_int),
        ("state", C.c_int64),
        ("is_rle", C.c_int),
        ("rle_value", C.c_int),
        ("rle_additional_bits", C.c_int),
    ]


class _RBits(C.Structure):
    _fields_ = [("data", C.c_void_p), ("len", C.c_int64), ("offset", C.c_int64)]


class _FBits(C.Structure):
    _fields_ = [
        ("data", C.c_void_p),
        ("len", C.c_size_t),
        ("pos", C.c_size_t),
        ("buffer", C.c_uint8),
        ("offset", C.c_uint),
    ]


class _HufTable(C.Structure):
    _fields_ = [
        ("max_bits", C.c_int),
        ("size", C.c_int),
        ("number_of_bits", C.POINTER(C.c_int)),
        ("symbols", C.POINTER(C.c_int)),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.szo_strerror.restype = C.c_char_p
        L.szo_strerror.argtypes = [C.c_int]
        L.szo_decode_frame.restype = C.c_int
        L.szo_decode_frame.argtypes = [
            C.c_void_p,
            C.c_size_t,
            C.POINTER(C.c_void_p),
            C.POINTER(C.c_size_t),
            C.POINTER(_Trace),
        ]
        L.szo_trace_free.argtypes = [C.POINTER(_Trace)]
        L.szo_decode_batch_mt.restype = C.c_int
        L.szo_decode_batch_mt.argtypes = [
            C.c_void_p,
            C.c_void_p,
            C.c_void_p,
            C.c_uint32,
            C.c_void_p,
            C.c_size_t,
            C.c_void_p,
            C.c_void_p,
            C.c_void_p,
            C.c_int,
        ]
        L.szo_rbits_init.argtypes = [C.POINTER(_RBits), C.c_void_p, C.c_size_t]
        L.szo_rbits_read.restype = C.c_uint64
        L.szo_rbits_read.argtypes = [C.POINTER(_RBits), C.c_int]
        L.szo_rbits_bits_still_in_stream.restype = C.c_int64
        L.szo_rbits_bits_still_in_stream.argtypes = [C.POINTER(_RBits)]
        L.szo_fbits_init.argtypes = [C.POINTER(_FBits), C.c_void_p, C.c_size_t]
        L.szo_fbits_read.restype = C.c_int
        L.szo_fbits_read.argtypes = [C.POINTER(_FBits), C.c_int, C.POINTER(C.c_uint64)]
        L.szo_fbits_unwind_bit.restype = C.c_int
        L.szo_fbits_unwind_bit.argtypes = [C.POINTER(_FBits)]
        for name in ("szo_fse_build_ll_table", "szo_fse_build_ml_table", "szo_fse_build_of_table"):
            getattr(L, name).restype = C.c_int
            getattr(L, name).argtypes = [C.POINTER(_FseTable)]
        L.szo_fse_table_free.argtypes = [C.POINTER(_FseTable)]
        L.szo_fse_read_table_description.restype = C.c_int
        L.szo_fse_read_table_description.argtypes = [C.POINTER(_FseTable), C.c_void_p, C.c_size_t, C.POINTER(C.c_int)]
        L.szo_fse_build_decoding_table.restype = C.c_int
        L.szo_fse_build_decoding_table.argtypes = [
            C.POINTER(_FseTable),
            C.c_void_p,
            C.c_int,
            C.c_void_p,
            C.c_int,
        ]
        L.szo_huf_decode_tree_desc.restype = C.c_int
        L.szo_huf_decode_tree_desc.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.szo_huf_build.restype = C.c_int
        L.szo_huf_build.argtypes = [C.c_void_p, C.c_int, C.POINTER(_HufTable)]
        L.szo_huf_table_free.argtypes = [C.POINTER(_HufTable)]
        L.szo_huf_decode_stream.restype = C.c_int
        L.szo_huf_decode_stream.argtypes = [
            C.POINTER(_HufTable),
            C.c_void_p,
            C.c_size_t,
            C.c_void_p,
            C.c_size_t,
            C.POINTER(C.c_int),
        ]
        L.szo_match_copy.restype = C.c_int
        L.szo_match_copy.argtypes = [C.c_void_p, C.POINTER(C.c_size_t), C.c_size_t, C.c_int64, C.c_int64]
        L.szo_next_offset.restype = C.c_int64
        L.szo_next_offset.argtypes = [C.POINTER(C.c_int64 * 3), C.POINTER(_Sequence)]
        L.free = C.CDLL(None).free
        L.free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


class OracleError(Exception):
    def __init__(self, code: int):
        self.code = code
        super().__init__(f"oracle error {code}: {lib().szo_strerror(code).decode()}")


@dataclass
class BlockTrace:
    type: int
    last: int
    block_size: int
    out_off: int
    out_len: int
    lit_type: int = 0
    lit_streams: int = 0
    lit_regen: int = 0
    literals: bytes = b""
    nseq: int = 0
    sequences: list = field(default_factory=list)  # (ll, ml, offset_value)
    real_offsets: list = field(default_factory=list)
    modes: tuple = (-1, -1, -1)
    huf_max_bits: int = 0
    hist_after: tuple = (1, 4, 8)


@dataclass
class FrameTrace:
    blocks: list
    window_size: int
    frame_content_size: int
    has_fcs: bool
    single_segment: bool
    bytes_consumed: int


def decode_frame(data: bytes, want_trace: bool = False, dictionary: bytes = None):
    """Decode one frame.  Returns bytes, or (bytes, FrameTrace) when want_trace.  dictionary (raw content or formatted): NOT a
    reference behaviour -- the oracle's restatement of RFC 8878 section 5 (szo_decode_frame_dict)."""
    L = lib()
    out = C.c_void_p()
    n = C.c_size_t()
    tr = _Trace() if want_trace else None
    buf = (C.c_uint8 * max(len(data), 1)).from_buffer_copy(data.ljust(1, b"\0")) if True else None
    if dictionary:
        L.szo_decode_frame_dict.restype = C.c_int
        L.szo_decode_frame_dict.argtypes = [C.c_void_p, C.c_size_t, C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_void_p]
        rc = L.szo_decode_frame_dict(buf, len(data), dictionary, len(dictionary), C.byref(out), C.byref(n), C.byref(tr) if want_trace else None)
    else:
        rc = L.szo_decode_frame(buf, len(data), C.byref(out), C.byref(n), C.byref(tr) if want_trace else None)
    if rc != 0:
        if want_trace:
            L.szo_trace_free(C.byref(tr))
        raise OracleError(rc)
    res = C.string_at(out, n.value) if n.value else b""
    L.free(out)
    if not want_trace:
        return res
    blocks = []
    for i in range(tr.nblocks):
        b = tr.blocks[i]
        bt = BlockTrace(b.type, b.last, b.block_size, b.out_off, b.out_len)
        if b.type == 2:
            bt.lit_type = b.lit_type
            bt.lit_streams = b.lit_streams
            bt.lit_regen = b.lit_regen
            bt.literals = C.string_at(C.addressof(tr.literals.contents) + b.lit_off, b.lit_regen) if b.lit_regen else b""
            bt.nseq = b.nseq
            bt.sequences = [
                (tr.sequences[b.seq_off + k].literal_length, tr.sequences[b.seq_off + k].match_length, tr.sequences[b.seq_off + k].offset)
                for k in range(b.nseq)
            ]
            bt.real_offsets = [tr.real_offsets[b.seq_off + k] for k in range(b.nseq)]
            bt.modes = (b.ll_mode, b.of_mode, b.ml_mode)
            bt.huf_max_bits = b.huf_max_bits
        bt.hist_after = tuple(b.hist_after)
        blocks.append(bt)
    ft = FrameTrace(blocks, tr.window_size, tr.frame_content_size, bool(tr.has_fcs), bool(tr.single_segment), tr.bytes_consumed)
    L.szo_trace_free(C.byref(tr))
    return res, ft


def decode_batch_mt(src, frame_off, frame_len, nthreads: int, dst=None, dst_off=None):
    """numpy-array front end of szo_decode_batch_mt (CPU baseline).  Returns (out_len, status)."""
    import numpy as np

    L = lib()
    n = len(frame_off)
    out_len = np.zeros(n, dtype=np.uint64)
    status = np.zeros(n, dtype=np.int32)
    fo = np.ascontiguousarray(frame_off, dtype=np.uint64)
    fl = np.ascontiguousarray(frame_len, dtype=np.uint64)
    dptr = dst.ctypes.data if dst is not None else None
    dcap = dst.nbytes if dst is not None else 0
    doff = np.ascontiguousarray(dst_off, dtype=np.uint64) if dst_off is not None else None
    L.szo_decode_batch_mt(
        src.ctypes.data,
        fo.ctypes.data,
        fl.ctypes.data,
        n,
        dptr,
        dcap,
        doff.ctypes.data if doff is not None else None,
        out_len.ctypes.data,
        status.ctypes.data,
        int(nthreads),
    )
    return out_len, status

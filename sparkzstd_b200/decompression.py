"""Host-side mirror of sparkzstd's `decompression` package over the C ABI.

Same names, argument meaning and error behaviour as the Go reference
(decompression/framereader.go, decompression/framedecompressor.go), so the parity tests read
like the reference's own usage (cmd/sparkzstd/main.go:22-40, :59-68).  The Go package that a
sparkzstd user would actually import lives under go/ (cgo over the same C ABI); this module is
the Python twin used by tests/ and bench.py because this image has no Go toolchain.

All decoding happens on the GPU through libszb200.so.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import io
import threading
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import BlockDesc, FrameDesc, SzbError


# --- error values, identity-comparable like the Go `Err*` variables ---------
class DecodeError(Exception):
    code = 0

    def __init__(self, msg: str = ""):
        super().__init__(msg or self.__class__.__name__)


def _mk(name: str, code: int):
    return type(name, (DecodeError,), {"code": code})


ErrWrongMagicnumber = _mk("ErrWrongMagicnumber", -1)
ErrCorruptSizes = _mk("ErrCorruptSizes", -2)
ErrOutOfBlocks = _mk("ErrOutOfBlocks", -3)
ErrIllegalBlockType = _mk("ErrIllegalBlockType", -7)
ErrIllegalBlockSize = _mk("ErrIllegalBlockSize", -8)
ErrCorruptedJumptable = _mk("ErrCorruptedJumptable", -10)
ErrNoHuffTableToCarryOver = _mk("ErrNoHuffTableToCarryOver", -14)
ErrStreamDidntDecodeToRightLength = _mk("ErrStreamDidntDecodeToRightLength", -15)
ErrWrongSumOfWeights = _mk("ErrWrongSumOfWeights", -16)
ErrCorruptedHuffTree = _mk("ErrCorruptedHuffTree", -17)
ErrBadPadding = _mk("ErrBadPadding", -18)
ErrDidntUseAllBitsToDecodeHuffman = _mk("ErrDidntUseAllBitsToDecodeHuffman", -19)
ErrNotAllBitsUsed = _mk("ErrNotAllBitsUsed", -20)
ErrNoLLTableToCarryOver = _mk("ErrNoLLTableToCarryOver", -21)
ErrNoMLTableToCarryOver = _mk("ErrNoMLTableToCarryOver", -22)
ErrNoOFTableToCarryOver = _mk("ErrNoOFTableToCarryOver", -23)
ErrDidntReadAllProbabilities = _mk("ErrDidntReadAllProbabilities", -25)
ErrDidntCopyAllLiteralBytes = _mk("ErrDidntCopyAllLiteralBytes", -28)
ErrCantRepeatBytes = _mk("ErrCantRepeatBytes", -30)
ErrUnexpectedEOF = _mk("ErrUnexpectedEOF", -32)

_BY_CODE = {c.code: c for c in list(globals().values()) if isinstance(c, type) and issubclass(c, DecodeError) and c.code}


def error_for(code: int, detail: str = "") -> Exception:
    cls = _BY_CODE.get(code)
    if cls is not None:
        return cls(_lib.load().szb_strerror(code).decode())
    return SzbError(code, detail)


# --- engine context ---------------------------------------------------------------------------
class Context:
    """One CUDA stream + scratch arenas on one GPU (szb_ctx).  One per thread."""

    def __init__(self, device: int = 0, stream: int = 0):
        self._L = _lib.load()
        h = C.c_void_p()
        rc = self._L.szb_ctx_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h))
        if rc != 0:
            raise SzbError(rc)
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._L.szb_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self) -> int:
        return int(self._L.szb_ctx_stream(self._h) or 0)

    def last_error(self) -> str:
        return (self._L.szb_ctx_last_error(self._h) or b"").decode()

    def launch_count(self) -> int:
        return int(self._L.szb_launch_count(self._h))

    def last_timing(self) -> dict:
        ms = (C.c_float * 10)()
        self._L.szb_last_timing(self._h, ms, 10)
        keys = ("total", "huffman_literals", "sequences", "scan", "execute", "h2d", "d2h", "seq_tables", "resolve")
        return dict(zip(keys, [float(x) for x in ms]))

    def _raise(self, rc: int):
        raise error_for(rc, self.last_error() if rc == -65 else "")

    # -- the batch entry point (szb_decode_batch) --
    def decode_batch_into(self, src: np.ndarray, frame_off: np.ndarray, frame_len: np.ndarray, dst: np.ndarray,
                          verify_checksum: bool = False):
        """src/dst: uint8 numpy arrays (host; pinned is better).  Returns (out_off, out_len, status).
        verify_checksum: also check each frame's content checksum on the GPU (status -68 on mismatch);
        the reference never does (it leaves the 4 bytes unread)."""
        n = len(frame_off)
        fo = np.ascontiguousarray(frame_off, dtype=np.uint64)
        fl = np.ascontiguousarray(frame_len, dtype=np.uint64)
        out_off = np.zeros(n, dtype=np.uint64)
        out_len = np.zeros(n, dtype=np.uint64)
        status = np.zeros(n, dtype=np.int32)
        rc = self._L.szb_decode_batch(
            self._h, src.ctypes.data, src.nbytes, fo.ctypes.data, fl.ctypes.data, n, dst.ctypes.data, dst.nbytes,
            out_off.ctypes.data, out_len.ctypes.data, status.ctypes.data, 4 if verify_checksum else 0,
        )
        if rc in (-65, -66, -34):
            self._raise(rc)
        return out_off, out_len, status

    def decode_stream(self, data: bytes, verify_checksum: bool = False) -> bytes:
        """szb_decode_stream: a `.zst` stream as the zstd tools write it -- frames back to back, skippable frames in between,
        content checksums -- decoded as one batch; returns the stream's content.  (The reference handles one frame per reader.)"""
        src = np.frombuffer(bytes(data) + b"\0" * 16, dtype=np.uint8)
        with Walk(src[: len(data)]) as w:
            frames = w.frames()
            known = w.known_output_size()
        n = len(frames)
        bad = [f.status for f in frames if f.status != 0]
        if bad:
            self._raise(int(bad[0]))
        if known is None:  # a streamed frame's size is in no header: the entropy stages tell
            fo = np.array([f.src_off for f in frames], dtype=np.uint64)
            fl = np.array([f.src_len for f in frames], dtype=np.uint64)
            known = self.output_size(src, fo, fl) if n else 0
        dst = np.empty(max(int(known), 1), dtype=np.uint8)
        out_off = np.zeros(max(n, 1), dtype=np.uint64)
        out_len = np.zeros(max(n, 1), dtype=np.uint64)
        status = np.zeros(max(n, 1), dtype=np.int32)
        found = C.c_uint32()
        total = C.c_uint64()
        rc = self._L.szb_decode_stream(self._h, src.ctypes.data, len(data), dst.ctypes.data, dst.nbytes, out_off.ctypes.data,
                                       out_len.ctypes.data, status.ctypes.data, max(n, 1), C.byref(found), C.byref(total),
                                       4 if verify_checksum else 0)
        if rc != 0:
            self._raise(rc)
        return dst[: total.value].tobytes()

    def decode_batch_dict(self, frames: Sequence[bytes], dictionary: bytes, capacity: Optional[int] = None) -> List[bytes]:
        """szb_decode_batch_dict: independent frames that were compressed with `dictionary` (raw content, or a formatted zstd
        dictionary).  Not a reference behaviour (the reference has no dictionary support); raises the first frame's error."""
        if not frames:
            return []
        d = C.c_void_p()
        dbuf = np.frombuffer(dictionary, dtype=np.uint8) if dictionary else np.zeros(1, np.uint8)
        rc = self._L.szb_dict_create(self._h, dbuf.ctypes.data, len(dictionary), C.byref(d))
        if rc != 0:
            self._raise(rc)
        try:
            lens = np.array([len(f) for f in frames], dtype=np.uint64)
            offs = np.zeros(len(frames), dtype=np.uint64)
            offs[1:] = np.cumsum(lens)[:-1]
            src = np.frombuffer(b"".join(frames) + b"\0" * 16, dtype=np.uint8)
            cap = capacity if capacity is not None else max(64, int(lens.sum()) * 24 + (1 << 20))
            dst = np.empty(cap, dtype=np.uint8)
            out_off = np.zeros(len(frames), dtype=np.uint64)
            out_len = np.zeros(len(frames), dtype=np.uint64)
            status = np.zeros(len(frames), dtype=np.int32)
            rc = self._L.szb_decode_batch_dict(self._h, d, src.ctypes.data, int(lens.sum()), offs.ctypes.data, lens.ctypes.data, len(frames),
                                               dst.ctypes.data, dst.nbytes, out_off.ctypes.data, out_len.ctypes.data, status.ctypes.data, 0)
            if rc in (-65, -66, -34):
                self._raise(rc)
            self.last_status = status
            bad = np.nonzero(status)[0]
            if len(bad):
                self._raise(int(status[bad[0]]))
            return [dst[int(o) : int(o) + int(n)].tobytes() for o, n in zip(out_off, out_len)]
        finally:
            self._L.szb_dict_destroy(d)

    def decode_batch(self, frames: Sequence[bytes], capacity: Optional[int] = None) -> List[bytes]:
        """Decodes independent frames; raises the first frame's error (like a loop of Decompress())."""
        if not frames:
            return []
        src = np.frombuffer(b"".join(frames), dtype=np.uint8)
        lens = np.array([len(f) for f in frames], dtype=np.uint64)
        offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64)
        if capacity is None:
            capacity = self.output_size(src, offs, lens)
        dst = np.empty(max(int(capacity), 1), dtype=np.uint8)
        out_off, out_len, status = self.decode_batch_into(src, offs, lens, dst)
        bad = np.nonzero(status)[0]
        if len(bad):
            self._raise(int(status[bad[0]]))
        return [dst[int(o) : int(o + l)].tobytes() for o, l in zip(out_off, out_len)]

    def output_size(self, src: np.ndarray, offs: np.ndarray, lens: np.ndarray) -> int:
        """Total decompressed size: from the frame headers when every frame declares it, else by running
        the entropy stages (a streamed frame's size is in no header)."""
        with Walk(src, offs, lens) as w:
            known = w.known_output_size()
        if known is not None:
            return known
        b = Batch(self, src, offs, lens)
        try:
            d_src = b.upload(src)
            b.decode_entropy(d_src)
            total, _, _ = b.sizes()
            return total
        finally:
            b.close()

    def decompress_frame(self, data: bytes) -> Tuple[bytes, int]:
        """szb_decompress_frame: one frame at data[0:]; returns (output, compressed bytes consumed)."""
        out = C.c_void_p()
        n = C.c_size_t()
        used = C.c_size_t()
        buf = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(1, dtype=np.uint8)
        rc = self._L.szb_decompress_frame(self._h, buf.ctypes.data, len(data), C.byref(out), C.byref(n), C.byref(used))
        if rc != 0:
            self._raise(rc)
        res = C.string_at(out, n.value) if n.value else b""
        self._L.szb_free(out)
        return res, used.value


    def decompress_reader(self, source, target, verify_checksum: bool = False, prefix: bytes = b"") -> Tuple[int, int, FrameDesc]:
        """szb_decompress_reader: one frame from `source` (.read(n)) to `target` (.write(bytes)) with the transfers overlapped
        -- input pieces go to the device while the next is being read, output pieces come back while `target` consumes the
        one before.  `prefix`: bytes already taken from the source (an eager header check).  Returns (compressed bytes the
        frame occupied, decoded bytes, the frame's header row)."""
        pending = bytearray(prefix)  # bytes taken from the source but not handed to the library yet
        err: List[BaseException] = []

        def _read(_user, buf, cap):
            try:
                if pending:
                    chunk = bytes(pending[:cap])
                    del pending[:cap]
                else:
                    chunk = source.read(cap)
                    if chunk and len(chunk) > cap:  # a reader that ignores n: keep the rest for the next call
                        pending.extend(chunk[cap:])
                        chunk = chunk[:cap]
                if not chunk:
                    return 0
                C.memmove(buf, bytes(chunk), len(chunk))
                return len(chunk)
            except BaseException as e:  # noqa: BLE001 -- carried across the C frame and re-raised below
                err.append(e)
                return -1

        def _write(_user, buf, n):
            try:
                target.write(C.string_at(buf, n))
                return 0
            except BaseException as e:  # noqa: BLE001
                err.append(e)
                return 1

        rcb = _READ_FN(_read)
        wcb = _WRITE_FN(_write)
        used = C.c_uint64()
        total = C.c_uint64()
        rc = self._L.szb_decompress_reader(self._h, rcb, None, wcb, None, C.byref(used), C.byref(total), 4 if verify_checksum else 0)
        if err:
            raise err[0]
        if rc != 0:
            self._raise(rc)
        fr = FrameDesc()
        self._L.szb_ctx_last_frame(self._h, C.byref(fr))
        return int(used.value), int(total.value), fr


_READ_FN = C.CFUNCTYPE(C.c_int64, C.c_void_p, C.c_void_p, C.c_size_t)
_WRITE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t)

_default_ctx = threading.local()


def default_context() -> Context:
    ctx = getattr(_default_ctx, "ctx", None)
    if ctx is None:
        ctx = Context(0)
        _default_ctx.ctx = ctx
    return ctx


# --- header walk (szb_walk_*) -----------------------------------------------------------------
class Walk:
    """Descriptor tables of a set of frames: the C++ twin of the Go header walker."""

    def __init__(self, src: np.ndarray, offs=None, lens=None):
        self._L = _lib.load()
        self._src = np.ascontiguousarray(src, dtype=np.uint8)
        h = C.c_void_p()
        if offs is None:
            rc = self._L.szb_walk_create(self._src.ctypes.data, self._src.nbytes, None, None, 0, C.byref(h))
        else:
            fo = np.ascontiguousarray(offs, dtype=np.uint64)
            fl = np.ascontiguousarray(lens, dtype=np.uint64)
            rc = self._L.szb_walk_create(self._src.ctypes.data, self._src.nbytes, fo.ctypes.data, fl.ctypes.data, len(fo), C.byref(h))
        if rc != 0:
            raise SzbError(rc)
        self._h = h

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def close(self):
        if self._h:
            self._L.szb_walk_destroy(self._h)
            self._h = None

    @property
    def nframes(self) -> int:
        return self._L.szb_walk_nframes(self._h)

    @property
    def nblocks(self) -> int:
        return self._L.szb_walk_nblocks(self._h)

    def frames(self) -> List[FrameDesc]:
        p = self._L.szb_walk_frames(self._h)
        return [FrameDesc.from_buffer_copy(p[i]) for i in range(self.nframes)]  # copies: outlive the walk

    def blocks(self) -> List[BlockDesc]:
        p = self._L.szb_walk_blocks(self._h)
        return [BlockDesc.from_buffer_copy(p[i]) for i in range(self.nblocks)]

    def frames_ptr(self):
        return self._L.szb_walk_frames(self._h)

    def blocks_ptr(self):
        return self._L.szb_walk_blocks(self._h)

    def known_output_size(self) -> Optional[int]:
        v = self._L.szb_walk_known_output_size(self._h)
        return None if v == 0xFFFFFFFFFFFFFFFF else int(v)


# --- staged batch (szb_batch_*) ---------------------------------------------------------------
class Batch:
    """Descriptor tables resident on the GPU; stages can be launched separately and timed."""

    def __init__(self, ctx: Context, src: np.ndarray, offs, lens):
        self.ctx = ctx
        self._L = ctx._L
        src = np.ascontiguousarray(src, dtype=np.uint8)
        fo = np.ascontiguousarray(offs, dtype=np.uint64)
        fl = np.ascontiguousarray(lens, dtype=np.uint64)
        h = C.c_void_p()
        rc = self._L.szb_batch_create(ctx._h, src.ctypes.data, src.nbytes, fo.ctypes.data, fl.ctypes.data, len(fo), C.byref(h))
        if rc != 0:
            ctx._raise(rc)
        self._h = h
        self.nframes = self._L.szb_batch_nframes(h)
        self.nblocks = self._L.szb_batch_nblocks(h)
        self._keep = []

    def close(self):
        if getattr(self, "_h", None):
            self._L.szb_batch_destroy(self._h)
            self._h = None
        self._keep = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, src: np.ndarray) -> int:
        """Copies src to a torch CUDA tensor (device memory plumbing) and returns its device pointer."""
        import torch

        host = np.ascontiguousarray(src)
        if not host.flags.writeable:
            host = host.copy()
        t = torch.from_numpy(host).to(f"cuda:{self.ctx.device}")
        pad = torch.zeros(16, dtype=torch.uint8, device=t.device)
        t = torch.cat([t, pad])
        torch.cuda.synchronize(t.device)
        self._keep.append(t)
        return t.data_ptr()

    def decode_entropy(self, d_src: int):
        rc = self._L.szb_batch_decode_entropy(self._h, C.c_void_p(d_src))
        if rc != 0:
            self.ctx._raise(rc)

    def sizes(self):
        total = C.c_uint64()
        off = np.zeros(self.nframes, dtype=np.uint64)
        ln = np.zeros(self.nframes, dtype=np.uint64)
        rc = self._L.szb_batch_sizes(self._h, C.byref(total), off.ctypes.data, ln.ctypes.data)
        if rc != 0:
            self.ctx._raise(rc)
        return int(total.value), off, ln

    def execute(self, d_src: int, d_dst: int, cap: int):
        rc = self._L.szb_batch_execute(self._h, C.c_void_p(d_src), C.c_void_p(d_dst), cap)
        if rc != 0:
            self.ctx._raise(rc)

    def run(self, d_src: int, d_dst: int, cap: int):
        rc = self._L.szb_batch_run(self._h, C.c_void_p(d_src), C.c_void_p(d_dst), cap)
        if rc != 0:
            self.ctx._raise(rc)

    def verify_checksums(self, d_dst: int):
        rc = self._L.szb_batch_verify_checksums(self._h, C.c_void_p(d_dst))
        if rc != 0:
            self.ctx._raise(rc)

    def finish(self) -> np.ndarray:
        st = np.zeros(max(self.nframes, 1), dtype=np.int32)
        rc = self._L.szb_batch_finish(self._h, st.ctypes.data)
        if rc in (-65, -66):
            self.ctx._raise(rc)
        return st[: self.nframes]

    def read_literals(self, block: int, n: int) -> bytes:
        buf = np.zeros(max(n, 1), dtype=np.uint8)
        rc = self._L.szb_batch_read_literals(self._h, block, buf.ctypes.data, buf.nbytes)
        if rc != 0:
            self.ctx._raise(rc)
        return buf[:n].tobytes()

    def read_sequences(self, block: int, n: int):
        ll = np.zeros(max(n, 1), dtype=np.uint32)
        ml = np.zeros(max(n, 1), dtype=np.uint32)
        of = np.zeros(max(n, 1), dtype=np.uint32)
        rc = self._L.szb_batch_read_sequences(self._h, block, ll.ctypes.data, ml.ctypes.data, of.ctypes.data, max(n, 1))
        if rc != 0:
            self.ctx._raise(rc)
        return ll[:n], ml[:n], of[:n]

    def read_block_results(self):
        size = np.zeros(max(self.nblocks, 1), dtype=np.uint64)
        st = np.zeros(max(self.nblocks, 1), dtype=np.int32)
        rc = self._L.szb_batch_read_block_results(self._h, size.ctypes.data, st.ctypes.data, max(self.nblocks, 1))
        if rc != 0:
            self.ctx._raise(rc)
        return size[: self.nblocks], st[: self.nblocks]


# --- the reference's public API ---------------------------------------------------------------
class _BlockHeader:
    def __init__(self):
        self.LastBlock = False
        self.Type = 0
        self.BlockSize = 0


class _Block:
    def __init__(self):
        self.Header = _BlockHeader()


class FrameDecompressor:
    """decompression.FrameDecompressor (framedecompressor.go:14-40): source io.Reader -> target io.Writer.

    source: any object with .read(); target: any object with .write().  The frame is decoded on
    the GPU in one go when the first block is asked for; the step API then walks the already
    decoded blocks so BlockCounter / CurrentBlock behave as in the reference.
    """

    def __init__(self, source, target, ctx: Optional[Context] = None):
        self._ctx = ctx
        self.Verbose = False
        self.Reset(source, target)

    # framedecompressor.go:42-52
    def Reset(self, newsource, newtarget):
        self._source = newsource
        self._target = newtarget
        self._data: Optional[bytes] = None
        self._pos = 0
        self._frame: Optional[FrameDesc] = None
        self._blocks: List[BlockDesc] = []
        self._out: Optional[bytes] = None
        self._block_out: Optional[np.ndarray] = None
        self._written = 0
        self.CurrentBlock = _Block()
        self.PreviousBlock = _Block()
        self.BlockCounter = 0

    def _load(self):
        if self._data is None:
            d = self._source.read()
            self._data = bytes(d) if d is not None else b""

    # framedecompressor.go:130-150
    def CheckMagicnum(self):
        self._load()
        if len(self._data) < 4:
            raise ErrUnexpectedEOF()
        if self._data[:4] != b"\x28\xb5\x2f\xfd":
            raise ErrWrongMagicnumber()
        self._pos = 4

    # framedecompressor.go:306-374
    def DecodeFrameHeader(self):
        self._load()
        src = np.frombuffer(self._data, dtype=np.uint8) if self._data else np.zeros(0, dtype=np.uint8)
        with Walk(src, [0], [len(self._data)]) as w:
            self._frame = w.frames()[0]
            self._blocks = w.blocks()
        # errors of the header itself surface here; block-level ones when the block is reached
        if self._frame.status != 0 and self._frame.nblocks == 0 and self._frame.status in (-1, -32):
            raise error_for(self._frame.status)

    @property
    def WindowSize(self) -> int:
        return int(self._frame.window_size) if self._frame else 0

    @property
    def FrameContentSize(self) -> int:
        return int(self._frame.content_size) if self._frame and self._frame.has_content_size else 0

    def _decode_all(self):
        if self._out is not None:
            return
        ctx = self._ctx or default_context()
        self._out, _ = ctx.decompress_frame(self._data)

    # framedecompressor.go:270-303
    def DecodeNextBlockHeader(self):
        if self._frame is None:
            self.DecodeFrameHeader()
        if self.BlockCounter >= len(self._blocks):
            if self._frame.status != 0:
                raise error_for(self._frame.status)
            raise ErrUnexpectedEOF()
        d = self._blocks[self.BlockCounter]
        self.PreviousBlock = self.CurrentBlock
        self.CurrentBlock = _Block()
        self.CurrentBlock.Header.LastBlock = bool(d.last)
        self.CurrentBlock.Header.Type = int(d.type)
        self.CurrentBlock.Header.BlockSize = int(d.block_size)

    # framedecompressor.go:198-244
    def DecodeNextBlock(self):
        if self.CurrentBlock.Header.LastBlock:
            raise ErrOutOfBlocks()
        self.DecodeNextBlockHeader()
        self._decode_all()

    # framedecompressor.go:153-170
    def Decompress(self):
        if self._data is None and self._out is None and hasattr(self._source, "read"):
            # nothing read yet: source -> GPU -> target with the transfers overlapped (szb_decompress_reader), instead of
            # read-all / decode / write-all.  Errors are the ones the step-by-step path raises.
            ctx = self._ctx or default_context()
            _, total, fr = ctx.decompress_reader(self._source, self._target)
            self._frame = fr
            self._written = total
            self._out = b""
            self.BlockCounter = int(fr.nblocks)
            self.CurrentBlock.Header.LastBlock = True
            return
        self.CheckMagicnum()
        self.DecodeFrameHeader()
        while not self.CurrentBlock.Header.LastBlock:
            self.DecodeNextBlock()
            self.BlockCounter += 1
        self._flush()

    def _flush(self):
        if self._out is not None and self._written < len(self._out):
            self._target.write(self._out[self._written :])
            self._written = len(self._out)


def NewFrameDecompressor(s, t, ctx: Optional[Context] = None) -> FrameDecompressor:
    """framedecompressor.go:55-61"""
    return FrameDecompressor(s, t, ctx)


class FrameReader:
    """decompression.FrameReader (framereader.go:9-109): an io.Reader over one zstd frame.

    The reference hands out bytes as its window ring evicts them, while it is still reading input.  Here the first Read starts
    szb_decompress_reader on a worker thread (input pieces travel to the device while the source is still being read; the
    output comes back in pinned pieces) and Read returns as soon as the first piece has landed, the rest of the
    device-to-host copy running behind it.  At most a few pieces are buffered (the queue is bounded)."""

    _HEADER_MAX = 18  # magic 4 + descriptor 1 + window 1 + dictionary id 4 + content size 8 (frame.go:23-127)

    def __init__(self, source=None, ctx: Optional[Context] = None):
        self.PrintStatus = False
        self._ctx = ctx
        self._fd = FrameDecompressor(None, None, ctx)
        self.Reset(source)

    def _eager_header_check(self, source) -> bytes:
        """framereader.go:22-31: the magic number and the frame header are checked when the reader is made.  Only the header's
        bytes are taken from the source; they are handed to the decoder in front of the rest."""
        head = b""
        while len(head) < self._HEADER_MAX:
            more = source.read(self._HEADER_MAX - len(head))
            if not more:
                break
            head += bytes(more)
        if len(head) < 4:
            raise ErrUnexpectedEOF()
        if head[:4] != b"\x28\xb5\x2f\xfd":
            raise ErrWrongMagicnumber()
        if len(head) < 5:
            raise ErrUnexpectedEOF()
        fhd = head[4]
        single = (fhd >> 5) & 1
        fcs = (1 if single else 0, 2, 4, 8)[fhd >> 6]
        need = 5 + (0 if single else 1) + (0, 1, 2, 4)[fhd & 3] + fcs
        if len(head) < need:
            raise ErrUnexpectedEOF()
        return head

    # framereader.go:35-49
    def Reset(self, source):
        self._stop_worker()
        self._source = source
        self._head = b""
        self._q = None
        self._worker = None
        self._cur = memoryview(b"")
        self._eof = source is None
        self._err = None
        self.readTotal = 0
        if source is not None:
            self._head = self._eager_header_check(source)

    def _stop_worker(self):
        w = getattr(self, "_worker", None)
        if w is not None and w.is_alive():
            self._abort = True
            try:
                while w.is_alive():
                    self._q.get(timeout=0.05)  # unblock a producer waiting on the full queue
            except Exception:
                pass
            w.join()
        self._abort = False

    def _start(self):
        import queue

        self._q = queue.Queue(maxsize=4)
        self._abort = False
        reader = self

        class _Sink:
            def write(self, b):
                if reader._abort:
                    raise RuntimeError("reader was reset")
                reader._q.put(bytes(b))

        def run():
            try:
                ctx = reader._ctx or default_context()
                ctx.decompress_reader(reader._source, _Sink(), prefix=reader._head)
                reader._q.put(None)
            except BaseException as e:  # noqa: BLE001 -- handed to the thread that calls Read
                reader._q.put(e)

        self._worker = threading.Thread(target=run, daemon=True)
        self._worker.start()

    # framereader.go:51-109.  Returns up to n bytes; b"" is io.EOF.
    def Read(self, n: int = -1) -> bytes:
        if self._err is not None:
            raise self._err
        if self._worker is None and not self._eof:
            self._start()
        out = []
        want = None if n is None or n < 0 else n
        while want is None or want > 0:
            if not len(self._cur):
                if self._eof:
                    break
                item = self._q.get()
                if item is None:
                    self._eof = True
                    break
                if isinstance(item, BaseException):
                    self._eof = True
                    self._err = item
                    if out:
                        break
                    raise item
                self._cur = memoryview(item)
            take = len(self._cur) if want is None else min(want, len(self._cur))
            out.append(bytes(self._cur[:take]))
            self._cur = self._cur[take:]
            if want is not None:
                want -= take
                break  # a Read returns what is there: short reads are legal (io.Reader)
        data = b"".join(out)
        if self.PrintStatus:
            print(f"Read bytes: {self.readTotal + len(data)}")
        self.readTotal += len(data)
        return data

    read = Read

    def readinto(self, b) -> int:
        chunk = self.Read(len(b))
        b[: len(chunk)] = chunk
        return len(chunk)

    def close(self):
        """Stops a decode whose output was not read to its end (the worker would otherwise wait on the full queue)."""
        self._stop_worker()
        self._eof = True

    def __del__(self):
        try:
            self._stop_worker()
        except Exception:
            pass


def NewFrameReader(source=None, ctx: Optional[Context] = None) -> FrameReader:
    """framereader.go:17-33.  NewFrameReader(None) is legal (cmd/sparkzstd/main.go:126)."""
    return FrameReader(source, ctx)

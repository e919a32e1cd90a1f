"""sparkzstd-b200: a B200-native (sm_100a) zstd decode engine behind sparkzstd's API.

    from sparkzstd_b200 import decompression
    r = decompression.NewFrameReader(open("x.zst", "rb"))   # io.Reader over one frame
    decompression.NewFrameDecompressor(src, dst).Decompress()
    decompression.default_context().decode_batch([frame0, frame1, ...])

The hot path (FSE tables, Huffman literals, FSE sequences, sequence execution) runs in
hand-written CUDA kernels in libszb200.so (C ABI: include/szb200.h).  No CPU fallback.
"""
from . import decompression  # noqa: F401
from ._lib import SzbError, load  # noqa: F401

__all__ = ["decompression", "SzbError", "load"]

"""Host-side split of independent frames across GPUs.

Frames share no state (fresh offset history, tables and window per frame:
decompression/framedecompressor.go:42-61), so multi-GPU decode is a partition of the frame
list: one process per GPU, each with its own context, input slice and output arena.  There is
no data-path collective and no NCCL traffic; torch.distributed is only used by callers to agree
on the partition and to reduce timings.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


def shard_frames(weights: Sequence[int], world_size: int) -> List[np.ndarray]:
    """szb_shard_frames (include/szb200.h): greedy longest-processing-time binning of frames by weight (decompressed size
    when the header declares it, else a multiple of the compressed size).  Deterministic: every rank computes the same
    partition from the same weights.  Returns world_size sorted index arrays."""
    from . import load

    w = np.ascontiguousarray(np.asarray(weights, dtype=np.int64).astype(np.uint64))
    if world_size <= 1:
        return [np.arange(len(w), dtype=np.int64)]
    shard = np.zeros(max(len(w), 1), dtype=np.uint32)
    rc = load().szb_shard_frames(w.ctypes.data if len(w) else None, len(w), world_size, shard.ctypes.data, None)
    if rc != 0:
        raise ValueError(f"szb_shard_frames: {rc}")
    shard = shard[: len(w)]
    return [np.nonzero(shard == r)[0].astype(np.int64) for r in range(world_size)]


def frame_weights(frames) -> np.ndarray:
    """Weight per frame from walker rows (szb_frame_desc): content size if declared, else 3x the
    compressed extent (typical zstd ratio) so streamed frames are not underweighted."""
    out = np.empty(len(frames), dtype=np.int64)
    for i, f in enumerate(frames):
        out[i] = int(f.content_size) if f.has_content_size else 3 * int(f.src_len)
    return out


def agree_on_partition(weights: Sequence[int], world_size: int, rank: int, group=None) -> np.ndarray:
    """Rank 0's weights are broadcast (torch.distributed, any backend) so that all ranks bin the
    same numbers; returns this rank's frame indices."""
    import torch
    import torch.distributed as dist

    w = torch.as_tensor(np.asarray(weights, dtype=np.int64))
    if dist.is_available() and dist.is_initialized() and world_size > 1:
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        w = w.to(dev)
        dist.broadcast(w, src=0, group=group)
        w = w.cpu()
    return shard_frames(w.numpy(), world_size)[rank]

"""Frame-range sharding of one stream over the GPUs of a box.

Frames depend only on the stream-constant delta frame (reference
fusion_power_video.cc:36-38, :1164), so a sequence is cut into contiguous frame
ranges, one rank / GPU each; there is no exchange step and no collective on the
data path.  What the host has to do is bookkeeping, and that is what lives here:

  * ``frame_range``   which frames a rank owns;
  * ``split_stream``  cut a rank's self-contained stream (header + frames +
                      footer, as fpvc::Encoder writes it) into its parts;
  * ``merge_shards``  concatenate the ranks' frame chunks in rank order behind
                      rank 0's header and rebuild the footer: frame offsets are
                      the prefix sum of the chunk sizes (reference
                      .cc:1179-1197), so the result is byte-identical to what a
                      single encoder would have written;
  * ``gather_stream`` the same over ``torch.distributed`` (any backend): chunk
                      sizes and bytes travel as control-plane objects to rank 0.
"""
from __future__ import annotations

import struct
from typing import List, Sequence, Tuple


def frame_range(n_frames: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous range [start, end) of rank `rank`; ranges differ by at most one frame."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad world / rank")
    base, extra = divmod(n_frames, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def split_stream(stream: bytes) -> Tuple[bytes, List[bytes]]:
    """(header incl. the delta chunk, [frame chunk, ...]) of a complete stream; validates the footer."""
    if len(stream) < 13:
        raise ValueError("stream too small")
    delta_size = struct.unpack_from("<I", stream, 8)[0]
    if stream[12] != 1 or delta_size < 5 or 8 + delta_size > len(stream):
        raise ValueError("stream does not start with a delta chunk")
    pos = 8 + delta_size
    header = stream[:pos]
    frames, offsets = [], []
    while True:
        if pos + 5 > len(stream):
            raise ValueError("stream ends without a frame index")
        size, flag = struct.unpack_from("<IB", stream, pos)
        if flag == 2:
            break
        if flag != 0 or size < 10 or pos + size > len(stream):
            raise ValueError("malformed frame chunk")
        offsets.append(pos)
        frames.append(stream[pos:pos + size])
        pos += size
    n = len(frames)
    if size != 13 + 8 * n or pos + size != len(stream):
        raise ValueError("malformed frame index")
    got = list(struct.unpack_from(f"<{n}Q", stream, pos + 5))
    if got != offsets or struct.unpack_from("<Q", stream, pos + 5 + 8 * n)[0] != n:
        raise ValueError("frame index does not match the frames")
    return header, frames


def build_footer(first_offset: int, sizes: Sequence[int]) -> bytes:
    offsets, pos = [], first_offset
    for s in sizes:
        offsets.append(pos)
        pos += s
    n = len(offsets)
    return struct.pack("<IB", 13 + 8 * n, 2) + struct.pack(f"<{n}Q", *offsets) + struct.pack("<Q", n)


def merge_shards(header: bytes, shards: Sequence[Sequence[bytes]]) -> bytes:
    """One stream from rank 0's header and every rank's frame chunks (in rank order)."""
    chunks = [c for shard in shards for c in shard]
    return header + b"".join(chunks) + build_footer(len(header), [len(c) for c in chunks])


def gather_stream(local_stream: bytes, group=None) -> bytes | None:
    """Every rank passes the self-contained stream of its own frame range (all ranks used the same
    delta frame).  Returns the merged stream on rank 0, None elsewhere."""
    import torch.distributed as dist

    header, frames = split_stream(local_stream)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((header, frames), gathered, dst=0, group=group)
    if rank != 0:
        return None
    for r in range(1, world):
        if gathered[r][0] != gathered[0][0]:
            raise ValueError(f"rank {r} encoded with a different delta frame / geometry")
    return merge_shards(gathered[0][0], [g[1] for g in gathered])

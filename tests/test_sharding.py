"""Host-side logic of the multi-GPU path (frame-range sharding, footer merge), on CPU:
two gloo ranks each encode their own frame range -- with the reference's CPU encoder standing in
for the GPU one, since this box has no GPU and the merge logic does not care who wrote the chunks --
and rank 0 must end up with exactly the stream a single encoder writes."""
import os
import socket
import sys

import numpy as np
import pytest

from fusion_power_video_b200 import sharding, synth

HERE = os.path.dirname(os.path.abspath(__file__))


def test_frame_ranges_partition_the_sequence():
    for n in (0, 1, 7, 100, 10000):
        for world in (1, 2, 3, 8):
            spans = [sharding.frame_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.frame_range(10, 2, 2)


def _golden_with_stream():
    from oracle_binding import Ref, ref_available

    if not ref_available():
        pytest.skip("oracle/_ref/libfpv_ref.so not built")
    W, H, shift, n = 64, 32, 4, 9
    frames = synth.plasma_frames(n, W, H, bits=12, seed=31).reshape(n, -1)
    ref = Ref()
    return ref, W, H, shift, frames, ref.encode_stream(frames, W, H, shift, 0, frames[0], threads=2).tobytes()


def test_split_and_merge_round_trip():
    ref, W, H, shift, frames, stream = _golden_with_stream()
    header, chunks = sharding.split_stream(stream)
    assert len(chunks) == frames.shape[0]
    assert sharding.merge_shards(header, [chunks]) == stream
    assert sharding.merge_shards(header, [chunks[:4], chunks[4:5], [], chunks[5:]]) == stream
    for bad in (stream[:-1], stream[:20], stream[:8] + b"\x00" + stream[9:]):
        with pytest.raises(ValueError):
            sharding.split_stream(bad)


def _worker(rank, world, port, out_path):
    import torch.distributed as dist

    sys.path.insert(0, HERE)
    from oracle_binding import Ref

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    W, H, shift, n = 64, 32, 4, 9
    frames = synth.plasma_frames(n, W, H, bits=12, seed=31).reshape(n, -1)
    a, b = sharding.frame_range(n, world, rank)
    # every rank uses the sequence's delta frame (frame 0), as every GPU would
    local = Ref().encode_stream(frames[a:b], W, H, shift, 0, frames[0], threads=1).tobytes()
    merged = sharding.gather_stream(local)
    if rank == 0:
        with open(out_path, "wb") as f:
            f.write(merged)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_merge_to_the_single_encoder_stream(tmp_path):
    ref, W, H, shift, frames, stream = _golden_with_stream()
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "merged.fpv")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    merged = open(out, "rb").read()
    assert merged == stream, "sharded encode + merge differs from the single-encoder stream"
    n, dec, wo, ho = ref.decode_stream(np.frombuffer(merged, np.uint8), frames.shape[0], W, H)
    assert n == frames.shape[0]
    ok, fr, pv, nf = ref.random_access_decode(np.frombuffer(merged, np.uint8), frames.shape[0] - 1, W, H)
    assert ok and nf == frames.shape[0] and np.array_equal(fr, dec[-1])

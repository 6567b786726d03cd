"""Two-GPU tests (skipped on a single-GPU box): frame-range sharding over devices gives exactly the
single-GPU result; the delta frame is the only thing that crosses GPUs (fpv_copy_delta_peer)."""
import numpy as np
import pytest

import fusion_power_video_b200 as fpv
from fusion_power_video_b200 import host, sharding, synth

pytestmark = pytest.mark.gpu


def _need_two():
    if fpv.device_count() < 2:
        pytest.skip("needs two CUDA devices")


def test_planes_sharded_over_two_devices_match_one_device():
    _need_two()
    W, H, shift, n = 1280, 160, 4, 10
    frames = synth.plasma_frames(n, W, H, bits=12, seed=5).reshape(n, -1)
    with fpv.Context(W, H, shift, False, max_batch=n, device=0) as c0:
        c0.set_delta_raw(frames[0])
        want = c0.encode(frames)
        with fpv.Context(W, H, shift, False, max_batch=n, device=1) as c1:
            c1.copy_delta_from(c0)          # peer copy of the resident delta planes
            a, b = sharding.frame_range(n, 2, 1)
            got1 = c1.encode(frames[a:b])
            got0 = c0.encode(frames[:a])
            for k in range(4):
                assert np.array_equal(np.concatenate([got0[k], got1[k]]), want[k])
            # decode on the other device
            c1.set_delta_image(None)
        with fpv.Context(W, H, shift, False, max_batch=n, device=1) as c1:
            c1.copy_delta_from(c0)
            raw = c1.decode(want[1], want[2], want[0], fpv.DEC_UNEXTRACT)
            assert np.array_equal(raw, frames)


def test_streams_sharded_over_two_devices_merge_to_the_single_stream():
    _need_two()
    W, H, shift, n = 640, 96, 0, 11
    frames = synth.plasma_frames(n, W, H, bits=16, seed=6).reshape(n, -1)
    single = host.encode_stream(frames, W, H, shift, threads=2, batch=4, device=0)
    parts = []
    for r in range(2):
        a, b = sharding.frame_range(n, 2, r)
        parts.append(host.encode_stream(frames[a:b], W, H, shift, threads=2, batch=4, delta=frames[0], device=r))
    header, chunks0 = sharding.split_stream(parts[0])
    header1, chunks1 = sharding.split_stream(parts[1])
    assert header == header1
    assert sharding.merge_shards(header, [chunks0, chunks1]) == single

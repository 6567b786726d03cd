"""Host-layer logic on the oracle-backed stand-in of the C ABI (tests/cpu_host.py): runs without a GPU.  The streams
the host layer writes here are compared with the compiled reference (oracle/_ref) where it is available."""
import numpy as np
import pytest

import cpu_host
from fusion_power_video_b200 import sharding, synth
from oracle_binding import Ref, ref_available


@pytest.fixture(scope="module")
def host():
    return cpu_host.host()


def test_encoder_stream_equals_reference_and_round_trips(host):
    W, H, shift, n = 128, 64, 4, 13
    frames = synth.plasma_frames(n, W, H, bits=12, seed=3).reshape(n, -1)
    stream = host.encode_stream(frames, W, H, shift, threads=3, batch=4)
    if ref_available():
        assert stream == Ref().encode_stream(frames, W, H, shift, 0, frames[0], 2).tobytes()
    back = host.decode_stream(stream, n, W, H, block=4096, batch=4, raw_shift=shift)
    assert np.array_equal(back, frames)


@pytest.mark.parametrize("devices", [(0, 1), (0, 1, 2, 3)])
def test_one_encoder_over_several_devices_is_byte_identical(host, devices):
    W, H, shift, n = 128, 64, 0, 41
    frames = synth.plasma_frames(n, W, H, bits=16, seed=4).reshape(n, -1)
    single = host.encode_stream(frames, W, H, shift, threads=4, batch=4)
    multi = host.encode_stream_multi(frames, W, H, shift, threads=4, batch=4, devices=devices)
    assert multi == single


def test_shards_of_a_sequence_merge_to_the_single_stream(host):
    W, H, shift, n, world = 64, 64, 0, 23, 3
    frames = synth.plasma_frames(n, W, H, bits=16, seed=5).reshape(n, -1)
    single = host.encode_stream(frames, W, H, shift, threads=2, batch=4)
    parts = []
    for r in range(world):
        a, b = sharding.frame_range(n, world, r)
        parts.append(host.encode_stream(frames[a:b], W, H, shift, threads=2, batch=4, delta=frames[0]))
    header, chunks0 = sharding.split_stream(parts[0])
    rest = [sharding.split_stream(p)[1] for p in parts[1:]]
    assert sharding.merge_shards(header, [chunks0] + rest) == single


def test_paced_ingest_counts_drops_and_latency(host):
    W, H, n = 64, 64, 16
    frames = synth.plasma_frames(n, W, H, bits=16, seed=6).reshape(n, -1)
    easy = host.ingest(frames, W, H, fps=200, seconds=0.5, threads=2, batch=4, ring_frames=32)
    assert easy["offered"] == 100 and easy["dropped"] == 0 and easy["encoded"] == 100
    assert 0 < easy["p50_ms"] <= easy["p99_ms"] <= easy["max_ms"]
    # an offered load no CPU stand-in can follow, with a one-frame ring: frames must be dropped, never blocked on
    hard = host.ingest(frames, W, H, fps=200000, seconds=0.05, threads=1, batch=4, ring_frames=1)
    assert hard["offered"] == 10000 and hard["dropped"] > 0 and hard["encoded"] + hard["dropped"] == hard["offered"]


def test_encoder_pipeline_is_race_free_under_thread_sanitizer():
    """tests/cpu_cabi/tsan_driver.cc: 40 Encoder runs over 1-4 devices and batch sizes 1-5 under ThreadSanitizer,
    every run must write the same bytes and TSAN must stay silent."""
    import os
    import subprocess

    if not os.path.exists("/usr/bin/g++"):
        pytest.skip("no system g++ with libtsan")
    r = subprocess.run(["make", "-C", os.path.join(cpu_host.ROOT, "tests", "cpu_cabi"), "tsan"], capture_output=True, text=True,
                       timeout=600)
    if "cannot find -ltsan" in r.stderr:
        pytest.skip("libtsan not installed")
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "all runs identical" in r.stdout and "ThreadSanitizer" not in r.stdout + r.stderr


def test_directory_walk_of_the_host_decoders_matches_the_restatement(host):
    """ScanCodedPlane (csrc/host/host_common.cc): what StreamingDecoder / RandomAccessDecoder use to find the chunks of a
    GPU-coded plane stream -- same offsets and length as the CPU restatement's scanner; libbrotli streams, truncated
    streams and damaged directories are not mistaken for one (the decoders then take the libbrotlidec path)."""
    import huffcoder_ref as href

    rng = np.random.default_rng(8)
    for size, kind in ((1, 0), (70000, 1), (65536, 2), (200000, 3), (131073, 1)):
        data = [np.full(size, 3, np.uint8), np.minimum(rng.geometric(0.2, size) - 1, 255).astype(np.uint8),
                rng.integers(0, 4, size).astype(np.uint8), rng.integers(0, 256, size).astype(np.uint8)][kind]
        stream = href.encode_plane(data)
        want = href.scan_plane(stream + b"\x11\x22\x33", size)
        assert want is not None
        got = host.scan_coded_plane(stream + b"\x11\x22\x33", size)
        assert got is not None and got[0] == want[0] and got[1] == want[1] == len(stream)
        assert host.scan_coded_plane(stream[:-1], size) is None                 # the final 0x03 is missing
        assert host.scan_coded_plane(stream[: len(stream) // 2], size) is None   # truncated
        assert host.scan_coded_plane(stream, size + 65536) is None               # one chunk short for this plane size
        bad = bytearray(stream)
        bad[2] ^= 0xFF                                                           # the directory's 'F'
        assert host.scan_coded_plane(bytes(bad), size) is None
    # libbrotli's own output (quality 1) carries no directory
    W, H, n = 64, 64, 2
    frames = synth.plasma_frames(n, W, H, bits=16, seed=2).reshape(n, -1)
    s = host.encode_stream(frames, W, H, 0, threads=0, batch=1)
    assert host.scan_coded_plane(s[24:], W * H) is None

"""GPU parity tests for the host layer (fpvc::Encoder / StreamingDecoder /
RandomAccessDecoder in libfusion_power_video_b200.so): the byte stream must be
the reference's, byte for byte, and both decoders must return what the
reference's decoders return."""
import glob
import hashlib
import os

import numpy as np
import pytest

from fusion_power_video_b200 import host, synth
from oracle_binding import Ref, ref_available

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = [p for p in sorted(glob.glob(os.path.join(GOLDEN, "case_*.npz"))) if "stream_sha256" in np.load(p)]


def sha(b):
    return np.frombuffer(hashlib.sha256(bytes(b)).digest(), np.uint8)


@pytest.mark.parametrize("threads,batch", [(0, 1), (3, 2), (4, 32)], ids=["sync", "t3b2", "t4b32"])
@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[5:-4] for p in CASES])
def test_stream_is_byte_identical_to_reference(path, threads, batch):
    g = np.load(path)
    W, H, shift, be = int(g["W"]), int(g["H"]), int(g["shift"]), int(g["be"])
    if W % 4 or H % 4:
        pytest.skip("the reference itself reads out of bounds for sizes not divisible by 4")
    stream = host.encode_stream(g["frames"], W, H, shift, be, threads=threads, batch=batch, delta=g["delta"])
    assert len(stream) == int(g["stream_size"])
    assert np.array_equal(sha(stream), g["stream_sha256"]), "compressed stream differs from the reference encoder's"


@pytest.mark.parametrize("block", [0, 1, 977, 65536])
@pytest.mark.parametrize("path", CASES[:8], ids=[os.path.basename(p)[5:-4] for p in CASES[:8]])
def test_streaming_decoder_matches_reference(path, block):
    g = np.load(path)
    W, H, shift, be = int(g["W"]), int(g["H"]), int(g["shift"]), int(g["be"])
    if W % 4 or H % 4:
        pytest.skip("not encodable")
    n = g["frames"].shape[0]
    stream = host.encode_stream(g["frames"], W, H, shift, be, threads=2, batch=3, delta=g["delta"])
    dec = host.decode_stream(stream, n + 2, W, H, block=block, batch=3)
    assert dec.shape[0] == n
    assert np.array_equal(dec, g["decoded"]), "decoded images differ from the reference StreamingDecoder's"
    raw = host.decode_stream(stream, n + 2, W, H, block=block, batch=2, raw_shift=shift, big_endian=be)
    assert np.array_equal(raw.view(np.uint8).reshape(n, -1), g["unextracted"]), "fused UnextractFrame differs"
    # the host-side UnextractFrame kept for API compatibility
    assert np.array_equal(host.unextract(dec[0], W, H, shift, be), g["unextracted"][0])


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[5:-4] for p in CASES])
def test_random_access_decoder_matches_reference(path):
    g = np.load(path)
    W, H, shift, be = int(g["W"]), int(g["H"]), int(g["shift"]), int(g["be"])
    if W % 4 or H % 4:
        pytest.skip("not encodable")
    n = g["frames"].shape[0]
    stream = host.encode_stream(g["frames"], W, H, shift, be, threads=2, batch=4, delta=g["delta"])
    nf, frames, preview = host.random_access(stream, n - 1, 1, W, H)
    assert nf == n
    assert np.array_equal(frames[0], g["decoded"][n - 1])
    assert np.array_equal(preview, g["last_preview_decoded"])
    nf, frames, _ = host.random_access(stream, 0, n, W, H, batch=3, want_preview=False)
    assert np.array_equal(frames, g["decoded"])


def test_malformed_streams_fail_like_the_reference():
    g = np.load(CASES[0])
    W, H, shift, be = int(g["W"]), int(g["H"]), int(g["shift"]), int(g["be"])
    stream = bytearray(host.encode_stream(g["frames"], W, H, shift, be, threads=1, batch=2, delta=g["delta"]))
    bad = bytearray(stream)
    bad[12] = 7                      # delta chunk flag
    with pytest.raises(host.HostError):
        host.decode_stream(bytes(bad), 4, W, H)
    bad = bytearray(stream)
    bad[0:4] = (0).to_bytes(4, "little")   # xsize 0
    with pytest.raises(host.HostError):
        host.decode_stream(bytes(bad), 4, W, H)
    with pytest.raises(host.HostError):
        host.random_access(bytes(stream[:-3]), 0, 1, W, H)   # truncated footer
    # a truncated stream still yields every complete frame (reference .cc:916-922)
    n = g["frames"].shape[0]
    dec = host.decode_stream(bytes(stream[: len(stream) - 40 - 8 * n]), n, W, H)
    assert 0 < dec.shape[0] <= n and np.array_equal(dec, g["decoded"][: dec.shape[0]])


@pytest.mark.parametrize("W,H,bits,shift,n", [(1280, 800, 12, 4, 9), (1024, 1024, 16, 0, 5)])
def test_full_size_against_the_compiled_reference(W, H, bits, shift, n):
    """Whole-file equality with the reference Encoder at BASELINE.json's geometries (needs oracle/_ref)."""
    if not ref_available():
        pytest.skip("oracle/_ref/libfpv_ref.so not present on this box")
    ref = Ref()
    frames = synth.plasma_frames(n, W, H, bits=bits, seed=77).reshape(n, -1)
    stream = host.encode_stream(frames, W, H, shift, False, threads=4, batch=4)
    expect = ref.encode_stream(frames, W, H, shift, 0, frames[0], threads=4)
    assert len(stream) == expect.size and np.array_equal(np.frombuffer(stream, np.uint8), expect)
    # our stream through the reference decoder, the reference's stream (== ours) through our decoder
    nd, dec_ref, wo, ho = ref.decode_stream(np.frombuffer(stream, np.uint8), n, W, H)
    assert nd == n and (wo, ho) == (W, H)
    dec = host.decode_stream(stream, n, W, H, block=65536, batch=4)
    assert np.array_equal(dec, dec_ref)
    raw = host.decode_stream(stream, n, W, H, batch=8, raw_shift=shift)
    assert np.array_equal(raw, frames), "round trip does not reproduce the input"


# ---- Encoder with the GPU entropy coder (GpuOptions::gpu_entropy / FPV_GPU_ENTROPY) ---------------------------
# The bytes differ from libbrotli's, so the checks are the north star's other two: the reference's own decoders
# (decode.cc's StreamingDecoder, RandomAccessDecoder) read the stream, and the round trip reproduces the input.

@pytest.mark.parametrize("threads,batch", [(0, 1), (2, 3), (4, 32)], ids=["sync", "t2b3", "t4b32"])
@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[5:-4] for p in CASES])
def test_gpu_entropy_stream_decodes_like_the_reference_stream(path, threads, batch):
    g = np.load(path)
    W, H, shift, be = int(g["W"]), int(g["H"]), int(g["shift"]), int(g["be"])
    if W % 4 or H % 4:
        pytest.skip("not encodable")
    n = g["frames"].shape[0]
    stream = host.encode_stream(g["frames"], W, H, shift, be, threads=threads, batch=batch, delta=g["delta"],
                                gpu_entropy=True)
    dec = host.decode_stream(stream, n + 2, W, H, block=4096, batch=3)
    assert dec.shape[0] == n and np.array_equal(dec, g["decoded"])
    nf, frames, preview = host.random_access(stream, n - 1, 1, W, H)
    assert nf == n and np.array_equal(frames[0], g["decoded"][n - 1])
    assert np.array_equal(preview, g["last_preview_decoded"])
    if ref_available():
        ref = Ref()
        nd, dec_ref, wo, ho = ref.decode_stream(np.frombuffer(stream, np.uint8), n + 2, W, H)
        assert nd == n and (wo, ho) == (W, H) and np.array_equal(dec_ref[:n], g["decoded"])
        ok, img, prev, nf = ref.random_access_decode(np.frombuffer(stream, np.uint8), n - 1, W, H)
        assert ok and nf == n and np.array_equal(img, g["decoded"][n - 1].reshape(-1))
        assert np.array_equal(prev, g["last_preview_decoded"].reshape(-1))


@pytest.mark.parametrize("W,H,bits,shift,n", [(1280, 800, 12, 4, 40), (1024, 1024, 16, 0, 9)])
def test_gpu_entropy_full_size_round_trip(W, H, bits, shift, n):
    frames = synth.plasma_frames(n, W, H, bits=bits, seed=78).reshape(n, -1)
    stream = host.encode_stream(frames, W, H, shift, False, threads=4, batch=16, gpu_entropy=True)
    brotli = host.encode_stream(frames, W, H, shift, False, threads=4, batch=16)
    assert len(stream) < 1.03 * len(brotli), "GPU entropy coder much worse than brotli quality 1"
    raw = host.decode_stream(stream, n, W, H, batch=8, raw_shift=shift)
    assert np.array_equal(raw, frames), "round trip does not reproduce the input"
    if ref_available():
        nd, dec_ref, wo, ho = Ref().decode_stream(np.frombuffer(stream, np.uint8), n, W, H)
        assert nd == n
        assert np.array_equal(dec_ref[:n] >> shift if shift else dec_ref[:n], frames)


def test_gpu_coded_stream_both_decoder_paths(monkeypatch):
    """A GPU-coded stream through StreamingDecoder twice: coded bytes decoded on the GPU from the chunk directories
    (the default for such streams), and through libbrotlidec on host threads (FPV_HOST_BROTLI_DECODE) -- the same
    images either way; a truncated stream yields the same leading frames on both paths."""
    W, H, shift, n = 1280, 160, 4, 21
    frames = synth.plasma_frames(n, W, H, bits=12, seed=5).reshape(n, -1)
    stream = host.encode_stream(frames, W, H, shift, False, threads=4, batch=8, gpu_entropy=True)
    gpu = host.decode_stream(stream, n, W, H, batch=8, raw_shift=shift)
    assert np.array_equal(gpu, frames)
    cut = stream[: len(stream) * 2 // 3]
    gpu_cut = host.decode_stream(cut, n, W, H, batch=8, raw_shift=shift)
    monkeypatch.setenv("FPV_HOST_BROTLI_DECODE", "1")
    cpu = host.decode_stream(stream, n, W, H, batch=8, raw_shift=shift)
    cpu_cut = host.decode_stream(cut, n, W, H, batch=8, raw_shift=shift)
    assert np.array_equal(cpu, frames)
    assert 0 < gpu_cut.shape[0] < n and np.array_equal(gpu_cut, cpu_cut)


@pytest.mark.parametrize("gpu_entropy", [False, True])
def test_pinned_input_is_uploaded_in_place_and_gives_the_same_stream(gpu_entropy):
    """Frames handed to CompressFrame in page-locked memory are not copied on the host (fpv_*_submit_v from the
    caller's buffers); the stream is byte for byte the one pageable input gives."""
    import ctypes as C

    import fusion_power_video_b200 as fpv

    W, H, shift, n = 640, 96, 4, 37
    frames = synth.plasma_frames(n, W, H, bits=12, seed=15).reshape(n, -1)
    L = fpv.lib()
    nbytes = frames.nbytes
    p = L.fpv_host_alloc(nbytes)
    assert p and L.fpv_host_is_pinned(C.c_void_p(p)) == 1
    assert L.fpv_host_is_pinned(C.c_void_p(frames.ctypes.data)) == 0
    try:
        pinned = np.frombuffer((C.c_uint8 * nbytes).from_address(p), np.uint16).reshape(n, -1)
        pinned[:] = frames
        a = host.encode_stream(frames, W, H, shift, False, threads=4, batch=8, gpu_entropy=gpu_entropy)
        b = host.encode_stream(pinned, W, H, shift, False, threads=4, batch=8, gpu_entropy=gpu_entropy)
        assert a == b
        assert np.array_equal(host.decode_stream(b, n, W, H, batch=8, raw_shift=shift), frames)
    finally:
        L.fpv_host_free(C.c_void_p(p))

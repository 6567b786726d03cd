"""ColumnarBatchEncoder / Batch / ColumnarBatchDecoder (csrc/host/columnar_batch.*), the mirror of the reference's
columnar_batch/ wrapper: frames -> batches of compressed planes -> images, wired like the reference's
columnar_batch_decoder_test.cc.  Expected images follow the reference's Frame semantics."""
import numpy as np
import pytest

from fusion_power_video_b200 import host, synth
from oracle_binding import Oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


def left_aligned(frames, shift, be):
    """What DecompressImage returns: (high << 8) | low of the split pixel (little-endian and plain big-endian cases)."""
    f = frames.byteswap() if be else frames
    return (f.astype(np.uint32) << shift).astype(np.uint16)


def raw_preview(q, W, H):
    high = (q >> 8).astype(np.uint32).reshape(-1, H // 4, 4, W // 4, 4)
    return ((high.sum(axis=(2, 4)) // 16) & 0xFE).astype(np.uint8).reshape(q.shape[0], -1)


@pytest.mark.parametrize("W,H,bits,shift,be,n,fpb", [
    (100, 100, 16, 0, False, 3, 2),       # the reference's own smoke test geometry (columnar_batch_decoder_test.cc:41)
    (1280, 160, 12, 4, False, 23, 10),    # partial last batch
    (256, 64, 16, 0, True, 7, 7),         # exactly one full batch, then Close sends nullptr
    (320, 48, 8, 8, False, 5, 2),         # no low plane
    (64, 32, 12, 4, False, 130, 70),      # batches larger than one GPU call (64)
])
def test_round_trip_all_image_types(W, H, bits, shift, be, n, fpb):
    frames = synth.plasma_frames(n, W, H, bits=bits, seed=9).reshape(n, -1)
    if be:
        frames = frames.byteswap()
    ts = np.arange(n, dtype=np.int64) * 1000 + 123456
    q = left_aligned(frames, shift, be)

    full, t, info = host.columnar_roundtrip(frames, ts, W, H, shift, be, fpb, host.IMAGE_FULL, unshift=False)
    assert full.shape[0] == n and np.array_equal(t, ts)
    assert np.array_equal(full, q), "FULL images differ from the left-aligned input"
    assert info["batches"] == (n + fpb - 1) // fpb
    assert info["encoder_close"] == ts[-1] and info["decoder_close"] == ts[-1]
    assert 0 < info["compressed_bytes"] < 2 * frames.size

    unshifted, _, _ = host.columnar_roundtrip(frames, ts, W, H, shift, be, fpb, host.IMAGE_FULL, unshift=True)
    # the reference only unshifts images of more than 8 significant bits (columnar_batch_decoder.cc:82)
    assert np.array_equal(unshifted, q >> shift if 16 - shift > 8 else q)

    msb, t, _ = host.columnar_roundtrip(frames, ts, W, H, shift, be, fpb, host.IMAGE_MSB8)
    assert np.array_equal(t, ts) and np.array_equal(msb, (q >> 8).astype(np.uint8))

    prev, t, _ = host.columnar_roundtrip(frames, ts, W, H, shift, be, fpb, host.IMAGE_PREVIEW)
    assert np.array_equal(t, ts) and np.array_equal(prev, raw_preview(q, W, H))


def test_reference_smoke_sequence():
    """columnar_batch_decoder_test.cc:41-56: 100x100 frames img[i] = i * k, k = 1, 2, 3, two frames per batch."""
    W = H = 100
    base = np.arange(W * H, dtype=np.uint16)
    frames = np.stack([base * k for k in (1, 2, 3)])
    ts = np.array([123456, 234567, 345678], np.int64)
    img, t, info = host.columnar_roundtrip(frames, ts, W, H, 0, False, 2, host.IMAGE_FULL)
    assert np.array_equal(img, frames) and np.array_equal(t, ts) and info["batches"] == 2


@pytest.mark.parametrize("W,H,bits,shift,be,n,fpb", [
    (100, 100, 16, 0, False, 5, 2),
    (1280, 160, 12, 4, False, 12, 5),
    (256, 64, 16, 0, True, 7, 7),
    (320, 48, 8, 8, False, 5, 2),
])
def test_batch_planes_match_oracle_predict(oracle, W, H, bits, shift, be, n, fpb):
    """The plane columns of a Batch, brotli-decoded, are exactly what the reference's Frame ctor + Predict leave in
    high() / low() / preview() with the first frame as the delta frame (columnar_batch_encoder.cc:37-46,
    columnar_batch.cc:65-90) -- checked against the oracle, not against a round trip."""
    frames = synth.plasma_frames(n, W, H, bits=bits, seed=21).reshape(n, -1)
    frames[n - 1] = frames[0]                 # a frame equal to the delta frame: planes of zeros
    if be:
        frames = frames.byteswap()
    ts = np.arange(n, dtype=np.int64) + 5
    flags, high, low, preview = host.columnar_planes(frames, ts, W, H, shift, be, fpb)
    seen = set()
    for i in range(n):
        f, h, l, p = oracle.predict(frames[i], W, H, shift, be, frames[0])
        seen.add(f)
        assert flags[i] == f, f"frame {i}: flags {flags[i]} != {f}"
        assert np.array_equal(high[i], h), f"frame {i}: high plane"
        assert np.array_equal(preview[i], p), f"frame {i}: preview"
        if not (f & 4):
            assert np.array_equal(low[i], l), f"frame {i}: low plane"
    assert len(seen) >= 1

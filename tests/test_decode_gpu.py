"""GPU parity tests for the inverse transform (DecompressImage's post-brotli
part, UnextractFrame, Frame::Uncompress's plane undo) through the C ABI."""
import glob
import os

import numpy as np
import pytest

import fusion_power_video_b200 as fpv
from fusion_power_video_b200 import synth
from oracle_binding import Oracle

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = [p for p in sorted(glob.glob(os.path.join(GOLDEN, "case_*.npz"))) if "decoded" in np.load(p).files]


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


@pytest.mark.parametrize("serial", [False, True], ids=["spec", "serial"])
@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[5:-4] for p in CASES])
def test_golden_case(path, serial, monkeypatch):
    if serial:
        monkeypatch.setenv("FPV_DECODE_SERIAL", "1")
    g = np.load(path)
    W, H, shift, be = int(g["W"]), int(g["H"]), int(g["shift"]), int(g["be"])
    low = g["low"] if shift != 8 else None
    with fpv.Context(W, H, shift, be, max_batch=8) as ctx:
        ctx.set_delta_image(g["delta_image"])
        img = ctx.decode(g["high"], low, g["flags"])
        assert np.array_equal(img, g["decoded"])
        raw = ctx.decode(g["high"], low, g["flags"], fpv.DEC_UNEXTRACT)
        assert np.array_equal(raw.view(np.uint8).reshape(raw.shape[0], -1), g["unextracted"])
        # the encoder-side delta entry point yields the same resident delta frame
        ctx.set_delta_raw(g["delta"])
        assert np.array_equal(ctx.decode(g["high"], low, g["flags"]), g["decoded"])


@pytest.mark.parametrize(
    "W,H",
    [(1, 1), (1, 9), (7, 1), (2, 2), (3, 5), (13, 7), (31, 4), (32, 3), (33, 6), (64, 5), (100, 10), (129, 3), (250, 9), (516, 4), (1000, 3)],
)
def test_arbitrary_geometry_random_planes(oracle, W, H):
    """The inverse is defined for ANY residual planes and any W, H >= 1 (no % 4 requirement)."""
    rng = np.random.default_rng(W * 1000 + H)
    n = 8
    high = rng.integers(0, 256, (n, W * H), dtype=np.uint8)
    high[1] = 0
    high[2] = rng.integers(0, 3, W * H)
    low = rng.integers(0, 256, (n, W * H), dtype=np.uint8)
    delta = rng.integers(0, 65536, W * H, dtype=np.uint16)
    flags = np.array([0, 1, 2, 3, 4, 5, 6, 7], np.uint8)
    with fpv.Context(W, H, 3, 0, max_batch=4) as ctx:
        ctx.set_delta_image(delta)
        img = ctx.decode(high, low, flags)
        raw = ctx.decode(high, low, flags, fpv.DEC_UNEXTRACT)
        for i in range(n):
            exp = oracle.inverse(high[i], None if flags[i] & 4 else low[i], delta, W, H, int(flags[i]))
            assert np.array_equal(img[i], exp), f"frame {i} flags {flags[i]}: first diff at {np.flatnonzero(img[i] != exp)[:8]}"
            assert np.array_equal(raw[i].view(np.uint8), oracle.unextract(exp, 3, 0))


@pytest.mark.parametrize("kernel", ["default", "pair", "simd"])
@pytest.mark.parametrize("kind", ["random", "smooth", "sparse", "sawtooth"])
@pytest.mark.parametrize("W,H", [(1280, 40), (1024, 64), (2048, 16), (4096, 8)])
def test_speculation_adversarial(oracle, W, H, kind, kernel, monkeypatch):
    """Residual planes chosen to make the incoming-west guess wrong as often as possible."""
    if kernel != "default":
        monkeypatch.setenv("FPV_DECODE_KERNEL", kernel)
    rng = np.random.default_rng(7)
    n = 3
    if kind == "random":
        high = rng.integers(0, 256, (n, W * H), dtype=np.uint8)
    elif kind == "smooth":
        img = synth.plasma_frames(n, W, H, bits=16, seed=4).reshape(n, -1)
        high = np.stack([oracle.cg_forward((f >> 8).astype(np.uint8), W) for f in img])
    elif kind == "sparse":
        high = (rng.random((n, W * H)) < 0.01).astype(np.uint8) * rng.integers(1, 256, (n, W * H), dtype=np.uint8)
    else:
        high = np.tile((np.arange(W * H) % 7 == 0).astype(np.uint8) * 200, (n, 1)).astype(np.uint8)
        high[:, ::2] += 1
    flags = np.full(n, 2 | 4, np.uint8)
    with fpv.Context(W, H, 0, 0, max_batch=4) as ctx:
        img = ctx.decode(high, None, flags)
        for i in range(n):
            exp = oracle.inverse(high[i], None, None, W, H, 6)
            assert np.array_equal(img[i], exp), f"frame {i}: first diff at {np.flatnonzero(img[i] != exp)[:8]}"


PAIR_GEOMS = [(64, 5), (80, 7), (128, 3), (256, 9), (320, 6), (512, 4), (768, 5), (1024, 6), (1040, 3),
              (1280, 1), (1280, 2), (1280, 9),
              # wider than 1280: the pair kernel's split mode (one frame as a left and a right half)
              (1312, 5), (1536, 4), (1920, 7), (2016, 2), (2048, 1), (2048, 6), (2560, 3)]


@pytest.mark.parametrize("kernel", ["fused", "pair", "simd", "spec"])
@pytest.mark.parametrize("W,H", PAIR_GEOMS)
def test_every_kernel_mixed_flags(oracle, W, H, kernel, monkeypatch):
    """All four row kernels on the same planes; neighbouring frames differ in every flag, so the
    pair kernel sees frame pairs that mix delta / ClampedGradient / low-plane use, and an odd tail."""
    monkeypatch.setenv("FPV_DECODE_KERNEL", kernel)
    rng = np.random.default_rng(W * 31 + H)
    flags = np.array([3, 1, 2, 0, 7, 3, 5, 2, 6, 4, 3], np.uint8)
    n = len(flags)
    high = rng.integers(0, 256, (n, W * H), dtype=np.uint8)
    img = synth.plasma_frames(n, W, H, bits=16, seed=W + H).reshape(n, -1)
    for i in range(0, n, 3):  # realistic residuals (long healing distances) next to white noise
        high[i] = oracle.cg_forward((img[i] >> 8).astype(np.uint8), W)
    low = rng.integers(0, 256, (n, W * H), dtype=np.uint8)
    delta = rng.integers(0, 65536, W * H, dtype=np.uint16)
    delta[: W * H // 2] |= 0x80FF  # carries out of both bytes of the delta add
    for shift, be in ((4, 0), (0, 1)):
        with fpv.Context(W, H, shift, be, max_batch=16) as ctx:
            ctx.set_delta_image(delta)
            out = ctx.decode(high, low, flags)
            raw = ctx.decode(high, low, flags, fpv.DEC_UNEXTRACT)
            for i in range(n):
                exp = oracle.inverse(high[i], None if flags[i] & 4 else low[i], delta, W, H, int(flags[i]))
                assert np.array_equal(out[i], exp), f"frame {i} flags {flags[i]}: first diff at {np.flatnonzero(out[i] != exp)[:8]}"
                assert np.array_equal(raw[i].view(np.uint8), oracle.unextract(exp, shift, be)), f"unextract, frame {i}"


@pytest.mark.parametrize("kernel", ["fused", "pair"])
def test_pair_kernel_without_delta_and_low(oracle, kernel, monkeypatch):
    monkeypatch.setenv("FPV_DECODE_KERNEL", kernel)
    W, H, n = 1280, 24, 5
    img = synth.plasma_frames(n, W, H, bits=8, seed=77).reshape(n, -1)
    high = np.stack([oracle.cg_forward(f.astype(np.uint8), W) for f in img])
    flags = np.full(n, 2 | 4, np.uint8)
    with fpv.Context(W, H, 8, 0, max_batch=8) as ctx:
        out = ctx.decode(high, None, flags, fpv.DEC_UNEXTRACT)
        assert np.array_equal(out.reshape(n, -1), img)


def test_big_endian_unextract(oracle):
    W, H, n = 64, 16, 2
    rng = np.random.default_rng(3)
    high = rng.integers(0, 256, (n, W * H), dtype=np.uint8)
    low = rng.integers(0, 256, (n, W * H), dtype=np.uint8)
    flags = np.array([2, 0], np.uint8)
    with fpv.Context(W, H, 4, 1, max_batch=4) as ctx:
        raw = ctx.decode(high, low, flags, fpv.DEC_UNEXTRACT)
        for i in range(n):
            exp = oracle.unextract(oracle.inverse(high[i], low[i], None, W, H, int(flags[i])), 4, 1)
            assert np.array_equal(raw[i].view(np.uint8), exp)


@pytest.mark.parametrize("shift", [9, 12, 16])
def test_big_endian_shift_above_8_is_a_decode_only_context(oracle, shift):
    """The reference's Frame constructor is undefined for big-endian data with shift > 8 (it shifts by 8 - shift,
    .cc:407-412) but UnextractFrame (.cc:850-862) is not: such a context decodes and refuses to encode."""
    W, H, n = 96, 8, 3
    rng = np.random.default_rng(shift)
    high = rng.integers(0, 256, (n, W * H), dtype=np.uint8)
    low = rng.integers(0, 256, (n, W * H), dtype=np.uint8)
    flags = np.array([2, 0, 2], np.uint8)
    with fpv.Context(W, H, shift, 1, max_batch=4) as ctx:
        raw = ctx.decode(high, low, flags, fpv.DEC_UNEXTRACT)
        for i in range(n):
            exp = oracle.unextract(oracle.inverse(high[i], low[i], None, W, H, int(flags[i])), shift, 1)
            assert np.array_equal(raw[i].view(np.uint8), exp)
        with pytest.raises(fpv.FpvError) as e:
            ctx.encode(np.zeros((1, W * H), np.uint16))
        assert e.value.code == 3


def test_use_delta_without_delta_frame_is_an_error():
    with fpv.Context(16, 16, 0, 0, max_batch=1) as ctx:
        with pytest.raises(fpv.FpvError) as e:
            ctx.decode(np.zeros((1, 256), np.uint8), np.zeros((1, 256), np.uint8), np.array([1], np.uint8))
        assert e.value.code == 4  # "delta frame not given", .cc:310


@pytest.mark.parametrize("W,H,n", [(128, 32, 4), (100, 100, 3), (36, 8, 5), (1280, 64, 3), (2052, 16, 3)])
def test_unpredict_planes_vs_oracle(oracle, W, H, n):
    """Frame::Uncompress's undo on byte planes (k_decode_spec in planes mode, in place): high planes of every
    alignment class and previews whose width W / 4 is odd."""
    frames = synth.plasma_frames(n, W, H, bits=16, seed=31).reshape(n, -1)
    yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    frames[2] = ((xx * 300 + yy * 500) & 0xFFFF).astype(np.uint16).reshape(-1)
    dh, dl = oracle.delta_planes(frames[0], 0, 0)
    with fpv.Context(W, H, 0, 0, max_batch=2) as ctx:
        ctx.set_delta_raw(frames[0])
        flags, high, low, preview = ctx.encode(frames)
        h2, l2, p2 = ctx.unpredict_planes(high, low, preview, flags)
        for i in range(n):
            eh, el, ep = oracle.unpredict_planes(high[i], low[i], preview[i], dh, dl, W, H, int(flags[i]))
            assert np.array_equal(h2[i], eh) and np.array_equal(l2[i], el) and np.array_equal(p2[i], ep)
            # and it really undoes the prediction
            assert np.array_equal((h2[i].astype(np.uint16) << 8) | l2[i], frames[i])


@pytest.mark.parametrize("W,H,n,skew", [(1024, 256, 6, 0), (1024, 64, 5, 8), (2048, 32, 3, 8), (512, 32, 7, 8), (1008, 24, 5, 0)])
def test_device_pointer_decode(oracle, W, H, n, skew):
    """Device pointers on a caller's stream.  skew: the output starts 16 bytes into its allocation, so it is not
    32-byte aligned and the fused kernel must leave its 256-bit-store path for the bulk-store one; 1008 columns:
    the last lane with columns stores half of its 32 (one 256-bit store)."""
    import torch

    frames = synth.plasma_frames(n, W, H, bits=16, seed=13).reshape(n, -1)
    dev = torch.device("cuda:0")
    with fpv.Context(W, H, 0, 0, max_batch=8) as ctx:
        ctx.set_delta_raw(frames[0])
        flags, high, low, preview = ctx.encode(frames)
        d_high = torch.from_numpy(high).to(dev)
        d_low = torch.from_numpy(low).to(dev)
        d_flags = torch.from_numpy(flags).to(dev)
        d_buf = torch.zeros(n * W * H + 64, dtype=torch.int16, device=dev)
        d_out = d_buf[skew:skew + n * W * H]
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            ctx.decode_device(d_high.data_ptr(), d_low.data_ptr(), d_flags.data_ptr(), n, d_out.data_ptr(), stream=s.cuda_stream)
        s.synchronize()
        assert np.array_equal(d_out.cpu().numpy().view(np.uint16).reshape(n, -1), frames)
        assert not d_buf[:skew].any() and not d_buf[skew + n * W * H:].any(), "wrote outside the output"


@pytest.mark.parametrize("W,H,bits,shift,n", [(1280, 800, 12, 4, 5), (1024, 1024, 16, 0, 5), (2048, 2048, 16, 0, 3)])
def test_benchmark_geometries_decode_vs_oracle(oracle, W, H, bits, shift, n):
    """BASELINE.json's frame sizes through encode -> decode: decoded images equal the oracle's DecompressImage
    (pair kernel, its L = 32 layout at 1024, its split mode at 2048) and the round trip reproduces the raw file."""
    frames = synth.plasma_frames(n, W, H, bits=bits, seed=W ^ H).reshape(n, -1)
    with fpv.Context(W, H, shift, 0, max_batch=8) as ctx:
        ctx.set_delta_raw(frames[0])
        flags, high, low, _ = ctx.encode(frames)
        out = ctx.decode(high, low, flags)
        raw = ctx.decode(high, low, flags, fpv.DEC_UNEXTRACT)
    delta = oracle.delta_image(frames[0], shift, 0)
    for i in range(n):
        exp = oracle.inverse(high[i], None if flags[i] & 4 else low[i], delta, W, H, int(flags[i]))
        assert np.array_equal(out[i], exp), f"frame {i}: first diff at {np.flatnonzero(out[i] != exp)[:8]}"
    assert np.array_equal(raw, frames), "encode -> decode does not reproduce the input"


@pytest.mark.parametrize("W,H,shift,be,bits", [(1888, 64, 8, 1, 8), (2208, 56, 0, 0, 16), (1312, 76, 8, 0, 8), (2336, 80, 8, 1, 8),
                                               (2528, 36, 3, 1, 13), (1952, 8, 0, 0, 16), (2496, 12, 8, 0, 8)])
def test_split_mode_partial_last_lane(oracle, W, H, shift, be, bits):
    """Regression (found by scripts/gpu_fuzz.py): split mode with a half width that is not a multiple of the lane
    width hands x[last_t] of a partial lane to the right half; a repair round that 'settles' after that pixel must
    still exchange it."""
    n = 6
    for seed in (1, 2, 3):
        frames = synth.plasma_frames(n, W, H, bits=bits, seed=seed * 1000 + W).reshape(n, -1)
        if be:
            frames = frames.byteswap()
        with fpv.Context(W, H, shift, be, max_batch=4) as ctx:
            ctx.set_delta_raw(frames[0])
            flags, high, low, _ = ctx.encode(frames)
            out = ctx.decode(high, low, flags)
        delta = oracle.delta_image(frames[0], shift, be)
        for i in range(n):
            exp = oracle.inverse(high[i], None if (flags[i] & 4) or low is None else low[i], delta, W, H, int(flags[i]))
            assert np.array_equal(out[i], exp), f"seed {seed} frame {i}: first diff at {np.flatnonzero(out[i] != exp)[:8]}"


def test_randomised_sweep_short():
    """A short run of the randomised parity sweep (scripts/gpu_fuzz.py: encode, decode, entropy coder vs the oracle)."""
    import importlib.util
    import sys

    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts", "gpu_fuzz.py")
    spec = importlib.util.spec_from_file_location("gpu_fuzz", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    argv = sys.argv
    sys.argv = ["gpu_fuzz.py", "12", "7"]
    try:
        assert mod.main() == 0
    finally:
        sys.argv = argv

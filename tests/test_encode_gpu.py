"""GPU parity tests for the encode transform (Frame ctor + Frame::Predict),
through the C ABI, against the golden vectors and the oracle."""
import glob
import os

import numpy as np
import pytest

import fusion_power_video_b200 as fpv
from fusion_power_video_b200 import synth
from oracle_binding import Oracle

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(glob.glob(os.path.join(GOLDEN, "case_*.npz")))


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


def check_planes(got, exp_flags, exp_high, exp_low, exp_preview, what=""):
    flags, high, low, preview = got
    assert flags.tolist() == [int(x) for x in exp_flags], f"{what} flags"
    for i in range(len(exp_flags)):
        assert np.array_equal(high[i], exp_high[i]), f"{what} high plane of frame {i}: first diff at {np.flatnonzero(high[i] != exp_high[i])[:8]}"
        if exp_low is not None:
            assert np.array_equal(low[i], exp_low[i]), f"{what} low plane of frame {i}: first diff at {np.flatnonzero(low[i] != exp_low[i])[:8]}"
        assert np.array_equal(preview[i], exp_preview[i]), f"{what} preview of frame {i}: first diff at {np.flatnonzero(preview[i] != exp_preview[i])[:8]}"


@pytest.mark.parametrize("options", [fpv.ENC_DEFAULT, fpv.ENC_GENERIC], ids=["fast", "generic"])
@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[5:-4] for p in CASES])
def test_golden_case(path, options):
    g = np.load(path)
    W, H, shift, be = int(g["W"]), int(g["H"]), int(g["shift"]), int(g["be"])
    with fpv.Context(W, H, shift, be, max_batch=8) as ctx:
        if int(g["has_delta"]):
            ctx.set_delta_raw(g["delta"])
        got = ctx.encode(g["frames"], options)
        check_planes(got, g["flags"], g["high"], g["low"] if shift != 8 else None, g["preview"])
        # a second call on the same context (adaptive flag guess now primed) must not change anything
        got = ctx.encode(g["frames"], options)
        check_planes(got, g["flags"], g["high"], g["low"] if shift != 8 else None, g["preview"], "second call")


def oracle_batch(oracle, frames, W, H, shift, be, delta):
    out = [oracle.predict(f, W, H, shift, be, delta) for f in frames]
    flags = [o[0] for o in out]
    high = [o[1] for o in out]
    low = None if shift == 8 else [o[2] for o in out]
    preview = [o[3] for o in out]
    return flags, high, low, preview


@pytest.mark.parametrize(
    "W,H,bits,shift,n",
    [(1280, 800, 12, 4, 6), (1024, 1024, 16, 0, 5), (2048, 2048, 16, 0, 2), (256, 64, 16, 0, 9), (1288, 36, 12, 4, 3), (4096, 16, 16, 0, 2)],
)
def test_benchmark_geometries_vs_oracle(oracle, W, H, bits, shift, n):
    frames = synth.plasma_frames(n, W, H, bits=bits, seed=W + H).reshape(n, -1)
    exp = oracle_batch(oracle, frames, W, H, shift, 0, frames[0])
    with fpv.Context(W, H, shift, 0, max_batch=4) as ctx:  # n > max_batch: exercises chunking
        ctx.set_delta_raw(frames[0])
        check_planes(ctx.encode(frames), *exp, what="fast")
        check_planes(ctx.encode(frames, fpv.ENC_GENERIC), *exp, what="generic")


def mixed_batch(W, H):
    """Frames whose reference decisions differ inside one batch: forces the redo passes."""
    rng = np.random.default_rng(99)
    plasma = synth.plasma_frames(3, W, H, bits=16, seed=5).reshape(3, -1)
    const = np.full(W * H, 0x4000, np.uint16) + rng.integers(0, 256, W * H).astype(np.uint16)  # delta NOT chosen
    noise = rng.integers(0, 65536, W * H, dtype=np.uint16)
    yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    ramp = ((xx * 37 + yy * 91) & 0xFFFF).astype(np.uint16).reshape(-1)  # CG chosen
    zero = np.zeros(W * H, np.uint16)
    return np.stack([plasma[0], const, noise, ramp, plasma[1], zero, ramp[::-1].copy(), const, plasma[2]])


@pytest.mark.parametrize("W,H", [(256, 96), (640, 128), (1280, 64)])
@pytest.mark.parametrize("band", [None, "8", "16"])
def test_mixed_decisions_redo_passes(oracle, W, H, band, monkeypatch):
    if band:
        monkeypatch.setenv("FPV_BAND_ROWS", band)
    frames = mixed_batch(W, H)
    delta = synth.plasma_frames(1, W, H, bits=16, seed=77).reshape(-1)
    exp = oracle_batch(oracle, frames, W, H, 0, 0, delta)
    assert len(set(exp[0])) >= 3, f"test data should cover several flag combinations, got {exp[0]}"
    with fpv.Context(W, H, 0, 0, max_batch=16) as ctx:
        ctx.set_delta_raw(delta)
        check_planes(ctx.encode(frames), *exp, what="fast")
        # primed guess: rerun in a different order
        order = [3, 2, 1, 0, 8, 7, 6, 5, 4]
        exp2 = tuple([e[i] for i in order] if e is not None else None for e in exp)
        check_planes(ctx.encode(frames[order]), *exp2, what="fast reordered")
        check_planes(ctx.encode(frames, fpv.ENC_GENERIC), *exp, what="generic")


def test_no_delta_option_matches_empty_delta(oracle):
    """Encoder::Init encodes the delta frame itself with Predict(EMPTY) (.cc:1099-1100)."""
    W, H = 320, 64
    frames = synth.plasma_frames(2, W, H, bits=16, seed=3).reshape(2, -1)
    exp = oracle_batch(oracle, frames, W, H, 0, 0, None)
    with fpv.Context(W, H, 0, 0, max_batch=4) as ctx:
        ctx.set_delta_raw(frames[0])
        check_planes(ctx.encode(frames, fpv.ENC_NO_DELTA), *exp)
        ctx.set_delta_raw(None)
        check_planes(ctx.encode(frames), *exp)


@pytest.mark.parametrize("n", [2368, 2500, 3100, 700])
def test_large_batches_frame_tasks_and_split_launch(oracle, n):
    """Batches of more frames than persistent CTAs: whole-frame tasks (the decision taken in shared memory), and a
    batch that does not fill its last wave cut in two launches (whole waves as frame tasks, the rest as band tasks) --
    flags, planes and previews of EVERY frame against the oracle; the mixed content forces redo passes in both parts."""
    W, H = 256, 16
    base = mixed_batch(W, H)
    rng = np.random.default_rng(n)
    pick = rng.integers(0, base.shape[0], n)
    frames = base[pick].copy()
    frames[:, 0] ^= (np.arange(n) & 0xFF).astype(np.uint16)        # no two frames exactly alike
    delta = synth.plasma_frames(1, W, H, bits=16, seed=77).reshape(-1)
    exp = oracle_batch(oracle, frames, W, H, 0, 0, delta)
    assert len(set(exp[0])) >= 3
    with fpv.Context(W, H, 0, 0, max_batch=n) as ctx:
        ctx.set_delta_raw(delta)
        check_planes(ctx.encode(frames), *exp, what=f"{n} frames")
        check_planes(ctx.encode(frames[::-1]), *tuple(e[::-1] if e is not None else None for e in exp), what="reversed")


@pytest.mark.parametrize("stages", ["2", "5"])
def test_ring_depths(oracle, stages, monkeypatch):
    monkeypatch.setenv("FPV_STAGES", stages)
    W, H, n = 512, 200, 7
    frames = synth.plasma_frames(n, W, H, bits=12, seed=8).reshape(n, -1)
    exp = oracle_batch(oracle, frames, W, H, 4, 0, frames[0])
    with fpv.Context(W, H, 4, 0, max_batch=8) as ctx:
        ctx.set_delta_raw(frames[0])
        check_planes(ctx.encode(frames), *exp)


def test_device_pointer_api_and_stream(oracle):
    import torch

    W, H, n = 1280, 800, 5
    frames = synth.plasma_frames(n, W, H, bits=12, seed=21).reshape(n, -1)
    exp = oracle_batch(oracle, frames, W, H, 4, 0, frames[0])
    dev = torch.device("cuda:0")
    d_frames = torch.from_numpy(frames.view(np.int16)).to(dev)
    d_high = torch.zeros((n, W * H), dtype=torch.uint8, device=dev)
    d_low = torch.zeros((n, W * H), dtype=torch.uint8, device=dev)
    d_prev = torch.zeros((n, W * H // 16), dtype=torch.uint8, device=dev)
    d_flags = torch.zeros(n, dtype=torch.uint8, device=dev)
    stream = torch.cuda.Stream()
    with fpv.Context(W, H, 4, 0, max_batch=2) as ctx:
        with torch.cuda.stream(stream):
            ctx.set_delta_raw_device(d_frames.data_ptr(), stream.cuda_stream)
            before = ctx.kernel_launches
            ctx.encode_device(d_frames.data_ptr(), n, d_flags.data_ptr(), d_high.data_ptr(), d_low.data_ptr(),
                              d_prev.data_ptr(), stream=stream.cuda_stream)
            assert ctx.kernel_launches > before
        stream.synchronize()
        got = (d_flags.cpu().numpy(), d_high.cpu().numpy(), d_low.cpu().numpy(), d_prev.cpu().numpy())
        check_planes(got, *exp)


def test_unsupported_geometry_is_an_error():
    with fpv.Context(30, 16, 0, 0, max_batch=1) as ctx:
        with pytest.raises(fpv.FpvError) as e:
            ctx.encode(np.zeros((1, 30 * 16), np.uint16))
        assert e.value.code == 3


def test_full_size_round_trip_property():
    """BASELINE config 1 at full frame size: encode -> decode reproduces the input (size-independent property)."""
    W, H, n = 1280, 800, 12
    frames = synth.plasma_frames(n, W, H, bits=12, seed=1).reshape(n, -1)
    with fpv.Context(W, H, 4, 0, max_batch=16) as ctx:
        ctx.set_delta_raw(frames[0])
        flags, high, low, preview = ctx.encode(frames)
        assert not np.any(flags & 4)
        raw = ctx.decode(high, low, flags, fpv.DEC_UNEXTRACT)
        assert np.array_equal(raw, frames)
        img = ctx.decode(high, low, flags)
        assert np.array_equal(img, frames << 4)

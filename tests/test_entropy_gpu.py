"""GPU entropy coder (csrc/fpv_entropy.cu): its container chunks must (a) equal the CPU restatement of the
bitstream (tests/huffcoder_ref.py) byte for byte and (b) decode with libbrotlidec -- the library the reference's
decoder uses (fusion_power_video.cc:186-214) -- back to exactly the planes the transform produced."""
import struct

import numpy as np
import pytest

import huffcoder_ref as href

pytestmark = pytest.mark.gpu


def expected_chunk(flags, high, low, preview):
    bp = href.encode_plane(preview)
    core = bytes([flags]) + (b"" if (flags & 4) or low is None else href.encode_plane(low)) + href.encode_plane(high)
    total = 10 + len(bp) + len(core)
    return struct.pack("<IBIB", total, 0, len(bp) + 1, (flags & 2) | 4) + bp + core


def split_chunk(chunk, P, PP, has_low):
    """Container chunk -> flags, planes, using libbrotlidec the way the reference's DecompressImage does."""
    total, kind, bp1, pflags = struct.unpack_from("<IBIB", chunk, 0)
    assert total == len(chunk) and kind == 0
    bp = chunk[10:10 + bp1 - 1]
    preview = href.brotli_decode(bp, PP)
    core = chunk[10 + bp1 - 1:]
    flags = core[0]
    rest = core[1:]
    # the low and high streams are concatenated without a length: decode the first, find where it ended
    streams = []
    for _ in range(2 if (has_low and not flags & 4) else 1):
        out, used = href.brotli_decode_prefix(rest, P)
        streams.append(out)
        rest = rest[used:]
    assert len(rest) == 0
    low = streams[0] if len(streams) == 2 else None
    return flags, pflags, streams[-1], low, preview


CASES = [
    # W, H, bits, shift, big_endian, n
    (1280, 160, 12, 4, False, 5),
    (1024, 128, 16, 0, False, 3),
    (256, 64, 8, 8, False, 3),      # no low plane
    (520, 36, 16, 0, True, 2),      # ragged widths, unaligned preview planes
    (12, 8, 16, 0, False, 4),       # tiny: preview of 6 bytes
    (2048, 96, 16, 0, False, 2),
]


@pytest.mark.parametrize("W,H,bits,shift,be,n", CASES)
def test_stream_matches_cpu_restatement_and_decodes(W, H, bits, shift, be, n):
    import fusion_power_video_b200 as fpv
    from fusion_power_video_b200 import synth

    P, PP = W * H, (W // 4) * (H // 4)
    frames = synth.plasma_frames(n, W, H, bits=bits, seed=7).reshape(n, -1)
    if be:
        frames = frames.byteswap()
    with fpv.Context(W, H, shift, be, max_batch=n) as ctx:
        ctx.set_delta_raw(frames[0])
        flags, high, low, preview = ctx.encode(frames)
        sflags, chunks = ctx.encode_stream(frames)
    assert np.array_equal(flags, sflags)
    for i in range(n):
        exp = expected_chunk(int(flags[i]), high[i], None if low is None else low[i], preview[i])
        assert chunks[i] == exp, f"frame {i}: GPU chunk ({len(chunks[i])} B) differs from the CPU restatement ({len(exp)} B)"
        f2, pflags, h2, l2, p2 = split_chunk(chunks[i], P, PP, low is not None)
        assert f2 == int(flags[i]) and pflags == ((f2 & 2) | 4)
        assert h2 == high[i].tobytes() and p2 == preview[i].tobytes()
        if l2 is not None:
            assert l2 == low[i].tobytes()


@pytest.mark.parametrize("kind", ["constant", "two_values", "noise", "zero_low"])
def test_degenerate_planes(kind):
    import fusion_power_video_b200 as fpv

    W, H, n = 512, 256, 2
    P, PP = W * H, (W // 4) * (H // 4)
    rng = np.random.default_rng(3)
    if kind == "constant":
        frames = np.full((n, P), 0x1234, np.uint16)
    elif kind == "two_values":
        frames = rng.integers(0, 2, (n, P)).astype(np.uint16) * 0x0101
    elif kind == "noise":
        frames = rng.integers(0, 65536, (n, P)).astype(np.uint16)
    else:
        frames = (rng.integers(0, 256, (n, P)).astype(np.uint16) << 8)
    with fpv.Context(W, H, 0, False, max_batch=n) as ctx:
        flags, high, low, preview = ctx.encode(frames, fpv.ENC_NO_DELTA)
        sflags, chunks = ctx.encode_stream(frames, fpv.ENC_NO_DELTA)
    assert np.array_equal(flags, sflags)
    for i in range(n):
        assert chunks[i] == expected_chunk(int(flags[i]), high[i], low[i], preview[i])
        f2, _, h2, l2, p2 = split_chunk(chunks[i], P, PP, True)
        assert h2 == high[i].tobytes() and p2 == preview[i].tobytes()
        assert (l2 is None) == bool(flags[i] & 4)


def test_full_size_batch_round_trip():
    """C2 geometry, a batch of 16: every plane stream decodes with libbrotlidec; sizes beat or match brotli q1 loosely."""
    import fusion_power_video_b200 as fpv
    from fusion_power_video_b200 import synth

    W, H, n = 1280, 800, 16
    P, PP = W * H, (W // 4) * (H // 4)
    frames = synth.plasma_frames(n, W, H, bits=12, seed=1).reshape(n, -1)
    with fpv.Context(W, H, 4, False, max_batch=n) as ctx:
        ctx.set_delta_raw(frames[0])
        flags, high, low, preview = ctx.encode(frames)
        sflags, chunks = ctx.encode_stream(frames)
    assert np.array_equal(flags, sflags)
    for i in range(n):
        f2, _, h2, l2, p2 = split_chunk(chunks[i], P, PP, True)
        assert h2 == high[i].tobytes() and l2 == low[i].tobytes() and p2 == preview[i].tobytes()
        assert len(chunks[i]) < 0.55 * 2 * P      # 12-bit frames: well under 9 bits per pixel


def test_crafted_planes_depth_limit_and_ragged_chunks():
    """Planes uploaded as they are (fpv_entropy_device): Fibonacci symbol counts force the depth-limit loop of the
    Huffman construction (an unlimited tree would be 21 deep), a plane size that is not a multiple of the chunk
    size leaves a ragged last chunk, a two-symbol preview uses a one-bit code."""
    import torch
    import fusion_power_video_b200 as fpv

    W, H, n = 1280, 68, 3          # P = 87040 = 65536 + 21504
    P, PP = W * H, (W // 4) * (H // 4)
    rng = np.random.default_rng(21)
    fib = [1, 1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144, 233, 377, 610, 987, 1597, 2584, 4181, 6765, 10946, 17711]
    deep = np.concatenate([np.full(v, i * 7 % 256, np.uint8) for i, v in enumerate(fib)])
    high = np.zeros((n, P), np.uint8)
    for i in range(n):
        rng.shuffle(deep)
        high[i, :deep.size] = deep                      # first chunk: 46367 skewed symbols + zeros
        high[i, 65536:] = rng.integers(0, 256, P - 65536)
    low = np.minimum(rng.geometric(0.05, (n, P)) - 1, 255).astype(np.uint8)
    preview = rng.integers(0, 2, (n, PP)).astype(np.uint8) * 254
    flags = np.array([3, 4, 0], np.uint8)               # frame 1 has no low plane
    dev = torch.device("cuda", 0)
    t = {k: torch.from_numpy(v).to(dev) for k, v in dict(high=high, low=low, preview=preview, flags=flags).items()}
    with fpv.Context(W, H, 0, False, max_batch=n) as ctx:
        cap = ctx.stream_bound(n)
        out = torch.zeros(cap, dtype=torch.uint8, device=dev)
        off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        ctx.entropy_device(t["flags"].data_ptr(), t["high"].data_ptr(), t["low"].data_ptr(), t["preview"].data_ptr(), n,
                           out.data_ptr(), cap, off.data_ptr())
        torch.cuda.synchronize()
    off = off.cpu().numpy()
    out = out.cpu().numpy()
    assert max(href.huffman_lengths(np.bincount(high[0, :65536], minlength=256).tolist(), 30)) > 15
    for i in range(n):
        chunk = out[off[i]:off[i + 1]].tobytes()
        assert chunk == expected_chunk(int(flags[i]), high[i], low[i], preview[i]), f"frame {i}"
        f2, _, h2, l2, p2 = split_chunk(chunk, P, PP, True)
        assert h2 == high[i].tobytes() and p2 == preview[i].tobytes() and (l2 is None) == bool(flags[i] & 4)


def test_more_frames_than_layout_threads():
    """More than 1024 frames in one call: k_entropy_layout scans them in passes with a carry."""
    import fusion_power_video_b200 as fpv
    from fusion_power_video_b200 import synth

    W, H, n = 64, 16, 1100
    frames = synth.plasma_frames(n, W, H, bits=12, seed=3).reshape(n, -1)
    with fpv.Context(W, H, 4, False, max_batch=n) as ctx:
        ctx.set_delta_raw(frames[0])
        flags, high, low, preview = ctx.encode(frames)
        sflags, chunks = ctx.encode_stream(frames)
    assert np.array_equal(flags, sflags) and len(chunks) == n
    for i in range(n):
        assert int.from_bytes(chunks[i][:4], "little") == len(chunks[i]), f"frame {i}: container size field"
    for i in (0, 1, 511, 1023, 1024, 1025, n - 1):
        assert chunks[i] == expected_chunk(int(flags[i]), high[i], low[i], preview[i]), f"frame {i}"


def _chunk_table(chunks, P):
    """Container chunks of encode_stream -> (blob, chunk table, flags) for fpv_decode_coded, walking the directories
    with the CPU restatement's scanner."""
    blob, table, flags = bytearray(), [], []
    for f, ch in enumerate(chunks):
        total, kind, bp1 = struct.unpack_from("<IBI", ch)
        assert total == len(ch) and kind == 0
        core = ch[10 + bp1 - 1:]
        fl = core[0]
        flags.append(fl)
        pos = 1
        for plane in ([1, 0] if not fl & 4 else [0]):
            offs, length = href.scan_plane(core[pos:], P)
            table += [(len(blob) + pos + o, f, plane, k) for k, o in enumerate(offs)]
            pos += length
        assert pos == len(core)
        blob += core
        blob += bytes(-len(blob) % 16)
    return bytes(blob), table, np.array(flags, np.uint8)


@pytest.mark.parametrize("W,H,bits,shift,n", [(1280, 160, 12, 4, 6), (1024, 136, 16, 0, 5), (320, 48, 8, 8, 4), (2048, 72, 16, 0, 3),
                                              (100, 100, 16, 0, 3), (512, 132, 12, 4, 50), (256, 64, 16, 0, 17)])
def test_gpu_entropy_decoder_round_trip(W, H, bits, shift, n):
    """fpv_decode_coded: coded container chunks -> (k_entropy_decode: one warp per chunk from the directories) ->
    inverse transform == the raw input; same images as the planes through fpv_decode."""
    import fusion_power_video_b200 as fpv
    from fusion_power_video_b200 import synth

    frames = synth.plasma_frames(n, W, H, bits=bits, seed=W + n).reshape(n, -1)
    frames[n - 1] = frames[0]                  # planes of zeros: constant chunks
    # (batches of 17 and 50 frames are cut into 2 and 4 pieces that run on the context's slots side by side)
    with fpv.Context(W, H, shift, False, max_batch=max(8, n)) as ctx:
        ctx.set_delta_raw(frames[0])
        flags, chunks = ctx.encode_stream(frames)
        pf, high, low, _ = ctx.encode(frames)
        want = ctx.decode(high, low, pf)
        blob, table, fl = _chunk_table(chunks, W * H)
        assert np.array_equal(fl, flags)
        got = ctx.decode_coded(blob, table, fl)
        assert np.array_equal(got, want)
        raw = ctx.decode_coded(blob, table, fl, fpv.DEC_UNEXTRACT)
        assert np.array_equal(raw, frames)


def test_gpu_entropy_decoder_crafted_planes():
    """Planes straight into the coder (deep trees: codes longer than the 10-bit table; incompressible: raw chunks;
    constant; a ragged last chunk), decoded again by k_entropy_decode; flags 0 keeps the inverse transform out of it."""
    import torch

    import fusion_power_video_b200 as fpv

    W, H, n = 1024, 68, 6            # 69632 bytes per plane: one full chunk and a ragged one
    P, PP = W * H, (W // 4) * (H // 4)
    rng = np.random.default_rng(3)
    wts = 1.7 ** -np.arange(40)
    planes = [rng.choice(40, P, p=wts / wts.sum()).astype(np.uint8),                      # depth-limited deep tree
              rng.integers(0, 256, P).astype(np.uint8),                                 # raw chunks
              np.full(P, 9, np.uint8),                                                  # constant
              np.minimum(rng.geometric(0.05, P) - 1, 255).astype(np.uint8),
              (np.cumsum(rng.integers(-2, 3, P)) & 255).astype(np.uint8),
              rng.integers(0, 2, P).astype(np.uint8)]
    high = np.stack(planes)
    low = np.stack(planes[::-1])
    prev = np.zeros((n, PP), np.uint8)
    flags = np.zeros(n, np.uint8)
    dev = torch.device("cuda", 0)
    th, tl, tp, tf = (torch.from_numpy(a).to(dev) for a in (high, low, prev, flags))
    with fpv.Context(W, H, 0, False, max_batch=n) as ctx:
        cap = ctx.stream_bound(n)
        out = torch.zeros(cap, dtype=torch.uint8, device=dev)
        off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        ctx.entropy_device(tf.data_ptr(), th.data_ptr(), tl.data_ptr(), tp.data_ptr(), n, out.data_ptr(), cap, off.data_ptr())
        torch.cuda.synchronize()
        off, out = off.cpu().numpy(), out.cpu().numpy()
        chunks = [out[off[i]:off[i + 1]].tobytes() for i in range(n)]
        blob, table, fl = _chunk_table(chunks, P)
        img = ctx.decode_coded(blob, table, fl)
        assert np.array_equal(img, (high.astype(np.uint16) << 8) | low)
        # a damaged directory is refused, not decoded to garbage: span 1's bit position of the first chunk
        bad = bytearray(blob)
        at = table[0][0] + 2 + 138 + 3
        bad[at] ^= 0x10
        with pytest.raises(fpv.FpvError):
            ctx.decode_coded(bytes(bad), table, fl)
        # and a chunk table that misses a chunk is refused on the host
        with pytest.raises(fpv.FpvError):
            ctx.decode_coded(blob, table[:-1], fl)

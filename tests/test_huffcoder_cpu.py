"""CPU checks of the entropy coder's bitstream restatement (tests/huffcoder_ref.py): every stream it writes must
decode with libbrotlidec -- the decoder library the reference links (fusion_power_video.cc:186-214) -- to the input.
The GPU coder is compared with this restatement byte for byte in tests/test_entropy_gpu.py."""
import numpy as np
import pytest

import huffcoder_ref as href


def _cases():
    rng = np.random.default_rng(11)
    fib = [1, 1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144, 233, 377, 610, 987, 1597, 2584, 4181, 6765, 10946, 17711, 28657]
    deep = np.concatenate([np.full(v, i, np.uint8) for i, v in enumerate(fib)])
    rng.shuffle(deep)
    return {
        "one_byte": np.array([5], np.uint8),
        "three_bytes": rng.integers(0, 256, 3).astype(np.uint8),
        "constant": np.full(1000, 7, np.uint8),
        "two_symbols": rng.integers(0, 2, 5000).astype(np.uint8),
        "three_symbols_two_chunks": rng.integers(0, 3, 70000).astype(np.uint8),
        "uniform_noise_raw_fallback": rng.integers(0, 256, 200000).astype(np.uint8),
        "all_lengths_equal": np.tile(np.arange(256, dtype=np.uint8), 256),
        "geometric": np.minimum(rng.geometric(0.3, 300000) - 1, 255).astype(np.uint8),
        "skewed_ragged_tail": np.minimum(rng.geometric(0.02, 65536 * 3 + 17) - 1, 255).astype(np.uint8),
        "fibonacci_depth_limit": deep,
        "chunk_boundary_minus_1": rng.integers(0, 40, 65535).astype(np.uint8),
        "chunk_boundary_plus_1": rng.integers(0, 40, 65537).astype(np.uint8),
    }


CASES = _cases()


@pytest.mark.parametrize("name", sorted(CASES))
def test_stream_decodes_with_libbrotlidec(name):
    data = CASES[name]
    stream = href.encode_plane(data)
    assert href.brotli_decode(stream, data.size) == data.tobytes()
    out, used = href.brotli_decode_prefix(stream + b"\x55\xaa", data.size)   # trailing bytes belong to the next stream
    assert out == data.tobytes() and used == len(stream)


@pytest.mark.parametrize("name", sorted(CASES))
def test_directories_alone_reconstruct_the_plane(name):
    """What the GPU entropy decoder reads: chunk sizes, code lengths and span bit positions from the directory
    meta-blocks, then only the literal bits -- restated in huffcoder_ref.decode_chunk."""
    data = CASES[name]
    stream = href.encode_plane(data)
    scan = href.scan_plane(stream + b"\x55\xaa", data.size)
    assert scan is not None
    offs, length = scan
    assert length == len(stream) and len(offs) == (data.size + href.CHUNK - 1) // href.CHUNK
    ends = offs[1:] + [length - 1]
    back = b"".join(href.decode_chunk(stream[a:b]) for a, b in zip(offs, ends))
    assert back == data.tobytes()


def test_libbrotli_streams_carry_no_directory():
    # a stream that starts like brotli quality 1 output is not mistaken for a directory-carrying one
    assert href.scan_plane(bytes([0x1b, 0x03, 0x00, 0xf8, 0x25, 0x00, 0xa2, 0x90, 0xa8, 0x00]), 4) is None


def test_depth_limit_is_enforced():
    fib = [1, 1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144, 233, 377, 610, 987, 1597, 2584, 4181, 6765, 10946, 17711, 28657]
    assert max(href.huffman_lengths(fib, 15)) <= 15
    assert max(href.huffman_lengths(fib, 30)) > 15          # the unlimited tree is deeper
    lens = href.huffman_lengths(fib, 15)
    assert sum(2.0 ** -l for l in lens if l) == 1.0          # still a complete prefix code


def test_compressible_planes_shrink_and_noise_does_not_grow():
    rng = np.random.default_rng(5)
    skew = np.minimum(rng.geometric(0.4, 1 << 18) - 1, 255).astype(np.uint8)
    assert len(href.encode_plane(skew)) < 0.4 * skew.size
    noise = rng.integers(0, 256, 1 << 18).astype(np.uint8)
    assert len(href.encode_plane(noise)) <= noise.size + (6 + href.DIR_BLOCK) * 4 + 2

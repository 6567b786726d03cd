"""CPU restatement of the GPU entropy coder's bitstream (csrc/fpv_entropy.cu) -- test infrastructure.

The GPU coder emits, per plane, a valid RFC 7932 (brotli) stream made of independent, byte-aligned
chunks: one compressed meta-block per chunk with a single literal prefix code (canonical Huffman,
max length 15, built from the chunk's histogram), one insert-and-copy command that inserts the whole
chunk, no backward references, followed by an empty metadata meta-block that pads to a byte boundary.
Any brotli decoder (the reference's decode.cc uses libbrotlidec) decodes it.

This module states the same bitstream in plain Python / numpy so that tests can (a) check the format
against libbrotlidec without a GPU and (b) compare the GPU output byte for byte.
"""
from __future__ import annotations

import ctypes as C
import heapq

import numpy as np

CHUNK = 65536
MAX_BITS = 15
# RFC 7932 section 5: insert length code -> (base, extra bits)
INSERT_BASE = [0, 1, 2, 3, 4, 5, 6, 8, 10, 14, 18, 26, 34, 50, 66, 98, 130, 194, 322, 578, 1090, 2114, 6210, 22594]
INSERT_EXTRA = [0, 0, 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 7, 8, 9, 10, 12, 14, 24]
# RFC 7932 section 3.5: order in which code length code lengths are stored, and their fixed code
CL_ORDER = [1, 2, 3, 4, 0, 5, 17, 6, 16, 7, 8, 9, 10, 11, 12, 13, 14, 15]
CLCL_BITS = [(0, 2), (7, 4), (3, 3), (2, 2), (1, 2), (15, 4)]   # value -> (bits LSB-first, nbits)


class BitWriter:
    def __init__(self):
        self.acc, self.n, self.out = 0, 0, bytearray()

    def put(self, value, nbits):
        assert 0 <= value < (1 << nbits) or nbits == 0
        self.acc |= value << self.n
        self.n += nbits
        while self.n >= 8:
            self.out.append(self.acc & 0xFF)
            self.acc >>= 8
            self.n -= 8

    def align(self):
        if self.n:
            self.out.append(self.acc & 0xFF)
            self.acc, self.n = 0, 0


def huffman_lengths(counts, max_bits):
    """Code lengths of a Huffman code for `counts` (>= 2 non-zero entries), depth limited to max_bits by
    re-running the construction with counts floored at a doubling limit (the same heuristic the GPU uses).
    Ties are broken deterministically: by (weight, node id), leaves in symbol order first."""
    n = len(counts)
    limit = 1
    while True:
        heap = [(max(int(c), limit), i) for i, c in enumerate(counts) if c]
        assert len(heap) >= 2
        heapq.heapify(heap)
        parent = {}
        nxt = n
        while len(heap) > 1:
            a = heapq.heappop(heap)
            b = heapq.heappop(heap)
            parent[a[1]] = nxt
            parent[b[1]] = nxt
            heapq.heappush(heap, (a[0] + b[0], nxt))
            nxt += 1
        depth = [0] * n
        for i, c in enumerate(counts):
            if c:
                d, k = 0, i
                while k in parent:
                    k = parent[k]
                    d += 1
                depth[i] = d
        if max(depth) <= max_bits:
            return depth
        limit *= 2


def canonical_codes(lengths):
    """Canonical code of RFC 7932 3.2, bit-reversed for LSB-first writing."""
    max_len = max(lengths) if len(lengths) else 0
    bl_count = [0] * (max_len + 2)
    for l in lengths:
        if l:
            bl_count[l] += 1
    code, next_code = 0, [0] * (max_len + 2)
    for b in range(1, max_len + 1):
        code = (code + bl_count[b - 1]) << 1
        next_code[b] = code
    out = [0] * len(lengths)
    for s, l in enumerate(lengths):
        if l:
            c = next_code[l]
            next_code[l] += 1
            out[s] = int(format(c, "0%db" % l)[::-1], 2)
    return out


def write_complex_code(bw, lengths):
    """Complex prefix code (RFC 7932 3.5) without run-length symbols: every code length up to the last
    non-zero one is written with the code length code."""
    last = max(i for i, l in enumerate(lengths) if l)
    seq = lengths[:last + 1]
    cl_hist = [0] * 18
    for l in seq:
        cl_hist[l] += 1
    used = [i for i in range(18) if cl_hist[i]]
    if len(used) == 1:
        cl_len = [0] * 18
        cl_len[used[0]] = 1          # stored as 1, costs 0 bits per symbol (decoder's single-code case)
        to_store = 18
        cl_code = [0] * 18
        cl_eff = [0] * 18
    else:
        cl_len = huffman_lengths(cl_hist, 5)
        to_store = 18
        while cl_len[CL_ORDER[to_store - 1]] == 0:
            to_store -= 1
        cl_code = canonical_codes(cl_len)
        cl_eff = cl_len
    skip = 0
    if cl_len[CL_ORDER[0]] == 0 and cl_len[CL_ORDER[1]] == 0:
        skip = 3 if cl_len[CL_ORDER[2]] == 0 else 2
    bw.put(skip, 2)
    for i in range(skip, to_store):
        v, nb = CLCL_BITS[cl_len[CL_ORDER[i]]]
        bw.put(v, nb)
    for l in seq:
        bw.put(cl_code[l], cl_eff[l])


def encode_chunk(data: np.ndarray, first: bool) -> bytes:
    n = int(data.size)
    assert 1 <= n <= (1 << 24)
    bw = BitWriter()
    if first:
        bw.put(0, 1)                       # WBITS = 16
    hist = np.bincount(data, minlength=256)
    used = np.flatnonzero(hist)
    bw.put(0, 1)                           # ISLAST = 0
    nib = 4 if n - 1 < (1 << 16) else 5 if n - 1 < (1 << 20) else 6
    bw.put(nib - 4, 2)
    bw.put(n - 1, 4 * nib)                 # MLEN - 1
    bw.put(0, 1)                           # ISUNCOMPRESSED = 0
    bw.put(0, 1); bw.put(0, 1); bw.put(0, 1)   # NBLTYPESL = NBLTYPESI = NBLTYPESD = 1
    bw.put(0, 2); bw.put(0, 4)             # NPOSTFIX = 0, NDIRECT = 0
    bw.put(0, 2)                           # literal context mode of block type 0
    bw.put(0, 1); bw.put(0, 1)             # NTREESL = NTREESD = 1
    # literal prefix code
    if used.size == 1:
        bw.put(1, 2); bw.put(0, 2); bw.put(int(used[0]), 8)          # simple code, NSYM = 1
        lengths, codes = [0] * 256, [0] * 256
    else:
        lengths = huffman_lengths(hist.tolist(), MAX_BITS)
        codes = canonical_codes(lengths)
        write_complex_code(bw, lengths)
    # insert-and-copy prefix code: one symbol = (insert code ic, copy code 0)
    ic = max(i for i in range(24) if INSERT_BASE[i] <= n)
    cell = 128 if ic < 8 else 256 if ic < 16 else 448
    bw.put(1, 2); bw.put(0, 2); bw.put(cell + ((ic & 7) << 3), 10)
    # distance prefix code: one symbol (0), alphabet 64
    bw.put(1, 2); bw.put(0, 2); bw.put(0, 6)
    # the command: symbol costs 0 bits; insert extra bits; copy code 0 has none
    bw.put(n - INSERT_BASE[ic], INSERT_EXTRA[ic])
    lit_start = 8 * len(bw.out) + bw.n
    lit_bits = int(sum(int(hist[i]) * lengths[i] for i in range(256)))
    if (lit_start + lit_bits + 6 + 7) // 8 > n + 4:
        # Huffman coding does not pay: uncompressed meta-block (header up to ISUNCOMPRESSED = 1, pad, raw bytes)
        bw = BitWriter()
        if first:
            bw.put(0, 1)
        bw.put(0, 1); bw.put(nib - 4, 2); bw.put(n - 1, 4 * nib); bw.put(1, 1)
        bw.align()
        return bytes(bw.out) + data.tobytes()
    for b in data.tolist():
        bw.put(codes[b], lengths[b])
    # empty metadata meta-block: pads to the byte boundary
    bw.put(0, 1); bw.put(3, 2); bw.put(0, 1); bw.put(0, 2)
    bw.align()
    return bytes(bw.out)


def encode_plane(data, chunk=CHUNK) -> bytes:
    data = np.ascontiguousarray(np.asarray(data, dtype=np.uint8).reshape(-1))
    out = bytearray()
    if data.size == 0:
        return bytes([0x06])               # WBITS 16, ISLAST, ISLASTEMPTY
    for ci, off in enumerate(range(0, data.size, chunk)):
        out += encode_chunk(data[off:off + chunk], ci == 0)
    out.append(0x03)                       # ISLAST = 1, ISLASTEMPTY = 1
    return bytes(out)


_dec = None


def brotli_decode(stream: bytes, expect: int) -> bytes:
    """libbrotlidec one-shot decode (the library the reference's decoder links)."""
    global _dec
    if _dec is None:
        _dec = C.CDLL("/usr/lib/x86_64-linux-gnu/libbrotlidec.so.1")
        _dec.BrotliDecoderDecompress.argtypes = [C.c_size_t, C.c_char_p, C.POINTER(C.c_size_t), C.c_char_p]
        _dec.BrotliDecoderDecompress.restype = C.c_int
    out = C.create_string_buffer(max(expect, 1))
    n = C.c_size_t(expect)
    r = _dec.BrotliDecoderDecompress(len(stream), stream, C.byref(n), out)
    if r != 1:
        raise ValueError(f"brotli decoder result {r}")
    return out.raw[:n.value]


def brotli_decode_prefix(data: bytes, expect: int):
    """Decodes ONE brotli stream from the front of `data` (streaming API, as the reference's BrotliDecompress does,
    fusion_power_video.cc:186-214).  Returns (decoded bytes, number of input bytes the stream occupied)."""
    L = C.CDLL("/usr/lib/x86_64-linux-gnu/libbrotlidec.so.1")
    L.BrotliDecoderCreateInstance.restype = C.c_void_p
    L.BrotliDecoderCreateInstance.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.BrotliDecoderDestroyInstance.argtypes = [C.c_void_p]
    L.BrotliDecoderDecompressStream.restype = C.c_int
    L.BrotliDecoderDecompressStream.argtypes = [C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_void_p),
                                                C.POINTER(C.c_size_t), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    st = L.BrotliDecoderCreateInstance(None, None, None)
    inbuf = C.create_string_buffer(data, len(data))
    out = C.create_string_buffer(max(expect, 1))
    avail_in, avail_out = C.c_size_t(len(data)), C.c_size_t(expect)
    next_in = C.c_void_p(C.addressof(inbuf))
    next_out = C.c_void_p(C.addressof(out))
    r = L.BrotliDecoderDecompressStream(st, C.byref(avail_in), C.byref(next_in), C.byref(avail_out), C.byref(next_out), None)
    L.BrotliDecoderDestroyInstance(st)
    if r != 1:
        raise ValueError(f"brotli stream decoder result {r}")
    return out.raw[:expect - avail_out.value], len(data) - avail_in.value

"""CPU restatement of the GPU entropy coder's bitstream (csrc/fpv_entropy.cu) -- test infrastructure.

The GPU coder emits, per plane, a valid RFC 7932 (brotli) stream made of independent, byte-aligned
chunks: one compressed meta-block per chunk with a single literal prefix code (canonical Huffman,
max length 15, built from the chunk's histogram), one insert-and-copy command that inserts the whole
chunk, no backward references, followed by an empty metadata meta-block that pads to a byte boundary.
Every chunk starts with a metadata meta-block -- skipped by every brotli decoder -- that carries a DIRECTORY for
the GPU entropy decoder (k_entropy_decode): the chunk's size, its code lengths and the bit positions of its 2048-byte
spans, so that a warp decodes a chunk without parsing a prefix-code header and with every lane busy.
Any brotli decoder (the reference's decode.cc uses libbrotlidec) decodes the stream.

This module states the same bitstream in plain Python / numpy so that tests can (a) check the format
against libbrotlidec without a GPU and (b) compare the GPU output byte for byte.
"""
from __future__ import annotations

import ctypes as C
import heapq

import numpy as np

CHUNK = 65536
MAX_BITS = 15
SPAN = 2048            # bytes of a chunk one decoder lane reconstructs
DIR_BYTES = 234        # payload of the directory meta-block
DIR_BLOCK = 2 + DIR_BYTES   # its header is 14 (first chunk of a stream: 15) bits, padded to two bytes
KIND_HUFFMAN, KIND_RAW, KIND_CONSTANT = 0, 1, 2
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


def directory(kind, chunk_bytes, n, constant, lengths, spans):
    """The directory payload: 'F' 'D' | version 1 | kind | chunk bytes u24 | n - 1 u16 | constant symbol |
    256 code lengths as nibbles | 32 span bit positions u24 (relative to the start of the chunk)."""
    d = bytearray(DIR_BYTES)
    d[0], d[1], d[2], d[3] = 0x46, 0x44, 1, kind
    d[4:7] = chunk_bytes.to_bytes(3, "little")
    d[7:9] = (n - 1).to_bytes(2, "little")
    d[9] = constant
    for i in range(128):
        d[10 + i] = lengths[2 * i] | (lengths[2 * i + 1] << 4)
    for k, v in enumerate(spans):
        d[138 + 3 * k:141 + 3 * k] = v.to_bytes(3, "little")
    return bytes(d)


def encode_chunk(data: np.ndarray, first: bool) -> bytes:
    n = int(data.size)
    assert 1 <= n <= CHUNK
    head = BitWriter()
    if first:
        head.put(0, 1)                     # WBITS = 16
    # the directory: a metadata meta-block (ISLAST 0, MNIBBLES coded 3, reserved 0, MSKIPBYTES 1, MSKIPLEN - 1)
    head.put(0, 1); head.put(3, 2); head.put(0, 1); head.put(1, 2); head.put(DIR_BYTES - 1, 8)
    head.align()
    assert len(head.out) == 2
    bw = BitWriter()                       # the compressed meta-block starts on the byte after the directory
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
        bw.put(0, 1); bw.put(nib - 4, 2); bw.put(n - 1, 4 * nib); bw.put(1, 1)
        bw.align()
        body = bytes(bw.out) + data.tobytes()
        spans = [8 * (DIR_BLOCK + len(bw.out))] + [0] * 31
        return bytes(head.out) + directory(KIND_RAW, DIR_BLOCK + len(body), n, 0, [0] * 256, spans) + body
    spans = [0] * 32
    for i, b in enumerate(data.tolist()):
        if i % SPAN == 0:
            spans[i // SPAN] = 8 * DIR_BLOCK + 8 * len(bw.out) + bw.n
        bw.put(codes[b], lengths[b])
    # empty metadata meta-block: pads to the byte boundary
    bw.put(0, 1); bw.put(3, 2); bw.put(0, 1); bw.put(0, 2)
    bw.align()
    kind = KIND_CONSTANT if used.size == 1 else KIND_HUFFMAN
    return (bytes(head.out) + directory(kind, DIR_BLOCK + len(bw.out), n, int(used[0]) if used.size == 1 else 0, lengths, spans)
            + bytes(bw.out))


def encode_plane(data, chunk=CHUNK) -> bytes:
    data = np.ascontiguousarray(np.asarray(data, dtype=np.uint8).reshape(-1))
    out = bytearray()
    if data.size == 0:
        return bytes([0x06])               # WBITS 16, ISLAST, ISLASTEMPTY
    for ci, off in enumerate(range(0, data.size, chunk)):
        out += encode_chunk(data[off:off + chunk], ci == 0)
    out.append(0x03)                       # ISLAST = 1, ISLASTEMPTY = 1
    return bytes(out)


def scan_plane(stream: bytes, plane_bytes: int):
    """Walks the directories of a plane stream: returns (chunk offsets, stream length) or None if the stream does not
    carry them (a libbrotli stream)."""
    offs, pos = [], 0
    for _ in range((plane_bytes + CHUNK - 1) // CHUNK):
        if pos + DIR_BLOCK > len(stream) or stream[pos + 2:pos + 5] != b"FD\x01":
            return None
        offs.append(pos)
        pos += int.from_bytes(stream[pos + 6:pos + 9], "little")
    if pos >= len(stream) or stream[pos] != 0x03:
        return None
    return offs, pos + 1


def decode_chunk(chunk: bytes) -> bytes:
    """What k_entropy_decode does with one chunk: only the directory and the literal bits are read."""
    d = chunk[2:2 + DIR_BYTES]
    assert d[0:3] == b"FD\x01"
    kind, n, constant = d[3], int.from_bytes(d[7:9], "little") + 1, d[9]
    spans = [int.from_bytes(d[138 + 3 * k:141 + 3 * k], "little") for k in range(32)]
    if kind == KIND_RAW:
        o = spans[0] // 8
        return chunk[o:o + n]
    if kind == KIND_CONSTANT:
        return bytes([constant]) * n
    lengths = [(d[10 + i // 2] >> (4 * (i & 1))) & 15 for i in range(256)]
    codes = canonical_codes(lengths)
    table = {(codes[s], lengths[s]): s for s in range(256) if lengths[s]}
    bits = int.from_bytes(chunk, "little")
    out = bytearray()
    for k in range((n + SPAN - 1) // SPAN):
        pos = spans[k]
        for _ in range(min(SPAN, n - k * SPAN)):
            for l in range(1, MAX_BITS + 1):
                sym = table.get(((bits >> pos) & ((1 << l) - 1), l))
                if sym is not None:
                    out.append(sym)
                    pos += l
                    break
            else:
                raise ValueError("no code matches")
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

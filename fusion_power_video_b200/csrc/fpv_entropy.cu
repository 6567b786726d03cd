// fpv_entropy.cu -- chunk-parallel GPU entropy coder for the three byte planes of a frame.
//
// The reference hands each plane to libbrotli at quality 1 (fusion_power_video.cc:643-688,
// BrotliEncoderCompress(1, 22, GENERIC)); its decoder accepts ANY valid RFC 7932 stream
// (BrotliDecoderDecompressStream, .cc:186-214).  The reference notes that "only the entropy
// coding matters, not the LZ77" (.cc:166-169), so this coder emits a brotli stream that uses the
// format's entropy stage only:
//
//   plane stream := chunk* 0x03                       (0x03 = ISLAST, ISLASTEMPTY)
//   chunk        := [WBITS bit, first chunk only]
//                   metadata meta-block of kDirBytes bytes: the DIRECTORY for k_entropy_decode (below); brotli
//                     decoders skip it
//                   compressed meta-block: MLEN = chunk size (<= 65536), one block type per
//                     category, NTREES = 1, one literal prefix code (canonical Huffman from the
//                     chunk's own histogram, max depth 15, stored as a "complex" code without
//                     run-length symbols), a one-symbol insert-and-copy code whose only command
//                     inserts the whole chunk, a one-symbol distance code, the literals
//                   empty metadata meta-block, which pads the chunk to a byte boundary
//                 | uncompressed meta-block when Huffman coding would not shrink the chunk
//
// Chunks are byte aligned and independent, so one CTA codes one chunk and a gather pass
// concatenates them; the frame container bytes (.cc:820-846: total | 0 | 1+|bp| | pflags | bp |
// flags | low? | high) are written on the device too, so the host only forwards pointers.
// tests/huffcoder_ref.py states the same bitstream on the CPU; tests check it against libbrotlidec
// and against this file byte for byte.
#include <cuda_runtime.h>
#include <stdint.h>

#include "fpv_internal.h"

namespace fpv {

namespace {

constexpr int kET = 256;                       // threads per CTA
// The chunk's bit buffer in shared memory is a WINDOW of 32 KiB of the coded chunk: the directory, the prefix-code
// header and -- for chunks that compress to less than half, the usual case -- all literal bits sit in window 0; larger
// chunks are packed window by window (a thread walks its span once per window it touches).  A buffer for the whole
// chunk (66 KB) left room for three CTAs per SM only, and the kernel's time goes into barriers around its serial
// sections (one thread merges the Huffman tree, writes the header): more resident CTAs hide exactly that.
constexpr uint32_t kWinWords = 8192;
constexpr uint32_t kOutWords = kWinWords + 8;

// RFC 7932 section 5: insert length code -> base, extra bits
__constant__ uint32_t kInsBase[24] = {0, 1, 2, 3, 4, 5, 6, 8, 10, 14, 18, 26, 34, 50, 66, 98, 130, 194, 322, 578, 1090, 2114, 6210, 22594};
__constant__ uint8_t kInsExtra[24] = {0, 0, 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 7, 8, 9, 10, 12, 14, 24};
// RFC 7932 section 3.5: storage order of the code length code lengths and their fixed code
__constant__ uint8_t kClOrder[18] = {1, 2, 3, 4, 0, 5, 17, 6, 16, 7, 8, 9, 10, 11, 12, 13, 14, 15};
__constant__ uint8_t kClclBits[6] = {0, 7, 3, 2, 1, 15};
__constant__ uint8_t kClclLen[6] = {2, 4, 3, 2, 2, 4};

struct Shared {
  uint32_t out[kOutWords];       // the chunk's bit buffer
  uint32_t hist[256];
  uint32_t keys[256];            // sort keys (weight << 8 | symbol), then scratch
  uint32_t wt[512];              // Huffman nodes: leaves in sorted order, then internal nodes
  uint16_t parent[512];
  uint32_t lc[256];              // per symbol: bit-reversed code | length << 16
  uint8_t len[256];
  uint32_t bl_count[16], next_code[16];
  uint32_t cl_hist[18], cl_lc[18];
  uint32_t warp_sum[8];
  uint32_t pos;                  // running bit position of the serial writer
  uint32_t maxd, lit_start, misc;
};

__device__ __forceinline__ void put_bits(Shared& s, uint32_t& pos, uint32_t value, uint32_t nbits) {
  if (nbits == 0) return;
  const uint32_t w = pos >> 5, sh = pos & 31;
  atomicOr(&s.out[w], value << sh);
  if (sh + nbits > 32) atomicOr(&s.out[w + 1], value >> (32 - sh));
  pos += nbits;
}

// `nbytes` little-endian bytes of `value` at byte `idx` of the directory payload (each byte is written once)
__device__ __forceinline__ void dir_put(Shared& s, uint32_t idx, uint32_t value, uint32_t nbytes) {
  for (uint32_t i = 0; i < nbytes; i++) {
    const uint32_t b = 2 + idx + i;
    atomicOr(&s.out[b >> 2], ((value >> (8 * i)) & 255u) << (8 * (b & 3u)));
  }
}

// exclusive scan of one value per thread over the CTA; returns the prefix, *total = sum
__device__ __forceinline__ uint32_t block_excl_scan(Shared& s, uint32_t v, uint32_t* total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  __syncthreads();
  if (lane == 31) s.warp_sum[w] = x;
  __syncthreads();
  uint32_t base = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint32_t t = s.warp_sum[i];
    if (i < w) base += t;
    tot += t;
  }
  *total = tot;
  return base + x - v;
}

// Serial Huffman over <= 18 symbols with a depth limit (thread 0 only): the code length code.
// Same construction as the 256-symbol one below: leaves sorted by (max(count, floor), symbol),
// two queues, leaves win ties, the floor doubles until the tree is shallow enough.
__device__ void small_huffman(const uint32_t* counts, int n, int max_bits, uint8_t* depth) {
  uint32_t key[18], wt[36];
  uint8_t par[36];
  for (uint32_t floor_ = 1;; floor_ <<= 1) {
    int m = 0;
    for (int i = 0; i < n; i++)
      if (counts[i]) {
        const uint32_t w = counts[i] > floor_ ? counts[i] : floor_;
        uint32_t k = (w << 8) | (uint32_t)i;
        int j = m++;
        while (j > 0 && key[j - 1] > k) { key[j] = key[j - 1]; j--; }
        key[j] = k;
      }
    for (int i = 0; i < m; i++) wt[i] = key[i] >> 8;
    int li = 0, ni = m, nn = m;
    for (int k = 0; k + 1 < m; k++) {
      int a, b;
      if (li < m && (ni >= nn || wt[li] <= wt[ni])) a = li++; else a = ni++;
      if (li < m && (ni >= nn || wt[li] <= wt[ni])) b = li++; else b = ni++;
      wt[nn] = wt[a] + wt[b];
      par[a] = par[b] = (uint8_t)nn;
      nn++;
    }
    int maxd = 0;
    for (int i = 0; i < n; i++) depth[i] = 0;
    for (int i = 0; i < m; i++) {
      int d = 0, k = i;
      while (k != 2 * m - 2) { k = par[k]; d++; }
      depth[key[i] & 255u] = (uint8_t)d;
      if (d > maxd) maxd = d;
    }
    if (maxd <= max_bits) return;
  }
}

// One byte of the chunk -> histogram
__device__ __forceinline__ void hist4(Shared& s, uint32_t w) {
  atomicAdd(&s.hist[w & 255u], 1u);
  atomicAdd(&s.hist[(w >> 8) & 255u], 1u);
  atomicAdd(&s.hist[(w >> 16) & 255u], 1u);
  atomicAdd(&s.hist[w >> 24], 1u);
}

}  // namespace

__global__ void __launch_bounds__(kET) k_entropy_chunk(const EntropyParams p) {
  extern __shared__ __align__(16) uint8_t esm[];
  Shared& s = *reinterpret_cast<Shared*>(esm);
  const uint32_t tid = threadIdx.x;
  const uint32_t c = blockIdx.x;
  const uint32_t f = c / p.cpf, r = c % p.cpf;
  const uint8_t* base;
  uint64_t plen;
  uint32_t ci;
  if (r < p.cpp) { base = p.preview + (uint64_t)f * p.PP; plen = p.PP; ci = r; }
  else if (r < p.cpp + p.cpl) { base = p.low ? p.low + (uint64_t)f * p.P : nullptr; plen = p.P; ci = r - p.cpp; }
  else { base = p.high + (uint64_t)f * p.P; plen = p.P; ci = r - p.cpp - p.cpl; }
  const bool is_low = r >= p.cpp && r < p.cpp + p.cpl;
  if (is_low && (base == nullptr || (p.flags[f] & kFlagNoLow))) {
    if (tid == 0) p.chunk_bytes[c] = 0;     // the low stream is absent (.cc:658-659)
    return;
  }
  const uint64_t off = (uint64_t)ci * kEntropyChunk;
  const uint32_t n = (uint32_t)(plen - off < kEntropyChunk ? plen - off : kEntropyChunk);
  const bool first = ci == 0, last = off + n == plen;
  const uint8_t* src = base + off;
  const bool aligned = (reinterpret_cast<uintptr_t>(src) & 15u) == 0;

  for (uint32_t i = tid; i < kOutWords; i += kET) s.out[i] = 0;
  s.hist[tid] = 0;
  if (tid < 18) s.cl_hist[tid] = 0;
  if (tid < 16) s.bl_count[tid] = 0;
  __syncthreads();

  // ---- histogram -----------------------------------------------------------------------------
  if (aligned) {
    const uint32_t n16 = n / 16;
    for (uint32_t i = tid; i < n16; i += kET) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(src) + i);
      hist4(s, v.x); hist4(s, v.y); hist4(s, v.z); hist4(s, v.w);
    }
    for (uint32_t i = n16 * 16 + tid; i < n; i += kET) atomicAdd(&s.hist[src[i]], 1u);
  } else {
    for (uint32_t i = tid; i < n; i += kET) atomicAdd(&s.hist[src[i]], 1u);
  }
  __syncthreads();
  const uint32_t cnt = s.hist[tid];
  const int nz = __syncthreads_count(cnt != 0);

  // ---- literal code lengths ------------------------------------------------------------------
  if (nz >= 2) {
    for (uint32_t floor_ = 1;; floor_ <<= 1) {
      s.keys[tid] = cnt ? (((cnt > floor_ ? cnt : floor_) << 8) | tid) : 0xffffffffu;
      if (tid == 0) s.maxd = 0;
      __syncthreads();
      // bitonic sort of 256 keys, ascending
      for (uint32_t k = 2; k <= 256; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
          const uint32_t ixj = tid ^ j;
          if (ixj > tid) {
            const uint32_t a = s.keys[tid], b = s.keys[ixj];
            const bool up = (tid & k) == 0;
            if ((a > b) == up) { s.keys[tid] = b; s.keys[ixj] = a; }
          }
          __syncthreads();
        }
      if (tid < (uint32_t)nz) s.wt[tid] = s.keys[tid] >> 8;
      __syncthreads();
      if (tid == 0) {
        const int m = nz;
        int li = 0, ni = m, nn = m;
        for (int k = 0; k + 1 < m; k++) {
          int a, b;
          if (li < m && (ni >= nn || s.wt[li] <= s.wt[ni])) a = li++; else a = ni++;
          if (li < m && (ni >= nn || s.wt[li] <= s.wt[ni])) b = li++; else b = ni++;
          s.wt[nn] = s.wt[a] + s.wt[b];
          s.parent[a] = s.parent[b] = (uint16_t)nn;
          nn++;
        }
      }
      s.len[tid] = 0;
      __syncthreads();
      if (tid < (uint32_t)nz) {
        uint32_t d = 0, k = tid;
        const uint32_t root = 2 * (uint32_t)nz - 2;
        while (k != root) { k = s.parent[k]; d++; }
        s.len[s.keys[tid] & 255u] = (uint8_t)d;
        atomicMax(&s.maxd, d);
      }
      __syncthreads();
      const uint32_t maxd = s.maxd;
      __syncthreads();
      if (maxd <= 15) break;
    }
    // canonical code (RFC 7932 3.2), bit-reversed for the LSB-first writer
    const uint32_t L = s.len[tid];
    if (L) atomicAdd(&s.bl_count[L], 1u);
    __syncthreads();
    if (tid == 0) {
      uint32_t code = 0;
      s.next_code[0] = 0;
      for (int b = 1; b <= 15; b++) {
        code = (code + s.bl_count[b - 1]) << 1;
        s.next_code[b] = code;
      }
    }
    __syncthreads();
    uint32_t rank = 0;
    for (uint32_t j = 0; j < tid; j++) rank += (s.len[j] == L);
    s.lc[tid] = L ? ((__brev(s.next_code[L] + rank) >> (32 - L)) | (L << 16)) : 0u;
  } else {
    s.len[tid] = 0;
    s.lc[tid] = 0;
  }
  __syncthreads();

  // ---- sizes: literal bits, last used symbol ---------------------------------------------------
  uint32_t lit_bits;
  block_excl_scan(s, cnt * s.len[tid], &lit_bits);
  const uint32_t mylen = s.len[tid];
  if (tid == 0) s.misc = 0;
  __syncthreads();
  if (cnt) atomicMax(&s.misc, tid);
  __syncthreads();
  const uint32_t last_sym = s.misc;     // largest used symbol
  if (nz >= 2 && tid <= last_sym) atomicAdd(&s.cl_hist[mylen], 1u);
  __syncthreads();

  // ---- header, written by thread 0 up to the literal code lengths -----------------------------
  uint32_t ic = 0;
  for (int i = 23; i >= 0; i--)
    if (kInsBase[i] <= n) { ic = (uint32_t)i; break; }
  if (tid == 0) {
    uint32_t pos = 0;
    if (first) put_bits(s, pos, 0, 1);                 // WBITS = 16
    // the directory: metadata meta-block (ISLAST 0, MNIBBLES coded 3, reserved 0, MSKIPBYTES 1, MSKIPLEN - 1), padded
    // to two bytes; its payload is filled in as the numbers become known
    put_bits(s, pos, 0, 1); put_bits(s, pos, 3, 2); put_bits(s, pos, 0, 1); put_bits(s, pos, 1, 2);
    put_bits(s, pos, kDirBytes - 1, 8);
    dir_put(s, 0, 0x46u | (0x44u << 8) | (1u << 16), 3);                 // 'F' 'D' version 1
    dir_put(s, 7, n - 1, 2);
    pos = kDirBits;                                    // the compressed meta-block starts on the next byte
    put_bits(s, pos, 0, 1);                            // ISLAST = 0
    const uint32_t nib = (n - 1) < (1u << 16) ? 4 : (n - 1) < (1u << 20) ? 5 : 6;
    put_bits(s, pos, nib - 4, 2);
    put_bits(s, pos, n - 1, 4 * nib);                  // MLEN - 1
    s.misc = pos;                                      // position of the ISUNCOMPRESSED bit
    put_bits(s, pos, 0, 1);                            // ISUNCOMPRESSED = 0
    put_bits(s, pos, 0, 3);                            // NBLTYPESL = NBLTYPESI = NBLTYPESD = 1
    put_bits(s, pos, 0, 6);                            // NPOSTFIX = 0, NDIRECT = 0
    put_bits(s, pos, 0, 2);                            // context mode of literal block type 0
    put_bits(s, pos, 0, 2);                            // NTREESL = NTREESD = 1
    if (nz < 2) {
      put_bits(s, pos, 1, 2); put_bits(s, pos, 0, 2); put_bits(s, pos, last_sym, 8);   // simple code, NSYM = 1
    } else {
      // the code length code: Huffman (depth <= 5) over the code lengths 0..15 in use
      uint8_t cl_len[18];
      uint32_t used = 0, only = 0;
      for (int i = 0; i < 18; i++)
        if (s.cl_hist[i]) { used++; only = (uint32_t)i; }
      int to_store = 18;
      if (used == 1) {
        for (int i = 0; i < 18; i++) cl_len[i] = 0;
        cl_len[only] = 1;        // stored as 1; a single code length costs 0 bits per symbol
        for (int i = 0; i < 18; i++) s.cl_lc[i] = 0;
      } else {
        small_huffman(s.cl_hist, 18, 5, cl_len);
        while (cl_len[kClOrder[to_store - 1]] == 0) to_store--;
        uint32_t blc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, nc[8];
        for (int i = 0; i < 18; i++) blc[cl_len[i]]++;
        blc[0] = 0;
        uint32_t code = 0;
        nc[0] = 0;
        for (int b = 1; b <= 5; b++) { code = (code + blc[b - 1]) << 1; nc[b] = code; }
        for (int i = 0; i < 18; i++) {
          const uint32_t l = cl_len[i];
          s.cl_lc[i] = l ? ((__brev(nc[l]++) >> (32 - l)) | (l << 16)) : 0u;
        }
      }
      uint32_t skip = 0;
      if (cl_len[kClOrder[0]] == 0 && cl_len[kClOrder[1]] == 0) skip = cl_len[kClOrder[2]] == 0 ? 3 : 2;
      put_bits(s, pos, skip, 2);
      for (int i = (int)skip; i < to_store; i++) {
        const uint32_t v = cl_len[kClOrder[i]];
        put_bits(s, pos, kClclBits[v], kClclLen[v]);
      }
    }
    s.pos = pos;
  }
  __syncthreads();
  // the literal code lengths 0..last_sym, in parallel
  {
    const uint32_t e = (nz >= 2 && tid <= last_sym) ? s.cl_lc[mylen] : 0u;
    uint32_t total;
    const uint32_t o = block_excl_scan(s, e >> 16, &total);
    uint32_t pos = s.pos + o;
    put_bits(s, pos, e & 0xffffu, e >> 16);
    __syncthreads();
    if (tid == 0) {
      pos = s.pos + total;
      // insert-and-copy code: one symbol = (insert code ic, copy code 0); distance code: symbol 0 of 64
      const uint32_t cell = ic < 8 ? 128u : ic < 16 ? 256u : 448u;
      put_bits(s, pos, 1, 2); put_bits(s, pos, 0, 2); put_bits(s, pos, cell + ((ic & 7u) << 3), 10);
      put_bits(s, pos, 1, 2); put_bits(s, pos, 0, 2); put_bits(s, pos, 0, 6);
      put_bits(s, pos, n - kInsBase[ic], kInsExtra[ic]);     // the one command's insert extra bits
      s.lit_start = pos;
    }
    __syncthreads();
  }
  const uint32_t lit_start = s.lit_start;
  const uint32_t comp_bytes = (lit_start - kDirBits + lit_bits + 6 + 7) / 8;
  const bool raw = comp_bytes > n + 4;       // Huffman coding does not pay: uncompressed meta-block

  uint32_t bytes;
  uint8_t* dst = p.scratch + (uint64_t)c * kEntropyChunkCap;
  if (!raw) {
    // directory: the code lengths as nibbles, the kind
    if (tid < 128) dir_put(s, kDirLengths + tid, (uint32_t)s.len[2 * tid] | ((uint32_t)s.len[2 * tid + 1] << 4), 1);
    if (tid == 0 && nz < 2) { dir_put(s, 3, kKindConstant, 1); dir_put(s, 9, last_sym, 1); }
    // ---- literals: thread t codes the bytes [t S, (t+1) S); S is a power of two so that the decoder's spans of
    //      kEntropySpan bytes start where a thread starts ------------------------------------------
    uint32_t S = 16;
    while (S * kET < n) S <<= 1;
    const uint32_t b0 = tid * S < n ? tid * S : n, b1 = b0 + S < n ? b0 + S : n;
    uint32_t mybits = 0;
    if (aligned) {
      uint32_t i = b0;
      for (; i + 16 <= b1; i += 16) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + i));
        const uint32_t ww[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++)
          mybits += s.len[ww[k] & 255u] + s.len[(ww[k] >> 8) & 255u] + s.len[(ww[k] >> 16) & 255u] + s.len[ww[k] >> 24];
      }
      for (; i < b1; i++) mybits += s.len[src[i]];
    } else {
      for (uint32_t i = b0; i < b1; i++) mybits += s.len[src[i]];
    }
    uint32_t total;
    const uint32_t o = block_excl_scan(s, mybits, &total);
    const uint32_t bitpos = lit_start + o;
    if (b0 < n && b0 % kEntropySpan == 0) dir_put(s, kDirSpans + 3 * (b0 / kEntropySpan), bitpos, 3);
    // after the literals: an empty metadata meta-block (ISLAST = 0, MNIBBLES = 0 coded 3, reserved 0, MSKIPBYTES = 0)
    // pads to a byte; then, for a plane's last chunk, 0x03 (ISLAST, ISLASTEMPTY)
    const uint32_t pad_pos = lit_start + lit_bits;
    const uint32_t nbytes = (pad_pos + 6 + 7) / 8;
    bytes = nbytes + (last ? 1u : 0u);
    if (tid == 0) dir_put(s, 4, nbytes, 3);                 // chunk bytes (without the stream's final 0x03)
    const uint32_t my_w0 = bitpos >> 5, my_w1 = (bitpos + mybits + 31) >> 5;   // words this thread's bits touch
    for (uint32_t lo = 0; lo * 4 < bytes; lo += kWinWords) {
      const uint32_t hi = lo + kWinWords;
      if (lo > 0) {
        for (uint32_t i = tid; i < kWinWords; i += kET) s.out[i] = 0;
        __syncthreads();
      }
      auto or_word = [&](uint32_t w, uint32_t v) {
        if (w >= lo && w < hi) atomicOr(&s.out[w - lo], v);
      };
      if (mybits && my_w0 < hi && my_w1 > lo) {
        uint32_t wi = bitpos >> 5, nb = bitpos & 31u;
        unsigned long long acc = 0;
        auto emit = [&](uint32_t byte) {
          const uint32_t e = s.lc[byte];
          acc |= (unsigned long long)(e & 0xffffu) << nb;
          nb += e >> 16;
          if (nb >= 32) {
            or_word(wi, (uint32_t)acc);
            wi++;
            acc >>= 32;
            nb -= 32;
          }
        };
        if (aligned) {
          uint32_t i = b0;
          for (; i + 16 <= b1; i += 16) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + i));
            const uint32_t ww[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
              emit(ww[k] & 255u); emit((ww[k] >> 8) & 255u); emit((ww[k] >> 16) & 255u); emit(ww[k] >> 24);
            }
          }
          for (; i < b1; i++) emit(src[i]);
        } else {
          for (uint32_t i = b0; i < b1; i++) emit(src[i]);
        }
        if (nb) or_word(wi, (uint32_t)acc);
      }
      if (tid == 0) {
        // the six bits of the empty metadata block: 0, 11, 0, 00 = value 6 at pad_pos (may straddle two words)
        const uint32_t w = pad_pos >> 5, sh = pad_pos & 31u;
        or_word(w, 6u << sh);
        if (sh + 6 > 32) or_word(w + 1, 6u >> (32 - sh));
        if (last) or_word(nbytes >> 2, 0x03u << (8 * (nbytes & 3u)));
      }
      __syncthreads();
      const uint32_t w_end = (bytes + 3) / 4 < hi ? (bytes + 3) / 4 : hi;
      for (uint32_t i = lo + tid; i < w_end; i += kET) reinterpret_cast<uint32_t*>(dst)[i] = s.out[i - lo];
      __syncthreads();
    }
  } else {
    // ---- uncompressed meta-block: header up to ISUNCOMPRESSED = 1, pad, raw bytes -------------
    const uint32_t hdr_bits = s.misc + 1 - kDirBits;   // bits of the meta-block header up to ISUNCOMPRESSED
    const uint32_t hb = (hdr_bits + 7) / 8;
    __syncthreads();
    if (tid == 0) {
      dir_put(s, 3, kKindRaw, 1);
      dir_put(s, 4, kDirBlock + hb + n, 3);
      dir_put(s, kDirSpans, 8 * (kDirBlock + hb), 3);     // where the raw bytes start
    }
    __syncthreads();
    // the directory block as it is; of the compressed header everything after the ISUNCOMPRESSED bit is dropped
    for (uint32_t i = tid; i < kDirBlock; i += kET) dst[i] = reinterpret_cast<const uint8_t*>(s.out)[i];
    if (tid == 0) {
      const uint32_t pos = s.misc - kDirBits;             // kDirBits is a multiple of 32: the header is in one word
      uint32_t w0 = s.out[kDirBits / 32] & ((1u << pos) - 1u);
      w0 |= 1u << pos;                                    // hdr_bits <= 1 + 2 + 24 + 1 = 28
      for (uint32_t i = 0; i < hb; i++) dst[kDirBlock + i] = (uint8_t)(w0 >> (8 * i));
      if (last) dst[kDirBlock + hb + n] = 0x03;
    }
    for (uint32_t i = tid; i < n; i += kET) dst[kDirBlock + hb + i] = src[i];
    bytes = kDirBlock + hb + n + (last ? 1u : 0u);
  }
  if (tid == 0) p.chunk_bytes[c] = bytes;
}

// Frame sizes -> frame offsets (exclusive scan); one CTA.  frame_off[n] = total bytes.
__global__ void __launch_bounds__(1024) k_entropy_layout(const EntropyParams p, uint64_t* frame_off) {
  __shared__ uint64_t wsum[32];
  __shared__ uint64_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (uint32_t f0 = 0; f0 < p.n; f0 += blockDim.x) {
    const uint32_t f = f0 + threadIdx.x;
    uint64_t sz = 0;
    if (f < p.n) {
      sz = 11;     // total(4) kind(1) 1+|bp|(4) pflags(1) ... flags(1)
      for (uint32_t k = 0; k < p.cpf; k++) sz += p.chunk_bytes[(uint64_t)f * p.cpf + k];
    }
    uint64_t x = sz;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint64_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wsum[w] = x;
    __syncthreads();
    uint64_t base = carry_s;
    for (int i = 0; i < w; i++) base += wsum[i];
    if (f < p.n) frame_off[f] = base + x - sz;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry_s = base + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) frame_off[p.n] = carry_s;
}

// Copies every chunk to its place in the frame's container chunk and writes the container header
// (fusion_power_video.cc:820-846).
__global__ void __launch_bounds__(kET) k_entropy_gather(const EntropyParams p, const uint64_t* frame_off, uint8_t* out,
                                                        uint64_t capacity, uint32_t* overflow) {
  __shared__ uint64_t dst_s;
  const uint32_t c = blockIdx.x, f = c / p.cpf, r = c % p.cpf;
  const uint32_t* cb = p.chunk_bytes + (uint64_t)f * p.cpf;
  if (frame_off[p.n] > capacity) {
    if (c == 0 && threadIdx.x == 0) *overflow = 1;
    return;
  }
  if (threadIdx.x == 0) {
    uint64_t o = frame_off[f] + 10;
    uint32_t bp = 0, core = 0;
    for (uint32_t k = 0; k < p.cpf; k++) {
      if (k < r) o += cb[k];
      if (k < p.cpp) bp += cb[k]; else core += cb[k];
    }
    if (r >= p.cpp) o += 1;          // the core's flags byte sits between the preview stream and the low stream
    dst_s = o;
    if (r == 0) {
      uint8_t* h = out + frame_off[f];
      const uint32_t total = 11 + bp + core, fl = p.flags[f];
      const uint32_t bp1 = bp + 1;
      for (int i = 0; i < 4; i++) { h[i] = (uint8_t)(total >> (8 * i)); h[5 + i] = (uint8_t)(bp1 >> (8 * i)); }
      h[4] = 0;                                          // frame chunk
      h[9] = (uint8_t)((fl & kFlagCG) | kFlagNoLow);      // preview flags (.cc:842)
      h[10 + bp] = (uint8_t)fl;                          // core := flags | low? | high
    }
  }
  __syncthreads();
  const uint32_t nb = cb[r];
  const uint8_t* src = p.scratch + (uint64_t)c * kEntropyChunkCap;
  uint8_t* dst = out + dst_s;
  // destination alignment is arbitrary: bytes up to the first 4-byte boundary, then words assembled
  // from two aligned source words
  const uint32_t head = (uint32_t)((4 - (reinterpret_cast<uintptr_t>(dst) & 3u)) & 3u);
  if (nb <= head + 8) {
    for (uint32_t i = threadIdx.x; i < nb; i += kET) dst[i] = src[i];
    return;
  }
  if (threadIdx.x < head) dst[threadIdx.x] = src[threadIdx.x];
  const uint32_t words = (nb - head) / 4;
  const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
  uint32_t* d32 = reinterpret_cast<uint32_t*>(dst + head);
  const uint32_t sh = head * 8;
  for (uint32_t i = threadIdx.x; i < words; i += kET) {
    const uint32_t a = s32[i], b = sh ? s32[i + 1] : 0u;
    d32[i] = sh ? __funnelshift_r(a, b, sh) : a;
  }
  for (uint32_t i = head + words * 4 + threadIdx.x; i < nb; i += kET) dst[i] = src[i];
}

// =====================================================================================
// k_entropy_decode -- the inverse of k_entropy_chunk for streams that carry directories.
//
// One warp per chunk.  Only the directory meta-block and the literal bits are read: the code lengths come from
// the directory's nibbles (canonical code as in RFC 7932 3.2, the same construction k_entropy_chunk uses), a
// 10-bit lookup table per warp in shared memory resolves most symbols in one step (longer codes: canonical
// search over the lengths 11..15), and lane k starts at the directory's bit position of span k (2048 plane
// bytes), so all 32 lanes decode at once.  Raw and constant chunks are copied / filled.  Every read is bounded
// by the blob size and every write by the plane: a malformed chunk sets *err and is abandoned.
// =====================================================================================
namespace {

constexpr int kDecWarps = 4;
constexpr uint32_t kLutBits = 10;

struct DecWarp {
  uint16_t lut[1u << kLutBits];    // symbol | length << 8; 0: a code longer than kLutBits
  uint8_t len[256];
  uint8_t sorted[256];             // symbols ordered by (length, symbol)
  uint16_t first[16];              // canonical code of the first symbol of each length (MSB-first)
  uint16_t count[16];
  uint16_t offset[16];             // index of that symbol in sorted[]
};

__device__ __forceinline__ uint32_t load_u24(const uint8_t* p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
}

}  // namespace

__global__ void __launch_bounds__(32 * kDecWarps) k_entropy_decode(const EntropyDecodeParams p) {
  __shared__ DecWarp dsm[kDecWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t ci = blockIdx.x * kDecWarps + warp;
  if (ci >= p.n_chunks) return;
  DecWarp& d = dsm[warp];
  const CodedChunk ch = p.chunks[ci];
  auto bad = [&](uint32_t why) { if (lane == 0) atomicMax(p.err, why); };
  if (ch.offset > p.blob_bytes || p.blob_bytes - ch.offset < kDirBlock + 1) { bad(1); return; }
  const uint8_t* base = p.blob + ch.offset;
  const uint8_t* dir = base + 2;
  if (dir[0] != 0x46 || dir[1] != 0x44 || dir[2] != 1) { bad(2); return; }
  const uint32_t kind = dir[3], cb = load_u24(dir + 4), n = ((uint32_t)dir[7] | ((uint32_t)dir[8] << 8)) + 1u;
  const uint64_t plane_off = (uint64_t)ch.index * kEntropyChunk;
  if (cb < kDirBlock || cb > p.blob_bytes - ch.offset || ch.frame < p.frame0 || ch.frame - p.frame0 >= p.n_frames || ch.plane > 1 ||
      plane_off + n > p.P || (plane_off + n != p.P && n != kEntropyChunk)) { bad(3); return; }
  uint8_t* plane = ch.plane ? p.low : p.high;
  if (plane == nullptr) { bad(4); return; }
  uint8_t* dst = plane + (uint64_t)(ch.frame - p.frame0) * p.P + plane_off;
  const bool dst4 = (reinterpret_cast<uintptr_t>(dst) & 3u) == 0;

  if (kind == kKindConstant) {
    const uint32_t v = dir[9];
    for (uint32_t i = lane; i < n; i += 32) dst[i] = (uint8_t)v;
    return;
  }
  if (kind == kKindRaw) {
    const uint32_t o = load_u24(dir + kDirSpans) >> 3;
    if (o > cb || cb - o < n) { bad(5); return; }
    const uint8_t* src = base + o;
    if (dst4) {
      for (uint32_t i = 4 * lane; i + 4 <= n; i += 128)
        *reinterpret_cast<uint32_t*>(dst + i) = (uint32_t)src[i] | ((uint32_t)src[i + 1] << 8) | ((uint32_t)src[i + 2] << 16) |
                                                ((uint32_t)src[i + 3] << 24);
      for (uint32_t i = (n & ~3u) + lane; i < n; i += 32) dst[i] = src[i];
    } else {
      for (uint32_t i = lane; i < n; i += 32) dst[i] = src[i];
    }
    return;
  }
  if (kind != kKindHuffman) { bad(6); return; }

  // ---- the canonical code from the directory's code lengths ------------------------------------------
  {
    const uint32_t w = (uint32_t)dir[kDirLengths + 4 * lane] | ((uint32_t)dir[kDirLengths + 4 * lane + 1] << 8) |
                       ((uint32_t)dir[kDirLengths + 4 * lane + 2] << 16) | ((uint32_t)dir[kDirLengths + 4 * lane + 3] << 24);
#pragma unroll
    for (int j = 0; j < 8; j++) d.len[8 * lane + j] = (uint8_t)((w >> (4 * j)) & 15u);
  }
  __syncwarp();
  if (lane == 0) {
    uint32_t cnt[16];
    for (int i = 0; i < 16; i++) cnt[i] = 0;
    for (int i = 0; i < 256; i++) cnt[d.len[i]]++;
    cnt[0] = 0;
    uint32_t code = 0, off = 0, kraft = 0;
    for (int b = 1; b <= 15; b++) {
      code = (code + cnt[b - 1]) << 1;
      d.first[b] = (uint16_t)code;
      d.count[b] = (uint16_t)cnt[b];
      d.offset[b] = (uint16_t)off;
      off += cnt[b];
      kraft += cnt[b] << (15 - b);
    }
    d.first[0] = d.count[0] = d.offset[0] = 0;
    uint32_t next[16];
    for (int b = 0; b < 16; b++) next[b] = d.offset[b];
    for (int i = 0; i < 256; i++) {
      const uint32_t l = d.len[i];
      if (l) d.sorted[next[l]++] = (uint8_t)i;
    }
    if (kraft != (1u << 15)) d.count[0] = 1;       // not a complete prefix code
  }
  __syncwarp();
  if (d.count[0]) { bad(7); return; }
  // canonical decode of the low bits of `v` (stream order: the code's first bit is bit 0), lengths lo..hi
  auto search = [&](uint32_t v, int lo, int hi, uint32_t& sym, uint32_t& len) -> bool {
    const uint32_t r = __brev(v);
    for (int L = lo; L <= hi; L++) {
      const uint32_t c = (r >> (32 - L)) - d.first[L];
      if (c < d.count[L]) { sym = d.sorted[d.offset[L] + c]; len = (uint32_t)L; return true; }
    }
    return false;
  };
  for (uint32_t idx = lane; idx < (1u << kLutBits); idx += 32) {
    uint32_t sym = 0, len = 0;
    d.lut[idx] = search(idx, 1, (int)kLutBits, sym, len) ? (uint16_t)(sym | (len << 8)) : (uint16_t)0;
  }
  __syncwarp();

  // ---- lane k decodes span k -------------------------------------------------------------------------
  const uint32_t k = (uint32_t)lane;
  if ((uint64_t)k * kEntropySpan >= n) return;
  uint32_t count = n - k * kEntropySpan < kEntropySpan ? n - k * kEntropySpan : kEntropySpan;
  const uint32_t pos = load_u24(dir + kDirSpans + 3 * k);
  if (pos < kDirBits || pos > 8 * cb) { bad(8); return; }
  // aligned 32-bit reads of the blob; nothing past its last word is touched
  const uintptr_t b0 = reinterpret_cast<uintptr_t>(base);
  const uint32_t* words = reinterpret_cast<const uint32_t*>(b0 & ~(uintptr_t)3);
  const uint64_t abs_bit = (uint64_t)(b0 & 3u) * 8 + pos;
  uint64_t wi = abs_bit >> 5;
  const uint64_t last_word = (uint64_t)((reinterpret_cast<uintptr_t>(p.blob) + p.blob_bytes - 1 - (b0 & ~(uintptr_t)3)) >> 2);
  auto next_word = [&]() -> uint32_t {
    const uint32_t v = wi <= last_word ? __ldg(words + wi) : 0u;
    wi++;
    return v;
  };
  unsigned long long buf = (unsigned long long)next_word() >> (abs_bit & 31u);
  uint32_t nbits = 32 - (uint32_t)(abs_bit & 31u);
  uint8_t* o = dst + (uint64_t)k * kEntropySpan;
  uint32_t acc = 0, na = 0;
  const bool vec = (reinterpret_cast<uintptr_t>(o) & 15u) == 0;
  uint32_t q[4];
  uint32_t nq = 0;
  while (count) {
    if (nbits < 32) { buf |= (unsigned long long)next_word() << nbits; nbits += 32; }
    uint32_t sym, len;
    const uint32_t e = d.lut[(uint32_t)buf & ((1u << kLutBits) - 1u)];
    if (e) { sym = e & 255u; len = e >> 8; }
    else if (!search((uint32_t)buf, (int)kLutBits + 1, 15, sym, len)) { atomicMax(p.err, 9u); return; }
    buf >>= len;
    nbits -= len;
    count--;
    acc |= sym << (8 * na);
    if (++na == 4) {
      if (vec) {
        q[nq++] = acc;
        if (nq == 4) { *reinterpret_cast<uint4*>(o) = make_uint4(q[0], q[1], q[2], q[3]); o += 16; nq = 0; }
      } else if (dst4) {
        *reinterpret_cast<uint32_t*>(o) = acc; o += 4;
      } else {
        o[0] = (uint8_t)acc; o[1] = (uint8_t)(acc >> 8); o[2] = (uint8_t)(acc >> 16); o[3] = (uint8_t)(acc >> 24); o += 4;
      }
      acc = 0; na = 0;
    }
  }
  // tails: whole words still queued, then single bytes
  for (uint32_t i = 0; i < nq; i++) { *reinterpret_cast<uint32_t*>(o) = q[i]; o += 4; }
  for (uint32_t i = 0; i < na; i++) o[i] = (uint8_t)(acc >> (8 * i));
  // the directory is redundant with the bit stream: a span must end where the next one starts (a damaged
  // directory would otherwise decode to garbage without anyone noticing), the last one inside the chunk
  const uint64_t end_pos = wi * 32 - nbits - (uint64_t)(b0 & 3u) * 8;
  const bool last_span = (uint64_t)(k + 1) * kEntropySpan >= n;
  if (last_span ? end_pos > 8ull * cb : end_pos != load_u24(dir + kDirSpans + 3 * (k + 1))) atomicMax(p.err, 10u);
}

int enqueue_entropy_decode(const EntropyDecodeParams& p, cudaStream_t stream, cudaError_t* err) {
  if (p.n_chunks == 0) { *err = cudaSuccess; return 0; }
  k_entropy_decode<<<(p.n_chunks + kDecWarps - 1) / kDecWarps, 32 * kDecWarps, 0, stream>>>(p);
  *err = cudaGetLastError();
  return *err == cudaSuccess ? 1 : -1;
}

size_t entropy_smem_bytes() { return sizeof(Shared); }

int enqueue_entropy(const EntropyParams& p, uint64_t* frame_off, uint8_t* out, uint64_t capacity, uint32_t* overflow,
                    cudaStream_t stream, cudaError_t* err) {
  // per launch: the attribute is per device and does not survive cudaDeviceReset
  *err = cudaFuncSetAttribute(k_entropy_chunk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Shared));
  if (*err != cudaSuccess) return -1;
  const uint32_t chunks = p.n * p.cpf;
  k_entropy_chunk<<<chunks, kET, sizeof(Shared), stream>>>(p);
  k_entropy_layout<<<1, 1024, 0, stream>>>(p, frame_off);
  k_entropy_gather<<<chunks, kET, 0, stream>>>(p, frame_off, out, capacity, overflow);
  *err = cudaGetLastError();
  return *err == cudaSuccess ? 3 : -1;
}

}  // namespace fpv

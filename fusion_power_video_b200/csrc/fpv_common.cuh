// fpv_common.cuh -- device-side building blocks shared by all kernels.
//
// Citations are file:line into the reference's fusion_power_video.cc.
//
// Register packing conventions.  One 32-bit register always carries TWO pixels
// in its two 16-bit lanes (pixel 2j in bits 0..15, pixel 2j+1 in bits 16..31):
//   "lane form"   each lane holds a byte value in [0,255]           (generic kernels)
//   "q form"      each lane holds (high << 8) | low, i.e. the left-aligned
//                 pixel the reference splits into its two planes     (fast kernel)
//   "S form"      each lane holds (byte << 8) | junk, junk <= 1      (fast kernel)
// sm_100a has native packed-u16 min/max, also three-input (VIMNMX3.U16x2), but
// only multi-instruction emulations of packed-u8 arithmetic; 16-bit lanes let
// plain 32-bit IADD3 do the byte arithmetic of two pixels with the low byte of
// each lane acting as a guard against cross-lane carries.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace fpv {

constexpr uint32_t kLaneMask = 0x00ff00ffu;
constexpr uint32_t kLaneBias = 0x01000100u;  // +256 per lane: keeps lane subtractions non-negative

constexpr int kFlagDelta = 1;       // fusion_power_video.h:68-73
constexpr int kFlagCG = 2;
constexpr int kFlagNoLow = 4;

// The six loop variants of Frame::Frame(u16) (.cc:388-445) plus the
// shift > 8 little-endian case of the generic branch.
enum SplitMode : int {
  kLE0 = 0,    // .cc:421-427
  kLE8 = 1,    // .cc:431-433  (no low plane)
  kLEs = 2,    // .cc:437-443, 1 <= shift <= 7
  kBE0 = 3,    // .cc:391-397
  kBE8 = 4,    // .cc:401-403  (no low plane)
  kBEs = 5,    // .cc:407-415, 1 <= shift <= 7
  kLEbig = 6,  // .cc:437-443, 9 <= shift <= 16
  kNumSplitModes = 7
};

__host__ __device__ inline bool mode_has_low(int mode) { return mode != kLE8 && mode != kBE8; }

// Host helper: which variant handles (shift, big_endian); -1 if the reference
// itself is undefined there (big-endian generic branch shifts by 8 - shift,
// negative for shift > 8, .cc:407-412).
inline int pick_split_mode(int shift, int big_endian) {
  if (shift < 0 || shift > 16) return -1;
  if (big_endian) {
    if (shift == 0) return kBE0;
    if (shift == 8) return kBE8;
    if (shift < 8) return kBEs;
    return -1;
  }
  if (shift == 0) return kLE0;
  if (shift == 8) return kLE8;
  if (shift < 8) return kLEs;
  return kLEbig;
}

// ---- scalar forms (one pixel) ------------------------------------------------

// p = the uint16 as loaded natively.  Returns high, low in [0,255].
template <int MODE>
__device__ __forceinline__ void split1(uint32_t p, int s, uint32_t& h, uint32_t& l) {
  if (MODE == kLE0) { h = p >> 8; l = p & 0xffu; }
  else if (MODE == kLE8) { h = p & 0xffu; l = 0; }
  else if (MODE == kLEs || MODE == kLEbig) { uint32_t q = (p << s) & 0xffffu; h = q >> 8; l = q & 0xffu; }
  else if (MODE == kBE0) { h = p & 0xffu; l = p >> 8; }
  else if (MODE == kBE8) { h = p >> 8; l = 0; }
  else { h = ((p << s) | (p >> (16 - s))) & 0xffu; l = (p >> (8 - s)) & 0xffu; }
}

// .cc:247-252 on values already in [0,255]:
// CG = n + w - clamp(nw, min(n,w), max(n,w)).
__device__ __forceinline__ uint32_t cg1(uint32_t n, uint32_t w, uint32_t nw) {
  uint32_t lo = min(n, w), hi = max(n, w);
  return n + w - min(max(nw, lo), hi);
}

// ---- lane forms (two pixels per register) -----------------------------------

// x = two raw uint16 pixels as loaded (pixel 0 in the low half).
template <int MODE>
__device__ __forceinline__ void split2(uint32_t x, int s, uint32_t& h2, uint32_t& l2) {
  if (MODE == kLE0) {
    h2 = __byte_perm(x, 0u, 0x4341);           // (x >> 8) & M
    l2 = x & kLaneMask;
  } else if (MODE == kLE8) {
    h2 = x & kLaneMask; l2 = 0;
  } else if (MODE == kLEs) {
    h2 = (x >> (8 - s)) & kLaneMask;           // ((p << s) >> 8) & 0xff
    l2 = (x & ((0xffu >> s) * 0x00010001u)) << s;
  } else if (MODE == kLEbig) {
    int t = s - 8;                             // 1..8
    h2 = (x & ((0xffu >> t) * 0x00010001u)) << t;
    l2 = 0;                                    // (p << s) & 0xff == 0 for s >= 8
  } else if (MODE == kBE0) {
    h2 = x & kLaneMask;
    l2 = __byte_perm(x, 0u, 0x4341);
  } else if (MODE == kBE8) {
    h2 = __byte_perm(x, 0u, 0x4341); l2 = 0;
  } else {  // kBEs
    uint32_t t1 = (x & ((0xffu >> s) * 0x00010001u)) << s;
    uint32_t t2 = (x >> (16 - s)) & (((1u << s) - 1u) * 0x00010001u);
    h2 = t1 | t2;
    l2 = (x >> (8 - s)) & kLaneMask;
  }
}

// Delta planes live on the device in "image form": one uint16 per pixel,
// (delta_high << 8) | delta_low, i.e. the decoded delta image itself
// (.cc:337-338 reads it exactly this way).
__device__ __forceinline__ void split2_delta(uint32_t d, uint32_t& dh2, uint32_t& dl2) {
  dh2 = __byte_perm(d, 0u, 0x4341);
  dl2 = d & kLaneMask;
}

// Packed u16x2 min / max: single VIMNMX.U16x2 on sm_100a.
__device__ __forceinline__ uint32_t vmin2(uint32_t a, uint32_t b) { return __vminu2(a, b); }
__device__ __forceinline__ uint32_t vmax2(uint32_t a, uint32_t b) { return __vmaxu2(a, b); }

// ClampedGradient on two pixels.  Inputs are lane-form values in [0,255];
// result is in [0,255] per lane (n + w - clamp >= 0 and <= 255 per lane, so the
// 32-bit add/sub never carries across lanes).
__device__ __forceinline__ uint32_t cg2(uint32_t n2, uint32_t w2, uint32_t nw2) {
  uint32_t lo = vmin2(n2, w2), hi = vmax2(n2, w2);
  uint32_t c = vmin2(vmax2(nw2, lo), hi);
  return n2 + w2 - c;
}

// (a - b) mod 256 per lane, for lane-form a, b in [0,255].
__device__ __forceinline__ uint32_t sub2(uint32_t a2, uint32_t b2) {
  return (a2 + kLaneBias - b2) & kLaneMask;
}

// Packs four lane-form registers (8 pixels) into 8 bytes.
__device__ __forceinline__ uint2 pack8(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  return make_uint2(__byte_perm(a, b, 0x6420), __byte_perm(c, d, 0x6420));
}

// ClampedGradient residual of four consecutive preview pixels (one word) on the preview's flat array (.cc:575-586):
// c = the word, cm = the word before it in flat order (only its last pixel is used), n / nm = the same one row up.
__device__ __forceinline__ uint32_t finalize_word(uint32_t c, uint32_t cm, uint32_t n, uint32_t nm) {
  const uint32_t wv = __funnelshift_l(cm, c, 8);    // bytes i-1 .. i+2
  const uint32_t nwv = __funnelshift_l(nm, n, 8);
  // lane form: pixels (0,1) and (2,3)
  const uint32_t c01 = __byte_perm(c, 0u, 0x4140), c23 = __byte_perm(c, 0u, 0x4342);
  const uint32_t n01 = __byte_perm(n, 0u, 0x4140), n23 = __byte_perm(n, 0u, 0x4342);
  const uint32_t w01 = __byte_perm(wv, 0u, 0x4140), w23 = __byte_perm(wv, 0u, 0x4342);
  const uint32_t q01 = __byte_perm(nwv, 0u, 0x4140), q23 = __byte_perm(nwv, 0u, 0x4342);
  const uint32_t r01 = sub2(c01, cg2(n01, w01, q01)), r23 = sub2(c23, cg2(n23, w23, q23));
  return __byte_perm(r01, r23, 0x6420);
}

// ---- q form / S form (fast encode kernel) -------------------------------------

constexpr uint32_t kHiBytes = 0xff00ff00u;
constexpr uint32_t kLoBytes = 0x00ff00ffu;

// Loop-invariant constants of make_q2 for one (mode, shift).
struct QConst {
  uint32_t sh_a, sh_b;   // shift amounts
  uint32_t m_a, m_b;     // masks
};

__host__ __device__ inline QConst make_qconst(int mode, int s) {
  QConst c = {0, 0, 0, 0};
  if (mode == kLE8 || mode == kLEs || mode == kLEbig) {
    c.sh_a = (uint32_t)s;                                      // q = (x << s) & m_a
    c.m_a = ((0xffffu << s) & 0xffffu) * 0x00010001u;
    c.sh_b = s >= 32 ? 0u : 1u << s;                           // the same shift as a multiplier (make_hs_lo)
    c.m_b = c.m_a & 0xff00ff00u;                               // mask of the high bytes of q
  } else if (mode == kBEs) {
    c.sh_a = (uint32_t)(s + 8);                                // ((x & m_a) << (s + 8))
    c.m_a = (0xffu >> s) * 0x00010001u;
    c.sh_b = (uint32_t)(8 - s);                                // | ((x >> (8 - s)) & m_b)
    c.m_b = ((((1u << s) - 1u) << 8) | 0xffu) * 0x00010001u;
  }
  return c;
}

// Two raw uint16 pixels as loaded -> q form: per lane (high << 8) | low with
// exactly the high / low bytes of the six reference expressions (.cc:388-445;
// table in SURVEY.md 9.1).  For the two no-low modes the low byte is 0.
template <int MODE>
__device__ __forceinline__ uint32_t make_q2(uint32_t x, const QConst& c) {
  if (MODE == kLE0) return x;                                           // high = p >> 8, low = p & 0xff
  if (MODE == kBE0) return __byte_perm(x, 0u, 0x2301);                  // high = p & 0xff, low = p >> 8
  if (MODE == kBE8) return x & kHiBytes;                                // high = (p >> 8) & 0xff
  if (MODE == kBEs)                                                     // high = ((p << s) | (p >> (16 - s))) & 0xff
    return ((x & c.m_a) << c.sh_a) | ((x >> c.sh_b) & c.m_b);           // low  = (p >> (8 - s)) & 0xff
  return (x << c.sh_a) & c.m_a;                                         // q = (uint16_t)(p << s), s = 1..16
}

// Bytes 1 and 3 of a and of b -> 4 consecutive output bytes (S form -> plane).
__device__ __forceinline__ uint32_t pack_hi(uint32_t a, uint32_t b) { return __byte_perm(a, b, 0x7531); }
// Bytes 0 and 2 of a and of b.
__device__ __forceinline__ uint32_t pack_lo(uint32_t a, uint32_t b) { return __byte_perm(a, b, 0x6420); }

// ClampedGradient residual of two pixels in S form:
//   CG(n, w, nw) = n + w - median(n, w, nw) = min3 + max3 - nw          (.cc:247-252)
//   residual     = h - CG = h + nw - min3 - max3                        (.cc:570)
// The +0x0080 per lane keeps the (junk) low bytes in [124, 131] whatever the
// cross-lane carries are, so bytes 1 and 3 of the result are exact mod 256.
__device__ __forceinline__ uint32_t cg_residual_s(uint32_t h, uint32_t n, uint32_t w, uint32_t nw) {
  const uint32_t mn = __vimin3_u16x2(n, w, nw), mx = __vimax3_u16x2(n, w, nw);
  uint32_t t = h + nw + 0x00800080u;   // one IADD3 ...
  asm("" : "+r"(t));                   // (keeps the compiler from re-associating into three adds)
  return t - mn - mx;                  // ... and a second one
}

// ---- integer entropy heuristic ------------------------------------------------

// Per-frame device-side record shared by every encode kernel.
struct FrameStat {
  uint32_t hist_d[256];  // raw high bytes at i % 15 == 0            (.cc:526-531)
  uint32_t hist_a[256];  // post-delta high at i = W+1+31k           (.cc:554-562)
  uint32_t hist_b[256];  // same samples minus ClampedGradient
  uint32_t low_or;       // OR of all split low bytes                (.cc:447-449)
  uint32_t assumed;      // flags (bit0 delta, bit1 cg) the last transform pass assumed
  uint32_t final_flags;  // decided FrameFlags byte
  uint32_t done;         // 1 once outputs in memory match final_flags
  // Fast path: instead of hist_d, the number of delta-decision samples whose
  // raw high byte has bit k set.  Enough to decide USE_DELTA in all but
  // near-constant frames (see k_decide), with no shared-memory atomics.
  uint32_t dbits[8];
  uint32_t delta_known;  // USE_DELTA decision already taken (redo passes keep it)            (generic path)
  uint32_t delta_dec;
  uint32_t tickets;      // fast path: warps that have published their part of the frame in this pass
  uint32_t pad_;
};

// EstimateEntropy (.cc:235-244) evaluated by one 256-thread block, one bin per
// thread.  Returns floor(1024 * S / sum) with the reference's `int`
// accumulator truncation: sums are formed mod 2^32 and reinterpreted as int32.
// `red` is 8 words of shared scratch.  All threads receive the result.
__device__ __forceinline__ uint64_t block_entropy256(uint32_t v, uint32_t* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // pass 1: sum
  uint32_t s = v;
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __syncthreads();
  if (lane == 0) red[wid] = s;
  __syncthreads();
  uint32_t sum32 = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) sum32 += red[i];
  int64_t sum = (int64_t)(int32_t)sum32;
  if (sum == 0) return 0;  // uniform across the block
  // approxLog2(sum) on the sign-extended 64-bit value
  int log2sum = 63 - __clzll((long long)sum);
  // pass 2: acc -= v * (log2(v) - log2sum), v == 0 contributes 0
  uint32_t term = 0;
  if (v) term = v * (uint32_t)(log2sum - (31 - __clz(v)));
#pragma unroll
  for (int o = 16; o; o >>= 1) term += __shfl_xor_sync(0xffffffffu, term, o);
  __syncthreads();
  if (lane == 0) red[wid] = term;
  __syncthreads();
  uint32_t acc32 = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) acc32 += red[i];
  uint64_t sum_of_logs = (uint64_t)(int64_t)(int32_t)acc32;
  return (1024ull * sum_of_logs) / (uint64_t)sum;
}

}  // namespace fpv

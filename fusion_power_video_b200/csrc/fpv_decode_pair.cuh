// fpv_decode_pair.cuh -- the warp-specialised decode kernel (k_decode_pair): round 1's default, since round 2 the
// fallback behind k_decode_fused (fpv_decode_fused.cuh), which reuses this file's chain_step, pair_ddup_word and
// PairParams.  FPV_DECODE_KERNEL=pair selects it; the parity tests run it on every geometry.
//
// Same job as the rest of fpv_decode.cu: the post-brotli part of DecompressImage
// (fusion_power_video.cc:326-344) fused with UnextractFrame (.cc:850-862).
//
// The inverse ClampedGradient (.cc:327-332) is a serial chain in flat pixel
// order, so rows of a frame are processed strictly one after the other and the
// only freedom is (a) frames are independent, (b) a row can be cut into
// segments whose incoming west value is speculated and repaired (see the top of
// fpv_decode.cu).  Measured on plasma-like frames a wrong incoming value heals
// with a probability of only ~0.23 per pixel, so short segments pay a full
// second pass; this kernel therefore uses
//
//   * TWO PAIRS OF FRAMES PER CTA, six warps; per pair:
//       chain warp  Lane l owns the segment of L = 8*LW2 columns [l*L, (l+1)*L) of BOTH frames: frame A
//                   in the low 16-bit lane of a register, frame B in the high one ("pair form", byte
//                   values in [0,255]).  One step for both frames is
//                       x = (c + min3(n,w,nw) + max3(n,w,nw)) & 0x00ff00ff
//                   with c = r + 256 - nw, because CG = n + w - median(n,w,nw)
//                   = min3 + max3 - nw (.cc:247-252): 2 VIMNMX3.U16x2 + IADD3 +
//                   LOP3.  32 segments per frame (instead of 64) are long enough
//                   for a cheap look-ahead: pass 0 runs only the last K0 pixels
//                   of every segment from a guess and hands the result to the
//                   right neighbour as its incoming west value (right in ~99 %
//                   of the segments), pass 1 runs every segment in full, and a
//                   repair loop re-runs segments whose input turned out wrong
//                   until nothing changes.  Lane 0's input is exact (last pixel
//                   of the previous row: flat indexing, .cc:327-332), so by
//                   induction over lanes the fixed point is the serial result.
//       helper warp issues the TMA row loads and turns the residual bytes of row y+1 into pair form (PRE).
//       IO warp     turns the finished row y-1 (POST) into output pixels: delta add with independent
//                   byte wrap (.cc:337-338), high/low recombination, UnextractFrame shift and byte
//                   swap -- all on two pixels per register, the delta image coming in pair form
//                   (d | d << 16) so that one word serves the same column of both frames -- and
//                   issues the TMA row stores.
//     Roles are placed by SM sub-partition (%warpid): chains on 0 / 1, IO and helper warps on 2 / 3.
//   * SPLIT MODE (widths 1281..2560): the two 16-bit lanes hold the left and the right half of ONE
//     frame; see pair_chain_row.
//   * TMA both ways: rows are fetched with 1-D cp.async.bulk into 3-deep rings
//     signalled by mbarriers (one elected thread issues), finished rows leave
//     through a shared-memory row buffer and cp.async.bulk stores, so global
//     traffic is fully coalesced and costs no LSU instructions.
//   * Shared-memory accesses of the IO / helper warps are bank-conflict free for every L: lane-rotated
//     chunk order (chunk()) and a permuted delta row (pair_ddup_word).
//   * one named barrier per row couples the three warps of a pair; pre/post buffers are
//     double-buffered so the chain warp never waits for the IO warps' work of
//     the same row.
//
// Algorithmic traffic 4 B/px (1 + 1 in, 2 out); the pair-form delta image (4 B per column for two
// frames) is L2-resident.
#pragma once

#include <type_traits>

#include "fpv_internal.h"
#include "fpv_ptx.cuh"

namespace fpv {

struct PairParams {
  const uint8_t* high;
  const uint8_t* low;      // may be nullptr
  const uint8_t* flags;
  const uint32_t* ddup;    // delta image in pair form, rows permuted and padded (pair_ddup_word); may be nullptr
  uint16_t* out;
  uint32_t W, H;           // W: columns of one "frame" of the pair (split mode: half the frame's width)
  uint32_t stride;         // elements from one row to the next (== W, split mode: 2 W)
  uint64_t P;
  int shift, big_endian, unextract;
  uint32_t n;
};

#ifndef FPV_PAIR_K0
#define FPV_PAIR_K0 16
#endif
constexpr int kPairThreads = 192;   // per CTA: two pairs of frames x (chain, IO, helper) warps
constexpr int kPairRing = 3;

// -DFPV_PAIR_PROF: per-role cycle accounting (profiling builds only, see scripts/gpu_pair_prof.py).
// g_pair_prof: [0..6] chain: load+c, pass 0, pass 1, repair, store, barrier wait, repair rounds;
// [8..10] IO: work, TMA-store read wait, barrier wait; [12..14] helper: issue, pre_row (incl. TMA wait), barrier wait;
// [15] rows counted (chain warps).
#ifdef FPV_PAIR_PROF
__device__ unsigned long long g_pair_prof[16];
#define PROF_DECL(n) uint32_t prof_acc[n] = {}; uint32_t prof_t = (uint32_t)clock()
#define PROF_MARK(i) do { const uint32_t t__ = (uint32_t)clock(); prof_acc[i] += t__ - prof_t; prof_t = t__; } while (0)
#define PROF_FLUSH(base, n) do { if (lane == 0) for (int i__ = 0; i__ < (n); i__++) atomicAdd(&g_pair_prof[(base) + i__], (unsigned long long)prof_acc[i__]); } while (0)
#define PROF_PARAMS , uint32_t (&prof_acc)[8], uint32_t& prof_t
#define PROF_ARGS , prof_acc, prof_t
#define PROF_COUNT(i) prof_acc[i]++
#else
#define PROF_PARAMS
#define PROF_ARGS
#define PROF_COUNT(i)
#define PROF_DECL(n)
#define PROF_MARK(i)
#define PROF_FLUSH(base, n)
#endif

// Shared-memory plan (RB = 32 * L bytes = one padded byte row):
//   R1   ring x { residual A | residual B }                      2 RB each
//   R2   ring x { low A | low B | duplicated delta (4 B/col) }   6 RB each
//   PRE  2 x pair-form residual row  (L words per chain lane)    4 RB each
//   POST 2 x pair-form finished row                              4 RB each
//   OUT  { output row A | output row B }  (uint16 pixels)        4 RB
//   6 mbarriers
// A CTA holds two such regions (two pairs of frames) and a 4-word role table.
static inline size_t pair_smem_bytes(int LW2) {
  const size_t RB = 32 * 8 * (size_t)LW2;
  const size_t per_pair = RB * (kPairRing * 2 + kPairRing * 6 + 8 + 8 + 4) + 128;   // + mbarriers (padded)
  return 2 * per_pair + 32;                                                      // + role table
}

// Layout of one row of the duplicated delta image (global memory and, copied verbatim by TMA, shared
// memory): 32 L words, the word of column c = lane L + 8 chunk + 4 half + j (lane = c / L) sits at
//     ((2 chunk + half) 32 + lane) 4 + j
// so that the 32 lanes of an LDS.128 read 32 consecutive 16-byte slots (conflict free), like the
// PRE / POST buffers.  A linear row would put the lanes 4 L bytes apart: 8-way conflicts for L = 32.
__host__ __device__ inline uint32_t pair_ddup_word(uint32_t c, uint32_t L) {
  const uint32_t lane = c / L, r = c % L;
  return ((r >> 2) * 32u + lane) * 4u + (r & 3u);
}

__device__ __forceinline__ uint2 lds64(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
  return v;
}

// (a & m) | (b & ~m) in one LOP3.
__device__ __forceinline__ uint32_t bitselect(uint32_t a, uint32_t b, uint32_t m) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(r) : "r"(a), "r"(b), "r"(m));
  return r;
}

// One step of the chain for two pixels in S FORM: each 16-bit lane holds (byte << 8) | guard.
//     x = c + min3(n,w,nw) + max3(n,w,nw),   c = r - nw   (plain 32-bit arithmetic)
// because CG = n + w - median(n,w,nw) = min3 + max3 - nw (.cc:247-252).  With the byte in the TOP half of its
// lane the mod-256 wrap of the reference's uint8 arithmetic (.cc:330) happens by itself: what overflows
// the high lane leaves the register, what overflows the low lane (0..2 per step: the lane sum r - nw + min3 +
// max3 = r + n + w - median lies in [0, 765]) lands in the GUARD byte of the high lane.  The low lane's guard
// stays 0 (nothing carries into bit 0).  The high lane's guard only ever grows by that carry: n and nw are
// clean (guard 0), so if the 16-bit compare picks w it propagates w's guard once (w cannot be both the
// strict minimum and the strict maximum; if all three tie w equals the clean n), hence
//     guard(x) <= guard(w) + 2   and   guard <= 2 L <= 80 < 256 within a segment of L <= 40 pixels,
// starting from a clean incoming value.  A non-zero guard never changes a byte: it can only break ties
// between equal bytes, and either choice has the same top byte.  The finished row is cleaned once
// (& 0xff00ff00) before it becomes the next row's n / nw and the IO warp's input.  Against the masked
// lane-form step (min3, max3, add, and) this is one ALU-pipe instruction less per step and a dependent
// depth of 2 instead of 3; two chain warps share a sub-partition's ALU pipe, so both count.
__device__ __forceinline__ uint32_t chain_step(uint32_t c, uint32_t n, uint32_t w, uint32_t nw) {
  return c + __vimin3_u16x2(n, w, nw) + __vimax3_u16x2(n, w, nw);
}

// One row of the chain warp: residual row (pair form, in PRE) -> finished row in x[] and in POST.
// n[] is the finished previous row; the caller alternates two register arrays between rows so
// that nothing is copied.
// SPLIT: the two 16-bit lanes are not two frames but the left and the right half of ONE frame (widths
// up to 2560).  The chain then runs through the low lane's 32 segments and on through the high
// lane's: segment 0 of the right half takes its incoming values from the last segment of the left
// half of the same row (speculated and repaired like every other segment boundary), segment 0 of the
// left half from the right half of the previous row (exact).
template <int LW2, bool FULL, int K0T, int G, bool SPLIT>
__device__ __forceinline__ void pair_chain_row(const uint32_t (&n)[8 * LW2], uint32_t (&x)[8 * LW2], const uint32_t y,
                                               const uint32_t pre, const uint32_t post, const uint32_t cgmask,
                                               const uint32_t vmask, const int lane, const uint32_t last_lane,
                                               const uint32_t last_t, uint32_t& last_prev, uint32_t& last_prev2,
                                               const int bar_id PROF_PARAMS) {
  constexpr int L = 8 * LW2;
  constexpr int K0 = K0T < L ? K0T : L / 2;   // look-ahead pixels of pass 0
  // word quad k of this lane sits at 16-byte slot k*32 + lane
  const uint32_t a_lo = (uint32_t)lane * 16, a_hi = a_lo;
#pragma unroll
  for (int k = 0; k < 2 * LW2; k++) {
    const uint4 v = lds128(pre + k * 512 + (k < LW2 ? a_lo : a_hi));
    x[4 * k + 0] = v.x; x[4 * k + 1] = v.y; x[4 * k + 2] = v.z; x[4 * k + 3] = v.w;
  }
  // PRE holds the residual row in S form (see the helper's pre_row)
  if (cgmask != 0 && y > 0) {
    uint32_t c[L];
    uint32_t nw_in = __shfl_up_sync(0xffffffffu, n[L - 1], 1);
    // lane 0: the pixel before column 0 in flat order.  Pair mode: the last pixel of row y-2 of each
    // frame.  Split mode: left half <- last pixel of row y-2 (right half), right half <- the left
    // half's last pixel of row y-1.
    if (lane == 0) nw_in = SPLIT ? ((last_prev2 >> 16) | (last_prev << 16)) : last_prev2;
    // flat index W (row 1, column 0) is not predicted (.cc:327 starts at W+1)
    const bool copy_first = (y == 1) && (lane == 0);
    const uint32_t copy_mask = SPLIT ? 0x0000ffffu : 0xffffffffu;   // split: only the left half has a column 0
    const uint32_t r_first = x[0];
#pragma unroll
    for (int t = 0; t < L; t++) c[t] = x[t] - (t == 0 ? nw_in : n[t - 1]);
    PROF_MARK(0);

    // pass 0: estimate this segment's last pixel from a guess K0 pixels back
    uint32_t w_in;
#ifdef FPV_ABL_NO_PASS0
    w_in = __shfl_up_sync(0xffffffffu, n[L - 1], 1);
    if (lane == 0) w_in = last_prev;
    if (false)
#endif
    {
      uint32_t nw = n[L - K0 - 1], w = nw;          // the guess: west == north-west
#pragma unroll
      for (int t = L - K0; t < L; t++) {
        const uint32_t nn = n[t];
        w = chain_step(c[t], nn, w, nw);
        nw = nn;
      }
      w &= kHiBytes;                                // drop the guard byte before handing the value on
      w_in = __shfl_up_sync(0xffffffffu, w, 1);
      if (SPLIT) {
        // the left half's last segment feeds the right half's first (a guess, repaired below)
        const uint32_t wl = __shfl_sync(0xffffffffu, w, (int)last_lane);
        if (lane == 0) w_in = (last_prev >> 16) | (wl << 16);
      } else if (lane == 0) {
        w_in = last_prev;                           // exact for segment 0
      }
    }
    PROF_MARK(1);
    // pass 1: every segment in full
    {
      uint32_t w = w_in, nw = nw_in;
#pragma unroll
      for (int t = 0; t < L; t++) {
        const uint32_t nn = n[t];
        uint32_t v = chain_step(c[t], nn, w, nw);
        if (t == 0 && copy_first) v = (r_first & copy_mask) | (v & ~copy_mask);
        x[t] = v;
        w = v;
        nw = nn;
      }
    }
    PROF_MARK(2);
    // repair: re-run segments whose incoming value was wrong until nothing changes
#ifdef FPV_ABL_NO_REPAIR
    if (false)
#endif
    for (;;) {
      uint32_t w_new = __shfl_up_sync(0xffffffffu, x[L - 1] & kHiBytes, 1);
      if (SPLIT) {
        uint32_t xl = x[L - 1];
        if (!FULL) {
#pragma unroll
          for (int k = 0; k < 2 * LW2; k++)
            if ((uint32_t)(4 * k + 3) == last_t) xl = x[4 * k + 3];
        }
        xl = __shfl_sync(0xffffffffu, xl, (int)last_lane);   // only its (guard-free) low lane is used
        if (lane == 0) w_new = (last_prev >> 16) | (xl << 16);
      } else if (lane == 0) {
        w_new = last_prev;
      }
      const bool changed = ((w_new ^ w_in) & vmask) != 0;
      if (!__any_sync(0xffffffffu, changed)) break;
      PROF_COUNT(6);
      w_in = w_new;
      uint32_t w = w_in, nw = nw_in;
      bool settled = false;     // the chains met their old values before the segment ends: no end changed
#pragma unroll
      for (int k = 0; k < L / G; k++) {
        bool same = true;
#pragma unroll
        for (int j = 0; j < G; j++) {
          const int t = G * k + j;
          const uint32_t nn = n[t];
          uint32_t v = chain_step(c[t], nn, w, nw);
          if (t == 0 && copy_first) v = (r_first & copy_mask) | (v & ~copy_mask);
          if (j == G - 1) same = ((v ^ x[t]) & vmask) == 0;
          x[t] = v;
          w = v;
          nw = nn;
        }
        // every chain met its previous values: the rest of the segment is unchanged
        if (k + 1 < L / G && __all_sync(0xffffffffu, same)) { settled = true; break; }
      }
      // Nothing to hand on: skip the exchange and the vote of another round.  Not in split mode with a
      // partial last lane: its hand-over pixel x[last_t] lies before the point where the chains met
      // again and may have changed.
      if (settled && !(SPLIT && !FULL)) break;
    }
    PROF_MARK(3);
    if (cgmask == 0xffffffffu) {
      // clean the guard bytes: this row is the next row's n / nw and the IO warp's input
#pragma unroll
      for (int t = 0; t < L; t++) x[t] &= kHiBytes;
    } else {
      // one of the two frames is not ClampedGradient-predicted: its row is the residual row
      const uint32_t keep = cgmask & kHiBytes;
#pragma unroll
      for (int k = 0; k < 2 * LW2; k++) {
        const uint4 v = lds128(pre + k * 512 + (k < LW2 ? a_lo : a_hi));
        x[4 * k + 0] = (x[4 * k + 0] & keep) | (v.x & ~cgmask);
        x[4 * k + 1] = (x[4 * k + 1] & keep) | (v.y & ~cgmask);
        x[4 * k + 2] = (x[4 * k + 2] & keep) | (v.z & ~cgmask);
        x[4 * k + 3] = (x[4 * k + 3] & keep) | (v.w & ~cgmask);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 2 * LW2; k++)
    sts128(post + k * 512 + (k < LW2 ? a_lo : a_hi), x[4 * k + 0], x[4 * k + 1], x[4 * k + 2], x[4 * k + 3]);
  {
    uint32_t v = x[L - 1];
    if (!FULL) {
#pragma unroll
      for (int k = 0; k < 2 * LW2; k++)
        if ((uint32_t)(4 * k + 3) == last_t) v = x[4 * k + 3];   // W % 4 == 0: the row ends a quad
    }
    last_prev2 = last_prev;
    last_prev = __shfl_sync(0xffffffffu, v, (int)last_lane);
  }
  PROF_MARK(4);
  pair_bar_sync(bar_id);
  PROF_MARK(5);
  PROF_COUNT(7);
}

// LW2:   the chain lane's segment is L = 8 LW2 px (LW2 = ceil(W / 256)).
// FULL:  W == 32 L (every lane owns a complete segment).
// SHIFT: UnextractFrame with a non-zero shift is fused into the write-out.
// Repair groups: the chains are re-run G pixels at a time with a vote in between.  8 measured best for
// L = 40 (4 and 16 do not even divide it evenly), 16 for L = 32 (2048x2048: 69.9 -> 71.5 %).
#ifndef FPV_PAIR_G
#define FPV_PAIR_G 0
#endif
template <int LW2, bool FULL, bool SHIFT, bool SPLIT = false, int K0T = FPV_PAIR_K0,
          int G = (FPV_PAIR_G ? FPV_PAIR_G : (LW2 == 4 ? 16 : 8))>
__global__ void __launch_bounds__(kPairThreads, 2) k_decode_pair(const PairParams p) {
  extern __shared__ __align__(128) uint8_t psm[];
  constexpr int L = 8 * LW2;
  constexpr uint32_t RB = 32 * L;
  constexpr uint32_t kR1 = 0, kR1Slot = 2 * RB;
  constexpr uint32_t kR2 = kR1 + kPairRing * kR1Slot, kR2Slot = 6 * RB;
  constexpr uint32_t kPre = kR2 + kPairRing * kR2Slot, kBuf = 4 * RB;
  constexpr uint32_t kPost = kPre + 2 * kBuf;
  constexpr uint32_t kOut = kPost + 2 * kBuf;
  constexpr uint32_t kBars = kOut + 4 * RB;
  constexpr uint32_t kPairBytes = kBars + 128;      // per-pair region; the role table follows the two regions
  const uint32_t W = p.W, H = p.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- roles.  A latency-bound chain warp runs ~1.4x slower (measured) when it shares its SM
  //      sub-partition with throughput warps, and two chain warps on one sub-partition do not
  //      slow each other down.  The hardware gives the four warps of a CTA four consecutive warp
  //      slots (rotated from CTA to CTA), sub-partition = slot % 4, so roles are taken from
  //      %warpid: sub-partitions 0 / 1 run the chains of pair 0 / 1 of every resident CTA,
  //      sub-partitions 2 / 3 their IO warps.  The claim table makes the assignment a bijection
  //      whatever %warpid says (it is only a placement hint).
  uint32_t* claim = reinterpret_cast<uint32_t*>(psm + 2 * kPairBytes);
  if (threadIdx.x < 6) claim[threadIdx.x] = 0xffffffffu;
  __syncthreads();
  // roles: 0, 1 chain of pair 0 / 1 (sub-partitions 0 / 1); 2, 3 IO (sub-partitions 2 / 3); 4, 5 helper
  uint32_t role = 0xffffffffu;
  if (lane == 0) {
    uint32_t slot;
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(slot));
    if (atomicCAS(&claim[slot & 3u], 0xffffffffu, (uint32_t)warp) == 0xffffffffu) role = slot & 3u;
  }
  __syncthreads();
  if (lane == 0 && role == 0xffffffffu) {
    for (uint32_t r = 4; r < 10; r++)          // helpers first, then whatever is left
      if (atomicCAS(&claim[r % 6], 0xffffffffu, (uint32_t)warp) == 0xffffffffu) { role = r % 6; break; }
  }
  role = __shfl_sync(0xffffffffu, role, 0);
  const uint32_t pair = role & 1u;
  const bool is_chain = role < 2, is_helper = role >= 4;
  const int bar_id = 1 + (int)pair;                 // named barrier of this pair's three warps

  const uint32_t sm0 = smem_u32(psm) + pair * kPairBytes;
  const uint32_t full1 = sm0 + kBars, full2 = full1 + 8 * kPairRing;
  // pair mode: frames (fA, fA + 1); split mode: the two halves of frame fA
  const uint32_t fA = SPLIT ? 2 * blockIdx.x + pair : 4 * blockIdx.x + 2 * pair;
  if (fA >= p.n) return;                            // odd number of pairs: this half of the CTA has nothing to do
  const uint32_t fB = SPLIT ? fA : (fA + 1 < p.n ? fA + 1 : fA);   // odd tail: the pair is (A, A), B is not stored
  const uint32_t offB = SPLIT ? W : 0u;             // split mode: "frame B" starts W columns into the row
  const uint32_t flA = p.flags[fA], flB = p.flags[fB];
  const bool lowA = !(flA & kFlagNoLow) && p.low != nullptr, lowB = !(flB & kFlagNoLow) && p.low != nullptr;
  const bool delA = (flA & kFlagDelta) && p.ddup != nullptr, delB = (flB & kFlagDelta) && p.ddup != nullptr;
  const uint32_t cgmask = ((flA & kFlagCG) ? 0x0000ffffu : 0u) | ((flB & kFlagCG) ? 0xffff0000u : 0u);
  const uint32_t col0 = (uint32_t)lane * L;         // first column of this lane (both roles)

  if (is_chain) {
    // =============================== chain warp ===================================
    const bool lane_valid = FULL || col0 < W;
    const uint32_t vmask = lane_valid ? (cgmask & kHiBytes) : 0u;   // compares look at the bytes, not the guards
    const uint32_t last_lane = FULL ? 31u : (W - 1) / L, last_t = FULL ? (uint32_t)(L - 1) : (W - 1) % L;
    uint32_t ra[L], rb[L];                    // finished rows, alternating roles
#pragma unroll
    for (int t = 0; t < L; t++) rb[t] = 0;
    uint32_t last_prev = 0, last_prev2 = 0;   // h[y-1][W-1], h[y-2][W-1] of both frames

    pair_bar_sync(bar_id);               // mbarriers are initialised, the ring is zero-filled where needed
    pair_bar_sync(bar_id);               // row 0 is in PRE[0] (the IO warp's prologue)
    PROF_DECL(8);
    for (uint32_t y = 0; y < H; y += 2) {
      pair_chain_row<LW2, FULL, K0T, G, SPLIT>(rb, ra, y, sm0 + kPre, sm0 + kPost, cgmask, vmask, lane, last_lane, last_t,
                                        last_prev, last_prev2, bar_id PROF_ARGS);
      if (y + 1 < H)
        pair_chain_row<LW2, FULL, K0T, G, SPLIT>(ra, rb, y + 1, sm0 + kPre + kBuf, sm0 + kPost + kBuf, cgmask, vmask, lane,
                                          last_lane, last_t, last_prev, last_prev2, bar_id PROF_ARGS);
    }
#ifdef FPV_PAIR_PROF
    if (lane == 0) {
      for (int i = 0; i < 7; i++) atomicAdd(&g_pair_prof[i], (unsigned long long)prof_acc[i]);
      atomicAdd(&g_pair_prof[15], (unsigned long long)prof_acc[7]);
    }
#endif
    return;
  }

  // ============================ IO warp and helper warp ==============================
  // Lane l serves the chain lane l: columns [l L, (l+1) L) of both frames.  The helper issues
  // the TMA loads and turns residual bytes into pair form (PRE); the IO warp turns finished
  // rows (POST) into output pixels and issues the TMA stores.
  const bool elected = lane == 0;
  const uint32_t slot0 = (uint32_t)lane * 16;       // this lane's 16-byte slot in every quad row of PRE / POST
  const bool do_swap = p.unextract && p.big_endian;
  const uint32_t shmul = 1u << ((32 - p.shift) & 31), um = (0xffffu >> (p.shift & 31)) * 0x00010001u;
  // output words: two consecutive pixels of one frame out of two pair-form registers; the
  // UnextractFrame byte swap (.cc:857-860) is folded into the selector
  const uint32_t selA = do_swap ? 0x4501u : 0x5410u, selB = do_swap ? 0x6723u : 0x7632u;
  const uint32_t dmask = (delA ? 0x0000ffffu : 0u) | (delB ? 0xffff0000u : 0u);
  const uint32_t r2_bytes = (lowA ? W : 0u) + (lowB ? W : 0u) + (dmask ? 4 * RB : 0u);

  if (is_helper && elected) {
    for (int i = 0; i < 2 * kPairRing; i++) mbar_init(full1 + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // The write-out code is branch-free: a frame without a low plane reads zeros from ring rows that
  // TMA never writes; without delta the delta words are masked off (mh = ml = 0).
  if (is_helper && (!lowA || !lowB)) {
    for (uint32_t i = lane; i < kPairRing * RB / 4; i += 32) {
      const uint32_t s = i / (RB / 4), o = i % (RB / 4);
      if (!lowA) sts32(sm0 + kR2 + s * kR2Slot + 4 * o, 0u);
      if (!lowB) sts32(sm0 + kR2 + s * kR2Slot + RB + 4 * o, 0u);
    }
  }
  if (is_helper && !dmask)
    for (uint32_t i = lane; i < kPairRing * RB; i += 32)
      sts32(sm0 + kR2 + (i / RB) * kR2Slot + 2 * RB + 4 * (i % RB), 0u);
  pair_bar_sync(bar_id);

  // ---- TMA issue state.  It is warp-uniform (which lets the compiler hold it in uniform
  //      registers); only the elected lane executes the copy instructions.
  const uint32_t stride = p.stride;
  const uint8_t* s1A = p.high + (uint64_t)fA * p.P;   // next residual row to fetch
  const uint8_t* s1B = p.high + (uint64_t)fB * p.P + offB;
  const uint8_t* s2A = p.low + (uint64_t)fA * p.P;    // next low row (dereferenced only if lowA / lowB)
  const uint8_t* s2B = p.low + (uint64_t)fB * p.P + offB;
  const uint32_t* s2D = p.ddup;                       // 32 L words per row (pair_ddup_word)
  uint32_t i1_row = 0, i1_slot = 0, i2_row = 0, i2_slot = 0;
  auto issue_r1 = [&]() {     // residual rows of both frames
    if (i1_row < H && elected) {
      const uint32_t dst = sm0 + kR1 + i1_slot * kR1Slot, bar = full1 + 8 * i1_slot;
      mbar_arrive_expect_tx(bar, 2 * W);
      bulk_g2s(dst, s1A, W, bar);
      bulk_g2s(dst + RB, s1B, W, bar);
    }
    s1A += stride; s1B += stride;
    i1_row++;
    if (++i1_slot == kPairRing) i1_slot = 0;
  };
  auto issue_r2 = [&]() {     // low rows and the duplicated delta row
    if (i2_row < H && r2_bytes && elected) {
      const uint32_t dst = sm0 + kR2 + i2_slot * kR2Slot, bar = full2 + 8 * i2_slot;
      mbar_arrive_expect_tx(bar, r2_bytes);
      if (lowA) bulk_g2s(dst, s2A, W, bar);
      if (lowB) bulk_g2s(dst + RB, s2B, W, bar);
      if (dmask) bulk_g2s(dst + 2 * RB, s2D, 4 * RB, bar);    // a whole permuted row: 32 L words
    }
    s2A += stride; s2B += stride; s2D += RB;
    i2_row++;
    if (++i2_slot == kPairRing) i2_slot = 0;
  };

  // ---- consumer state: ring slot and mbarrier parity of the next row of each kind ------------
  uint32_t c1_slot = 0, c1_par = 0, c2_slot = 0, c2_par = 0;
  uint16_t* oA = p.out + (uint64_t)fA * p.P;        // next output row
  uint16_t* oB = p.out + (uint64_t)fB * p.P + offB;

  // Order in which a lane walks its LW2 chunks of 8 columns.  With L = 32 (LW2 = 4) the lanes sit
  // 32 / 64 bytes apart in the TMA-filled rows and in the output row: taking the chunks in a
  // lane-dependent rotation makes the 64-bit loads of a half-warp and the 128-bit stores of a
  // quarter-warp hit distinct banks ((l & 3, chunk) resp. (l & 1, chunk) are then all different).
  // Only addresses depend on it; register arrays stay indexed by the loop counter.
  const uint32_t rot = ((uint32_t)lane >> 2) + 2u * (((uint32_t)lane >> 1) & 1u);
  auto chunk = [&](int k) -> uint32_t {
    if constexpr (LW2 == 4) return ((uint32_t)k + rot) & 3u;
    else return (uint32_t)k;
  };
  // residual bytes of the next row -> pair form for the chain warp, into PRE[buf]
  auto pre_row = [&](uint32_t buf) {
    mbar_wait(full1 + 8 * c1_slot, c1_par);
    const uint32_t ra = sm0 + kR1 + c1_slot * kR1Slot + col0, dst = sm0 + kPre + buf * kBuf + slot0;
    if (++c1_slot == kPairRing) { c1_slot = 0; c1_par ^= 1u; }
    uint2 A[LW2], B[LW2];
#pragma unroll
    for (int k = 0; k < LW2; k++) { A[k] = lds64(ra + 8 * chunk(k)); B[k] = lds64(ra + RB + 8 * chunk(k)); }
    // PRE holds the residuals in S form, bytes [0, a_j, 0, b_j]; the zero bytes come out of the shifted
    // B words t = [0, 0, b0, b1] and u = [b2, b3, 0, 0] (the shifts are FMA-pipe multiplies)
#pragma unroll
    for (int k = 0; k < LW2; k++) {
      uint32_t t = B[k].x * 65536u, u = __umulhi(B[k].x, 65536u);
      sts128(dst + (2 * chunk(k)) * 512, __byte_perm(A[k].x, t, 0x6404), __byte_perm(A[k].x, t, 0x7414),
             __byte_perm(A[k].x, u, 0x4626), __byte_perm(A[k].x, u, 0x5636));
      t = B[k].y * 65536u; u = __umulhi(B[k].y, 65536u);
      sts128(dst + (2 * chunk(k) + 1) * 512, __byte_perm(A[k].y, t, 0x6404), __byte_perm(A[k].y, t, 0x7414),
             __byte_perm(A[k].y, u, 0x4626), __byte_perm(A[k].y, u, 0x5636));
    }
  };
  // finished row in POST[buf] (S form) -> output pixels (.cc:335-344 and .cc:850-862), then one bulk store per
  // frame.  DALL: both frames of the pair add the delta image (the common case, no per-lane delta mask).
  auto post_row = [&](uint32_t buf, auto dall_tag) {
    constexpr bool DALL = decltype(dall_tag)::value;
    if (r2_bytes) mbar_wait(full2 + 8 * c2_slot, c2_par);
    const uint32_t src = sm0 + kPost + buf * kBuf + slot0;
    const uint32_t la = sm0 + kR2 + c2_slot * kR2Slot + col0, da = sm0 + kR2 + c2_slot * kR2Slot + 2 * RB + slot0;
    if (++c2_slot == kPairRing) { c2_slot = 0; c2_par ^= 1u; }
    const uint32_t oa = sm0 + kOut + col0 * 2;
    uint4 X[2 * LW2], D[2 * LW2];
    uint2 A[LW2], B[LW2];
#pragma unroll
    for (int k = 0; k < LW2; k++) {
      X[2 * k] = lds128(src + (2 * chunk(k)) * 512);
      X[2 * k + 1] = lds128(src + (2 * chunk(k) + 1) * 512);
    }
#pragma unroll
    for (int k = 0; k < LW2; k++) { A[k] = lds64(la + 8 * chunk(k)); B[k] = lds64(la + RB + 8 * chunk(k)); }
#pragma unroll
    for (int k = 0; k < LW2; k++) {
      D[2 * k] = lds128(da + (2 * chunk(k)) * 512);          // pair_ddup_word: slot (2 chunk + half) 32 + lane
      D[2 * k + 1] = lds128(da + (2 * chunk(k) + 1) * 512);
    }
#pragma unroll
    for (int k = 0; k < LW2; k++) {     // 8 columns per step
      uint32_t Z[8], V[8];
      // low bytes of both frames in lane form [lA, 0, lB, 0]; the shifts are FMA-pipe multiplies
      uint32_t t = B[k].x * 65536u, u = __umulhi(B[k].x, 65536u);
      Z[0] = __byte_perm(A[k].x, t, 0x4640); Z[1] = __byte_perm(A[k].x, t, 0x4741);
      Z[2] = __byte_perm(A[k].x, u, 0x6462); Z[3] = __byte_perm(A[k].x, u, 0x6563);
      t = B[k].y * 65536u; u = __umulhi(B[k].y, 65536u);
      Z[4] = __byte_perm(A[k].y, t, 0x4640); Z[5] = __byte_perm(A[k].y, t, 0x4741);
      Z[6] = __byte_perm(A[k].y, u, 0x6462); Z[7] = __byte_perm(A[k].y, u, 0x6563);
      const uint32_t Xs[8] = {X[2 * k].x, X[2 * k].y, X[2 * k].z, X[2 * k].w,
                              X[2 * k + 1].x, X[2 * k + 1].y, X[2 * k + 1].z, X[2 * k + 1].w};
      const uint32_t Ds[8] = {D[2 * k].x, D[2 * k].y, D[2 * k].z, D[2 * k].w,
                              D[2 * k + 1].x, D[2 * k + 1].y, D[2 * k + 1].z, D[2 * k + 1].w};
#pragma unroll
      for (int j = 0; j < 8; j++) {
        // per 16-bit lane: ((x + dh) & 0xff) << 8 | ((l + dl) & 0xff) with packed 16-bit adds (VIADD.16x2:
        // no carry between the lanes).  x arrives as x << 8, so the first sum has the high byte in place
        // (its low byte is dl: dropped by the select); the second has the low byte in place (its carry
        // lands in the high byte: dropped) -- the bytes wrap independently, .cc:337-338.  dmask switches
        // the delta off for a frame of the pair that does not use it.
#ifdef FPV_ABL_IO_LIGHT
        V[j] = Xs[j] ^ Ds[j] ^ Z[j];
#else
        const uint32_t d = DALL ? Ds[j] : (Ds[j] & dmask);
        V[j] = bitselect(__vadd2(Xs[j], d), __vadd2(Z[j], d), kHiBytes);
        if (SHIFT) V[j] = __umulhi(V[j], shmul) & um;     // per lane: (pixel >> shift), .cc:855
#endif
      }
      sts128(oa + 16 * chunk(k), __byte_perm(V[0], V[1], selA), __byte_perm(V[2], V[3], selA),
             __byte_perm(V[4], V[5], selA), __byte_perm(V[6], V[7], selA));               // frame A: 8 pixels
      sts128(oa + 2 * RB + 16 * chunk(k), __byte_perm(V[0], V[1], selB), __byte_perm(V[2], V[3], selB),
             __byte_perm(V[4], V[5], selB), __byte_perm(V[6], V[7], selB));               // frame B
    }
    fence_proxy_async();
    __syncwarp();
    if (elected) {
      bulk_s2g(oA, sm0 + kOut, 2 * W);
      if (SPLIT || fB != fA) bulk_s2g(oB, sm0 + kOut + 2 * RB, 2 * W);
      bulk_commit();
    }
    oA += stride; oB += stride;
  };

  if (is_helper) {
    issue_r1(); issue_r1(); issue_r1();
    issue_r2();
    pre_row(0);
    pair_bar_sync(bar_id);
    PROF_DECL(3);
    for (uint32_t y = 0; y < H; y++) {
      issue_r1();          // row y + 3 into slot y % 3: its row y was consumed in iteration y - 1
      issue_r2();          // row y + 1 into slot (y + 1) % 3: its row y - 2 was consumed (by the IO warp) in iteration y - 1
      PROF_MARK(0);
      if (y + 1 < H) pre_row((y + 1) & 1u);
      PROF_MARK(1);
      pair_bar_sync(bar_id);
      PROF_MARK(2);
    }
    PROF_FLUSH(12, 3);
    return;
  }
  pair_bar_sync(bar_id);
  PROF_DECL(3);
  const bool dall = dmask == 0xffffffffu;
  for (uint32_t y = 0; y < H; y++) {
    if (y >= 1) {
      if (dall) post_row((y - 1) & 1u, std::true_type{});
      else post_row((y - 1) & 1u, std::false_type{});
    }
    PROF_MARK(0);
    if (elected) bulk_wait_read0();   // the output row buffer may be rewritten after the barrier
    PROF_MARK(1);
    pair_bar_sync(bar_id);
    PROF_MARK(2);
  }
  PROF_FLUSH(8, 3);
  if (dall) post_row((H - 1) & 1u, std::true_type{});
  else post_row((H - 1) & 1u, std::false_type{});
  if (elected) bulk_wait0();
}

template <int LW2, bool SPLIT = false>
static cudaError_t launch_pair(const PairParams& p, bool full, int blocks, cudaStream_t stream) {
  const size_t smem = pair_smem_bytes(LW2);
  const bool shift = p.unextract && p.shift != 0;
  cudaError_t e = cudaSuccess;
#define FPV_LAUNCH_PAIR(F, S)                                                                                            \
  do {                                                                                                                   \
    e = cudaFuncSetAttribute(k_decode_pair<LW2, F, S, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
    if (e == cudaSuccess) k_decode_pair<LW2, F, S, SPLIT><<<blocks, kPairThreads, smem, stream>>>(p);                    \
  } while (0)
  if (full && shift) FPV_LAUNCH_PAIR(true, true);
  else if (full) FPV_LAUNCH_PAIR(true, false);
  else if (shift) FPV_LAUNCH_PAIR(false, true);
  else FPV_LAUNCH_PAIR(false, false);
#undef FPV_LAUNCH_PAIR
  return e == cudaSuccess ? cudaGetLastError() : e;
}

}  // namespace fpv

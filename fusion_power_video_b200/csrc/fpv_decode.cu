// fpv_decode.cu -- decode-side kernels: the post-brotli part of DecompressImage
// (fusion_power_video.cc:326-344), UnextractFrame (.cc:850-862) and the
// plane-level undo of Frame::Uncompress (.cc:595-641).
//
// The inverse ClampedGradient is a strict serial chain in flat pixel order
// (.cc:327-332): h[i] += CG(h[i-W], h[i-1], h[i-W-1]) uses the just-written
// west neighbour, and the flat index wraps across row ends, so the exact
// dependency DAG has critical path W*H per frame.  What is exploited here:
//
//  * frames are independent                      -> one warp per frame;
//  * within a row (row y-1 final), x_i = f_i(x_{i-1}) with
//        f_i(w) = r_i + n_i + w - clamp(w, min(n_i,nw_i), max(n_i,nw_i))
//    (== r_i + CG(n_i, w, nw_i)), which is CONSTANT in w whenever w lies in
//    [min(n,nw), max(n,nw)].  Each of the 32 lanes owns a contiguous segment
//    of the row and runs its chain from a guessed incoming west value; lanes
//    then exchange their last pixel (shuffle) and re-run only until the new
//    values meet the old ones.  Lane 0's input is exact, so the fixed point is
//    the serial answer (induction over lanes); worst case 32 rounds.
//
// Rows are staged with cp.async (residual row, low row, delta row) a few rows
// ahead so that the chain never waits on HBM; the delta add, the high/low
// recombination and UnextractFrame are fused into the row write-out.
//
// Kernels, in the order pick_decode_kernel prefers them:
//   k_decode_fused (fpv_decode_fused.cuh)  one warp per pair of frames, TMA rings, write-out interleaved with the chain
//   k_decode_pair  (fpv_decode_pair.cuh)   round 1's three-warps-per-pair kernel (FPV_DECODE_KERNEL=pair)
//   k_decode_simd  (below)                 widths the two above do not take (W % 4 == 0, W <= 2048)
//   k_decode_spec  (below)                 everything else; in planes mode also fpv_unpredict_planes
//   k_cg_inverse_serial                    rows too wide for shared memory only
#include <stdlib.h>
#include <string.h>

#include "fpv_internal.h"
#include "fpv_decode_pair.cuh"
#include "fpv_decode_fused.cuh"

namespace fpv {

namespace {

__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4_a(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16_a(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// byte b of word v, moved to bits 24..31 (the chain runs on values scaled by
// 2^24 so that 32-bit wrap-around IS the mod-256 wrap and unsigned compares
// order bytes correctly).
__device__ __forceinline__ uint32_t byte_hi(uint32_t v, int b) {
  switch (b) {
    case 0: return __byte_perm(v, 0u, 0x0444);
    case 1: return __byte_perm(v, 0u, 0x1444);
    case 2: return __byte_perm(v, 0u, 0x2444);
    default: return v & 0xff000000u;
  }
}

struct DecodeParams {
  const uint8_t* high;
  const uint8_t* low;      // may be nullptr (all frames flags & 4)
  const uint8_t* flags;
  const uint16_t* delta;   // image form, may be nullptr
  uint16_t* out;
  uint8_t* plane_out;      // planes mode (fpv_unpredict_planes): the reconstructed byte plane goes here (must be
                           // `high`: in place) and nothing else is done -- no low plane, no delta, no `out`
  uint32_t W, H;
  uint64_t P;
  int shift, big_endian, unextract;
  uint32_t n;
  uint32_t L;              // segment length in pixels (multiple of 4), 32*L >= W
  uint32_t Lw;             // L / 4
  uint32_t SW;             // slot stride in words (odd -> conflict-free)
  uint32_t Wp;             // W rounded up to 16
  uint32_t nst;            // staged rows in flight (1..3)
  uint32_t div_magic;      // ceil(2^32 / L) for col / L
  uint32_t warp_smem_words;
};

// ALIGN: 16 -> W % 16 == 0 (16-byte cp.async for low/delta), 4 -> W % 4 == 0,
//        1  -> anything (synchronous byte staging, scalar stores).
template <int ALIGN>
__global__ void __launch_bounds__(128) k_decode_spec(const DecodeParams p) {
  extern __shared__ __align__(16) uint32_t dsm[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint32_t W = p.W, H = p.H, L = p.L, Lw = p.Lw, SW = p.SW, Wp = p.Wp, NST = p.nst;
  uint32_t* wsm = dsm + (size_t)wib * p.warp_smem_words;
  // per-warp layout (words): hbuf[2][32*SW] | stage[NST] x { r[32*SW] | low[Wp/4] | delta[Wp/2] }
  const uint32_t slot_words = 32 * SW;
  const uint32_t stage_words = slot_words + Wp / 4 + Wp / 2;
  uint32_t* hbuf0 = wsm;
  uint32_t* hbuf1 = wsm + slot_words;
  uint32_t* stage0 = wsm + 2 * slot_words;

  const uint32_t cs = (uint32_t)lane * L;                 // first column of this lane's segment
  const bool lane_active = cs < W;
  const uint32_t seg_px = lane_active ? min(L, W - cs) : 0;
  const uint32_t seg_words = (seg_px + 3) / 4;
  const uint32_t last_seg = (W - 1) / L, last_off = (W - 1) - last_seg * L;

  const uint32_t warps_total = gridDim.x * (blockDim.x >> 5);
  for (uint32_t f = blockIdx.x * (blockDim.x >> 5) + wib; f < p.n; f += warps_total) {
    const uint32_t fl = p.flags[f];
    const bool use_delta = (fl & kFlagDelta) && p.delta != nullptr;
    const bool use_cg = (fl & kFlagCG) != 0;
    const bool has_low = !(fl & kFlagNoLow) && p.low != nullptr;
    const uint8_t* fh = p.high + (uint64_t)f * p.P;
    const uint8_t* flow = has_low ? p.low + (uint64_t)f * p.P : nullptr;
    uint16_t* fout = p.out + (uint64_t)f * p.P;

    // ---- row staging -------------------------------------------------------
    auto issue_row = [&](uint32_t y) {
      if (y < H) {
        uint32_t* st = stage0 + (size_t)(y % NST) * stage_words;
        uint32_t* rb = st;
        uint8_t* lb = reinterpret_cast<uint8_t*>(st + slot_words);
        uint8_t* db = reinterpret_cast<uint8_t*>(st + slot_words + Wp / 4);
        const uint8_t* src = fh + (uint64_t)y * W;
        if (ALIGN >= 4) {
          for (uint32_t k = 0; k < seg_words; k++) cp_async4(rb + lane * SW + k, src + cs + 4 * k);
          if (ALIGN == 16) {
            if (has_low)
              for (uint32_t q = lane; q < W / 16; q += 32) cp_async16(lb + 16 * q, flow + (uint64_t)y * W + 16 * q);
            if (use_delta)
              for (uint32_t q = lane; q < W / 8; q += 32)
                cp_async16(db + 16 * q, reinterpret_cast<const uint8_t*>(p.delta + (uint64_t)y * W) + 16 * q);
          } else {
            if (has_low)
              for (uint32_t q = lane; q < W / 4; q += 32) cp_async4(lb + 4 * q, flow + (uint64_t)y * W + 4 * q);
            if (use_delta)
              for (uint32_t q = lane; q < W / 2; q += 32)
                cp_async4(db + 4 * q, reinterpret_cast<const uint8_t*>(p.delta + (uint64_t)y * W) + 4 * q);
          }
        } else {
          uint8_t* rb8 = reinterpret_cast<uint8_t*>(rb);
          for (uint32_t c = lane; c < W; c += 32) {
            uint32_t sg = c / L, off = c - sg * L;
            rb8[(size_t)sg * SW * 4 + off] = src[c];
            if (has_low) lb[c] = flow[(uint64_t)y * W + c];
            if (use_delta) reinterpret_cast<uint16_t*>(db)[c] = p.delta[(uint64_t)y * W + c];
          }
        }
      }
      cp_async_commit();  // one group per row, empty groups keep the count uniform
    };

    __syncwarp();
    for (uint32_t y = 0; y + 1 < NST; y++) issue_row(y);

    uint32_t last_prev = 0, last_prev2 = 0;  // h[y-1][W-1], h[y-2][W-1], scaled by 2^24

    for (uint32_t y = 0; y < H; y++) {
      // keep NST-1 rows in flight beyond the one being processed
      issue_row(y + NST - 1);
      if (NST == 1) cp_async_wait<0>();
      else if (NST == 2) cp_async_wait<1>();
      else cp_async_wait<2>();
      __syncwarp();

      uint32_t* st = stage0 + (size_t)(y % NST) * stage_words;
      const uint32_t* rb = st + lane * SW;
      uint32_t* hcur = (y & 1) ? hbuf1 : hbuf0;
      const uint32_t* hprev = (y & 1) ? hbuf0 : hbuf1;
      uint32_t* myh = hcur + lane * SW;
      const uint32_t* myn = hprev + lane * SW;

      if (!use_cg || y == 0) {
        for (uint32_t k = 0; k < seg_words; k++) myh[k] = rb[k];
      } else {
        // incoming west / north-west values for the first pixel of the segment
        uint32_t w_in, nw_in;
        if (lane == 0) {
          w_in = last_prev;
          nw_in = last_prev2;
        } else {
          // pixel above-left of my first pixel: last pixel of the previous lane's
          // segment in row y-1.  It is also the guess for the incoming west value.
          uint32_t v = lane_active ? hprev[(lane - 1) * SW + Lw - 1] : 0;
          nw_in = v & 0xff000000u;
          w_in = nw_in;
        }
        uint32_t out_px = 0;
        bool need = lane_active, first = true;
        for (;;) {
          if (need) {
            uint32_t x = w_in, nwv = nw_in;
            bool ran_to_end = true;
            for (uint32_t k = 0; k < seg_words; k++) {
              const uint32_t rw = rb[k], nd = myn[k], old = myh[k];
              uint32_t xs[4];
#pragma unroll
              for (int b = 0; b < 4; b++) {
                uint32_t r = byte_hi(rw, b), nn = byte_hi(nd, b);
                uint32_t lo = min(nn, nwv), hi = max(nn, nwv);
                uint32_t t = min(max(x, lo), hi);
                uint32_t xn = r + nn + x - t;
                // flat index W (row 1, column 0) is not predicted (.cc:327 starts at W+1)
                if (y == 1 && lane == 0 && k == 0 && b == 0) xn = r;
                nwv = nn;
                x = xn;
                xs[b] = xn;
              }
              uint32_t word = __byte_perm(__byte_perm(xs[0], xs[1], 0x0073),
                                          __byte_perm(xs[2], xs[3], 0x0073), 0x5410);
              myh[k] = word;
              // pixels past the end of the row inside the last word are scratch
              if (!first && k + 1 < seg_words && (word >> 24) == (old >> 24)) {
                ran_to_end = false;  // met the previous values: the rest is unchanged
                break;
              }
            }
            if (ran_to_end) {
              // last valid pixel of the segment
              uint32_t lastw = myh[seg_words - 1];
              out_px = byte_hi(lastw, (int)((seg_px - 1) & 3u));
            }
          }
          first = false;
          uint32_t prev_out = __shfl_up_sync(0xffffffffu, out_px, 1);
          bool changed = lane > 0 && lane_active && prev_out != w_in;
          if (!__any_sync(0xffffffffu, changed)) break;
          need = changed;
          if (changed) w_in = prev_out;
        }
      }
      __syncwarp();
      last_prev2 = last_prev;
      last_prev = byte_hi(hcur[last_seg * SW + (last_off >> 2)], (int)(last_off & 3u));

      // ---- write-out: delta add (bytes wrap independently, .cc:337-338),
      //      recombination, optional UnextractFrame ---------------------------
      const uint32_t* lb = st + slot_words;
      const uint32_t* db = st + slot_words + Wp / 4;
      uint16_t* orow = fout + (uint64_t)y * W;
      if (p.plane_out != nullptr) {
        // planes mode: the row of reconstructed bytes, in place (row y was staged before it is overwritten;
        // rows further down are only read)
        uint8_t* prow = p.plane_out + (uint64_t)f * p.P + (uint64_t)y * W;
        if (use_cg && y > 0)
          for (uint32_t q = lane; q < (W + 3) / 4; q += 32) {
            uint32_t col = 4 * q;
            uint32_t sg = __umulhi(col, p.div_magic);
            if ((sg + 1) * L <= col) sg++;
            else if (sg * L > col) sg--;
            const uint32_t hw = hcur[sg * SW + ((col - sg * L) >> 2)];
            if (ALIGN >= 4) {
              *reinterpret_cast<uint32_t*>(prow + col) = hw;
            } else {
#pragma unroll
              for (uint32_t b = 0; b < 4; b++)
                if (col + b < W) prow[col + b] = (uint8_t)(hw >> (8 * b));
            }
          }
      } else
      for (uint32_t q = lane; q < (W + 3) / 4; q += 32) {
        uint32_t col = 4 * q;
        uint32_t sg = __umulhi(col, p.div_magic);
        if ((sg + 1) * L <= col) sg++;            // magic is exact up to one step
        else if (sg * L > col) sg--;
        uint32_t off = col - sg * L;
        uint32_t hw = hcur[sg * SW + (off >> 2)];
        uint32_t lw = has_low ? lb[q] : 0u;
        uint32_t v01 = __byte_perm(lw, hw, 0x5140);   // (h0<<8|l0) | (h1<<8|l1)<<16
        uint32_t v23 = __byte_perm(lw, hw, 0x7362);
        if (use_delta) {
          v01 = __vadd4(v01, db[2 * q]);
          v23 = __vadd4(v23, db[2 * q + 1]);
        }
        if (p.unextract) {
          uint32_t m = (0xffffu >> p.shift) * 0x00010001u;
          v01 = (v01 >> p.shift) & m;
          v23 = (v23 >> p.shift) & m;
          if (p.big_endian) {
            v01 = __byte_perm(v01, 0u, 0x2301);
            v23 = __byte_perm(v23, 0u, 0x2301);
          }
        }
        if (ALIGN >= 4) {
          *reinterpret_cast<uint2*>(orow + col) = make_uint2(v01, v23);
        } else {
          if (col + 0 < W) orow[col + 0] = (uint16_t)(v01 & 0xffffu);
          if (col + 1 < W) orow[col + 1] = (uint16_t)(v01 >> 16);
          if (col + 2 < W) orow[col + 2] = (uint16_t)(v23 & 0xffffu);
          if (col + 3 < W) orow[col + 3] = (uint16_t)(v23 >> 16);
        }
      }
      __syncwarp();  // stage (y % NST) and hprev may be overwritten from here on
    }
    cp_async_wait<0>();
    __syncwarp();
  }
}


// =====================================================================================
// SIMD row kernel (the default): one warp per frame, 64 segments per row.
//
// Lane l owns the 2L contiguous columns [2L*l, 2L*(l+1)) as two half-segments of
// L = 4*LW pixels.  The two halves are two independent chains run in the two
// 16-bit lanes of one register ("lane form", values in [0,255]):
//     x = (c + min3(n, w, nw) + max3(n, w, nw)) & 0x00ff00ff,   c = r + 256 - nw
// because CG(n, w, nw) = n + w - median = min3 + max3 - nw (.cc:247-252) and the
// reconstructed byte is r + CG mod 256 (.cc:328-332).  That is two VIMNMX3.U16x2,
// one IADD3 and one LOP3 per step for TWO pixels, with a dependent depth of 3.
// The previous row, the current row and the c terms of a lane's 40-odd pixels all
// stay in registers; shared memory only stages the input rows (cp.async, a few
// rows ahead).  Segment inputs are speculated and repaired exactly as described
// at the top of this file: 64 chains start from guessed west values, then rounds
// of "shuffle the segment ends, re-run until the new values meet the old ones"
// until no segment's input changes.  Segment 0's input is exact, so by induction
// over segments the fixed point is the serial result.
// =====================================================================================

struct SimdParams {
  const uint8_t* high;
  const uint8_t* low;      // may be nullptr
  const uint8_t* flags;
  const uint16_t* delta;   // image form, may be nullptr
  uint16_t* out;
  uint32_t W, H;
  uint64_t P;
  int shift, big_endian, unextract;
  uint32_t n;
  uint32_t nst;            // staged rows in flight (2..4)
};

constexpr uint32_t kLaneM = 0x00ff00ffu;

// A16:  W % 16 == 0 (16-byte cp.async); otherwise W % 4 == 0 (4-byte cp.async).
// FULL: W == 64 L, every half-segment of every lane is complete (no column tests).
template <int LW, bool A16, bool FULL>
__global__ void __launch_bounds__(32) k_decode_simd(const SimdParams p) {
  extern __shared__ __align__(16) uint32_t dsm[];
  constexpr int L = 4 * LW;
  constexpr uint32_t RW = 16 * L;            // words per staged byte row (64 L columns)
  constexpr uint32_t STAGE = 4 * RW;         // residual | low | delta (2 RW)
  constexpr int CB = A16 ? 16 : 4;           // cp.async chunk bytes
  constexpr int NJ = (64 * L / CB + 31) / 32;   // chunks per lane per byte row (upper bound)
  const int lane = threadIdx.x;
  const uint32_t W = p.W, H = p.H, NST = p.nst;
  const uint32_t col0 = (uint32_t)lane * 2 * L;
  const bool v0 = FULL || col0 < W, v1 = FULL || col0 + L < W;   // half-segment holds real pixels
  const uint32_t vmask = (v0 ? 0x0000ffffu : 0u) | (v1 ? 0xffff0000u : 0u);
  // where the last pixel of a row lives (W % 4 == 0, L % 4 == 0: it ends a 4-step group)
  const uint32_t last_seg = FULL ? 63u : (W - 1) / L, last_t = FULL ? (uint32_t)(L - 1) : (W - 1) % L;
  const int last_lane = (int)(last_seg >> 1);
  const bool last_hi = (last_seg & 1u) != 0;
  const bool do_shift = p.unextract && p.shift != 0, do_swap = p.unextract && p.big_endian;
  const uint32_t ush = (uint32_t)p.shift, um = (0xffffu >> p.shift) * 0x00010001u;
  // cp.async chunk validity of this lane: bit j <=> chunk (lane + 32 j) lies inside the row
  uint32_t cmask1 = 0, cmask2 = 0;   // byte rows (W bytes) / delta rows (2 W bytes)
#pragma unroll
  for (int j = 0; j < 2 * NJ; j++) {
    if (j < NJ && (uint32_t)(lane + 32 * j) * CB < W) cmask1 |= 1u << j;
    if ((uint32_t)(lane + 32 * j) * CB < 2 * W) cmask2 |= 1u << j;
  }
  const uint32_t dsm_a = smem_addr(dsm) + (uint32_t)lane * CB;

  for (uint32_t f = blockIdx.x; f < p.n; f += gridDim.x) {
    const uint32_t fl = p.flags[f];
    const bool use_delta = (fl & kFlagDelta) && p.delta != nullptr;
    const bool use_cg = (fl & kFlagCG) != 0;
    const bool has_low = !(fl & kFlagNoLow) && p.low != nullptr;
    // running source pointers of the next row to stage (this lane's first chunk)
    const uint8_t* src_r = p.high + (uint64_t)f * p.P + (uint32_t)lane * CB;
    const uint8_t* src_l = has_low ? p.low + (uint64_t)f * p.P + (uint32_t)lane * CB : nullptr;
    const uint8_t* src_d = reinterpret_cast<const uint8_t*>(p.delta) + (uint32_t)lane * CB;
    uint16_t* orow = p.out + (uint64_t)f * p.P + col0;
    uint32_t issue_y = 0, issue_slot = 0;

    auto issue_row = [&]() {
      if (issue_y < H) {
        const uint32_t st = dsm_a + issue_slot * (STAGE * 4);
#pragma unroll
        for (int j = 0; j < NJ; j++)
          if (cmask1 & (1u << j)) {
            if (A16) cp_async16_a(st + 32 * CB * j, src_r + 32 * CB * j);
            else cp_async4_a(st + 32 * CB * j, src_r + 32 * CB * j);
          }
        if (has_low) {
#pragma unroll
          for (int j = 0; j < NJ; j++)
            if (cmask1 & (1u << j)) {
              if (A16) cp_async16_a(st + RW * 4 + 32 * CB * j, src_l + 32 * CB * j);
              else cp_async4_a(st + RW * 4 + 32 * CB * j, src_l + 32 * CB * j);
            }
          src_l += W;
        }
        if (use_delta) {
#pragma unroll
          for (int j = 0; j < 2 * NJ; j++)
            if (cmask2 & (1u << j)) {
              if (A16) cp_async16_a(st + 2 * RW * 4 + 32 * CB * j, src_d + 32 * CB * j);
              else cp_async4_a(st + 2 * RW * 4 + 32 * CB * j, src_d + 32 * CB * j);
            }
          src_d += 2 * W;
        }
        src_r += W;
      }
      cp_async_commit();  // one group per row; empty groups keep the count uniform
      issue_y++;
      if (++issue_slot == NST) issue_slot = 0;
    };

    __syncwarp();
    for (uint32_t y = 0; y + 1 < NST; y++) issue_row();

    uint32_t nrow[L], xrow[L], c[L];
#pragma unroll
    for (int t = 0; t < L; t++) nrow[t] = 0;
    uint32_t last_prev = 0, last_prev2 = 0;  // h[y-1][W-1], h[y-2][W-1]
    uint32_t slot = 0;

    for (uint32_t y = 0; y < H; y++) {
      issue_row();
      if (NST == 2) cp_async_wait<1>();
      else if (NST == 3) cp_async_wait<2>();
      else cp_async_wait<3>();
      __syncwarp();
      const uint32_t* st = dsm + (size_t)slot * STAGE + (uint32_t)lane * 2 * LW;   // this lane's chunk
      if (++slot == NST) slot = 0;

      // ---- residual bytes -> lane-form pairs (half 0 in bits 0-15, half 1 in bits 16-31)
#pragma unroll
      for (int k = 0; k < LW; k++) {
        const uint32_t A = st[k], B = st[LW + k];
        const uint32_t Ae = A & kLaneM, Ao = __byte_perm(A, 0u, 0x4341);
        const uint32_t Be = B & kLaneM, Bo = __byte_perm(B, 0u, 0x4341);
        xrow[4 * k + 0] = __byte_perm(Ae, Be, 0x5410);
        xrow[4 * k + 1] = __byte_perm(Ao, Bo, 0x5410);
        xrow[4 * k + 2] = __byte_perm(Ae, Be, 0x7632);
        xrow[4 * k + 3] = __byte_perm(Ao, Bo, 0x7632);
      }

      if (use_cg && y > 0) {
        // north-west of each half's first pixel: last pixel of the segment to the left, row y-1
        const uint32_t ln = nrow[L - 1];
        uint32_t nw_in = __funnelshift_l(__shfl_up_sync(0xffffffffu, ln, 1), ln, 16);
        if (lane == 0) nw_in = (nw_in & 0xffff0000u) | last_prev2;
        uint32_t w_in = nw_in;                       // the guess: west == north-west
        if (lane == 0) w_in = (w_in & 0xffff0000u) | last_prev;   // exact for segment 0
        // flat index W (row 1, column 0) is not predicted (.cc:327 starts at W+1)
        const bool copy_first = (y == 1) && (lane == 0);
        const uint32_t r_first = xrow[0];
#pragma unroll
        for (int t = 0; t < L; t++) c[t] = xrow[t] + 0x01000100u - (t == 0 ? nw_in : nrow[t - 1]);

        // round 1: every chain runs its whole segment
        {
          uint32_t w = w_in, nw = nw_in;
#pragma unroll
          for (int t = 0; t < L; t++) {
            const uint32_t n = nrow[t];
            uint32_t x = (c[t] + __vimin3_u16x2(n, w, nw) + __vimax3_u16x2(n, w, nw)) & kLaneM;
            if (t == 0 && copy_first) x = (x & 0xffff0000u) | (r_first & 0x0000ffffu);
            xrow[t] = x;
            w = x;
            nw = n;
          }
        }
        // repair rounds
        for (;;) {
          const uint32_t o = xrow[L - 1];
          uint32_t w_new = __funnelshift_l(__shfl_up_sync(0xffffffffu, o, 1), o, 16);
          if (lane == 0) w_new = (w_new & 0xffff0000u) | last_prev;
          const bool changed = ((w_new ^ w_in) & vmask) != 0;
          if (!__any_sync(0xffffffffu, changed)) break;
          w_in = w_new;
          uint32_t w = w_in, nw = nw_in;
#pragma unroll
          for (int k = 0; k < LW; k++) {
            bool same = true;
#pragma unroll
            for (int j = 0; j < 4; j++) {
              const int t = 4 * k + j;
              const uint32_t n = nrow[t];
              uint32_t x = (c[t] + __vimin3_u16x2(n, w, nw) + __vimax3_u16x2(n, w, nw)) & kLaneM;
              if (t == 0 && copy_first) x = (x & 0xffff0000u) | (r_first & 0x0000ffffu);
              if (j == 3) same = ((x ^ xrow[t]) & vmask) == 0;
              xrow[t] = x;
              w = x;
              nw = n;
            }
            // every chain met its previous values: the rest of the segment is unchanged
            if (k + 1 < LW && __all_sync(0xffffffffu, same)) break;
          }
        }
      }

      // ---- the row is final: carry it to the next row ------------------------------
      {
        uint32_t v = xrow[L - 1];
        if (!FULL) {
#pragma unroll
          for (int k = 0; k < LW; k++)
            if ((uint32_t)(4 * k + 3) == last_t) v = xrow[4 * k + 3];
        }
        v = last_hi ? (v >> 16) : (v & 0xffffu);
        last_prev2 = last_prev;
        last_prev = __shfl_sync(0xffffffffu, v, last_lane);
      }
#pragma unroll
      for (int t = 0; t < L; t++) nrow[t] = xrow[t];

      // ---- write-out: delta add (bytes wrap independently, .cc:337-338),
      //      recombination, optional UnextractFrame (.cc:850-862) -------------------
      const uint32_t* lbw = st + RW;                                        // this lane's low bytes
      const uint2* dbw = reinterpret_cast<const uint2*>(st + 2 * RW + (uint32_t)lane * 2 * LW);  // delta: 2x the offset
#pragma unroll
      for (int k = 0; k < LW; k++) {
        const uint32_t t01 = __byte_perm(xrow[4 * k + 0], xrow[4 * k + 1], 0x6240);  // a0 a1 b0 b1
        const uint32_t t23 = __byte_perm(xrow[4 * k + 2], xrow[4 * k + 3], 0x6240);  // a2 a3 b2 b3
        const uint32_t hw2[2] = {__byte_perm(t01, t23, 0x5410), __byte_perm(t01, t23, 0x7632)};
#pragma unroll
        for (int h = 0; h < 2; h++) {
          if (FULL || col0 + (uint32_t)(h * L + 4 * k) < W) {
            const uint32_t lw = has_low ? lbw[h * LW + k] : 0u;
            uint32_t v01 = __byte_perm(lw, hw2[h], 0x5140);   // (h0<<8|l0) | (h1<<8|l1)<<16
            uint32_t v23 = __byte_perm(lw, hw2[h], 0x7362);
            if (use_delta) {
              const uint2 d = dbw[h * LW + k];
              v01 = __vadd4(v01, d.x);
              v23 = __vadd4(v23, d.y);
            }
            if (do_shift) {
              v01 = (v01 >> ush) & um;
              v23 = (v23 >> ush) & um;
            }
            if (do_swap) {
              v01 = __byte_perm(v01, 0u, 0x2301);
              v23 = __byte_perm(v23, 0u, 0x2301);
            }
            *reinterpret_cast<uint2*>(orow + h * L + 4 * k) = make_uint2(v01, v23);
          }
        }
      }
      orow += W;
      __syncwarp();  // the consumed stage may be overwritten from here on
    }
    cp_async_wait<0>();
    __syncwarp();
  }
}

template <int LW>
static cudaError_t launch_simd(const SimdParams& p, bool a16, bool full, int blocks, size_t smem,
                               cudaStream_t stream) {
  cudaError_t e = cudaSuccess;
#define FPV_LAUNCH_SIMD(A, F)                                                                                \
  do {                                                                                                       \
    e = cudaFuncSetAttribute(k_decode_simd<LW, A, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e == cudaSuccess) k_decode_simd<LW, A, F><<<blocks, 32, smem, stream>>>(p);                          \
  } while (0)
  if (a16 && full) FPV_LAUNCH_SIMD(true, true);
  else if (a16) FPV_LAUNCH_SIMD(true, false);
  else FPV_LAUNCH_SIMD(false, false);
#undef FPV_LAUNCH_SIMD
  return e == cudaSuccess ? cudaGetLastError() : e;
}

// ---- trivially serial fallback -------------------------------------------------
// One thread per frame runs .cc:327-332 literally on a scratch copy of the high
// plane; a second, fully parallel kernel does .cc:335-344 (+ .cc:850-862).
__global__ void k_cg_inverse_serial(uint8_t* planes, const uint8_t* flags, uint32_t W, uint64_t n_px,
                                    uint32_t n) {
  uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n || !(flags[f] & kFlagCG)) return;
  uint8_t* h = planes + (uint64_t)f * n_px;
  for (uint64_t i = (uint64_t)W + 1; i < n_px; i++)
    h[i] = (uint8_t)(h[i] + cg1(h[i - W], h[i - 1], h[i - W - 1]));
}

__global__ void k_combine(const uint8_t* high, const uint8_t* low, const uint8_t* flags,
                          const uint16_t* delta, uint16_t* out, uint64_t P, int shift, int big_endian,
                          int unextract) {
  uint32_t f = blockIdx.y;
  uint32_t fl = flags[f];
  bool use_delta = (fl & kFlagDelta) && delta != nullptr;
  bool has_low = !(fl & kFlagNoLow) && low != nullptr;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P;
       i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t h = high[(uint64_t)f * P + i];
    uint32_t l = has_low ? low[(uint64_t)f * P + i] : 0u;
    if (use_delta) {
      uint32_t d = delta[i];
      h = (h + (d >> 8)) & 0xffu;
      l = (l + (d & 0xffu)) & 0xffu;
    }
    uint32_t v = (h << 8) | l;
    if (unextract) {
      v >>= shift;
      if (big_endian) v = ((v & 0xffu) << 8) | (v >> 8);
    }
    out[(uint64_t)f * P + i] = (uint16_t)v;
  }
}

// high/low planes += delta planes, for frames with flags & 1 (.cc:600-603).
__global__ void k_planes_add_delta(uint8_t* high, uint8_t* low, const uint8_t* flags,
                                   const uint16_t* delta, uint64_t P) {
  uint32_t f = blockIdx.y;
  if (!(flags[f] & kFlagDelta)) return;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P;
       i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t d = delta[i];
    high[(uint64_t)f * P + i] = (uint8_t)(high[(uint64_t)f * P + i] + (d >> 8));
    if (low) low[(uint64_t)f * P + i] = (uint8_t)(low[(uint64_t)f * P + i] + (d & 0xffu));
  }
}

// delta image -> pair form for k_decode_pair, one padded and permuted row of 32 L words per image row
// (pair_ddup_word).  Pair mode: (d | d << 16), the same column of two frames.  Split mode (one frame,
// left half in the low lane, right half in the high lane): d[row][c] | d[row][c + W/2] << 16.
__global__ void k_delta_dup(const uint16_t* delta, uint32_t* ddup, uint32_t W, uint32_t H, uint32_t L, int split) {
  const uint32_t Wp = split ? W / 2 : W;          // columns of the pair kernel's "frame"
  const uint64_t total = (uint64_t)Wp * H;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t row = i / Wp;
    const uint32_t c = (uint32_t)(i % Wp);
    const uint32_t a = delta[row * W + c], b = split ? delta[row * W + c + Wp] : a;
    ddup[row * 32u * L + pair_ddup_word(c, L)] = a | (b << 16);
  }
}

}  // namespace

// Widths the pair kernel takes two frames at a time / one frame as two halves.
static bool pair_width_ok(uint32_t W) { return W % 16 == 0 && W >= 64 && W <= 1280; }
static bool split_width_ok(uint32_t W) { return W % 32 == 0 && W > 1280 && W <= 2560; }

// Columns per lane / 8 of the pair kernel for a geometry (the duplicated delta image depends on it).
static int pair_lw2(const Geom& g, bool* split_out) {
  const bool split = split_width_ok(g.W);
  const uint32_t Wp = split ? g.W / 2 : g.W;
  int LW2 = (int)((Wp + 255) / 256);
  if (const char* v = getenv("FPV_PAIR_LW2")) { const int k = atoi(v); if (k >= LW2 && k <= 5) LW2 = k; }
  if (split_out) *split_out = split;
  return LW2;
}

size_t delta_dup_bytes(const Geom& g) { return ((size_t)g.W + 256) * g.H * 4; }

int enqueue_delta_dup(const Geom& g, const uint16_t* delta_image, uint32_t* ddup, cudaStream_t stream,
                      cudaError_t* err) {
  if (!pair_width_ok(g.W) && !split_width_ok(g.W)) { *err = cudaSuccess; return 0; }   // the pair kernel is not used
  bool split = false;
  const uint32_t L = 8u * (uint32_t)pair_lw2(g, &split);
  unsigned gx = (unsigned)((g.P + 255) / 256);
  if (gx > 1184) gx = 1184;
  *err = cudaMemsetAsync(ddup, 0, (size_t)32 * L * g.H * 4, stream);      // padding columns
  if (*err != cudaSuccess) return -1;
  k_delta_dup<<<gx, 256, 0, stream>>>(delta_image, ddup, g.W, g.H, L, split ? 1 : 0);
  *err = cudaGetLastError();
  return *err == cudaSuccess ? 1 : -1;
}

// Which decode kernel handles a geometry; FPV_DECODE_KERNEL=pair|simd|spec forces one
// (tests run every kernel on every case it supports).
enum DecodeKernel { kDecFused, kDecPair, kDecSimd, kDecSpec };

static DecodeKernel pick_decode_kernel(const Geom& g, const uint16_t* delta, const uint32_t* ddup) {
  const bool pair_ok = (pair_width_ok(g.W) || split_width_ok(g.W)) && (delta == nullptr || ddup != nullptr);
  const bool simd_ok = g.W % 4 == 0 && g.W <= 64 * 32 && g.W >= 64;
  if (const char* v = getenv("FPV_DECODE_KERNEL")) {
    if (!strcmp(v, "fused") && pair_ok) return kDecFused;
    if (!strcmp(v, "pair") && pair_ok) return kDecPair;
    if (!strcmp(v, "simd") && simd_ok) return kDecSimd;
    if (!strcmp(v, "spec")) return kDecSpec;
  }
  if (getenv("FPV_DECODE_SEGMENTED")) return kDecSpec;
  if (pair_ok) return kDecFused;
  if (simd_ok) return kDecSimd;
  return kDecSpec;
}

// k_decode_spec for the geometry in p (W, H, P, n and the pointers set by the caller).
static int launch_spec(DecodeParams p, int num_sms, cudaStream_t stream, cudaError_t* err, const TimingHook* hook) {
  uint32_t L = (p.W + 31) / 32;
  L = (L + 3) / 4 * 4;
  p.L = L; p.Lw = L / 4; p.SW = p.Lw | 1u;
  p.Wp = (p.W + 15) / 16 * 16;
  p.div_magic = (uint32_t)((0x100000000ull + L - 1) / L);
  const uint32_t slot_words = 32 * p.SW;
  const uint32_t stage_words = slot_words + p.Wp / 4 + p.Wp / 2;
  const int warps_per_block = 4;
  const size_t limit = 200 * 1024;
  int nst = 3;
  while (nst > 1 && (size_t)(2 * slot_words + nst * stage_words) * 4 * warps_per_block > limit) nst--;
  int wpb = warps_per_block;
  while (wpb > 1 && (size_t)(2 * slot_words + nst * stage_words) * 4 * wpb > limit) wpb--;
  p.nst = nst;
  p.warp_smem_words = 2 * slot_words + nst * stage_words;
  size_t smem = (size_t)p.warp_smem_words * 4 * wpb;
  if (smem > limit) {
    *err = cudaErrorInvalidConfiguration;  // geometry not supported: the caller falls back to enqueue_decode_serial
    return -2;
  }
  int blocks = (int)((p.n + wpb - 1) / wpb);
  int max_blocks = num_sms * 16;
  if (blocks > max_blocks) blocks = max_blocks;
  int align = (p.W % 16 == 0) ? 16 : (p.W % 4 == 0 ? 4 : 1);
  cudaError_t e = cudaSuccess;
  if (hook) cudaEventRecord(hook->start, stream);
  if (align == 16) {
    e = cudaFuncSetAttribute(k_decode_spec<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit);
    if (e == cudaSuccess) k_decode_spec<16><<<blocks, wpb * 32, smem, stream>>>(p);
  } else if (align == 4) {
    e = cudaFuncSetAttribute(k_decode_spec<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit);
    if (e == cudaSuccess) k_decode_spec<4><<<blocks, wpb * 32, smem, stream>>>(p);
  } else {
    e = cudaFuncSetAttribute(k_decode_spec<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit);
    if (e == cudaSuccess) k_decode_spec<1><<<blocks, wpb * 32, smem, stream>>>(p);
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  if (hook) cudaEventRecord(hook->stop, stream);
  *err = e;
  return e == cudaSuccess ? 1 : -1;
}


int enqueue_decode(const Geom& g, int num_sms, const uint8_t* high, const uint8_t* low,
                   const uint8_t* flags, const uint16_t* delta, const uint32_t* ddup, uint32_t n,
                   bool unextract, uint16_t* out, cudaStream_t stream, cudaError_t* err,
                   const TimingHook* hook) {
  const DecodeKernel which = pick_decode_kernel(g, delta, ddup);
  if (which == kDecPair || which == kDecFused) {
    PairParams pp;
    pp.high = high; pp.low = low; pp.flags = flags; pp.ddup = delta ? ddup : nullptr; pp.out = out;
    const bool split = split_width_ok(g.W);
    pp.W = split ? g.W / 2 : g.W; pp.stride = g.W; pp.H = g.H; pp.P = g.P; pp.shift = g.shift;
    pp.big_endian = g.big_endian;
    pp.unextract = unextract ? 1 : 0; pp.n = n;
    // A lane owns L = 8 LW2 contiguous columns.
    const int LW2 = pair_lw2(g, nullptr);
    const bool full = pp.W == 256u * (uint32_t)LW2;
    // two pairs of frames per CTA; split mode: two frames (each a pair of halves)
    const int blocks = split ? (int)((n + 1) / 2) : (int)((n + 3) / 4);
    cudaError_t e = cudaSuccess;
    if (hook) cudaEventRecord(hook->start, stream);
    if (which == kDecFused) {
      // one warp per pair of frames (split mode: per frame)
      if (split) {
        if (LW2 == 3) e = launch_fused<3, true>(pp, full, stream);
        else if (LW2 == 4) e = launch_fused<4, true>(pp, full, stream);
        else e = launch_fused<5, true>(pp, full, stream);
      } else
      switch (LW2) {
        case 1: e = launch_fused<1>(pp, full, stream); break;
        case 2: e = launch_fused<2>(pp, full, stream); break;
        case 3: e = launch_fused<3>(pp, full, stream); break;
        case 4: e = launch_fused<4>(pp, full, stream); break;
        default: e = launch_fused<5>(pp, full, stream); break;
      }
    } else if (split) {
      // half widths of 641..1280 columns
      if (LW2 == 3) e = launch_pair<3, true>(pp, full, blocks, stream);
      else if (LW2 == 4) e = launch_pair<4, true>(pp, full, blocks, stream);
      else e = launch_pair<5, true>(pp, full, blocks, stream);
    } else
    switch (LW2) {
      case 1: e = launch_pair<1>(pp, full, blocks, stream); break;
      case 2: e = launch_pair<2>(pp, full, blocks, stream); break;
      case 3: e = launch_pair<3>(pp, full, blocks, stream); break;
      case 4: e = launch_pair<4>(pp, full, blocks, stream); break;
      default: e = launch_pair<5>(pp, full, blocks, stream); break;
    }
    if (hook) cudaEventRecord(hook->stop, stream);
    *err = e;
    return e == cudaSuccess ? 1 : -1;
  }
  if (which == kDecSimd) {
    // SIMD row kernel: smallest segment length L = 4 LW with 64 L >= W
    SimdParams sp;
    sp.high = high; sp.low = low; sp.flags = flags; sp.delta = delta; sp.out = out;
    sp.W = g.W; sp.H = g.H; sp.P = g.P; sp.shift = g.shift; sp.big_endian = g.big_endian;
    sp.unextract = unextract ? 1 : 0; sp.n = n;
    const int LW = (int)((g.W + 255) / 256);
    sp.nst = 3;
    if (const char* v = getenv("FPV_DECODE_STAGES")) { int k = atoi(v); if (k >= 2 && k <= 4) sp.nst = (uint32_t)k; }
    const size_t smem = (size_t)sp.nst * 4 * 16 * 4 * LW * 4;
    int per_sm = (int)((size_t)(220 * 1024) / (smem + 1024));
    if (per_sm > 32) per_sm = 32;
    if (per_sm < 1) per_sm = 1;
    int blocks = (int)n;
    if (blocks > num_sms * per_sm) blocks = num_sms * per_sm;
    const bool a16 = g.W % 16 == 0;
    const bool full = g.W == 256u * (uint32_t)LW;
    cudaError_t e = cudaSuccess;
    if (hook) cudaEventRecord(hook->start, stream);
    switch (LW) {
      case 1: e = launch_simd<1>(sp, a16, full, blocks, smem, stream); break;
      case 2: e = launch_simd<2>(sp, a16, full, blocks, smem, stream); break;
      case 3: e = launch_simd<3>(sp, a16, full, blocks, smem, stream); break;
      case 4: e = launch_simd<4>(sp, a16, full, blocks, smem, stream); break;
      case 5: e = launch_simd<5>(sp, a16, full, blocks, smem, stream); break;
      case 6: e = launch_simd<6>(sp, a16, full, blocks, smem, stream); break;
      case 7: e = launch_simd<7>(sp, a16, full, blocks, smem, stream); break;
      default: e = launch_simd<8>(sp, a16, full, blocks, smem, stream); break;
    }
    if (hook) cudaEventRecord(hook->stop, stream);
    *err = e;
    return e == cudaSuccess ? 1 : -1;
  }
  DecodeParams p;
  p.high = high; p.low = low; p.flags = flags; p.delta = delta; p.out = out; p.plane_out = nullptr;
  p.W = g.W; p.H = g.H; p.P = g.P; p.shift = g.shift; p.big_endian = g.big_endian;
  p.unextract = unextract ? 1 : 0; p.n = n;
  return launch_spec(p, num_sms, stream, err, hook);
}

// Serial fallback: `scratch_high` must hold a writable copy of the high planes.
int enqueue_decode_serial(const Geom& g, uint8_t* scratch_high, const uint8_t* low,
                          const uint8_t* flags, const uint16_t* delta, uint32_t n, bool unextract,
                          uint16_t* out, cudaStream_t stream, cudaError_t* err) {
  k_cg_inverse_serial<<<(n + 31) / 32, 32, 0, stream>>>(scratch_high, flags, g.W, g.P, n);
  *err = cudaGetLastError();
  if (*err != cudaSuccess) return -1;
  unsigned gx = (unsigned)((g.P + 255) / 256); if (gx > 2048) gx = 2048; if (gx < 1) gx = 1;
  k_combine<<<dim3(gx, n), 256, 0, stream>>>(scratch_high, low, flags, delta, out, g.P, g.shift,
                                             g.big_endian, unextract ? 1 : 0);
  *err = cudaGetLastError();
  return *err == cudaSuccess ? 2 : -1;
}

// Frame::Uncompress's undo of Predict on byte planes, in place (.cc:595-641): inverse ClampedGradient of the high
// plane (width W) and of the preview (width W / 4) with the segmented speculative row kernel in planes mode, then
// the delta planes are added back.
int enqueue_unpredict_planes(const Geom& g, int num_sms, uint8_t* high, uint8_t* low,
                             uint8_t* preview, const uint8_t* flags, const uint16_t* delta,
                             uint32_t n, cudaStream_t stream, cudaError_t* err) {
  int launches = 0;
  DecodeParams p;
  p.low = nullptr; p.flags = flags; p.delta = nullptr; p.out = nullptr;
  p.shift = 0; p.big_endian = 0; p.unextract = 0; p.n = n;
  p.high = high; p.plane_out = high; p.W = g.W; p.H = g.H; p.P = g.P;
  if (launch_spec(p, num_sms, stream, err, nullptr) < 0) return -1;
  launches++;
  if (preview) {
    p.high = preview; p.plane_out = preview; p.W = g.PW; p.H = g.H / 4; p.P = g.PP;
    if (launch_spec(p, num_sms, stream, err, nullptr) < 0) return -1;
    launches++;
  }
  if (delta) {
    unsigned gx = (unsigned)((g.P + 255) / 256); if (gx > 2048) gx = 2048; if (gx < 1) gx = 1;
    k_planes_add_delta<<<dim3(gx, n), 256, 0, stream>>>(high, low, flags, delta, g.P);
    launches++;
  }
  *err = cudaGetLastError();
  return *err == cudaSuccess ? launches : -1;
}

}  // namespace fpv

#ifdef FPV_FUSED_PROF
// Profiling builds only: reads and clears k_decode_fused's histogram of repair rounds per row.
extern "C" int fpv_debug_fused_rounds(unsigned long long* out8) {
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(out8, fpv::g_fused_rounds, 8 * sizeof(unsigned long long)) != cudaSuccess) return 1;
  unsigned long long zero[8] = {};
  return cudaMemcpyToSymbol(fpv::g_fused_rounds, zero, sizeof zero) == cudaSuccess ? 0 : 1;
}
#endif

#ifdef FPV_PAIR_PROF
// Profiling builds only: reads and clears the per-role cycle counters of k_decode_pair.
extern "C" int fpv_debug_pair_prof(unsigned long long* out16) {
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(out16, fpv::g_pair_prof, 16 * sizeof(unsigned long long)) != cudaSuccess) return 1;
  unsigned long long zero[16] = {};
  return cudaMemcpyToSymbol(fpv::g_pair_prof, zero, sizeof zero) == cudaSuccess ? 0 : 1;
}
#endif

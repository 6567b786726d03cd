// fpv_decode_fused.cuh -- k_decode_fused: the default decode kernel (one warp per pair of frames).
//
// Same job as k_decode_pair (fpv_decode_pair.cuh): the post-brotli part of DecompressImage
// (fusion_power_video.cc:326-344) fused with UnextractFrame (.cc:850-862), the inverse
// ClampedGradient run as 32 speculated-and-repaired segment chains per row (see the top of
// fpv_decode.cu and pair_chain_row), two frames -- or the two halves of one wide frame -- in the
// two 16-bit lanes of every register, in S form (chain_step).
//
// What is different: k_decode_pair splits the work of a pair over three warps (chain, IO,
// helper) that hand rows to each other through shared memory (PRE / POST buffers) and meet at a
// named barrier every row.  Per-role cycle counters showed what that costs: shared-memory
// bandwidth (50 KB per pair and row, 80 % of the SM's 128 B/clk at the measured row rate), the
// row time of the slowest of three warps, and warps starved of issue slots right after every
// barrier.  Here ONE warp does everything for its pair:
//
//   row y:   residual bytes of row y (TMA ring) -> S form in registers -> c terms -> pass 0
//            -> pass 1  INTERLEAVED WITH  the write-out of row y-1  -> repair loop -> clean-up
//
// The write-out of row y-1 (delta add, high / low recombination, UnextractFrame, staging in the
// OUT row, one bulk store per frame) reads the finished row straight from the registers that
// serve as pass 1's north row, so it is independent of the running chain and sits in the same
// basic block: the dependent chain (3 ALU-pipe instructions per step, issued every ~11 cycles)
// leaves most issue slots free and the write-out fills them.  No PRE / POST traffic (30 KB per
// pair and row instead of 50), no barriers between warps, no roles; 8 warps (16 frames) per SM.
//
// Shared memory per warp (RB = 32 L bytes = one padded byte row):
//   R1   2 x { residual A | residual B }      consumed at the start of a row, refilled at once (lead: 2 rows)
//   LOW  3 x { low A | low B }                consumed by the write-out one row later (lead: 2 rows)
//   DEL  2 x duplicated delta row (4 RB)      L2-resident, lead: 1 row
//   OUT  { output row A | output row B }      uint16 pixels, source of the bulk stores
//   7 mbarriers
// Algorithmic traffic 4 B/px (1 + 1 in, 2 out); the delta image comes from L2.
#pragma once

#include "fpv_decode_pair.cuh"

namespace fpv {

// -DFPV_FUSED_PROF (profiling builds only, scripts/gpu_fused_rounds.py): histogram of repair rounds per row of a pair,
// [0..6] rounds, [7] = 7 or more.
#ifdef FPV_FUSED_PROF
__device__ unsigned long long g_fused_rounds[8];
#endif

constexpr int kFusedWarps = 4;          // warps (pairs of frames) per CTA; two CTAs per SM
constexpr int kFusedR1 = 2, kFusedLow = 3, kFusedDel = 2;

// (no OUT row when the output goes out with 256-bit stores straight from registers: DIRECT)
static inline size_t fused_warp_bytes(int LW2, bool direct) {
  const size_t RB = 32 * 8 * (size_t)LW2;
  return RB * (2 * kFusedR1 + 2 * kFusedLow + 4 * kFusedDel + (direct ? 0 : 4)) + 64;
}
static inline size_t fused_smem_bytes(int LW2, bool direct) { return kFusedWarps * fused_warp_bytes(LW2, direct); }
// Resident CTAs per SM: two.  Three fit for L <= 32 with direct stores (shared memory: no OUT row) if the kernel is
// held to 168 registers, but that build is slower although it spills almost nothing (1024x1024: 0.77 at 2368 frames,
// 0.85 at 3552 against 0.88 with two CTAs of 216-254 registers; 2048x2048 0.76 against 0.89): with fewer registers the
// compiler can no longer keep the write-out's operands in flight beside the chain.  -DFPV_FUSED_CTAS3=1 builds it.
#ifndef FPV_FUSED_CTAS3
#define FPV_FUSED_CTAS3 0
#endif
constexpr int fused_min_ctas(int LW2, bool direct) { return (FPV_FUSED_CTAS3 && direct && LW2 <= 4) ? 3 : 2; }

// Per-warp state of the TMA rings and the write-out.  Everything is warp-uniform.
// DIRECT: the output row does not go through shared memory: every lane stores its own 2 L bytes per frame with
// 256-bit global stores (full 32-byte sectors; needs L % 16 == 0 and a 32-byte aligned output).  With L = 32 the
// lanes' pieces of the OUT row lie 64 bytes apart, a 4-way bank conflict on every 128-bit shared store, and the
// proxy fence in front of the bulk stores has to wait for them.  (L = 40 keeps the OUT row and the bulk stores: its
// 80 bytes per lane are not whole sectors, and 128-bit stores of half sectors measured 0.59 of the roofline
// against 0.86.)
template <int LW2, bool DIRECT = false>
struct FusedCtx {
  static constexpr uint32_t L = 8 * LW2, RB = 32 * L;
  static constexpr uint32_t kR1 = 0, kLow = kR1 + kFusedR1 * 2 * RB, kDel = kLow + kFusedLow * 2 * RB;
  static constexpr uint32_t kOut = kDel + kFusedDel * 4 * RB, kBars = kOut + (DIRECT ? 0u : 4 * RB);
  static constexpr uint32_t kBytes = kBars + 64;     // one warp's region
  uint32_t sm0;                       // this warp's region
  uint32_t W, H, stride;
  const uint8_t *s1A, *s1B;           // next residual row to fetch
  const uint8_t *s2A, *s2B;           // next low row to fetch
  const uint32_t* s2D;                // next duplicated delta row to fetch
  uint16_t *oA, *oB;                  // next output row
  uint32_t i1_row, i1_slot, il_row, il_slot, id_row, id_slot;   // issue state
  uint32_t c1_slot, c1_par, cl_slot, cl_par, cd_slot, cd_par;   // consumer state
  bool lowA, lowB, store_b;
  uint32_t dmask, cgmask;
  uint32_t shmul, um, selA, selB;
  uint32_t col0;                      // first column of this lane
  int lane;

  __device__ __forceinline__ uint32_t bar_r1(uint32_t s) const { return sm0 + kBars + 8 * s; }
  __device__ __forceinline__ uint32_t bar_low(uint32_t s) const { return sm0 + kBars + 8 * (kFusedR1 + s); }
  __device__ __forceinline__ uint32_t bar_del(uint32_t s) const { return sm0 + kBars + 8 * (kFusedR1 + kFusedLow + s); }

  // ---- TMA issue (lane 0 copies; the bookkeeping is warp-uniform) -----------------------------
  __device__ __forceinline__ void issue_r1() {
    if (i1_row < H && lane == 0) {
      const uint32_t dst = sm0 + kR1 + i1_slot * 2 * RB, bar = bar_r1(i1_slot);
      mbar_arrive_expect_tx(bar, 2 * W);
      bulk_g2s(dst, s1A, W, bar);
      bulk_g2s(dst + RB, s1B, W, bar);
    }
    s1A += stride; s1B += stride;
    i1_row++;
    if (++i1_slot == kFusedR1) i1_slot = 0;
  }
  __device__ __forceinline__ void issue_low() {
    if ((lowA || lowB) && il_row < H && lane == 0) {
      const uint32_t dst = sm0 + kLow + il_slot * 2 * RB, bar = bar_low(il_slot);
      mbar_arrive_expect_tx(bar, (lowA ? W : 0u) + (lowB ? W : 0u));
      if (lowA) bulk_g2s(dst, s2A, W, bar);
      if (lowB) bulk_g2s(dst + RB, s2B, W, bar);
    }
    s2A += stride; s2B += stride;
    il_row++;
    if (++il_slot == kFusedLow) il_slot = 0;
  }
  __device__ __forceinline__ void issue_del() {
    if (dmask && id_row < H && lane == 0) {
      const uint32_t dst = sm0 + kDel + id_slot * 4 * RB, bar = bar_del(id_slot);
      mbar_arrive_expect_tx(bar, 4 * RB);
      bulk_g2s(dst, s2D, 4 * RB, bar);               // a whole permuted row: 32 L words (pair_ddup_word)
    }
    s2D += RB;
    id_row++;
    if (++id_slot == kFusedDel) id_slot = 0;
  }

  // Everything lane 0 has to issue at the end of a row, in ONE divergent region (six separate `if (lane == 0)` blocks
  // per row cost 11 % of the kernel's stall samples in branches and reconvergence): the bulk stores of the row just
  // written out (`store`: not for direct-store variants), then the refills of the three rings.
  // with_r1: the residual ring is refilled here too (direct-store variants); the bulk-store variants (L = 40) refill
  // it right after the slot is consumed at the start of the row -- a lead of one row instead of two measured 0.847
  // against 0.862 there, while the direct-store variants gain a point from the single region.
  __device__ __forceinline__ void refill(bool store, bool with_r1) {
    const bool do1 = with_r1 && i1_row < H, dol = (lowA || lowB) && il_row < H, dod = dmask != 0 && id_row < H;
    if (lane == 0) {
      if (store) {
        bulk_s2g(oA, sm0 + kOut, 2 * W);
        if (store_b) bulk_s2g(oB, sm0 + kOut + 2 * RB, 2 * W);
        bulk_commit();
      }
      if (do1) {
        const uint32_t dst = sm0 + kR1 + i1_slot * 2 * RB, bar = bar_r1(i1_slot);
        mbar_arrive_expect_tx(bar, 2 * W);
        bulk_g2s(dst, s1A, W, bar);
        bulk_g2s(dst + RB, s1B, W, bar);
      }
      if (dol) {
        const uint32_t dst = sm0 + kLow + il_slot * 2 * RB, bar = bar_low(il_slot);
        mbar_arrive_expect_tx(bar, (lowA ? W : 0u) + (lowB ? W : 0u));
        if (lowA) bulk_g2s(dst, s2A, W, bar);
        if (lowB) bulk_g2s(dst + RB, s2B, W, bar);
      }
      if (dod) {
        const uint32_t dst = sm0 + kDel + id_slot * 4 * RB, bar = bar_del(id_slot);
        mbar_arrive_expect_tx(bar, 4 * RB);
        bulk_g2s(dst, s2D, 4 * RB, bar);             // a whole permuted row: 32 L words (pair_ddup_word)
      }
    }
    if (with_r1) {
      s1A += stride; s1B += stride; i1_row++;
      if (++i1_slot == kFusedR1) i1_slot = 0;
    }
    s2A += stride; s2B += stride; il_row++;
    if (++il_slot == kFusedLow) il_slot = 0;
    s2D += RB; id_row++;
    if (++id_slot == kFusedDel) id_slot = 0;
  }

  // ---- residual bytes of the next row -> S form [0, a, 0, b] per register ----------------------
  __device__ __forceinline__ void load_residuals(uint32_t (&r)[8 * LW2]) {
    mbar_wait(bar_r1(c1_slot), c1_par);
    const uint32_t ra = sm0 + kR1 + c1_slot * 2 * RB + (uint32_t)lane * L;
    if (++c1_slot == kFusedR1) { c1_slot = 0; c1_par ^= 1u; }
    uint2 A[LW2], B[LW2];
    if constexpr (LW2 % 2 == 0) {
      // 16-byte loads: lanes 32 bytes apart (L = 32) still collide in pairs, 8-byte loads would 4-way
#pragma unroll
      for (int k = 0; k < LW2; k += 2) {
        const uint4 a = lds128(ra + 8 * k), b = lds128(ra + RB + 8 * k);
        A[k] = make_uint2(a.x, a.y); A[k + 1] = make_uint2(a.z, a.w);
        B[k] = make_uint2(b.x, b.y); B[k + 1] = make_uint2(b.z, b.w);
      }
    } else {
#pragma unroll
      for (int k = 0; k < LW2; k++) { A[k] = lds64(ra + 8 * k); B[k] = lds64(ra + RB + 8 * k); }
    }
    // the zero bytes come out of the shifted B words t = [0, 0, b0, b1] and u = [b2, b3, 0, 0]
    // (the shifts are FMA-pipe multiplies)
#pragma unroll
    for (int k = 0; k < LW2; k++) {
      uint32_t t = B[k].x * 65536u, u = __umulhi(B[k].x, 65536u);
      r[8 * k + 0] = __byte_perm(A[k].x, t, 0x6404); r[8 * k + 1] = __byte_perm(A[k].x, t, 0x7414);
      r[8 * k + 2] = __byte_perm(A[k].x, u, 0x4626); r[8 * k + 3] = __byte_perm(A[k].x, u, 0x5636);
      t = B[k].y * 65536u; u = __umulhi(B[k].y, 65536u);
      r[8 * k + 4] = __byte_perm(A[k].y, t, 0x6404); r[8 * k + 5] = __byte_perm(A[k].y, t, 0x7414);
      r[8 * k + 6] = __byte_perm(A[k].y, u, 0x4626); r[8 * k + 7] = __byte_perm(A[k].y, u, 0x5636);
    }
    if constexpr (!DIRECT) {
      __syncwarp();            // every lane has read the slot: refill it (row + kFusedR1)
      issue_r1();
    }                          // (direct-store variants: refill() at the end of the row)
  }
};

// Staged inputs of one row's write-out: low bytes of both frames and the duplicated delta words.
template <int LW2>
struct FusedOutRegs {
  uint2 A[LW2], B[LW2];
  uint4 D[2 * LW2];
  uint32_t pa[4], pb[4];       // DIRECT: output words of the even chunk, waiting for the odd one (one 256-bit store)
};

template <int LW2, bool DIRECT>
__device__ __forceinline__ void fused_out_load(FusedCtx<LW2, DIRECT>& cx, FusedOutRegs<LW2>& o) {
  constexpr uint32_t L = 8 * LW2, RB = 32 * L;
  using C = FusedCtx<LW2, DIRECT>;
  if (cx.lowA || cx.lowB) mbar_wait(cx.bar_low(cx.cl_slot), cx.cl_par);
  if (cx.dmask) mbar_wait(cx.bar_del(cx.cd_slot), cx.cd_par);
  const uint32_t la = cx.sm0 + C::kLow + cx.cl_slot * 2 * RB + (uint32_t)cx.lane * L;
  const uint32_t da = cx.sm0 + C::kDel + cx.cd_slot * 4 * RB + (uint32_t)cx.lane * 16;
  if (++cx.cl_slot == kFusedLow) { cx.cl_slot = 0; cx.cl_par ^= 1u; }
  if (++cx.cd_slot == kFusedDel) { cx.cd_slot = 0; cx.cd_par ^= 1u; }
  if constexpr (LW2 % 2 == 0) {
#pragma unroll
    for (int k = 0; k < LW2; k += 2) {
      const uint4 a = lds128(la + 8 * k), b = lds128(la + RB + 8 * k);
      o.A[k] = make_uint2(a.x, a.y); o.A[k + 1] = make_uint2(a.z, a.w);
      o.B[k] = make_uint2(b.x, b.y); o.B[k + 1] = make_uint2(b.z, b.w);
    }
  } else {
#pragma unroll
    for (int k = 0; k < LW2; k++) { o.A[k] = lds64(la + 8 * k); o.B[k] = lds64(la + RB + 8 * k); }
  }
#pragma unroll
  for (int k = 0; k < 2 * LW2; k++) o.D[k] = lds128(da + k * 512);   // pair_ddup_word: slot (2 chunk + half) 32 + lane
  if constexpr (!DIRECT) {
    // the previous row's bulk stores must have read the OUT row before it is rewritten
    if (cx.lane == 0) bulk_wait_read0();
    __syncwarp();
  }
}

// Columns [8k, 8k + 8) of this lane: finished row (S form, in xs) -> output pixels in the OUT row.
// per 16-bit lane: ((x + dh) & 0xff) << 8 | ((l + dl) & 0xff) with packed 16-bit adds (VIADD.16x2: no carry
// between the lanes).  x arrives as x << 8, so the first sum has the high byte in place (its low byte is dl:
// dropped by the select); the second has the low byte in place (its carry lands in the high byte: dropped) --
// the bytes wrap independently, .cc:337-338.  dmask switches the delta off for a frame that does not use it.
template <int LW2, bool SHIFT, bool DALL, bool DIRECT>
__device__ __forceinline__ void fused_out_chunk(const FusedCtx<LW2, DIRECT>& cx, FusedOutRegs<LW2>& o,
                                                const uint32_t (&xs)[8 * LW2], const int k) {
  constexpr uint32_t L = 8 * LW2, RB = 32 * L;
  using C = FusedCtx<LW2, DIRECT>;
  uint32_t Z[8], V[8];
  // low bytes of both frames in lane form [lA, 0, lB, 0]
  uint32_t t = o.B[k].x * 65536u, u = __umulhi(o.B[k].x, 65536u);
  Z[0] = __byte_perm(o.A[k].x, t, 0x4640); Z[1] = __byte_perm(o.A[k].x, t, 0x4741);
  Z[2] = __byte_perm(o.A[k].x, u, 0x6462); Z[3] = __byte_perm(o.A[k].x, u, 0x6563);
  t = o.B[k].y * 65536u; u = __umulhi(o.B[k].y, 65536u);
  Z[4] = __byte_perm(o.A[k].y, t, 0x4640); Z[5] = __byte_perm(o.A[k].y, t, 0x4741);
  Z[6] = __byte_perm(o.A[k].y, u, 0x6462); Z[7] = __byte_perm(o.A[k].y, u, 0x6563);
  const uint32_t Ds[8] = {o.D[2 * k].x, o.D[2 * k].y, o.D[2 * k].z, o.D[2 * k].w,
                          o.D[2 * k + 1].x, o.D[2 * k + 1].y, o.D[2 * k + 1].z, o.D[2 * k + 1].w};
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const uint32_t d = DALL ? Ds[j] : (Ds[j] & cx.dmask);
    V[j] = bitselect(__vadd2(xs[8 * k + j], d), __vadd2(Z[j], d), kHiBytes);
    if (SHIFT) V[j] = __umulhi(V[j], cx.shmul) & cx.um;     // per lane: (pixel >> shift), .cc:855
  }
  const uint32_t a0 = __byte_perm(V[0], V[1], cx.selA), a1 = __byte_perm(V[2], V[3], cx.selA),
                 a2 = __byte_perm(V[4], V[5], cx.selA), a3 = __byte_perm(V[6], V[7], cx.selA);   // frame A: 8 pixels
  const uint32_t b0 = __byte_perm(V[0], V[1], cx.selB), b1 = __byte_perm(V[2], V[3], cx.selB),
                 b2 = __byte_perm(V[4], V[5], cx.selB), b3 = __byte_perm(V[6], V[7], cx.selB);   // frame B
  if constexpr (DIRECT) {
    if ((k & 1) == 0) {
      o.pa[0] = a0; o.pa[1] = a1; o.pa[2] = a2; o.pa[3] = a3;
      o.pb[0] = b0; o.pb[1] = b1; o.pb[2] = b2; o.pb[3] = b3;
    } else if (cx.col0 + 8u * (uint32_t)(k + 1) <= cx.W) {    // W % 16 == 0: the 16 columns are inside the row or not
      // cx.oA / cx.oB point at this lane's first pixel of the row being written
      stg256(cx.oA + 8 * (k - 1), o.pa[0], o.pa[1], o.pa[2], o.pa[3], a0, a1, a2, a3);
      if (cx.store_b) stg256(cx.oB + 8 * (k - 1), o.pb[0], o.pb[1], o.pb[2], o.pb[3], b0, b1, b2, b3);
    }
  } else {
    const uint32_t oa = cx.sm0 + C::kOut + (uint32_t)cx.lane * 2 * L + 16 * k;
    sts128(oa, a0, a1, a2, a3);
    sts128(oa + 2 * RB, b0, b1, b2, b3);
  }
}

// OUT row -> global memory (one bulk store per frame), then refill the rings this row's write-out freed.
template <int LW2, bool DIRECT>
__device__ __forceinline__ void fused_out_store(FusedCtx<LW2, DIRECT>& cx) {
  constexpr uint32_t L = 8 * LW2, RB = 32 * L;
  using C = FusedCtx<LW2, DIRECT>;
  if constexpr (!DIRECT) fence_proxy_async();
  __syncwarp();            // every lane has read its R1 / LOW / DEL slots and written its piece of the OUT row
  cx.refill(!DIRECT, DIRECT);
  cx.oA += cx.stride; cx.oB += cx.stride;
}

// The write-out of a whole row at once (first / last rows and frames without ClampedGradient).
template <int LW2, bool SHIFT, bool DIRECT>
__device__ __forceinline__ void fused_out_row(FusedCtx<LW2, DIRECT>& cx, const uint32_t (&xs)[8 * LW2]) {
  FusedOutRegs<LW2> o;
  fused_out_load<LW2, DIRECT>(cx, o);
  if (cx.dmask == 0xffffffffu) {
#pragma unroll
    for (int k = 0; k < LW2; k++) fused_out_chunk<LW2, SHIFT, true, DIRECT>(cx, o, xs, k);
  } else {
#pragma unroll
    for (int k = 0; k < LW2; k++) fused_out_chunk<LW2, SHIFT, false, DIRECT>(cx, o, xs, k);
  }
  fused_out_store<LW2, DIRECT>(cx);
}

// One row y >= 1 of a pair with at least one ClampedGradient frame: chain of row y (into x, north row n)
// interleaved with the write-out of row y - 1 (= n).  See pair_chain_row for the chain itself.
template <int LW2, bool FULL, bool SHIFT, bool SPLIT, int K0T, int G, bool DALL, bool DIRECT>
__device__ __forceinline__ void fused_chain_row(FusedCtx<LW2, DIRECT>& cx, const uint32_t (&n)[8 * LW2], uint32_t (&x)[8 * LW2],
                                                const uint32_t y, const uint32_t vmask, const uint32_t last_lane,
                                                const uint32_t last_t, uint32_t& last_prev, uint32_t& last_prev2) {
  constexpr int L = 8 * LW2;
  constexpr int K0 = K0T < L ? K0T : L / 2;   // look-ahead pixels of pass 0
  const int lane = cx.lane;
  const uint32_t cgmask = cx.cgmask;
  uint32_t c[L];
  cx.load_residuals(c);                        // c holds r for now
  FusedOutRegs<LW2> o;
  fused_out_load<LW2, DIRECT>(cx, o);                  // staged early: the loads fly while pass 0 runs

  uint32_t nw_in = __shfl_up_sync(0xffffffffu, n[L - 1], 1);
  // lane 0: the pixel before column 0 in flat order.  Pair mode: the last pixel of row y-2 of each
  // frame.  Split mode: left half <- last pixel of row y-2 (right half), right half <- the left
  // half's last pixel of row y-1.
  if (lane == 0) nw_in = SPLIT ? ((last_prev2 >> 16) | (last_prev << 16)) : last_prev2;
  // flat index W (row 1, column 0) is not predicted (.cc:327 starts at W+1)
  const bool copy_first = (y == 1) && (lane == 0);
  const uint32_t copy_mask = SPLIT ? 0x0000ffffu : 0xffffffffu;   // split: only the left half has a column 0
  const uint32_t r_first = c[0];
#pragma unroll
  for (int t = L - 1; t >= 0; t--) c[t] = c[t] - (t == 0 ? nw_in : n[t - 1]);

  // pass 0: estimate this segment's last pixel from a guess K0 pixels back
  uint32_t w_in;
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
    uint32_t wl = 0;
    if (SPLIT) wl = __shfl_sync(0xffffffffu, w, (int)last_lane);
    // the first chunk of the write-out of row y-1 goes here: something to issue while the shuffles are in flight
    fused_out_chunk<LW2, SHIFT, DALL, DIRECT>(cx, o, n, 0);
    if (SPLIT) {
      // the left half's last segment feeds the right half's first (a guess, repaired below)
      if (lane == 0) w_in = (last_prev >> 16) | (wl << 16);
    } else if (lane == 0) {
      w_in = last_prev;                           // exact for segment 0
    }
  }
  // pass 1: every segment in full; the write-out of row y-1 (n) fills the chain's idle issue slots
  {
    uint32_t w = w_in, nw = nw_in;
#pragma unroll
    for (int k = 0; k < LW2; k++) {
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int t = 8 * k + j;
        const uint32_t nn = n[t];
        uint32_t v = chain_step(c[t], nn, w, nw);
        if (t == 0 && copy_first) v = (r_first & copy_mask) | (v & ~copy_mask);
        x[t] = v;
        w = v;
        nw = nn;
      }
      if (k + 1 < LW2) fused_out_chunk<LW2, SHIFT, DALL, DIRECT>(cx, o, n, k + 1);
    }
  }
  // repair: re-run segments whose incoming value was wrong until nothing changes
#ifdef FPV_FUSED_PROF
  uint32_t prof_rounds = 0;
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
#ifdef FPV_FUSED_PROF
    prof_rounds++;
#endif
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
#ifdef FPV_FUSED_PROF
  if (lane == 0) atomicAdd(&g_fused_rounds[prof_rounds < 7 ? prof_rounds : 7], 1ull);
#endif
  if (cgmask == 0xffffffffu) {
    // clean the guard bytes: this row is the next row's n / nw and the next write-out's input
#pragma unroll
    for (int t = 0; t < L; t++) x[t] &= kHiBytes;
  } else {
    // one of the two frames is not ClampedGradient-predicted: its row is the residual row r = c + nw
    const uint32_t keep = cgmask & kHiBytes;
#pragma unroll
    for (int t = 0; t < L; t++) x[t] = (x[t] & keep) | ((c[t] + (t == 0 ? nw_in : n[t - 1])) & ~cgmask);
  }
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
  // The bulk stores of row y-1 go last: fence.proxy.async waits for every shared-memory write of the warp, and by
  // now the OUT row's stores have long completed (right after pass 1 the fence took 8 % of all stall samples).
  fused_out_store<LW2, DIRECT>(cx);
}

// LW2:   a lane's segment is L = 8 LW2 px (LW2 = ceil(W / 256)).
// FULL:  W == 32 L (every lane owns a complete segment).
// SHIFT: UnextractFrame with a non-zero shift is fused into the write-out.
// SPLIT: the two 16-bit lanes are the left and the right half of ONE frame (widths 1281..2560).
template <int LW2, bool FULL, bool SHIFT, bool SPLIT = false, bool DIRECT = false, int K0T = FPV_PAIR_K0,
          int G = (FPV_PAIR_G ? FPV_PAIR_G : (LW2 == 4 ? 16 : 8))>
__global__ void __launch_bounds__(32 * kFusedWarps, fused_min_ctas(LW2, DIRECT)) k_decode_fused(const PairParams p) {
  extern __shared__ __align__(128) uint8_t fsm[];
  constexpr int L = 8 * LW2;
  constexpr uint32_t RB = 32 * L;
  using C = FusedCtx<LW2, DIRECT>;
  static_assert(!DIRECT || LW2 % 2 == 0, "256-bit stores need 16 columns per lane and store");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t unit = blockIdx.x * kFusedWarps + warp;
  // pair mode: frames (fA, fA + 1); split mode: the two halves of frame fA
  const uint32_t fA = SPLIT ? unit : 2 * unit;
  if (fA >= p.n) return;                            // warps are independent: no CTA-wide barrier anywhere
  const uint32_t fB = SPLIT ? fA : (fA + 1 < p.n ? fA + 1 : fA);   // odd tail: the pair is (A, A), B is not stored
  const uint32_t offB = SPLIT ? p.W : 0u;           // split mode: "frame B" starts W columns into the row
  const uint32_t flA = p.flags[fA], flB = p.flags[fB];
  const uint32_t W = p.W, H = p.H;

  C cx;
  cx.sm0 = smem_u32(fsm) + (uint32_t)warp * C::kBytes;
  cx.W = W; cx.H = H; cx.stride = p.stride; cx.lane = lane;
  cx.lowA = !(flA & kFlagNoLow) && p.low != nullptr;
  cx.lowB = !(flB & kFlagNoLow) && p.low != nullptr;
  const bool delA = (flA & kFlagDelta) && p.ddup != nullptr, delB = (flB & kFlagDelta) && p.ddup != nullptr;
  cx.dmask = (delA ? 0x0000ffffu : 0u) | (delB ? 0xffff0000u : 0u);
  cx.cgmask = ((flA & kFlagCG) ? 0x0000ffffu : 0u) | ((flB & kFlagCG) ? 0xffff0000u : 0u);
  cx.store_b = SPLIT || fB != fA;
  cx.s1A = p.high + (uint64_t)fA * p.P;
  cx.s1B = p.high + (uint64_t)fB * p.P + offB;
  cx.s2A = p.low + (uint64_t)fA * p.P;                // dereferenced only if lowA / lowB
  cx.s2B = p.low + (uint64_t)fB * p.P + offB;
  cx.s2D = p.ddup;
  cx.col0 = (uint32_t)lane * L;
  cx.oA = p.out + (uint64_t)fA * p.P + (DIRECT ? cx.col0 : 0u);
  cx.oB = p.out + (uint64_t)fB * p.P + offB + (DIRECT ? cx.col0 : 0u);
  cx.i1_row = cx.i1_slot = cx.il_row = cx.il_slot = cx.id_row = cx.id_slot = 0;
  cx.c1_slot = cx.c1_par = cx.cl_slot = cx.cl_par = cx.cd_slot = cx.cd_par = 0;
  const bool do_swap = p.unextract && p.big_endian;
  cx.shmul = 1u << ((32 - p.shift) & 31);
  cx.um = (0xffffu >> (p.shift & 31)) * 0x00010001u;
  // output words: two consecutive pixels of one frame out of two pair-form registers; the
  // UnextractFrame byte swap (.cc:857-860) is folded into the selector
  cx.selA = do_swap ? 0x4501u : 0x5410u;
  cx.selB = do_swap ? 0x6723u : 0x7632u;

  if (lane == 0) {
    for (int i = 0; i < kFusedR1 + kFusedLow + kFusedDel; i++) mbar_init(cx.sm0 + C::kBars + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // The write-out is branch-free: a frame without a low plane reads zeros from ring rows that TMA never
  // writes; without delta the delta words are masked off.
  if (!cx.lowA || !cx.lowB)
    for (uint32_t i = lane; i < kFusedLow * RB / 4; i += 32) {
      const uint32_t s = i / (RB / 4), o = i % (RB / 4);
      if (!cx.lowA) sts32(cx.sm0 + C::kLow + s * 2 * RB + 4 * o, 0u);
      if (!cx.lowB) sts32(cx.sm0 + C::kLow + s * 2 * RB + RB + 4 * o, 0u);
    }
  if (!cx.dmask)
    for (uint32_t i = lane; i < kFusedDel * RB; i += 32) sts32(cx.sm0 + C::kDel + 4 * i, 0u);
  __syncwarp();

  // prologue: rows 0, 1 of the residuals, rows 0, 1 of the low planes, row 0 of the delta image
  cx.issue_r1(); cx.issue_r1();
  cx.issue_low(); cx.issue_low();
  cx.issue_del();

  const bool lane_valid = FULL || (uint32_t)lane * L < W;
  const uint32_t vmask = lane_valid ? (cx.cgmask & kHiBytes) : 0u;   // compares look at the bytes, not the guards
  const uint32_t last_lane = FULL ? 31u : (W - 1) / L, last_t = FULL ? (uint32_t)(L - 1) : (W - 1) % L;
  uint32_t ra[L], rb[L];                    // finished rows, alternating roles
  uint32_t last_prev = 0, last_prev2 = 0;   // h[y-1][W-1], h[y-2][W-1] of both frames

  // row 0 is not predicted (.cc:327 starts at W + 1)
  cx.load_residuals(ra);
  {
    uint32_t v = ra[L - 1];
    if (!FULL) {
#pragma unroll
      for (int k = 0; k < 2 * LW2; k++)
        if ((uint32_t)(4 * k + 3) == last_t) v = ra[4 * k + 3];
    }
    last_prev = __shfl_sync(0xffffffffu, v, (int)last_lane);
  }
  // a refill of the three rings follows every write-out; row 0 has no write-out before it
  __syncwarp();
  cx.refill(false, DIRECT);

  if (cx.cgmask == 0) {
    // neither frame is ClampedGradient-predicted: rows are the residual rows
    for (uint32_t y = 1; y < H; y += 2) {
      cx.load_residuals(rb);
      fused_out_row<LW2, SHIFT, DIRECT>(cx, ra);
      if (y + 1 < H) {
        cx.load_residuals(ra);
        fused_out_row<LW2, SHIFT, DIRECT>(cx, rb);
      }
    }
  } else if (cx.dmask == 0xffffffffu) {
    for (uint32_t y = 1; y < H; y += 2) {
      fused_chain_row<LW2, FULL, SHIFT, SPLIT, K0T, G, true, DIRECT>(cx, ra, rb, y, vmask, last_lane, last_t, last_prev, last_prev2);
      if (y + 1 < H)
        fused_chain_row<LW2, FULL, SHIFT, SPLIT, K0T, G, true, DIRECT>(cx, rb, ra, y + 1, vmask, last_lane, last_t, last_prev,
                                                               last_prev2);
    }
  } else {
    for (uint32_t y = 1; y < H; y += 2) {
      fused_chain_row<LW2, FULL, SHIFT, SPLIT, K0T, G, false, DIRECT>(cx, ra, rb, y, vmask, last_lane, last_t, last_prev, last_prev2);
      if (y + 1 < H)
        fused_chain_row<LW2, FULL, SHIFT, SPLIT, K0T, G, false, DIRECT>(cx, rb, ra, y + 1, vmask, last_lane, last_t, last_prev,
                                                                last_prev2);
    }
  }
  // the last row's write-out: row H - 1 is in ra if H - 1 is even
  if ((H - 1) & 1u) fused_out_row<LW2, SHIFT, DIRECT>(cx, rb);
  else fused_out_row<LW2, SHIFT, DIRECT>(cx, ra);
  if (lane == 0) bulk_wait0();
}

template <int LW2, bool SPLIT = false>
static cudaError_t launch_fused(const PairParams& p, bool full, cudaStream_t stream) {
  const bool shift = p.unextract && p.shift != 0;
  const uint32_t units = SPLIT ? p.n : (p.n + 1) / 2;
  const int blocks = (int)((units + kFusedWarps - 1) / kFusedWarps);
  // 256-bit stores straight from registers where a lane's piece of the row is a whole number of 32-byte sectors
  const bool direct = LW2 % 2 == 0 && reinterpret_cast<uintptr_t>(p.out) % 32 == 0 && getenv("FPV_FUSED_NO_DIRECT") == nullptr;
  const size_t smem = fused_smem_bytes(LW2, direct);
  cudaError_t e = cudaSuccess;
#define FPV_LAUNCH_FUSED(F, S, D)                                                                                            \
  do {                                                                                                                       \
    e = cudaFuncSetAttribute(k_decode_fused<LW2, F, S, SPLIT, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
    if (e == cudaSuccess) k_decode_fused<LW2, F, S, SPLIT, D><<<blocks, 32 * kFusedWarps, smem, stream>>>(p);                \
  } while (0)
#define FPV_LAUNCH_FUSED_D(D)                       \
  do {                                              \
    if (full && shift) FPV_LAUNCH_FUSED(true, true, D);   \
    else if (full) FPV_LAUNCH_FUSED(true, false, D);      \
    else if (shift) FPV_LAUNCH_FUSED(false, true, D);     \
    else FPV_LAUNCH_FUSED(false, false, D);               \
  } while (0)
  if constexpr (LW2 % 2 == 0) {
    if (direct) FPV_LAUNCH_FUSED_D(true);
    else FPV_LAUNCH_FUSED_D(false);
  } else {
    (void)direct;
    FPV_LAUNCH_FUSED_D(false);
  }
#undef FPV_LAUNCH_FUSED_D
#undef FPV_LAUNCH_FUSED
  return e == cudaSuccess ? cudaGetLastError() : e;
}

}  // namespace fpv

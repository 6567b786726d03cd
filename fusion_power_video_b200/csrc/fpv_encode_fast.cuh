// fpv_encode_fast.cuh -- the fused, TMA-staged encode kernel (k_encode_fast).
//
// One launch = Frame ctor + Frame::Predict (fusion_power_video.cc:370-451,
// :491-593) for a list of frames under ASSUMED per-frame flags, plus the
// statistics the reference's two heuristics need (.cc:447-449, :522-531,
// :550-562).
//
// Work decomposition
//   task   = a whole frame when that keeps the persistent CTAs (at most 4 per SM) evenly loaded (the normal case),
//            else a band of `band_rows` rows (32 rows for short frame lists); CTAs take tasks round-robin.
//   stage  = up to 4 (or 2) consecutive rows of the band = ONE contiguous flat range
//            of the raw frame (and of the delta image), fetched with
//            cp.async.bulk (TMA, 1-D) into a shared-memory ring and signalled
//            through an mbarrier.  The range starts 8 pixels early: with the
//            reference's flat indexing (.cc:556-558) the west neighbour of
//            column 0 is the last pixel of the previous row, which is exactly
//            what precedes the row in memory.
//   strip  = 256 columns of a row, owned by one consumer warp (8 px per lane).
//            The warp walks down the rows of the band keeping the previous
//            row's (post-delta) high bytes in registers, so north / north-west
//            neighbours never touch memory again; a band that does not start at
//            row 0 is primed by a 1-row halo stage.
//   service warp = the last warp of the CTA.  Lane 0 is the TMA producer; all 32
//            lanes then take the delta-decision samples (every 15th pixel of the
//            flat frame, .cc:526-531) straight out of the landed stage, so the
//            consumer warps never see them.
//
// Arithmetic is done on two pixels per register ("q form" / "S form", see
// fpv_common.cuh) and balanced over the ALU and the FMA pipe (make_hs_lo, sub_fma).  Per row and lane:
// LDS.128 + LDS.32 of raw and of delta, ~54 ALU-pipe and ~24 FMA-pipe instructions, 2x STG.64 (+ 1 STG.16
// of preview every 4th row) and two predicated shared atomics for the ClampedGradient decision histograms.
// The delta decision needs no histogram in the common case: only 8 bit
// counters per frame (see frame_decide).
//
// Decisions are taken inside this kernel (round 2; before, k_encode_init and two k_decide launches sat between the
// passes).  Every warp adds its statistics of a task into a CTA-level table in shared memory.  With whole-frame
// tasks that table IS the frame's statistics: the last warp of the CTA to finish evaluates the reference's two
// integer heuristics right out of shared memory (decide_core), writes the flags byte when the assumption held or
// puts the frame on the next pass's list -- no global atomics, no fence, no init kernel.  With band tasks the
// tables are merged in global memory behind a ticket (task_done).  What this costs and what was tried first
// (a fence per warp: +25 %; the decision in a non-inlined function: +20 %; preview prediction in the service
// warp: slower than the separate k_finalize16) is in profiles/r02_encode.md.
#pragma once

#include "fpv_internal.h"
#include "fpv_ptx.cuh"

namespace fpv {

constexpr int kStripPx = 256;       // columns per consumer warp (8 pixels per lane)
constexpr int kMaxRowsPerStage = 4; // rows per stage: 4 (one preview row group) or 2
constexpr int kHaloPx = 8;          // pixels copied before a stage's first pixel (16 B)
constexpr int kWarpHistWords = 512; // hist_a, hist_b: 256 u32 counters each, warp-private
constexpr int kWarpScratchBytes = kWarpHistWords * 4;
constexpr int kCtaStatWords = 512 + 8 + 2;   // CTA-level statistics of one task: hist_a | hist_b | dbits | low_or | arrivals

struct FastParams {
  const uint16_t* frames;
  const uint16_t* delta;        // nullptr: no delta frame
  FrameStat* stats;
  const uint32_t* list;         // redo passes: the frames to redo; nullptr in pass 0 (all n frames, assumed = *guess)
  const uint32_t* count;        // redo passes: length of list
  uint32_t n;                   // frames of the batch
  uint32_t f0, nf;              // pass 0: this launch covers the frames [f0, f0 + nf) (a batch may be cut in two launches)
  uint32_t* next_list;          // where frame_decide puts frames whose assumption was wrong (nullptr: last pass)
  uint32_t* next_count;
  const uint32_t* guess_in;     // flags (bit 0 delta, bit 1 cg) pass 0 assumes for every frame ...
  uint32_t* guess_out;          // ... and where the batch's last frame leaves its own for the next call (another word:
                                // CTAs of this launch may still be reading guess_in)
  uint8_t* high;
  uint8_t* low;
  uint8_t* preview_raw;         // un-predicted preview planes (k_finalize16 predicts them)
  uint8_t* flags;
  uint32_t W, H;
  uint64_t P, PP;
  uint32_t PW;
  int shift;
  uint32_t band_rows;           // multiple of 4
  uint32_t bands;               // bands per frame
  uint32_t stages;              // ring depth
  uint32_t rows_per_stage;      // 2 or 4
  uint32_t stage_bytes;         // bytes of one plane of one stage: (rows_per_stage * W + 8) * 2
  uint32_t compute_warps;       // ceil(W / 256)
  QConst qc;                    // make_qconst(mode, shift)
};

// Walks the stage sequence of one CTA: tasks blockIdx.x, +gridDim.x, ...; per
// task an optional 1-row halo stage (row y0-1) then 4-row stages.  Producer,
// sampler and consumers each run their own copy and therefore agree on the
// running stage number.
struct StageCursor {
  uint32_t t, f, y0, y1, ys, rps;
  uint32_t tseq = 0;            // tasks this CTA has started before the current one (parity selects the CTA-level table)
  uint32_t band_rows, bands;    // of this launch (short frame lists use shorter bands, see the kernel)
  __device__ __forceinline__ bool load_task(const FastParams& p, uint32_t total) {
    if (t >= total) return false;
    f = p.list ? p.list[t / bands] : p.f0 + t / bands;
    const uint32_t b = t % bands;
    y0 = b * band_rows;
    y1 = min(p.H, y0 + band_rows);
    ys = y0 > 0 ? y0 - 1 : 0;   // one halo row: the north neighbours of the band's first row
    return true;
  }
  __device__ __forceinline__ bool halo() const { return ys < y0; }
  __device__ __forceinline__ uint32_t nrows() const {
    return ys < y0 ? min(rps, y0 - ys) : min(rps, y1 - ys);
  }
  __device__ __forceinline__ bool last_of_task() const { return ys >= y0 && ys + nrows() >= y1; }
  __device__ __forceinline__ bool advance(const FastParams& p, uint32_t total) {
    ys += nrows();
    if (ys < y1) return true;
    t += gridDim.x;
    tseq++;
    return load_task(p, total);
  }
};

// Position in the shared-memory ring: slot index and the parity of its current use.
struct RingPos {
  uint32_t slot = 0, phase = 0;
  __device__ __forceinline__ void next(uint32_t S) {
    if (++slot == S) { slot = 0; phase ^= 1u; }
  }
};

// Per-lane running state while walking down a strip (all S form).
struct StripState {
  uint32_t ph[4];   // previous row, post-delta high bytes
  uint32_t pw[4];   // previous row shifted one pixel west (= this row's north-west)
  uint32_t accA, accB;  // preview 4x4 box sums of the RAW high bytes (two preview pixels per lane)
  uint32_t orl;     // OR of the split pixels (low bytes = OR of low plane)
  uint32_t kc;      // offset (0..30) from this lane's first pixel to the next CG-decision sample
};

// Two raw pixels -> the two register forms a row is computed in, in as few ALU-pipe instructions
// as the mode allows (the kernel is bound by the integer ALU pipe, so shifts are done as IMAD on
// the FMA pipe and each mask is fused with the OR that follows it into one LOP3):
//   hs1  S form of the RAW high bytes with bit 16 set: (q & 0xff00ff00) | 0x10000.  Bit 16 is the
//        "+ 2^16" of the delta subtraction below; without delta it is junk <= 1 of lane 1.
//   lof  q | 0x01000100: bytes 0 / 2 are the low bytes, every lane is >= 256.
template <int MODE>
__device__ __forceinline__ void make_hs_lo(uint32_t x, const QConst& c, uint32_t& hs1, uint32_t& lof) {
  if (MODE == kLEs || MODE == kLE8 || MODE == kLEbig) {
    const uint32_t sh = x * c.sh_b;                       // sh_b = 1 << shift (mod 2^32): IMAD
    hs1 = (sh & c.m_b) | 0x00010000u;                     // m_b = m_a & 0xff00ff00
    lof = (sh & c.m_a) | 0x01000100u;
  } else {
    const uint32_t q = make_q2<MODE>(x, c);
    hs1 = (q & kHiBytes) | 0x00010000u;
    lof = q | 0x01000100u;
  }
  // one LOP3 each; opaque so that later uses (OR of the low bytes, box sums) take the finished
  // value instead of re-deriving it from the masks
  asm("" : "+r"(hs1));
  asm("" : "+r"(lof));
}

// a - b on the FMA pipe (IMAD); the compiler would pick the ALU pipe's IADD3.
__device__ __forceinline__ uint32_t sub_fma(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("mad.lo.u32 %0, %1, 0xffffffff, %2;" : "=r"(r) : "r"(b), "r"(a));
  return r;
}

// One row of one strip.  FIRSTROWS = true compiles the row-0 / row-1 special
// cases and the halo (not-owned) row; the steady state uses FIRSTROWS = false.
// DC = true: the frame is known to use delta and ClampedGradient (the common case), the two
// flags are compile-time constants.
template <int MODE, bool FIRSTROWS, bool FULL, bool DC>
__device__ __forceinline__ void fast_row(
    StripState& st, const uint32_t raw_a /* shared addr of this lane's 8 px */, const uint32_t del_a,
    const QConst qc, const bool use_delta_rt, const bool use_cg_rt, const bool active, const bool own,
    const uint32_t y, const bool emit_preview, const uint32_t c0, const uint32_t w31,
    uint8_t* __restrict__ out_high, uint8_t* __restrict__ out_low, uint8_t* __restrict__ out_prev,
    const uint32_t hist_a) {
  const bool use_delta = DC || use_delta_rt, use_cg = DC || use_cg_rt;
  uint32_t hr[4], hs[4], lo[4], w[4];
  uint32_t hw;  // S form of the two pixels west of this lane's first pixel
  {
    // lanes past the end of the row (FULL == false only) read whatever follows
    // in shared memory; nothing derived from it is ever stored
    const uint4 x = lds128(raw_a);
    const uint32_t xw = lds32(raw_a - 4);
    uint32_t unused;
    make_hs_lo<MODE>(x.x, qc, hr[0], lo[0]);
    make_hs_lo<MODE>(x.y, qc, hr[1], lo[1]);
    make_hs_lo<MODE>(x.z, qc, hr[2], lo[2]);
    make_hs_lo<MODE>(x.w, qc, hr[3], lo[3]);
    make_hs_lo<MODE>(xw, qc, hw, unused);
  }
  // statistics of the RAW planes: OR of the low bytes (bytes 0 / 2 of lof), 4x4 box sums of the
  // high bytes (bytes 1 / 3 of hs1)
  const bool counted = !FIRSTROWS || own;
  if (counted) {
    st.orl |= (lo[0] | lo[1]) | (lo[2] | lo[3]);
    st.accA = __dp4a(hr[0], 0x01000100u, __dp4a(hr[1], 0x01000100u, st.accA));
    st.accB = __dp4a(hr[2], 0x01000100u, __dp4a(hr[3], 0x01000100u, st.accB));
  }
  if (use_delta) {
    // Bytes wrap independently (.cc:534-537).  High: bits 8-15 / 24-31 of
    // (q & HI) + 2^16 - (d & HI); bit 16 is junk.  Low: bytes 0 / 2 of
    // (q | 0x0100 per lane) - (d & LO) = lof - d + (d & HI); the rest is junk the packing drops.
    const uint4 d = lds128(del_a);
    const uint32_t dw = lds32(del_a - 4);
    const uint32_t dd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint32_t dh = dd[j] & kHiBytes;
      hs[j] = sub_fma(hr[j], dh);
      lo[j] = lo[j] - dd[j] + dh;
    }
    hw = sub_fma(hw, dw & kHiBytes);
  } else {
#pragma unroll
    for (int j = 0; j < 4; j++) hs[j] = hr[j];
  }
  w[0] = __funnelshift_l(hw, hs[0], 16);
  w[1] = __funnelshift_l(hs[0], hs[1], 16);
  w[2] = __funnelshift_l(hs[1], hs[2], 16);
  w[3] = __funnelshift_l(hs[2], hs[3], 16);

  if (counted) {
    uint32_t res[4];
#pragma unroll
    for (int j = 0; j < 4; j++) res[j] = cg_residual_s(hs[j], st.ph[j], w[j], st.pw[j]);
    if (FIRSTROWS) {
      if (y == 0) {
#pragma unroll
        for (int j = 0; j < 4; j++) res[j] = hs[j];
      } else if (y == 1 && c0 == 0) {
        // flat index W (row 1, column 0) is copied, not predicted (.cc:566, :572)
        res[0] = (res[0] & 0xffff0000u) | (hs[0] & 0x0000ffffu);
      }
    }
    const uint2 res8 = make_uint2(pack_hi(res[0], res[1]), pack_hi(res[2], res[3]));
    const uint2 h8 = make_uint2(pack_hi(hs[0], hs[1]), pack_hi(hs[2], hs[3]));
    if (FULL || active) {
      stg64_cs(out_high, use_cg ? res8 : h8);
      if (mode_has_low(MODE)) stg64_cs(out_low, make_uint2(pack_lo(lo[0], lo[1]), pack_lo(lo[2], lo[3])));
    }
    // ---- ClampedGradient decision samples: flat index == W+1 (mod 31), >= W+1,
    //      a = post-delta high byte, b = a - CG (.cc:554-562).  A lane's 8 pixels
    //      hold at most one sample; it is picked out of the packed registers.
    if ((!FIRSTROWS || y >= 1) && st.kc < 8 && (FULL || active)) {
      const uint32_t a = __byte_perm(h8.x, h8.y, st.kc) & 0xffu;
      const uint32_t b = __byte_perm(res8.x, res8.y, st.kc) & 0xffu;
      red_shared_inc(hist_a + 4 * a);
      red_shared_inc(hist_a + 1024 + 4 * b);
    }
    if (emit_preview) {
      // the un-predicted preview pixels (.cc:500-512); k_finalize16 applies ClampedGradient to the preview plane
      if (FULL || active) {
        const uint32_t pv = ((st.accA >> 4) & 0xfeu) | (((st.accB >> 4) & 0xfeu) << 8);
        *reinterpret_cast<uint16_t*>(out_prev) = (uint16_t)pv;
      }
      st.accA = 0;
      st.accB = 0;
    }
  }
  st.kc = min(st.kc - w31, st.kc - w31 + 31u);   // (kc - W) mod 31; one of the two wrapped around
#pragma unroll
  for (int j = 0; j < 4; j++) { st.ph[j] = hs[j]; st.pw[j] = w[j]; }
}

// EstimateEntropy (.cc:235-244) on a 256-bin histogram spread over a warp (lane l holds bins 8 l .. 8 l + 7), with the
// reference's `int` accumulator truncation (sums are formed mod 2^32 and reinterpreted as int32; see block_entropy256).
__device__ __forceinline__ uint64_t warp_entropy256(const uint32_t (&v)[8]) {
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) s += v[j];
  const uint32_t sum32 = __reduce_add_sync(0xffffffffu, s);
  const int64_t sum = (int64_t)(int32_t)sum32;
  if (sum == 0) return 0;
  const int log2sum = 63 - __clzll((long long)sum);
  uint32_t term = 0;
#pragma unroll
  for (int j = 0; j < 8; j++)
    if (v[j]) term += v[j] * (uint32_t)(log2sum - (31 - __clz(v[j])));
  const uint32_t acc32 = __reduce_add_sync(0xffffffffu, term);
  const uint64_t sum_of_logs = (uint64_t)(int64_t)(int32_t)acc32;
  return (1024ull * sum_of_logs) / (uint64_t)sum;
}

// The reference's two decisions for frame f from its complete statistics (held by the calling warp: lane l has bins
// 8 l .. 8 l + 7 of both ClampedGradient histograms, lanes 0..7 the eight bit counters of the delta decision).
// USE_DELTA <=> EstimateEntropy(hist of every 15th raw high byte) > 0 (see the comment on k_decide in fpv_encode.cu),
// provable from the bit counters except for near-constant frames, where the exact histogram is rebuilt here.
// If the pass's assumption held, the frame is final: its flags byte is written.  Otherwise it goes onto the next
// pass's list with the corrected assumption.
template <int MODE>
__device__ __forceinline__ void decide_core(const FastParams& p, const uint32_t f, const uint32_t assumed, const int lane,
                                            const uint32_t dbit, const uint32_t (&va)[8], const uint32_t (&vb)[8],
                                            const uint32_t low_or) {
  FrameStat& st = p.stats[f];
  uint32_t dec_delta = 0;
  if (p.delta != nullptr) {
    const uint64_t N = (p.P + 14) / 15;
    const uint64_t c = lane < 8 ? (uint64_t)dbit : 0;
    const uint64_t m = c < N - c ? c : N - c;
    const bool proven = __any_sync(0xffffffffu, lane < 8 && 1024 * m >= N);
    const bool constant = __all_sync(0xffffffffu, lane >= 8 || m == 0);
    if (proven) dec_delta = 1;
    else if (constant) dec_delta = 0;
    else {
      // rare: a near-constant plane with a few strays; exact 256-bin histogram of the samples (hist_d is all zero)
      const uint16_t* img = p.frames + (uint64_t)f * p.P;
      for (uint64_t i = 15ull * (uint32_t)lane; i < p.P; i += 15ull * 32) {
        uint32_t h, l;
        split1<MODE>(img[i], p.shift, h, l);
        atomicAdd(&st.hist_d[h], 1u);
      }
      __threadfence();
      __syncwarp();
      uint32_t v[8];
#pragma unroll
      for (int j = 0; j < 8; j++) {
        v[j] = __ldcg(&st.hist_d[8 * lane + j]);
        st.hist_d[8 * lane + j] = 0;
      }
      dec_delta = warp_entropy256(v) > 0 ? 1u : 0u;
    }
  }
  uint32_t want = dec_delta | (assumed & 2u);
  if (dec_delta == (assumed & 1u)) {
    // the ClampedGradient histograms were taken on the right (post-delta) plane: .cc:564
    const uint64_t e_a = warp_entropy256(va), e_b = warp_entropy256(vb);
    want = dec_delta | ((e_b < e_a) ? 2u : 0u);
  }
  if (lane == 0) {
    if (want != assumed && p.next_list != nullptr) {
      st.assumed = want;
      p.next_list[atomicAdd(p.next_count, 1u)] = f;
    } else {
      const uint32_t nolow = mode_has_low(MODE) ? ((low_or & kLoBytes) == 0 ? kFlagNoLow : 0) : kFlagNoLow;
      p.flags[f] = (uint8_t)(want | nolow);
      if (f == p.n - 1) *p.guess_out = want;   // the next batch starts from this assumption
    }
  }
  __syncwarp();
}

// End of a task for one warp.  The warp's statistics are already in the CTA-level table of the task (shared memory);
// the LAST of the CTA's warps to get here (roll call: shared atomic, CTA-scope fence) closes the task.
//   * bands == 1 (the normal case: a task is a whole frame): the table IS the frame's statistics.  The decision is
//     taken right here, out of shared memory -- no global atomics, no fence, nothing to wait for.
//   * bands > 1 (few frames, or the redo passes): the table is added to the frame's global statistics with
//     value-returning atomics; once their results are back the additions have been performed at L2, and only then
//     the frame's ticket is taken; the CTA that takes the last of the `bands` tickets decides from the global record
//     and clears it.  (A memory fence per warp and task instead cost a quarter of the kernel's time, and any wait
//     of a pipeline warp stalls its whole CTA within one stage: see profiles/r02_encode.md.)
//   ctab: shared address of the task's table (hist_a[256] | hist_b[256] | dbits[8] | low_or | arrivals)
template <int MODE>
__device__ __forceinline__ void task_done(const FastParams& p, const uint32_t f, const uint32_t assumed, const uint32_t bands,
                                          const uint32_t ctab, const uint32_t warps, const int lane) {
#ifdef FPV_ABL_NO_TICKET
  return;                                              // timing experiment only
#endif
  __threadfence_block();
  __syncwarp();
  uint32_t arrived = 0;
  if (lane == 0) arrived = atoms_add(ctab + 4 * (512 + 8 + 1), 1u);
  arrived = __shfl_sync(0xffffffffu, arrived, 0);
  if (arrived + 1 != warps) return;
  // ---- last warp of the CTA on this task ----
  __threadfence_block();
  FrameStat& st = p.stats[f];
  uint32_t va[8], vb[8];
  if (bands == 1) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      va[j] = lds32(ctab + 4 * (8 * lane + j));
      vb[j] = lds32(ctab + 4 * (256 + 8 * lane + j));
      sts32(ctab + 4 * (8 * lane + j), 0);
      sts32(ctab + 4 * (256 + 8 * lane + j), 0);
    }
    const uint32_t dbit = lane < 8 ? lds32(ctab + 4 * (512 + lane)) : 0u;
    const uint32_t low_or = lds32(ctab + 4 * (512 + 8));
    __syncwarp();
    if (lane < 10) sts32(ctab + 4 * (512 + lane), 0);   // bit counters, low_or, arrivals
    decide_core<MODE>(p, f, assumed, lane, dbit, va, vb, low_or);
    return;
  }
  uint32_t dep = 0;
#pragma unroll 4
  for (uint32_t i = (uint32_t)lane; i < 512; i += 32) {
    const uint32_t v = lds32(ctab + 4 * i);
    if (v) {
      sts32(ctab + 4 * i, 0);
      dep |= atomicAdd(&st.hist_a[i], v);    // hist_a, hist_b are contiguous: 512 bins
    }
  }
  if (lane < 8) {
    const uint32_t v = lds32(ctab + 4 * (512 + lane));
    if (v) {
      sts32(ctab + 4 * (512 + lane), 0);
      dep |= atomicAdd(&st.dbits[lane], v);
    }
  }
  if (lane == 8) {
    const uint32_t v = lds32(ctab + 4 * (512 + 8));
    sts32(ctab + 4 * (512 + 8), 0);
    sts32(ctab + 4 * (512 + 8 + 1), 0);
    if (v) dep |= atomicOr(&st.low_or, v);
  }
  // the ticket is issued after every atomic above has returned (data dependency through `dep`)
  dep = __reduce_or_sync(0xffffffffu, dep);   // cannot issue before every lane's atomics have written their result
  uint32_t t = 0;
  if (lane == 0)
    asm volatile("atom.global.add.u32 %0, [%1], 1;   // after %2" : "=r"(t) : "l"(&st.tickets), "r"(dep) : "memory");
  t = __shfl_sync(0xffffffffu, t, 0);
  if (t + 1 != bands) return;
  // ---- last ticket of the frame: decide from the global record, and clear it for the next pass / call ----
  __threadfence();
#pragma unroll
  for (int j = 0; j < 8; j++) {
    va[j] = __ldcg(&st.hist_a[8 * lane + j]);
    vb[j] = __ldcg(&st.hist_b[8 * lane + j]);
    st.hist_a[8 * lane + j] = 0;
    st.hist_b[8 * lane + j] = 0;
  }
  const uint32_t dbit = lane < 8 ? __ldcg(&st.dbits[lane]) : 0u;
  const uint32_t low_or = __ldcg(&st.low_or);
  __syncwarp();
  if (lane < 8) st.dbits[lane] = 0;
  if (lane == 0) {
    st.low_or = 0;
    st.tickets = 0;
  }
  decide_core<MODE>(p, f, assumed, lane, dbit, va, vb, low_or);
}

// Shared memory: [ring: stages x {raw stage, delta stage}] [two CTA-level task tables] [full barriers] [empty barriers]
// FULL: xsize is a multiple of 256, every lane of every consumer warp owns pixels.
// PASS0: all n frames of the batch under the guessed flags; otherwise the frames of p.list under their own.
#ifndef FPV_FAST_MAXNREG
#define FPV_FAST_MAXNREG 80
#endif
// MAXR: register cap.  80 keeps four CTAs of up to six warps on an SM (widths up to 1280); frames of more than 1792
// columns (nine warps per CTA) get 72 so that THREE of their CTAs fit (288 threads x 72 registers).
template <int MODE, bool FULL, int RPS, bool PASS0, int MAXR = FPV_FAST_MAXNREG>
__global__ void __maxnreg__(MAXR) k_encode_fast(const FastParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  if (!PASS0 && *p.count == 0) return;   // a redo pass with nothing to redo (the normal case): gone in a microsecond
  const uint32_t S = p.stages;
  const uint32_t slot_bytes = 2 * p.stage_bytes;
  const int NW = (int)p.compute_warps;
  uint8_t* ctabs = smem + (size_t)S * slot_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ctabs + 2 * kCtaStatWords * 4);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + S);
  const uint32_t ring0 = smem_u32(smem);
  const uint32_t ctab0 = smem_u32(ctabs);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t W = p.W;
  const QConst qc = p.qc;   // from the host: each mask is a constant-bank operand of its LOP3, not a derived register

  for (uint32_t i = threadIdx.x; i < (2 * kCtaStatWords * 4) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(ctabs)[i] = 0;
  if (threadIdx.x == 0) {
    for (uint32_t i = 0; i < S; i++) {
      mbar_init(full0 + 8 * i, 1);
      mbar_init(empty0 + 8 * i, NW + 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // A short frame list (the redo passes: the few frames whose flags were guessed wrong) would leave
  // most CTAs without a task and the rest with one long one; it is cut into 32-row bands instead.
  const uint32_t nframes = PASS0 ? p.nf : *p.count;
  const bool short_list = nframes * p.bands < gridDim.x && p.band_rows > 32;
  StageCursor cur;
  cur.band_rows = short_list ? 32u : p.band_rows;
  cur.bands = short_list ? (p.H + 31u) / 32u : p.bands;
  const uint32_t total_tasks = nframes * cur.bands;
  const uint32_t guess = PASS0 ? (*p.guess_in & (p.delta != nullptr ? 3u : 2u)) : 0u;
  cur.t = blockIdx.x;
  cur.rps = RPS;
  bool more = cur.load_task(p, total_tasks);
  RingPos rp;        // ring position of the stage this warp works on

  if (warp == NW) {
    // ------------------- service warp: TMA producer + delta-decision sampler -------------------
    StageCursor ic = cur;  // issue cursor, runs S-1 stages ahead
    bool imore = more;
    RingPos ip;
    auto issue = [&]() {
      if (lane == 0) {
        const uint32_t slot = ip.slot;
        // wait until every reader of the slot's previous stage has released it
        mbar_wait(empty0 + 8 * slot, ip.phase ^ 1u);
        const uint32_t fl = PASS0 ? guess : p.stats[ic.f].assumed;
        const bool use_delta = p.delta != nullptr && (fl & 1u);
        const uint16_t* img = p.frames + (uint64_t)ic.f * p.P;
        // contiguous flat range [ys*W - 8, (ys+nrows)*W)
        const uint32_t px0 = ic.ys * W;  // < 2^30 (fpv_create bounds xsize * ysize)
        const uint32_t lead = ic.ys > 0 ? kHaloPx : 0;
        const uint32_t bytes = (ic.nrows() * W + lead) * 2;
        const uint32_t dst = ring0 + slot * slot_bytes + (kHaloPx - lead) * 2;
        mbar_arrive_expect_tx(full0 + 8 * slot, use_delta ? 2 * bytes : bytes);
        bulk_g2s(dst, img + px0 - lead, bytes, full0 + 8 * slot);
        if (use_delta) bulk_g2s(dst + p.stage_bytes, p.delta + px0 - lead, bytes, full0 + 8 * slot);
      }
      ip.next(S);
      imore = ic.advance(p, total_tasks);
    };
    for (uint32_t i = 0; i + 1 < S && imore; i++) issue();

    uint32_t acc[4] = {0, 0, 0, 0};  // per-lane counts of set bits {0,2},{1,3},{4,6},{5,7}, two u16 each
    while (more) {
      if (imore) issue();
      const uint32_t slot = rp.slot;
      mbar_wait(full0 + 8 * slot, rp.phase);
      if (!cur.halo()) {
        // samples of this stage: flat index % 15 == 0 (.cc:526-531), RAW high byte
        const uint32_t base = ring0 + slot * slot_bytes + kHaloPx * 2;
        const uint32_t npx = cur.nrows() * W;
        const uint32_t m = (cur.ys * W) % 15u;
        uint32_t c_lo = 0, c_hi = 0;  // four u8 counters each: bits 0..3, bits 4..7
        for (uint32_t pos = (m ? 15u - m : 0u) + 15u * (uint32_t)lane; pos < npx; pos += 15u * 32u) {
          const uint32_t v = make_q2<MODE>(lds16(base + 2 * pos), qc) >> 8;  // lane 1 is zero: v = high byte
          c_lo += ((v & 15u) * 0x00204081u) & 0x01010101u;
          c_hi += ((v >> 4) * 0x00204081u) & 0x01010101u;
        }
        acc[0] += c_lo & kLoBytes;
        acc[1] += (c_lo >> 8) & kLoBytes;
        acc[2] += c_hi & kLoBytes;
        acc[3] += (c_hi >> 8) & kLoBytes;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty0 + 8 * slot);
      if (cur.last_of_task()) {
        const uint32_t ctab = ctab0 + (cur.tseq & 1u) * (kCtaStatWords * 4);
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const uint32_t s0 = __reduce_add_sync(0xffffffffu, acc[k] & 0xffffu);
          const uint32_t s1 = __reduce_add_sync(0xffffffffu, acc[k] >> 16);
          // acc[0]: bits 0,2  acc[1]: bits 1,3  acc[2]: bits 4,6  acc[3]: bits 5,7
          const int b0 = (k & 1) + 4 * (k >> 1);
          if (lane == 0) {
            if (s0) atoms_add(ctab + 4 * (512 + b0), s0);
            if (s1) atoms_add(ctab + 4 * (512 + b0 + 2), s1);
          }
          acc[k] = 0;
        }
        task_done<MODE>(p, cur.f, PASS0 ? guess : p.stats[cur.f].assumed, cur.bands, ctab, (uint32_t)NW + 1, lane);
      }
      more = cur.advance(p, total_tasks);
      rp.next(S);
    }
    return;
  }

  // -------------------------------- consumers -----------------------------------
  const uint32_t sb = (uint32_t)warp * kStripPx;          // first column of this warp's strip
  const uint32_t c0 = sb + (uint32_t)lane * 8;
  const bool active = c0 < W;
  const uint32_t w31 = W % 31;
  const uint32_t rowb = W * 2;                            // bytes per row in a stage

  while (more) {
    // ---- start of a task ---------------------------------------------------------
    const uint32_t f = cur.f;
    const uint32_t assumed = PASS0 ? guess : p.stats[f].assumed;
    const bool use_delta = p.delta != nullptr && (assumed & 1u);
    const bool use_cg = (assumed & 2u) != 0;
    // The ClampedGradient decision samples (at most one per lane and row: 8 of a warp's 32 lanes) go straight into the
    // CTA-level table of the task: no warp-private histograms (round 1 / 2 kept 2 KB per warp and flushed them per
    // task; at 2048 columns that was the 16 KB that kept a third CTA off the SM).
    uint32_t ctab = ctab0 + (cur.tseq & 1u) * (kCtaStatWords * 4);
    asm volatile("" : "+r"(ctab));  // keep it in a register instead of recomputing it per row
    const uint32_t hist_a = ctab;
    const uint32_t bands = cur.bands;
    StripState st;
#pragma unroll
    for (int j = 0; j < 4; j++) { st.ph[j] = 0; st.pw[j] = 0; }
    st.accA = st.accB = st.orl = 0;
    uint32_t y = cur.ys;
    {
      // offset of the first pixel >= (y, c0) whose flat index is == W+1 (mod 31)
      const uint64_t i0 = (uint64_t)y * W + c0;
      const uint32_t r = (uint32_t)((i0 + 31ull * (W / 31 + 2) - (W + 1)) % 31);
      st.kc = r ? 31 - r : 0;
    }
    uint8_t* oh = p.high + (uint64_t)f * p.P + (uint64_t)y * W + c0;
    uint8_t* ol = mode_has_low(MODE) ? p.low + (uint64_t)f * p.P + (uint64_t)y * W + c0 : nullptr;
    uint8_t* op = p.preview_raw + (uint64_t)f * p.PP + (uint64_t)(y >> 2) * p.PW + (c0 >> 2);

    for (;;) {
      const uint32_t nrows = cur.nrows();
      const bool own = !cur.halo();
      const bool last = cur.last_of_task();
      const uint32_t slot = rp.slot;
      mbar_wait(full0 + 8 * slot, rp.phase);
      const uint32_t raw_lane = ring0 + slot * slot_bytes + kHaloPx * 2 + c0 * 2;  // pixel (y, c0)

      if (y >= 4 && nrows == RPS) {
        // steady state: RPS owned rows, none of them row 0 / row 1; stages start on even rows
        // (bands on multiples of 4), so a preview row group ends with the last row of a stage
        const bool group_end = RPS == 4 || (y & 2u);
        if (use_delta && use_cg) {
#pragma unroll
          for (int r = 0; r < RPS; r++) {
            fast_row<MODE, false, FULL, true>(st, raw_lane + r * rowb, raw_lane + r * rowb + p.stage_bytes, qc, true,
                                              true, active, true, y + r, r == RPS - 1 && group_end, c0, w31, oh, ol,
                                              op, hist_a);
            oh += W;
            if (mode_has_low(MODE)) ol += W;
          }
        } else {
#pragma unroll
          for (int r = 0; r < RPS; r++) {
            fast_row<MODE, false, FULL, false>(st, raw_lane + r * rowb, raw_lane + r * rowb + p.stage_bytes, qc,
                                               use_delta, use_cg, active, true, y + r, r == RPS - 1 && group_end, c0,
                                               w31, oh, ol, op, hist_a);
            oh += W;
            if (mode_has_low(MODE)) ol += W;
          }
        }
        if (group_end) op += p.PW;
        y += RPS;
      } else {
        for (uint32_t r = 0; r < nrows; r++) {
          fast_row<MODE, true, FULL, false>(st, raw_lane + r * rowb, raw_lane + r * rowb + p.stage_bytes, qc,
                                            use_delta, use_cg, active, own, y, (y & 3u) == 3u, c0, w31, oh, ol, op,
                                            hist_a);
          oh += W;
          if (mode_has_low(MODE)) ol += W;
          if ((y & 3u) == 3u) op += p.PW;
          y++;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty0 + 8 * slot);
      rp.next(S);
      more = cur.advance(p, total_tasks);
      if (last) break;
    }

    // ---- end of task: this warp's low-OR and histograms into the CTA-level table, then the task's roll call ----
    if (mode_has_low(MODE)) {
      // lanes past the end of the row accumulated garbage (see fast_row)
      const uint32_t orl = __reduce_or_sync(0xffffffffu, (FULL || active) ? (st.orl & kLoBytes) : 0u);
      if (lane == 0 && orl) atoms_or(ctab + 4 * (512 + 8), orl);
    }
    task_done<MODE>(p, f, assumed, bands, ctab, (uint32_t)NW + 1, lane);
  }
}

static inline size_t fast_smem_bytes(uint32_t W, int stages, int rows_per_stage) {
  const size_t stage_bytes = ((size_t)rows_per_stage * W + kHaloPx) * 2;
  const size_t warps = (W + kStripPx - 1) / kStripPx;
  (void)warps;
  return (size_t)stages * 2 * stage_bytes + 2 * kCtaStatWords * 4 + 2 * (size_t)stages * 8;
}

}  // namespace fpv

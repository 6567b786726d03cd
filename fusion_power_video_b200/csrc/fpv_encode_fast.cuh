// fpv_encode_fast.cuh -- the fused, TMA-staged encode kernel (k_encode_fast).
//
// One launch = Frame ctor + Frame::Predict (fusion_power_video.cc:370-451,
// :491-593) for a list of frames under ASSUMED per-frame flags, plus the
// statistics the reference's two heuristics need (.cc:447-449, :522-531,
// :550-562).
//
// Work decomposition
//   task   = (frame, band of `band_rows` rows; 32 rows for short frame lists); persistent CTAs (at most
//            4 per SM) take tasks round-robin.
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
// Everything per frame happens inside this kernel (round 2; before, k_encode_init, two k_decide launches and
// k_finalize16 cost 9 % of a step):
//   * decisions: every warp that finishes its part of a frame takes a ticket; the last one evaluates the
//     reference's two integer heuristics for the frame (frame_decide), writes the flags byte when the
//     assumption held, or puts the frame on the next pass's list, and clears the frame's statistics for the
//     next call (no init kernel);
//   * preview: the consumer warps write the un-predicted 4x4 box means into a small shared-memory ring, the
//     service warp applies ClampedGradient on the preview's own flat array of width W/4 (.cc:575-586) under the
//     same assumed flag and writes the final preview rows (no finalize kernel, no scratch plane in HBM).  A band
//     that does not start at row 0 is therefore primed by FOUR halo rows (one preview row), not one.
#pragma once

#include "fpv_internal.h"
#include "fpv_ptx.cuh"

namespace fpv {

constexpr int kStripPx = 256;       // columns per consumer warp (8 pixels per lane)
constexpr int kMaxRowsPerStage = 4; // rows per stage: 4 (one preview row group) or 2
constexpr int kHaloPx = 8;          // pixels copied before a stage's first pixel (16 B)
constexpr int kWarpHistWords = 512; // hist_a, hist_b: 256 u32 counters each, warp-private
constexpr int kWarpScratchBytes = kWarpHistWords * 4;
constexpr int kPrevRing = 8;        // preview rows kept in shared memory (consumers run at most a few rows ahead)

struct FastParams {
  const uint16_t* frames;
  const uint16_t* delta;        // nullptr: no delta frame
  FrameStat* stats;
  const uint32_t* list;         // redo passes: the frames to redo; nullptr in pass 0 (all n frames, assumed = *guess)
  const uint32_t* count;        // redo passes: length of list
  uint32_t n;                   // frames of the batch
  uint32_t* next_list;          // where frame_decide puts frames whose assumption was wrong (nullptr: last pass)
  uint32_t* next_count;
  const uint32_t* guess_in;     // flags (bit 0 delta, bit 1 cg) pass 0 assumes for every frame ...
  uint32_t* guess_out;          // ... and where the batch's last frame leaves its own for the next call (another word:
                                // CTAs of this launch may still be reading guess_in)
  uint8_t* high;
  uint8_t* low;
  uint8_t* preview;
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
  uint32_t prev_pitch;          // bytes per preview row in the shared-memory ring (PW rounded up to 16)
  QConst qc;                    // make_qconst(mode, shift)
};

// Walks the stage sequence of one CTA: tasks blockIdx.x, +gridDim.x, ...; per
// task an optional 1-row halo stage (row y0-1) then 4-row stages.  Producer,
// sampler and consumers each run their own copy and therefore agree on the
// running stage number.
struct StageCursor {
  uint32_t t, f, y0, y1, ys, rps;
  uint32_t pseq = 0;            // preview rows this CTA has completed before the current stage: ring slot = pseq % kPrevRing
  uint32_t band_rows, bands;    // of this launch (short frame lists use shorter bands, see the kernel)
  __device__ __forceinline__ bool load_task(const FastParams& p, uint32_t total) {
    if (t >= total) return false;
    f = p.list ? p.list[t / bands] : t / bands;
    const uint32_t b = t % bands;
    y0 = b * band_rows;
    y1 = min(p.H, y0 + band_rows);
    ys = y0 > 0 ? y0 - 4 : 0;   // four halo rows = one preview row (bands start on multiples of 4)
    return true;
  }
  __device__ __forceinline__ bool halo() const { return ys < y0; }
  __device__ __forceinline__ uint32_t nrows() const {
    return ys < y0 ? min(rps, y0 - ys) : min(rps, y1 - ys);
  }
  __device__ __forceinline__ bool last_of_task() const { return ys >= y0 && ys + nrows() >= y1; }
  __device__ __forceinline__ bool advance(const FastParams& p, uint32_t total) {
    ys += nrows();
    if ((ys & 3u) == 0) pseq++;   // the stage completed a group of four rows = one preview row
    if (ys < y1) return true;
    t += gridDim.x;
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
    uint8_t* __restrict__ out_high, uint8_t* __restrict__ out_low, const uint32_t prev_sm /* shared address of this
    lane's two preview pixels in the preview ring */,
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
  if (counted) st.orl |= (lo[0] | lo[1]) | (lo[2] | lo[3]);
  // (halo rows too: their preview row is the north neighbour of the band's first one)
  st.accA = __dp4a(hr[0], 0x01000100u, __dp4a(hr[1], 0x01000100u, st.accA));
  st.accB = __dp4a(hr[2], 0x01000100u, __dp4a(hr[3], 0x01000100u, st.accB));
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
  }
  if (emit_preview) {
    // the un-predicted preview pixels (.cc:500-512) go to the shared-memory ring; the service warp predicts them
    if (FULL || active) {
      const uint32_t pv = ((st.accA >> 4) & 0xfeu) | (((st.accB >> 4) & 0xfeu) << 8);
      sts16(prev_sm, pv);
    }
    st.accA = 0;
    st.accB = 0;
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

// Runs in the warp that finished the LAST part of frame f in this pass (ticket): the frame's statistics are complete.
// Evaluates the reference's two decisions exactly (see the comment on k_decide in fpv_encode.cu for the delta one:
// USE_DELTA <=> EstimateEntropy(hist of every 15th raw high byte) > 0, provable from 8 bit counters except for
// near-constant frames, where the exact histogram is rebuilt here).  If the pass's assumption held, the frame is
// final: its flags byte is written.  Otherwise it goes onto the next pass's list with the corrected assumption.
// Either way the statistics are cleared, so the next pass / the next call starts from zero without an init kernel.
template <int MODE>
__device__ __noinline__ void frame_decide(const FastParams& p, const uint32_t f, const uint32_t assumed, const int lane) {
  FrameStat& st = p.stats[f];
  __threadfence();
  uint32_t dec_delta = 0;
  if (p.delta != nullptr) {
    const uint64_t N = (p.P + 14) / 15;
    const uint64_t c = lane < 8 ? (uint64_t)__ldcg(&st.dbits[lane]) : 0;
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
  uint32_t va[8], vb[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    va[j] = __ldcg(&st.hist_a[8 * lane + j]);
    vb[j] = __ldcg(&st.hist_b[8 * lane + j]);
    st.hist_a[8 * lane + j] = 0;
    st.hist_b[8 * lane + j] = 0;
  }
  const uint32_t low_or = __ldcg(&st.low_or);
  uint32_t want = dec_delta | (assumed & 2u);
  if (dec_delta == (assumed & 1u)) {
    // the ClampedGradient histograms were taken on the right (post-delta) plane: .cc:564
    const uint64_t e_a = warp_entropy256(va), e_b = warp_entropy256(vb);
    want = dec_delta | ((e_b < e_a) ? 2u : 0u);
  }
  __syncwarp();
  if (lane < 8) st.dbits[lane] = 0;
  if (lane == 0) {
    st.low_or = 0;
    st.tickets = 0;
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

// Takes this warp's ticket for frame f after it has published its statistics; the last of `parts` runs the decision.
template <int MODE>
__device__ __forceinline__ void frame_ticket(const FastParams& p, const uint32_t f, const uint32_t assumed,
                                             const uint32_t parts, const int lane) {
  __threadfence();
  __syncwarp();
  uint32_t t = 0;
  if (lane == 0) t = atomicAdd(&p.stats[f].tickets, 1u);
  t = __shfl_sync(0xffffffffu, t, 0);
  if (t + 1 == parts) frame_decide<MODE>(p, f, assumed, lane);
}

// Shared memory: [ring: stages x {raw stage, delta stage}] [per consumer warp:
// hist_a | hist_b] [preview ring: kPrevRing rows] [full barriers] [empty barriers]
// FULL: xsize is a multiple of 256, every lane of every consumer warp owns pixels.
// PASS0: all n frames of the batch under the guessed flags; otherwise the frames of p.list under their own.
#ifndef FPV_FAST_MAXNREG
#define FPV_FAST_MAXNREG 80
#endif
template <int MODE, bool FULL, int RPS, bool PASS0>
__global__ void __maxnreg__(FPV_FAST_MAXNREG) k_encode_fast(const FastParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t S = p.stages;
  const uint32_t slot_bytes = 2 * p.stage_bytes;
  const int NW = (int)p.compute_warps;
  uint8_t* scratch = smem + (size_t)S * slot_bytes;
  uint8_t* prevring = scratch + (size_t)NW * kWarpScratchBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(prevring + (size_t)kPrevRing * p.prev_pitch);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + S);
  const uint32_t ring0 = smem_u32(smem);
  const uint32_t prev0 = smem_u32(prevring);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t W = p.W;
  const QConst qc = p.qc;   // from the host: each mask is a constant-bank operand of its LOP3, not a derived register

  for (uint32_t i = threadIdx.x; i < (uint32_t)NW * kWarpScratchBytes / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(scratch)[i] = 0;
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
  const uint32_t nframes = PASS0 ? p.n : *p.count;
  const bool short_list = nframes * p.bands < gridDim.x && p.band_rows > 32;
  StageCursor cur;
  cur.band_rows = short_list ? 32u : p.band_rows;
  cur.bands = short_list ? (p.H + 31u) / 32u : p.bands;
  const uint32_t total_tasks = nframes * cur.bands;
  const uint32_t parts = cur.bands * (uint32_t)(NW + 1);   // tickets per frame: every warp of every band
  const uint32_t guess = PASS0 ? (*p.guess_in & (p.delta != nullptr ? 3u : 2u)) : 0u;
  cur.t = blockIdx.x;
  cur.rps = RPS;
  bool more = cur.load_task(p, total_tasks);
  RingPos rp;        // ring position of the stage this warp works on

  if (warp == NW) {
    // ---------- service warp: TMA producer, delta-decision sampler, preview ClampedGradient ----------
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
    for (uint32_t i = 0; i < S && imore; i++) issue();    // the first S stages; stage k + S is issued at the end of iteration k

    // ---- preview: one row per group of four image rows; the stage that was consumed one iteration ago -------
    const uint32_t PW = p.PW, pww = PW / 4;            // (PW % 4 == 0 on this path)
    uint32_t last1 = 0, last2 = 0;                     // last pixel of the two preview rows above the current one
    StageCursor pc = cur;
    RingPos pr;
    bool have_prev = false;
    auto preview_stage = [&]() {
      // every consumer warp has released the stage (its preview pixels are in the ring)
      mbar_wait(empty0 + 8 * pr.slot, pr.phase);
      const uint32_t yend = pc.ys + pc.nrows();
      if (yend & 3u) return;                           // (2-row stages: the row group is not complete yet)
      const uint32_t q = (yend >> 2) - 1;              // preview row just completed
      const uint32_t row_q = prev0 + (pc.pseq & (kPrevRing - 1)) * p.prev_pitch;
      if (pc.halo()) {
        // rows above a band: only their last pixels matter (west / north-west of the band's first preview pixel)
        last1 = lds8(row_q + PW - 1);
        last2 = 0;
        if (q >= 1) {
          // last pixel of preview row q - 1: 4x4 box of the raw frame, straight from global memory
          uint32_t sum = 0;
          if (lane < 4) {
            const uint2 x = *reinterpret_cast<const uint2*>(p.frames + (uint64_t)pc.f * p.P +
                                                            (uint64_t)(4 * (q - 1) + (uint32_t)lane) * W + W - 4);
            sum = ((make_q2<MODE>(x.x, qc) >> 8) & 0xffu) + (make_q2<MODE>(x.x, qc) >> 24) +
                  ((make_q2<MODE>(x.y, qc) >> 8) & 0xffu) + (make_q2<MODE>(x.y, qc) >> 24);
          }
          sum += __shfl_xor_sync(0xffffffffu, sum, 1);
          sum += __shfl_xor_sync(0xffffffffu, sum, 2);
          last2 = __shfl_sync(0xffffffffu, (sum >> 4) & 0xfeu, 0);
        }
        return;
      }
      const uint32_t fl = PASS0 ? guess : p.stats[pc.f].assumed;
      const bool use_cg = (fl & 2u) != 0 && q >= 1;
      const uint32_t row_n = prev0 + ((pc.pseq - 1) & (kPrevRing - 1)) * p.prev_pitch;   // the row completed just before
      uint32_t* out = reinterpret_cast<uint32_t*>(p.preview + (uint64_t)pc.f * p.PP + (uint64_t)q * PW);
      for (uint32_t w = (uint32_t)lane; w < pww; w += 32) {
        const uint32_t c = lds32(row_q + 4 * w);
        uint32_t o = c;
        if (use_cg) {
          const uint32_t n = lds32(row_n + 4 * w);
          // the four pixels to the west in flat order: column 0's are the ends of the rows above (.cc:580-582)
          const uint32_t cm = w > 0 ? lds32(row_q + 4 * w - 4) : last1 << 24;
          const uint32_t nm = w > 0 ? lds32(row_n + 4 * w - 4) : last2 << 24;
          o = finalize_word(c, cm, n, nm);
          if (q == 1 && w == 0) o = (o & 0xffffff00u) | (c & 0xffu);   // flat index PW is copied (.cc:578, :584)
        }
        out[w] = o;
      }
      last2 = last1;
      last1 = lds8(row_q + PW - 1);
    };

    uint32_t acc[4] = {0, 0, 0, 0};  // per-lane counts of set bits {0,2},{1,3},{4,6},{5,7}, two u16 each
    while (more) {
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
      const bool task_end = cur.last_of_task();
      const uint32_t task_f = cur.f;
      if (task_end) {
        uint32_t* gb = p.stats[cur.f].dbits;
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const uint32_t s0 = __reduce_add_sync(0xffffffffu, acc[k] & 0xffffu);
          const uint32_t s1 = __reduce_add_sync(0xffffffffu, acc[k] >> 16);
          // acc[0]: bits 0,2  acc[1]: bits 1,3  acc[2]: bits 4,6  acc[3]: bits 5,7
          const int b0 = (k & 1) + 4 * (k >> 1);
          if (lane == 0) {
            if (s0) atomicAdd(&gb[b0], s0);
            if (s1) atomicAdd(&gb[b0 + 2], s1);
          }
          acc[k] = 0;
        }
      }
      // The previous stage's preview row: every consumer released that stage before the stage sampled above could
      // be issued (ring depth), so this never waits, and it runs while the consumers are busy with the current stage.
      if (have_prev) preview_stage();
      pc = cur;
      pr = rp;
      have_prev = true;
      more = cur.advance(p, total_tasks);
      rp.next(S);
      // Next TMA issue LAST: lane 0 sleeps on the slot's empty barrier and fires the copy the moment the last
      // consumer lets go -- nothing may sit between that release and the issue (the kernel waits on stage data).
      if (imore) issue();
      if (task_end) frame_ticket<MODE>(p, task_f, PASS0 ? guess : p.stats[task_f].assumed, parts, lane);
    }
    if (have_prev) preview_stage();
    return;
  }

  // -------------------------------- consumers -----------------------------------
  const uint32_t sb = (uint32_t)warp * kStripPx;          // first column of this warp's strip
  const uint32_t c0 = sb + (uint32_t)lane * 8;
  const bool active = c0 < W;
  const uint32_t w31 = W % 31;
  const uint32_t rowb = W * 2;                            // bytes per row in a stage
  uint32_t hist_a = smem_u32(scratch + (size_t)warp * kWarpScratchBytes);
  asm volatile("" : "+r"(hist_a));  // keep it in a register instead of recomputing it per row
  const uint32_t prev_lane = prev0 + (c0 >> 2);           // this lane's two pixels within a preview row of the ring

  while (more) {
    // ---- start of a task ---------------------------------------------------------
    const uint32_t f = cur.f;
    const uint32_t assumed = PASS0 ? guess : p.stats[f].assumed;
    const bool use_delta = p.delta != nullptr && (assumed & 1u);
    const bool use_cg = (assumed & 2u) != 0;
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

    for (;;) {
      const uint32_t nrows = cur.nrows();
      const bool own = !cur.halo();
      const bool last = cur.last_of_task();
      const uint32_t slot = rp.slot;
      mbar_wait(full0 + 8 * slot, rp.phase);
      const uint32_t raw_lane = ring0 + slot * slot_bytes + kHaloPx * 2 + c0 * 2;  // pixel (y, c0)

      if (own && y >= 4 && nrows == RPS) {
        // steady state: RPS owned rows, none of them row 0 / row 1; stages start on even rows
        // (bands on multiples of 4), so a preview row group ends with the last row of a stage
        const bool group_end = RPS == 4 || (y & 2u);
        const uint32_t prev_sm = prev_lane + (cur.pseq & (kPrevRing - 1)) * p.prev_pitch;
        if (use_delta && use_cg) {
#pragma unroll
          for (int r = 0; r < RPS; r++) {
            fast_row<MODE, false, FULL, true>(st, raw_lane + r * rowb, raw_lane + r * rowb + p.stage_bytes, qc, true,
                                              true, active, true, y + r, r == RPS - 1 && group_end, c0, w31, oh, ol,
                                              prev_sm, hist_a);
            oh += W;
            if (mode_has_low(MODE)) ol += W;
          }
        } else {
#pragma unroll
          for (int r = 0; r < RPS; r++) {
            fast_row<MODE, false, FULL, false>(st, raw_lane + r * rowb, raw_lane + r * rowb + p.stage_bytes, qc,
                                               use_delta, use_cg, active, true, y + r, r == RPS - 1 && group_end, c0,
                                               w31, oh, ol, prev_sm, hist_a);
            oh += W;
            if (mode_has_low(MODE)) ol += W;
          }
        }
        y += RPS;
      } else {
        for (uint32_t r = 0; r < nrows; r++) {
          fast_row<MODE, true, FULL, false>(st, raw_lane + r * rowb, raw_lane + r * rowb + p.stage_bytes, qc,
                                            use_delta, use_cg, active, own, y, (y & 3u) == 3u, c0, w31, oh, ol,
                                            prev_lane + (cur.pseq & (kPrevRing - 1)) * p.prev_pitch, hist_a);
          oh += W;
          if (mode_has_low(MODE)) ol += W;
          y++;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty0 + 8 * slot);
      rp.next(S);
      more = cur.advance(p, total_tasks);
      if (last) break;
    }

    // ---- end of task: publish low-OR and this warp's histograms, take the frame's ticket ----------
    if (mode_has_low(MODE)) {
      // lanes past the end of the row accumulated garbage (see fast_row)
      const uint32_t orl = __reduce_or_sync(0xffffffffu, (FULL || active) ? (st.orl & kLoBytes) : 0u);
      if (lane == 0 && orl) atomicOr(&p.stats[f].low_or, orl);
    }
    uint32_t* gh = p.stats[f].hist_a;  // hist_a, hist_b are contiguous: 512 bins
    for (uint32_t i = lane; i < kWarpHistWords; i += 32) {
      const uint32_t v = lds32(hist_a + 4 * i);
      if (v) {
        sts32(hist_a + 4 * i, 0);
        atomicAdd(&gh[i], v);
      }
    }
    frame_ticket<MODE>(p, f, assumed, parts, lane);
  }
}

static inline size_t fast_smem_bytes(uint32_t W, int stages, int rows_per_stage) {
  const size_t stage_bytes = ((size_t)rows_per_stage * W + kHaloPx) * 2;
  const size_t warps = (W + kStripPx - 1) / kStripPx;
  const size_t prev_pitch = ((W / 4) + 15) / 16 * 16;
  return (size_t)stages * 2 * stage_bytes + warps * kWarpScratchBytes + kPrevRing * prev_pitch + 2 * (size_t)stages * 8;
}

}  // namespace fpv

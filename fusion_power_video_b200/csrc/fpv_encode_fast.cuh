// fpv_encode_fast.cuh -- the fused, TMA-staged encode kernel (k_encode_fast).
//
// One launch = Frame ctor + Frame::Predict (fusion_power_video.cc:370-451,
// :491-593) for a list of frames under ASSUMED per-frame flags, plus the
// three decision histograms and the low-byte OR the reference's heuristics
// need (.cc:447-449, :522-531, :550-562).
//
// Work decomposition
//   task   = (frame, band of `band_rows` rows); persistent CTAs take tasks
//            round-robin.
//   stage  = up to 4 consecutive rows of the band = ONE contiguous flat range
//            of the raw frame (and of the delta image), fetched by the producer
//            warp with cp.async.bulk (TMA, 1-D) into a shared-memory ring and
//            signalled through an mbarrier.  The range starts 8 pixels early:
//            with the reference's flat indexing (.cc:556-558) the west
//            neighbour of column 0 is the last pixel of the previous row, which
//            is exactly what precedes the row in memory.
//   strip  = 256 columns of a row, owned by one consumer warp (8 px per lane).
//            The warp walks down the rows of the band keeping the previous
//            row's (post-delta) high bytes in registers, so north / north-west
//            neighbours never touch memory again; a band that does not start at
//            row 0 is primed by a 1-row halo stage.
//
// Per row and lane: 1x LDS.128 raw, 1x LDS.128 delta, 2x LDS.32 for the west
// pixel, lane-form arithmetic (fpv_common.cuh), 2x STG.64 (+1 STG.16 of
// preview every 4th row).  Sampling for the histograms is done cooperatively
// per warp-row out of shared memory into warp-private packed-u16 histograms
// (no CTA-wide barrier anywhere in the steady state).
#pragma once

#include "fpv_internal.h"

namespace fpv {

// ---- PTX wrappers -------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Consumer-side wait: data is normally there already or lands within the
// hardware suspend window of try_wait.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// Producer-side wait: the producer runs ahead of the consumers by the depth of
// the ring, so it mostly waits; back off instead of burning issue slots.
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(256);
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on `bar`.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes,
                                         uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t a) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds8(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t a, uint2 v) {
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void red_shared_add(uint32_t a, uint32_t v) {
  asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
// Streaming stores: outputs are written once and not re-read by this kernel.
__device__ __forceinline__ void stg64_cs(void* p, uint2 v) {
  asm volatile("st.global.cs.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}

constexpr int kStripPx = 256;       // columns per consumer warp (8 pixels per lane)
constexpr int kRowsPerStage = 4;    // one preview row group
constexpr int kHaloPx = 8;          // pixels copied before a stage's first pixel (16 B)
constexpr int kWarpHistWords = 384; // 3 histograms x 256 bins, two u16 counters per word
constexpr int kWarpScratchBytes = kWarpHistWords * 4 + 256;  // + residual staging row

struct FastParams {
  const uint16_t* frames;
  const uint16_t* delta;        // nullptr: no delta frame
  FrameStat* stats;
  const uint32_t* list;
  const uint32_t* count;
  uint8_t* high;
  uint8_t* low;
  uint8_t* preview_raw;
  uint32_t W, H;
  uint64_t P, PP;
  uint32_t PW;
  int shift;
  uint32_t band_rows;           // multiple of 4, <= 1024
  uint32_t bands;               // bands per frame
  uint32_t stages;              // ring depth
  uint32_t stage_bytes;         // bytes of one plane of one stage: (4W + 8) * 2
  uint32_t compute_warps;       // ceil(W / 256)
};

// Per-lane running state while walking down a strip.
struct StripState {
  uint32_t ph[4];   // previous row, post-delta high bytes, lane form
  uint32_t pw[4];   // previous row shifted one pixel west (= this row's north-west)
  uint32_t acc0, acc1;  // preview 4x4 box sums (two preview pixels per lane), lane form
  uint32_t orl;     // OR of split low bytes
};

// One row of one strip.  FIRSTROWS = true compiles the row-0 / row-1 special
// cases of the first stage of a frame and the halo (not-owned) row; the steady
// state uses FIRSTROWS = false.
template <int MODE, bool FIRSTROWS>
__device__ __forceinline__ void fast_row(
    StripState& st, const uint32_t raw_a /* shared addr of this lane's 8 px */,
    const uint32_t del_a, const uint32_t raw_strip_a /* shared addr of the strip's first px */,
    const uint32_t del_strip_a, const int s, const bool use_delta, const bool use_cg,
    const bool active, const bool own, const uint32_t y, const bool emit_preview,
    const uint32_t c0, const int lane, const uint32_t span /* valid px in this strip */,
    const uint32_t m15 /* flat index of strip start mod 15 */, const uint32_t m31,
    uint8_t* __restrict__ out_high, uint8_t* __restrict__ out_low, uint8_t* __restrict__ out_prev,
    const uint32_t hist_a, const uint32_t stage_res_a) {
  uint32_t xh[4], xl[4], h[4], l[4];
  uint32_t hleft;  // lane form of the two pixels west of this lane's first pixel
  {
    uint4 x = make_uint4(0, 0, 0, 0);
    uint32_t xw = 0;
    if (active) {
      x = lds128(raw_a);
      xw = lds32(raw_a - 4);
    }
    split2<MODE>(x.x, s, xh[0], xl[0]);
    split2<MODE>(x.y, s, xh[1], xl[1]);
    split2<MODE>(x.z, s, xh[2], xl[2]);
    split2<MODE>(x.w, s, xh[3], xl[3]);
    uint32_t dummy;
    split2<MODE>(xw, s, hleft, dummy);
  }
  if (use_delta) {
    uint4 d = make_uint4(0, 0, 0, 0);
    uint32_t dw = 0;
    if (active) {
      d = lds128(del_a);
      dw = lds32(del_a - 4);
    }
    uint32_t dh, dl;
    split2_delta(d.x, dh, dl); h[0] = sub2(xh[0], dh); l[0] = xl[0] + kLaneBias - dl;
    split2_delta(d.y, dh, dl); h[1] = sub2(xh[1], dh); l[1] = xl[1] + kLaneBias - dl;
    split2_delta(d.z, dh, dl); h[2] = sub2(xh[2], dh); l[2] = xl[2] + kLaneBias - dl;
    split2_delta(d.w, dh, dl); h[3] = sub2(xh[3], dh); l[3] = xl[3] + kLaneBias - dl;
    split2_delta(dw, dh, dl);  hleft = sub2(hleft, dh);
  } else {
#pragma unroll
    for (int j = 0; j < 4; j++) { h[j] = xh[j]; l[j] = xl[j]; }
  }
  uint32_t w[4];
  w[0] = __funnelshift_l(hleft, h[0], 16);
  w[1] = __funnelshift_l(h[0], h[1], 16);
  w[2] = __funnelshift_l(h[1], h[2], 16);
  w[3] = __funnelshift_l(h[2], h[3], 16);

  if (!FIRSTROWS || own) {
    uint32_t res[4];
#pragma unroll
    for (int j = 0; j < 4; j++) res[j] = sub2(h[j], cg2(st.ph[j], w[j], st.pw[j]));
    if (FIRSTROWS) {
      if (y == 0) {
#pragma unroll
        for (int j = 0; j < 4; j++) res[j] = h[j];
      } else if (y == 1 && c0 == 0) {
        // flat index W (row 1, column 0) is copied, not predicted (.cc:566, :572)
        res[0] = (res[0] & 0xffff0000u) | (h[0] & 0x0000ffffu);
      }
    }
    const uint2 res8 = pack8(res[0], res[1], res[2], res[3]);
    if (active) {
      stg64_cs(out_high, use_cg ? res8 : pack8(h[0], h[1], h[2], h[3]));
      if (mode_has_low(MODE)) stg64_cs(out_low, pack8(l[0], l[1], l[2], l[3]));
      st.orl |= xl[0] | xl[1] | xl[2] | xl[3];
      st.acc0 += xh[0] + xh[1];
      st.acc1 += xh[2] + xh[3];
      sts64(stage_res_a + 8 * lane, res8);
    }
    // ---- decision histograms, sampled cooperatively per warp-row -------------
    // delta decision: RAW high byte at flat index % 15 == 0 (.cc:526-531)
    {
      const uint32_t pd = (m15 ? 15 - m15 : 0) + 15 * lane;
      if (pd < span) {
        uint32_t hv, lv;
        split1<MODE>(lds16(raw_strip_a + 2 * pd), s, hv, lv);
        red_shared_add(hist_a + ((hv >> 1) << 2), 1u << ((hv & 1) << 4));
      }
    }
    __syncwarp();
    // CG decision: flat index == W+1 (mod 31), >= W+1, post-delta plane (.cc:554-562)
    if (!FIRSTROWS || y >= 1) {
      const uint32_t pc = (m31 ? 31 - m31 : 0) + 31 * lane;
      if (pc < span) {
        uint32_t a, lv;
        split1<MODE>(lds16(raw_strip_a + 2 * pc), s, a, lv);
        if (use_delta) a = (a - (lds16(del_strip_a + 2 * pc) >> 8)) & 0xffu;
        const uint32_t b = lds8(stage_res_a + pc);
        red_shared_add(hist_a + 512 + ((a >> 1) << 2), 1u << ((a & 1) << 4));
        red_shared_add(hist_a + 1024 + ((b >> 1) << 2), 1u << ((b & 1) << 4));
      }
    }
    __syncwarp();
    if (emit_preview && active) {
      const uint32_t s0 = (st.acc0 & 0xffffu) + (st.acc0 >> 16);
      const uint32_t s1 = (st.acc1 & 0xffffu) + (st.acc1 >> 16);
      const uint32_t pv = ((s0 >> 4) & 0xfeu) | (((s1 >> 4) & 0xfeu) << 8);
      *reinterpret_cast<uint16_t*>(out_prev) = (uint16_t)pv;
      st.acc0 = 0;
      st.acc1 = 0;
    }
  }
#pragma unroll
  for (int j = 0; j < 4; j++) { st.ph[j] = h[j]; st.pw[j] = w[j]; }
}

// Shared memory: [ring: stages x {raw stage, delta stage}] [per consumer warp:
// packed histograms + residual staging row] [full barriers] [empty barriers]
template <int MODE>
__global__ void __launch_bounds__(544) k_encode_fast(const FastParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t S = p.stages;
  const uint32_t slot_bytes = 2 * p.stage_bytes;
  const int NW = (int)p.compute_warps;
  uint8_t* scratch = smem + (size_t)S * slot_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(scratch + (size_t)NW * kWarpScratchBytes);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + S);
  const uint32_t ring0 = smem_u32(smem);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t W = p.W;

  for (uint32_t i = threadIdx.x; i < (uint32_t)NW * kWarpScratchBytes / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(scratch)[i] = 0;
  if (threadIdx.x == 0) {
    for (uint32_t i = 0; i < S; i++) {
      mbar_init(full0 + 8 * i, 1);
      mbar_init(empty0 + 8 * i, NW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const uint32_t total_tasks = (*p.count) * p.bands;
  uint32_t seq = 0;  // running stage number (same sequence in producer and consumers)

  if (warp == NW) {
    // ------------------------------ producer ------------------------------------
    if (lane == 0) {
      for (uint32_t t = blockIdx.x; t < total_tasks; t += gridDim.x) {
        const uint32_t f = p.list[t / p.bands], b = t % p.bands;
        const uint32_t y0 = b * p.band_rows;
        const uint32_t y1 = min(p.H, y0 + p.band_rows);
        const bool use_delta = p.delta != nullptr && (p.stats[f].assumed & 1u);
        const uint16_t* img = p.frames + (uint64_t)f * p.P;
        // stage list: optional 1-row halo stage (row y0-1), then 4-row stages
        uint32_t ys = y0 > 0 ? y0 - 1 : 0;
        while (ys < y1) {
          const uint32_t nrows = (ys < y0) ? 1 : min((uint32_t)kRowsPerStage, y1 - ys);
          const uint32_t slot = seq % S, ph = (seq / S) & 1u;
          mbar_wait_backoff(empty0 + 8 * slot, ph ^ 1u);
          // contiguous flat range [ys*W - 8, (ys+nrows)*W)
          const uint64_t px0 = (uint64_t)ys * W;
          const uint32_t lead = ys > 0 ? kHaloPx : 0;
          const uint32_t bytes = (nrows * W + lead) * 2;
          const uint32_t dst = ring0 + slot * slot_bytes + (kHaloPx - lead) * 2;
          mbar_arrive_expect_tx(full0 + 8 * slot, use_delta ? 2 * bytes : bytes);
          bulk_g2s(dst, img + px0 - lead, bytes, full0 + 8 * slot);
          if (use_delta) bulk_g2s(dst + p.stage_bytes, p.delta + px0 - lead, bytes, full0 + 8 * slot);
          seq++;
          ys += nrows;
        }
      }
    }
    return;
  }

  // -------------------------------- consumers -----------------------------------
  const uint32_t sb = (uint32_t)warp * kStripPx;          // first column of this warp's strip
  const uint32_t c0 = sb + (uint32_t)lane * 8;
  const bool active = c0 < W;
  const uint32_t span = sb < W ? min((uint32_t)kStripPx, W - sb) : 0;
  const uint32_t w15 = W % 15, w31 = W % 31;
  const uint32_t rowb = W * 2;                            // bytes per row in a stage
  const int s = p.shift;
  const uint32_t hist_a = smem_u32(scratch + (size_t)warp * kWarpScratchBytes);
  const uint32_t res_a = hist_a + kWarpHistWords * 4;

  for (uint32_t t = blockIdx.x; t < total_tasks; t += gridDim.x) {
    const uint32_t f = p.list[t / p.bands], b = t % p.bands;
    const uint32_t y0 = b * p.band_rows;
    const uint32_t y1 = min(p.H, y0 + p.band_rows);
    const uint32_t assumed = p.stats[f].assumed;
    const bool use_delta = p.delta != nullptr && (assumed & 1u);
    const bool use_cg = (assumed & 2u) != 0;

    StripState st;
#pragma unroll
    for (int j = 0; j < 4; j++) { st.ph[j] = 0; st.pw[j] = 0; }
    st.acc0 = st.acc1 = st.orl = 0;
    uint32_t y = y0 > 0 ? y0 - 1 : 0;
    // residues of the flat index of the STRIP's first pixel in row y
    const uint64_t i0 = (uint64_t)y * W + sb;
    uint32_t m15 = (uint32_t)(i0 % 15);
    uint32_t m31 = (uint32_t)((i0 + 31ull * (W / 31 + 2) - (W + 1)) % 31);
    uint8_t* oh = p.high + (uint64_t)f * p.P + (uint64_t)y * W + c0;
    uint8_t* ol = mode_has_low(MODE) ? p.low + (uint64_t)f * p.P + (uint64_t)y * W + c0 : nullptr;
    uint8_t* op = p.preview_raw + (uint64_t)f * p.PP + (uint64_t)(y >> 2) * p.PW + (c0 >> 2);

    while (y < y1) {
      const uint32_t nrows = (y < y0) ? 1 : min((uint32_t)kRowsPerStage, y1 - y);
      const uint32_t slot = seq % S, phs = (seq / S) & 1u;
      mbar_wait(full0 + 8 * slot, phs);
      const uint32_t raw_strip = ring0 + slot * slot_bytes + kHaloPx * 2 + sb * 2;  // pixel (y, sb)
      const uint32_t raw_lane = raw_strip + (uint32_t)lane * 16;

      if (y >= 4 && nrows == kRowsPerStage) {
        // steady state: 4 owned rows, none of them row 0 / row 1
#pragma unroll
        for (int r = 0; r < kRowsPerStage; r++) {
          fast_row<MODE, false>(st, raw_lane + r * rowb, raw_lane + r * rowb + p.stage_bytes,
                                raw_strip + r * rowb, raw_strip + r * rowb + p.stage_bytes, s, use_delta,
                                use_cg, active, true, y + r, r == kRowsPerStage - 1, c0, lane, span, m15,
                                m31, oh, ol, op, hist_a, res_a);
          oh += W;
          if (mode_has_low(MODE)) ol += W;
          m15 += w15; if (m15 >= 15) m15 -= 15;
          m31 += w31; if (m31 >= 31) m31 -= 31;
        }
        op += p.PW;
        y += kRowsPerStage;
      } else {
        for (uint32_t r = 0; r < nrows; r++) {
          fast_row<MODE, true>(st, raw_lane + r * rowb, raw_lane + r * rowb + p.stage_bytes,
                               raw_strip + r * rowb, raw_strip + r * rowb + p.stage_bytes, s, use_delta,
                               use_cg, active, y >= y0, y, (y & 3u) == 3u, c0, lane, span, m15, m31, oh, ol,
                               op, hist_a, res_a);
          oh += W;
          if (mode_has_low(MODE)) ol += W;
          if ((y & 3u) == 3u) op += p.PW;
          m15 += w15; if (m15 >= 15) m15 -= 15;
          m31 += w31; if (m31 >= 31) m31 -= 31;
          y++;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty0 + 8 * slot);
      seq++;
    }

    // ---- end of task: publish low-OR and this warp's histograms -----------------
    uint32_t orl = st.orl;
#pragma unroll
    for (int o = 16; o; o >>= 1) orl |= __shfl_xor_sync(0xffffffffu, orl, o);
    if (lane == 0 && (orl & kLaneMask)) atomicOr(&p.stats[f].low_or, orl & kLaneMask);
    uint32_t* gh = p.stats[f].hist_d;  // hist_d, hist_a, hist_b are contiguous: 768 bins
    for (uint32_t i = lane; i < kWarpHistWords; i += 32) {
      const uint32_t v = lds32(hist_a + 4 * i);
      if (v) {
        sts32(hist_a + 4 * i, 0);
        if (v & 0xffffu) atomicAdd(&gh[2 * i], v & 0xffffu);
        if (v >> 16) atomicAdd(&gh[2 * i + 1], v >> 16);
      }
    }
    __syncwarp();
  }
}

static inline size_t fast_smem_bytes(uint32_t W, int stages) {
  const size_t stage_bytes = ((size_t)kRowsPerStage * W + kHaloPx) * 2;
  const size_t warps = (W + kStripPx - 1) / kStripPx;
  return (size_t)stages * 2 * stage_bytes + warps * kWarpScratchBytes + 2 * (size_t)stages * 8;
}

}  // namespace fpv

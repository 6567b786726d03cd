// fpv_encode.cu -- encode-side kernels: Frame ctor + Frame::Predict on the GPU.
//
// Replaces fusion_power_video.cc:370-451 (split), :491-515 (preview),
// :517-544 (delta decision + apply), :546-593 (ClampedGradient decision +
// forward, high plane and preview).  See DESIGN.md for the layout.
//
// Two paths produce identical bytes:
//
//  * FAST (k_encode_fast): persistent CTAs stream contiguous 4-row stages of a
//    frame band into a shared-memory ring with 1-D TMA bulk copies
//    (cp.async.bulk + mbarrier), one warp per 256-column strip walks down the
//    rows keeping the previous row in registers, everything fused into one
//    read of the raw frame and one write of each output plane.  The
//    reference's per-frame decisions (USE_DELTA, USE_CG) depend on whole-frame
//    histograms, so the pass runs with ASSUMED flags while accumulating the
//    histograms; k_decide then evaluates the integer heuristics exactly and
//    frames whose assumption was wrong are redone (at most twice) by the same
//    kernel.  In the common case compulsory HBM traffic is 4.0625 B/pixel.
//
//  * GENERIC (k_gen_*): statistics first, then transform; plain loads, any
//    xsize % 4 == 0.  Used for geometries the bulk-copy path cannot take
//    (xsize % 8 != 0, very wide rows) and as an in-GPU cross-check.
#include <stdio.h>

#include "fpv_internal.h"

namespace fpv {

// =====================================================================================
// Small PTX wrappers (mbarrier + bulk async copy)
// =====================================================================================
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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
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
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// =====================================================================================
// Per-batch bookkeeping kernels
// =====================================================================================

// One block (256 threads) per frame: zero the statistics, set the assumed
// flags, build the identity work list.
__global__ void k_encode_init(FrameStat* stats, uint32_t* lists, uint32_t* counts, uint32_t n,
                              uint32_t cap, int has_delta, int use_guess) {
  uint32_t f = blockIdx.x;
  if (f >= n) return;
  FrameStat& st = stats[f];
  st.hist_d[threadIdx.x] = 0;
  st.hist_a[threadIdx.x] = 0;
  st.hist_b[threadIdx.x] = 0;
  if (threadIdx.x == 0) {
    uint32_t guess = use_guess ? (counts[3] & 3u) : 0u;
    if (!has_delta) guess &= 2u;
    st.low_or = 0;
    st.assumed = guess;
    st.final_flags = 0;
    st.done = 0;
    lists[f] = f;
    if (f == 0) {
      counts[0] = n;
      counts[1] = 0;
      counts[2] = 0;
    }
  }
  (void)cap;
}

// Evaluates the reference's decisions for the frames of list `in`.
//  phase 0 (fast path): a transform pass just ran with st.assumed.  If the
//          delta decision differs, redo with the right delta (cg assumption
//          kept); else decide CG (its histograms were taken on the right
//          plane), fix final_flags, redo iff the cg assumption was wrong.
//  phase 1 (generic): delta decision only  -> st.assumed bit 0.
//  phase 2 (generic): CG decision only     -> st.final_flags, st.assumed.
__global__ void __launch_bounds__(256)
k_decide(FrameStat* stats, const uint32_t* in, const uint32_t* in_count, uint32_t* out,
         uint32_t* out_count, int phase, int has_delta, int has_low) {
  __shared__ uint32_t red[8];
  if (blockIdx.x >= *in_count) return;
  uint32_t f = in[blockIdx.x];
  FrameStat& st = stats[f];
  const int t = threadIdx.x;
  uint32_t assumed = st.assumed;
  uint32_t nolow = has_low ? (st.low_or == 0 ? kFlagNoLow : 0) : kFlagNoLow;

  uint32_t dec_delta = assumed & 1u;
  if (phase == 0 || phase == 1) {
    // countd is {0: N} in the reference (d = a - high_[i] with a == high_[i],
    // .cc:527-529), whose EstimateEntropy is 0: USE_DELTA <=> 0 < E(counta).
    uint64_t ea = block_entropy256(st.hist_d[t], red);
    dec_delta = (has_delta && ea > 0) ? 1u : 0u;
  }
  if (phase == 1) {
    if (t == 0) st.assumed = dec_delta;
    return;
  }
  if (phase == 0 && dec_delta != (assumed & 1u)) {
    __syncthreads();
    st.hist_d[t] = 0; st.hist_a[t] = 0; st.hist_b[t] = 0;
    if (t == 0) {
      st.assumed = dec_delta | (assumed & 2u);
      out[atomicAdd(out_count, 1u)] = f;
    }
    return;
  }
  uint64_t e_a = block_entropy256(st.hist_a[t], red);
  uint64_t e_b = block_entropy256(st.hist_b[t], red);
  uint32_t dec_cg = (e_b < e_a) ? 1u : 0u;                       // .cc:564
  uint32_t fin = dec_delta | (dec_cg << 1) | nolow;
  if (phase == 2) {
    if (t == 0) { st.final_flags = fin; st.assumed = fin & 3u; st.done = 1; }
    return;
  }
  __syncthreads();
  if (dec_cg != ((assumed >> 1) & 1u)) {
    st.hist_d[t] = 0; st.hist_a[t] = 0; st.hist_b[t] = 0;
    if (t == 0) {
      st.final_flags = fin;
      st.assumed = fin & 3u;
      out[atomicAdd(out_count, 1u)] = f;
    }
  } else if (t == 0) {
    st.final_flags = fin;
    st.done = 1;
  }
}

// Writes the flags byte and the final preview (ClampedGradient applied on the
// preview's own flat array of width W/4 iff USE_CG, .cc:575-586).
__global__ void k_finalize(const FrameStat* stats, const uint8_t* preview_raw, uint8_t* preview,
                           uint8_t* flags, uint32_t* counts, uint32_t n, uint32_t PW,
                           uint64_t PP, int has_low) {
  uint32_t f = blockIdx.y;
  const FrameStat& st = stats[f];
  // NO_LOW_BYTES is only known once every pass has OR-ed its low bytes in.
  uint32_t fin = (st.final_flags & 3u) |
                 (has_low ? (st.low_or == 0 ? kFlagNoLow : 0) : kFlagNoLow);
  const uint8_t* pr = preview_raw + (uint64_t)f * PP;
  uint8_t* po = preview + (uint64_t)f * PP;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < PP;
       i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t v = pr[i];
    if ((fin & kFlagCG) && i > PW) v = (v - cg1(pr[i - PW], pr[i - 1], pr[i - PW - 1])) & 0xffu;
    po[i] = (uint8_t)v;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    flags[f] = (uint8_t)fin;
    if (f == n - 1) counts[3] = fin & 3u;  // guess for the next batch
  }
}

// =====================================================================================
// GENERIC path
// =====================================================================================

template <int MODE>
__global__ void __launch_bounds__(256)
k_gen_stats_delta(const uint16_t* __restrict__ frames, FrameStat* stats, uint64_t P, int s,
                  uint32_t chunk) {
  __shared__ uint32_t sh[256];
  uint32_t f = blockIdx.y;
  sh[threadIdx.x] = 0;
  __syncthreads();
  const uint16_t* img = frames + (uint64_t)f * P;
  uint64_t beg = (uint64_t)blockIdx.x * chunk;
  uint64_t end = beg + chunk < P ? beg + chunk : P;
  uint32_t lor = 0;
  for (uint64_t i = beg + threadIdx.x; i < end; i += blockDim.x) {
    uint32_t h, l;
    split1<MODE>(img[i], s, h, l);
    lor |= l;
    if (i % 15 == 0) atomicAdd(&sh[h], 1u);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) lor |= __shfl_xor_sync(0xffffffffu, lor, o);
  if ((threadIdx.x & 31) == 0 && lor) atomicOr(&stats[f].low_or, lor);
  __syncthreads();
  uint32_t v = sh[threadIdx.x];
  if (v) atomicAdd(&stats[f].hist_d[threadIdx.x], v);
}

template <int MODE>
__device__ __forceinline__ uint32_t gen_high(const uint16_t* img, const uint16_t* delta,
                                             uint64_t i, int s, bool use_delta) {
  uint32_t h, l;
  split1<MODE>(img[i], s, h, l);
  if (use_delta) h = (h - (uint32_t)(delta[i] >> 8)) & 0xffu;
  return h;
}

template <int MODE>
__global__ void __launch_bounds__(256)
k_gen_stats_cg(const uint16_t* __restrict__ frames, const uint16_t* __restrict__ delta,
               FrameStat* stats, uint32_t W, uint64_t P, int s) {
  __shared__ uint32_t sa[256], sb[256];
  uint32_t f = blockIdx.y;
  sa[threadIdx.x] = 0;
  sb[threadIdx.x] = 0;
  __syncthreads();
  const uint16_t* img = frames + (uint64_t)f * P;
  bool use_delta = (stats[f].assumed & 1u) != 0;
  for (uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;; m += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t i = (uint64_t)W + 1 + 31 * m;
    if (i >= P) break;
    uint32_t a = gen_high<MODE>(img, delta, i, s, use_delta);
    uint32_t n = gen_high<MODE>(img, delta, i - W, s, use_delta);
    uint32_t w = gen_high<MODE>(img, delta, i - 1, s, use_delta);
    uint32_t nw = gen_high<MODE>(img, delta, i - W - 1, s, use_delta);
    uint32_t b = (a - cg1(n, w, nw)) & 0xffu;
    atomicAdd(&sa[a], 1u);
    atomicAdd(&sb[b], 1u);
  }
  __syncthreads();
  uint32_t va = sa[threadIdx.x], vb = sb[threadIdx.x];
  if (va) atomicAdd(&stats[f].hist_a[threadIdx.x], va);
  if (vb) atomicAdd(&stats[f].hist_b[threadIdx.x], vb);
}

// One thread per group of 4 consecutive pixels (W % 4 == 0: groups never
// straddle rows).  Flags are final here.
template <int MODE>
__global__ void __launch_bounds__(256)
k_gen_transform(const uint16_t* __restrict__ frames, const uint16_t* __restrict__ delta,
                const FrameStat* stats, uint8_t* __restrict__ high, uint8_t* __restrict__ low,
                uint32_t W, uint64_t P, int s) {
  uint32_t f = blockIdx.y;
  const uint16_t* img = frames + (uint64_t)f * P;
  uint32_t fl = stats[f].final_flags;
  bool use_delta = fl & kFlagDelta, use_cg = fl & kFlagCG;
  uint64_t groups = P / 4;
  for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups;
       g += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t i0 = g * 4;
    uint2 raw = *reinterpret_cast<const uint2*>(img + i0);
    uint32_t px[4] = {raw.x & 0xffffu, raw.x >> 16, raw.y & 0xffffu, raw.y >> 16};
    uint32_t h[4], l[4];
#pragma unroll
    for (int k = 0; k < 4; k++) split1<MODE>(px[k], s, h[k], l[k]);
    if (use_delta) {
      uint2 dr = *reinterpret_cast<const uint2*>(delta + i0);
      uint32_t d[4] = {dr.x & 0xffffu, dr.x >> 16, dr.y & 0xffffu, dr.y >> 16};
#pragma unroll
      for (int k = 0; k < 4; k++) {
        h[k] = (h[k] - (d[k] >> 8)) & 0xffu;
        l[k] = (l[k] - (d[k] & 0xffu)) & 0xffu;
      }
    }
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      uint64_t i = i0 + k;
      o[k] = h[k];
      if (use_cg && i > W) {
        uint32_t n = gen_high<MODE>(img, delta, i - W, s, use_delta);
        uint32_t nw = gen_high<MODE>(img, delta, i - W - 1, s, use_delta);
        uint32_t w = (k == 0) ? gen_high<MODE>(img, delta, i - 1, s, use_delta) : h[k - 1];
        o[k] = (h[k] - cg1(n, w, nw)) & 0xffu;
      }
    }
    *reinterpret_cast<uint32_t*>(high + (uint64_t)f * P + i0) =
        o[0] | (o[1] << 8) | (o[2] << 16) | (o[3] << 24);
    if (mode_has_low(MODE))
      *reinterpret_cast<uint32_t*>(low + (uint64_t)f * P + i0) =
          l[0] | (l[1] << 8) | (l[2] << 16) | (l[3] << 24);
  }
}

// One thread per preview pixel: 4x4 box of RAW high bytes (.cc:500-512).
template <int MODE>
__global__ void __launch_bounds__(256)
k_gen_preview(const uint16_t* __restrict__ frames, uint8_t* __restrict__ preview_raw, uint32_t W,
              uint64_t P, uint32_t PW, uint64_t PP, int s) {
  uint32_t f = blockIdx.y;
  const uint16_t* img = frames + (uint64_t)f * P;
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < PP;
       q += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t py = q / PW, px = q % PW;
    uint32_t sum = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      uint2 raw = *reinterpret_cast<const uint2*>(img + (py * 4 + j) * W + px * 4);
      uint32_t v[4] = {raw.x & 0xffffu, raw.x >> 16, raw.y & 0xffffu, raw.y >> 16};
#pragma unroll
      for (int k = 0; k < 4; k++) {
        uint32_t h, l;
        split1<MODE>(v[k], s, h, l);
        sum += h;
      }
    }
    preview_raw[(uint64_t)f * PP + q] = (uint8_t)((sum / 16) & 0xfe);
  }
}

// Delta frame: raw -> image form.
template <int MODE>
__global__ void __launch_bounds__(256)
k_delta_from_raw(const uint16_t* __restrict__ raw, uint16_t* __restrict__ image, uint64_t P, int s) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P;
       i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t h, l;
    split1<MODE>(raw[i], s, h, l);
    image[i] = (uint16_t)((h << 8) | l);
  }
}

// =====================================================================================
// FAST path
// =====================================================================================

constexpr int kStripPx = 256;       // columns per warp (8 pixels per lane)
constexpr int kRowsPerStage = 4;    // one preview row group
constexpr int kHaloPx = 8;          // pixels copied before a stage's first pixel (16 B)

// v[j] for a run-time j without spilling the array to local memory.
__device__ __forceinline__ uint32_t sel4(const uint32_t (&v)[4], uint32_t j) {
  uint32_t a = (j & 1u) ? v[1] : v[0];
  uint32_t b = (j & 1u) ? v[3] : v[2];
  return (j & 2u) ? b : a;
}

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
  uint32_t band_rows;           // multiple of 4
  uint32_t bands;               // bands per frame
  uint32_t stages;              // ring depth
  uint32_t stage_bytes;         // bytes of one plane of one stage: (4W + 8) * 2
  uint32_t compute_warps;       // ceil(W / 256)
};

// Shared memory: [ring: stages x {raw stage, delta stage}] [hist 3x256 u32]
// [full barriers] [empty barriers]
template <int MODE>
__global__ void __launch_bounds__(544) k_encode_fast(const FastParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t S = p.stages;
  const uint32_t slot_bytes = 2 * p.stage_bytes;
  uint32_t* hist = reinterpret_cast<uint32_t*>(smem + (size_t)S * slot_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(hist + 768);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + S);
  const uint32_t ring0 = smem_u32(smem);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NW = (int)p.compute_warps;
  const uint32_t W = p.W;

  for (uint32_t i = threadIdx.x; i < 768; i += blockDim.x) hist[i] = 0;
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
        uint32_t f = p.list[t / p.bands], b = t % p.bands;
        uint32_t y0 = b * p.band_rows;
        uint32_t y1 = min(p.H, y0 + p.band_rows);
        bool use_delta = p.delta != nullptr && (p.stats[f].assumed & 1u);
        const uint16_t* img = p.frames + (uint64_t)f * p.P;
        // stage list: optional 1-row halo stage (row y0-1), then 4-row stages
        uint32_t ys = y0 > 0 ? y0 - 1 : 0;
        while (ys < y1) {
          uint32_t nrows = (ys < y0) ? 1 : min((uint32_t)kRowsPerStage, y1 - ys);
          uint32_t slot = seq % S, ph = (seq / S) & 1u;
          mbar_wait(empty0 + 8 * slot, ph ^ 1u);
          // contiguous flat range [ys*W - 8, (ys+nrows)*W); the 8-pixel lead-in
          // holds the west neighbours of column 0 (flat indexing, .cc:556-558).
          uint64_t px0 = (uint64_t)ys * W;
          uint32_t lead = ys > 0 ? kHaloPx : 0;
          uint32_t bytes = (nrows * W + lead) * 2;
          uint32_t dst = ring0 + slot * slot_bytes + (kHaloPx - lead) * 2;
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
  const uint32_t c0 = (uint32_t)warp * kStripPx + (uint32_t)lane * 8;
  const bool active = c0 < W;
  const uint32_t w15 = W % 15, w31 = W % 31;
  const int s = p.shift;

  for (uint32_t t = blockIdx.x; t < total_tasks; t += gridDim.x) {
    const uint32_t f = p.list[t / p.bands], b = t % p.bands;
    const uint32_t y0 = b * p.band_rows;
    const uint32_t y1 = min(p.H, y0 + p.band_rows);
    const uint32_t assumed = p.stats[f].assumed;
    const bool use_delta = p.delta != nullptr && (assumed & 1u);
    const bool use_cg = (assumed & 2u) != 0;
    uint8_t* out_high = p.high + (uint64_t)f * p.P;
    uint8_t* out_low = mode_has_low(MODE) ? p.low + (uint64_t)f * p.P : nullptr;
    uint8_t* out_prev = p.preview_raw + (uint64_t)f * p.PP;

    uint32_t ph[4] = {0, 0, 0, 0}, pw[4] = {0, 0, 0, 0};  // previous row: high, west-shifted high
    uint32_t acc0 = 0, acc1 = 0, orl = 0;
    // running residues of the flat index of this lane's first pixel in row y
    uint32_t ystart = y0 > 0 ? y0 - 1 : 0;
    uint64_t i0 = (uint64_t)ystart * W + c0;
    uint32_t m15 = (uint32_t)(i0 % 15);
    // (i0 - (W+1)) mod 31, kept non-negative by adding a multiple of 31
    uint32_t m31 = (uint32_t)((i0 + 31ull * (W / 31 + 2) - (W + 1)) % 31);

    uint32_t y = ystart;
    while (y < y1) {
      const uint32_t nrows = (y < y0) ? 1 : min((uint32_t)kRowsPerStage, y1 - y);
      const uint32_t slot = seq % S, phs = (seq / S) & 1u;
      mbar_wait(full0 + 8 * slot, phs);
      const uint8_t* raw_s = smem + (size_t)slot * slot_bytes + kHaloPx * 2;  // pixel (y, 0)
      const uint8_t* del_s = raw_s + p.stage_bytes;

      for (uint32_t r = 0; r < nrows; r++, y++) {
        const bool own = y >= y0;
        const uint32_t roff = (r * W + c0) * 2;
        uint32_t xh[4], xl[4], h[4], l[4];
        {
          uint4 x = make_uint4(0, 0, 0, 0);
          if (active) x = *reinterpret_cast<const uint4*>(raw_s + roff);
          split2<MODE>(x.x, s, xh[0], xl[0]);
          split2<MODE>(x.y, s, xh[1], xl[1]);
          split2<MODE>(x.z, s, xh[2], xl[2]);
          split2<MODE>(x.w, s, xh[3], xl[3]);
        }
        if (use_delta) {
          uint4 d = make_uint4(0, 0, 0, 0);
          if (active) d = *reinterpret_cast<const uint4*>(del_s + roff);
          uint32_t dh, dl;
          split2_delta(d.x, dh, dl); h[0] = sub2(xh[0], dh); l[0] = xl[0] + kLaneBias - dl;
          split2_delta(d.y, dh, dl); h[1] = sub2(xh[1], dh); l[1] = xl[1] + kLaneBias - dl;
          split2_delta(d.z, dh, dl); h[2] = sub2(xh[2], dh); l[2] = xl[2] + kLaneBias - dl;
          split2_delta(d.w, dh, dl); h[3] = sub2(xh[3], dh); l[3] = xl[3] + kLaneBias - dl;
        } else {
#pragma unroll
          for (int j = 0; j < 4; j++) { h[j] = xh[j]; l[j] = xl[j]; }
        }
        // west neighbour of this lane's first pixel: previous lane's last pixel,
        // or (lane 0) the pixel before it in flat order, read from the stage.
        uint32_t left = __shfl_up_sync(0xffffffffu, h[3] >> 16, 1);
        if (lane == 0) {
          uint32_t hp, lp;
          split1<MODE>(*reinterpret_cast<const uint16_t*>(raw_s + roff - 2), s, hp, lp);
          if (use_delta)
            hp = (hp - (uint32_t)(*reinterpret_cast<const uint16_t*>(del_s + roff - 2) >> 8)) & 0xffu;
          left = hp;
        }
        uint32_t w[4];
        w[0] = (h[0] << 16) | left;
        w[1] = __funnelshift_l(h[0], h[1], 16);
        w[2] = __funnelshift_l(h[1], h[2], 16);
        w[3] = __funnelshift_l(h[2], h[3], 16);

        if (own) {
          uint32_t res[4];
          if (y == 0) {
#pragma unroll
            for (int j = 0; j < 4; j++) res[j] = h[j];
          } else {
#pragma unroll
            for (int j = 0; j < 4; j++) res[j] = sub2(h[j], cg2(ph[j], w[j], pw[j]));
            // flat index W (row 1, column 0) is copied, not predicted (.cc:566, :572)
            if (y == 1 && c0 == 0) res[0] = (res[0] & 0xffff0000u) | (h[0] & 0x0000ffffu);
          }
          if (active) {
            uint64_t o = (uint64_t)y * W + c0;
            uint2 hv = use_cg ? pack8(res[0], res[1], res[2], res[3]) : pack8(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint2*>(out_high + o) = hv;
            if (mode_has_low(MODE))
              *reinterpret_cast<uint2*>(out_low + o) = pack8(l[0], l[1], l[2], l[3]);
            orl |= xl[0] | xl[1] | xl[2] | xl[3];
            acc0 += xh[0] + xh[1];
            acc1 += xh[2] + xh[3];
            // delta-decision sample: flat index % 15 == 0 (.cc:526-531), RAW high byte
            uint32_t od = m15 ? 15 - m15 : 0;
            if (od < 8) {
              uint32_t v = sel4(xh, od >> 1);
              v = (od & 1) ? (v >> 16) : (v & 0xffffu);
              atomicAdd(&hist[v], 1u);
            }
            // CG-decision sample: flat index == W+1 (mod 31), >= W+1 (.cc:554-562)
            uint32_t oc = m31 ? 31 - m31 : 0;
            if (oc < 8 && y >= 1 && !(y == 1 && c0 == 0 && oc == 0)) {
              uint32_t a = sel4(h, oc >> 1), bb = sel4(res, oc >> 1);
              if (oc & 1) { a >>= 16; bb >>= 16; } else { a &= 0xffffu; bb &= 0xffffu; }
              atomicAdd(&hist[256 + a], 1u);
              atomicAdd(&hist[512 + bb], 1u);
            }
            if ((y & 3u) == 3u) {
              uint32_t s0 = (acc0 & 0xffffu) + (acc0 >> 16);
              uint32_t s1 = (acc1 & 0xffffu) + (acc1 >> 16);
              uint32_t pv = ((s0 >> 4) & 0xfeu) | (((s1 >> 4) & 0xfeu) << 8);
              *reinterpret_cast<uint16_t*>(out_prev + (uint64_t)(y >> 2) * p.PW + (c0 >> 2)) = (uint16_t)pv;
              acc0 = 0; acc1 = 0;
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) { ph[j] = h[j]; pw[j] = w[j]; }
        m15 += w15; if (m15 >= 15) m15 -= 15;
        m31 += w31; if (m31 >= 31) m31 -= 31;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty0 + 8 * slot);
      seq++;
    }

    // ---- end of task: publish low-OR and the three histograms -----------------
#pragma unroll
    for (int o = 16; o; o >>= 1) orl |= __shfl_xor_sync(0xffffffffu, orl, o);
    if (lane == 0 && (orl & kLaneMask)) atomicOr(&p.stats[f].low_or, orl & kLaneMask);
    named_bar_sync(1, NW * 32);
    uint32_t* gh = p.stats[f].hist_d;  // hist_d, hist_a, hist_b are contiguous
    for (uint32_t i = threadIdx.x; i < 768; i += NW * 32) {
      uint32_t v = hist[i];
      if (v) { atomicAdd(&gh[i], v); hist[i] = 0; }
    }
    named_bar_sync(1, NW * 32);
  }
}

// =====================================================================================
// Host-side launch logic
// =====================================================================================

static size_t fast_smem_bytes(uint32_t W, int stages) {
  size_t stage_bytes = ((size_t)kRowsPerStage * W + kHaloPx) * 2;
  return (size_t)stages * 2 * stage_bytes + 768 * 4 + 2 * (size_t)stages * 8;
}

bool encode_fast_supported(const Geom& g, const EncodeTuning& t) {
  if (g.W % 8 != 0 || g.W < 8) return false;
  if (g.W > 32 * 31 * 8) return false;  // at most 31 compute warps + 1 producer
  if (g.W > 4096) return false;
  int stages = t.stages < 2 ? 2 : t.stages;
  return fast_smem_bytes(g.W, stages) <= (size_t)t.max_smem_optin;
}

#define FPV_DISPATCH_MODE(mode, CALL)                       \
  switch (mode) {                                           \
    case kLE0: { constexpr int M = kLE0; CALL; } break;     \
    case kLE8: { constexpr int M = kLE8; CALL; } break;     \
    case kLEs: { constexpr int M = kLEs; CALL; } break;     \
    case kBE0: { constexpr int M = kBE0; CALL; } break;     \
    case kBE8: { constexpr int M = kBE8; CALL; } break;     \
    case kBEs: { constexpr int M = kBEs; CALL; } break;     \
    default:   { constexpr int M = kLEbig; CALL; } break;   \
  }

template <int MODE>
static cudaError_t launch_fast(const FastParams& fp, int grid, int threads, size_t smem,
                               cudaStream_t stream) {
  static bool attr_set[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 16 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(k_encode_fast<MODE>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  k_encode_fast<MODE><<<grid, threads, smem, stream>>>(fp);
  return cudaGetLastError();
}

int enqueue_delta_from_raw(const Geom& g, const uint16_t* raw, uint16_t* delta_image,
                           cudaStream_t stream, cudaError_t* err) {
  int blocks = (int)((g.P + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  FPV_DISPATCH_MODE(g.mode, (k_delta_from_raw<M><<<blocks, 256, 0, stream>>>(raw, delta_image, g.P, g.shift)));
  *err = cudaGetLastError();
  return *err == cudaSuccess ? 1 : -1;
}

int enqueue_encode(const Geom& g, const EncodeTuning& t, const EncodeScratch& s,
                   const uint16_t* frames, const uint16_t* delta, uint32_t n, bool force_generic,
                   uint8_t* flags, uint8_t* high, uint8_t* low, uint8_t* preview,
                   cudaStream_t stream, cudaError_t* err, const TimingHook* hook) {
  int launches = 0;
  const int has_delta = delta != nullptr;
  const int has_low = mode_has_low(g.mode);
  const bool fast = !force_generic && encode_fast_supported(g, t);
#define FPV_CHECK_LAUNCH()                         \
  do {                                             \
    launches++;                                    \
    *err = cudaGetLastError();                     \
    if (*err != cudaSuccess) return -1;            \
  } while (0)

  k_encode_init<<<n, 256, 0, stream>>>(s.stats, s.lists, s.counts, n, s.cap, has_delta, fast ? 1 : 0);
  FPV_CHECK_LAUNCH();

  if (fast) {
    FastParams fp;
    fp.frames = frames; fp.delta = delta; fp.stats = s.stats;
    fp.high = high; fp.low = low; fp.preview_raw = s.preview_raw;
    fp.W = g.W; fp.H = g.H; fp.P = g.P; fp.PP = g.PP; fp.PW = g.PW; fp.shift = g.shift;
    fp.stages = t.stages < 2 ? 2 : t.stages;
    fp.stage_bytes = (kRowsPerStage * g.W + kHaloPx) * 2;
    fp.compute_warps = (g.W + kStripPx - 1) / kStripPx;
    // Band height: whole multiples of 4 rows; shrink for small batches so that
    // there are at least ~4 tasks per SM.
    uint32_t band = (uint32_t)t.band_rows;
    band = (band / 4) * 4; if (band < 4) band = 4;
    while (band > 8 && (uint64_t)n * ((g.H + band - 1) / band) < (uint64_t)t.num_sms * 4) band = ((band / 2) / 4) * 4;
    if (band > g.H) band = g.H;
    fp.band_rows = band;
    fp.bands = (g.H + band - 1) / band;
    const int threads = (int)(fp.compute_warps + 1) * 32;
    const size_t smem = fast_smem_bytes(g.W, (int)fp.stages);
    int ctas_per_sm = (int)((size_t)(t.max_smem_optin + 1024) / (smem + 1024));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    int by_threads = 2048 / threads; if (by_threads < 1) by_threads = 1;
    if (ctas_per_sm > by_threads) ctas_per_sm = by_threads;
    uint64_t max_tasks = (uint64_t)n * fp.bands;
    int grid = t.num_sms * ctas_per_sm;
    if ((uint64_t)grid > max_tasks) grid = (int)max_tasks;
    for (int pass = 0; pass < 3; pass++) {
      fp.list = s.lists + (size_t)pass * s.cap;
      fp.count = s.counts + pass;
      // redo passes are almost always empty: a small grid is enough
      int gpass = pass == 0 ? grid : (grid < t.num_sms ? grid : t.num_sms);
      cudaError_t e = cudaSuccess;
      if (hook && pass == 0) cudaEventRecord(hook->start, stream);
      FPV_DISPATCH_MODE(g.mode, (e = launch_fast<M>(fp, gpass, threads, smem, stream)));
      if (hook && pass == 0) cudaEventRecord(hook->stop, stream);
      launches++;
      if (e != cudaSuccess) { *err = e; return -1; }
      if (pass < 2) {
        k_decide<<<n, 256, 0, stream>>>(s.stats, fp.list, fp.count, s.lists + (size_t)(pass + 1) * s.cap,
                                        s.counts + pass + 1, 0, has_delta, has_low);
        FPV_CHECK_LAUNCH();
      }
    }
  } else {
    const uint32_t chunk = 8192;
    dim3 gA((unsigned)((g.P + chunk - 1) / chunk), n);
    FPV_DISPATCH_MODE(g.mode, (k_gen_stats_delta<M><<<gA, 256, 0, stream>>>(frames, s.stats, g.P, g.shift, chunk)));
    FPV_CHECK_LAUNCH();
    k_decide<<<n, 256, 0, stream>>>(s.stats, s.lists, s.counts, s.lists + s.cap, s.counts + 1, 1, has_delta, has_low);
    FPV_CHECK_LAUNCH();
    uint64_t samples = g.P > (uint64_t)g.W + 1 ? (g.P - g.W - 1 + 30) / 31 : 0;
    unsigned gbx = (unsigned)((samples + 255) / 256); if (gbx < 1) gbx = 1; if (gbx > 1024) gbx = 1024;
    dim3 gB(gbx, n);
    FPV_DISPATCH_MODE(g.mode, (k_gen_stats_cg<M><<<gB, 256, 0, stream>>>(frames, delta, s.stats, g.W, g.P, g.shift)));
    FPV_CHECK_LAUNCH();
    k_decide<<<n, 256, 0, stream>>>(s.stats, s.lists, s.counts, s.lists + s.cap, s.counts + 1, 2, has_delta, has_low);
    FPV_CHECK_LAUNCH();
    unsigned gtx = (unsigned)((g.P / 4 + 255) / 256); if (gtx < 1) gtx = 1; if (gtx > 4096) gtx = 4096;
    dim3 gT(gtx, n);
    if (hook) cudaEventRecord(hook->start, stream);
    FPV_DISPATCH_MODE(g.mode, (k_gen_transform<M><<<gT, 256, 0, stream>>>(frames, delta, s.stats, high, low, g.W, g.P, g.shift)));
    if (hook) cudaEventRecord(hook->stop, stream);
    FPV_CHECK_LAUNCH();
    unsigned gpx = (unsigned)((g.PP + 255) / 256); if (gpx < 1) gpx = 1; if (gpx > 1024) gpx = 1024;
    dim3 gP(gpx, n);
    FPV_DISPATCH_MODE(g.mode, (k_gen_preview<M><<<gP, 256, 0, stream>>>(frames, s.preview_raw, g.W, g.P, g.PW, g.PP, g.shift)));
    FPV_CHECK_LAUNCH();
  }

  unsigned gfx = (unsigned)((g.PP + 255) / 256); if (gfx < 1) gfx = 1; if (gfx > 256) gfx = 256;
  dim3 gF(gfx, n);
  k_finalize<<<gF, 256, 0, stream>>>(s.stats, s.preview_raw, preview, flags, s.counts, n, g.PW, g.PP, has_low);
  FPV_CHECK_LAUNCH();
#undef FPV_CHECK_LAUNCH
  return launches;
}

}  // namespace fpv

// fpv_encode.cu -- encode-side kernels: Frame ctor + Frame::Predict on the GPU.
//
// Replaces fusion_power_video.cc:370-451 (split), :491-515 (preview),
// :517-544 (delta decision + apply), :546-593 (ClampedGradient decision +
// forward, high plane and preview).  See DESIGN.md for the layout.
//
// Two paths produce identical bytes:
//
//  * FAST (k_encode_fast): persistent CTAs stream contiguous 4-row stages of a
//    frame (or a band of one) into a shared-memory ring with 1-D TMA bulk copies
//    (cp.async.bulk + mbarrier), one warp per 256-column strip walks down the
//    rows keeping the previous row in registers, everything fused into one
//    read of the raw frame and one write of each output plane.  The
//    reference's per-frame decisions (USE_DELTA, USE_CG) depend on whole-frame
//    histograms, so the pass runs with ASSUMED flags while accumulating the
//    histograms; the CTA that completes a frame evaluates the integer
//    heuristics exactly inside the kernel (task_done / decide_core) and frames
//    whose assumption was wrong are redone (at most twice) by the same kernel;
//    k_finalize16 then predicts the preview planes.  In the common case
//    compulsory HBM traffic is 4.0625 B/pixel.
//
//  * GENERIC (k_gen_*): statistics first, then transform; plain loads, any
//    xsize % 4 == 0.  Used for geometries the bulk-copy path cannot take
//    (xsize % 8 != 0, very wide rows) and as an in-GPU cross-check.
#include <stdio.h>
#include <stdlib.h>

#include "fpv_internal.h"
#include "fpv_encode_fast.cuh"

namespace fpv {

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
    st.delta_known = 0;
    st.delta_dec = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) st.dbits[k] = 0;
    lists[f] = f;
    if (f == 0) {
      counts[0] = n;
      counts[1] = 0;
      counts[2] = 0;
    }
  }
  (void)cap;
}

// split1 with a run-time mode (cold paths only).
__device__ __forceinline__ uint32_t high1_rt(int mode, uint32_t p, int s) {
  uint32_t h, l;
  switch (mode) {
    case kLE0: split1<kLE0>(p, s, h, l); break;
    case kLE8: split1<kLE8>(p, s, h, l); break;
    case kLEs: split1<kLEs>(p, s, h, l); break;
    case kBE0: split1<kBE0>(p, s, h, l); break;
    case kBE8: split1<kBE8>(p, s, h, l); break;
    case kBEs: split1<kBEs>(p, s, h, l); break;
    default: split1<kLEbig>(p, s, h, l); break;
  }
  return h;
}

// Evaluates the reference's decisions for the frames of list `in`.
//  phase 0 (fast path): a transform pass just ran with st.assumed.  If the
//          delta decision differs, redo with the right delta (cg assumption
//          kept); else decide CG (its histograms were taken on the right
//          plane), fix final_flags, redo iff the cg assumption was wrong.
//  phase 1 (generic): delta decision only  -> st.assumed bit 0.
//  phase 2 (generic): CG decision only     -> st.final_flags, st.assumed.
//
// Delta decision.  countd is {0: N} in the reference (d = a - high_[i] with
// a == high_[i], .cc:527-529) and EstimateEntropy of it is 0, so
//     USE_DELTA  <=>  E(counta) > 0  <=>  1024 * S >= N,
// S = sum_v counta[v] * (floor(log2 N) - floor(log2 counta[v])) (.cc:235-244).
// At most one bin can reach 2^floor(log2 N) (a zero term); every other
// non-empty bin adds at least its count.  Hence for ANY split of the byte
// values into two groups A, B:  S >= min(|A|, |B|).  The fast kernel counts, per
// frame, the samples with bit k set (k = 0..7), i.e. eight such splits:
//   * some k with 1024 * min(c_k, N - c_k) >= N   ->  E > 0 is proven;
//   * every c_k in {0, N}: all samples equal      ->  E == 0 exactly;
//   * otherwise (a near-constant plane with a few strays) this block builds
//     the exact 256-bin histogram from the frame's samples.
__global__ void __launch_bounds__(256)
k_decide(FrameStat* stats, const uint32_t* in, const uint32_t* in_count, uint32_t* out,
         uint32_t* out_count, int phase, int has_delta, int has_low, const uint16_t* frames,
         uint64_t P, int mode, int shift) {
  __shared__ uint32_t red[8];
  __shared__ uint32_t exact[256];
  if (blockIdx.x >= *in_count) return;
  uint32_t f = in[blockIdx.x];
  FrameStat& st = stats[f];
  const int t = threadIdx.x;
  uint32_t assumed = st.assumed;
  uint32_t nolow = has_low ? (st.low_or == 0 ? kFlagNoLow : 0) : kFlagNoLow;

  uint32_t dec_delta = assumed & 1u;
  if (phase == 1) {
    uint64_t ea = block_entropy256(st.hist_d[t], red);
    dec_delta = (has_delta && ea > 0) ? 1u : 0u;
    if (t == 0) st.assumed = dec_delta;
    return;
  }
  if (phase == 0) {
    if (st.delta_known) {
      dec_delta = st.delta_dec;
    } else if (!has_delta) {
      dec_delta = 0;
    } else {
      const uint64_t N = (P + 14) / 15;
      bool proven = false, constant = true;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const uint64_t c = st.dbits[k];
        const uint64_t m = c < N - c ? c : N - c;
        proven = proven || (1024 * m >= N);
        constant = constant && (m == 0);
      }
      if (proven) dec_delta = 1;
      else if (constant) dec_delta = 0;
      else {
        exact[t] = 0;
        __syncthreads();
        const uint16_t* img = frames + (uint64_t)f * P;
        for (uint64_t i = 15ull * t; i < P; i += 15ull * 256) atomicAdd(&exact[high1_rt(mode, img[i], shift)], 1u);
        __syncthreads();
        dec_delta = block_entropy256(exact[t], red) > 0 ? 1u : 0u;
      }
    }
    __syncthreads();
    if (t == 0) { st.delta_known = 1; st.delta_dec = dec_delta; }
  }
  if (phase == 0 && dec_delta != (assumed & 1u)) {
    st.hist_a[t] = 0; st.hist_b[t] = 0;
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
    st.hist_a[t] = 0; st.hist_b[t] = 0;
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
// flags_in != nullptr (fast path): the flags byte is final already (frame_decide wrote it), only the preview is done.
__global__ void k_finalize(const FrameStat* stats, const uint8_t* preview_raw, uint8_t* preview,
                           uint8_t* flags, uint32_t* counts, uint32_t n, uint32_t PW,
                           uint64_t PP, int has_low, const uint8_t* flags_in) {
  uint32_t f = blockIdx.y;
  const FrameStat& st = stats[f];
  // NO_LOW_BYTES is only known once every pass has OR-ed its low bytes in.
  uint32_t fin = flags_in ? flags_in[f]
                          : (st.final_flags & 3u) | (has_low ? (st.low_or == 0 ? kFlagNoLow : 0) : kFlagNoLow);
  const uint8_t* pr = preview_raw + (uint64_t)f * PP;
  uint8_t* po = preview + (uint64_t)f * PP;
  if ((PP & 3) == 0 && (PW & 3) == 0) {
    // 4 preview pixels per thread: word loads, ClampedGradient in lane form.
    const uint32_t* pr32 = reinterpret_cast<const uint32_t*>(pr);
    uint32_t* po32 = reinterpret_cast<uint32_t*>(po);
    const uint64_t words = PP / 4, pww = PW / 4;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < words;
         q += (uint64_t)gridDim.x * blockDim.x) {
      uint32_t c = pr32[q];
      uint32_t outw = c;
      if ((fin & kFlagCG) && q >= pww) {          // rows >= 1; pixel PW itself is patched below
        uint32_t cm = pr32[q - 1];                // 4 pixels to the west (flat order)
        uint32_t n = pr32[q - pww];
        uint32_t nm = q > pww ? pr32[q - pww - 1] : 0u;
        uint32_t wv = __funnelshift_l(cm, c, 8);  // bytes i-1 .. i+2
        uint32_t nwv = __funnelshift_l(nm, n, 8);
        // lane form: pixels (0,1) and (2,3)
        uint32_t c01 = __byte_perm(c, 0u, 0x4140), c23 = __byte_perm(c, 0u, 0x4342);
        uint32_t n01 = __byte_perm(n, 0u, 0x4140), n23 = __byte_perm(n, 0u, 0x4342);
        uint32_t w01 = __byte_perm(wv, 0u, 0x4140), w23 = __byte_perm(wv, 0u, 0x4342);
        uint32_t q01 = __byte_perm(nwv, 0u, 0x4140), q23 = __byte_perm(nwv, 0u, 0x4342);
        uint32_t r01 = sub2(c01, cg2(n01, w01, q01)), r23 = sub2(c23, cg2(n23, w23, q23));
        outw = __byte_perm(r01, r23, 0x6420);
        if (q == pww) outw = (outw & 0xffffff00u) | (c & 0xffu);  // index PW is copied (.cc:578, :584)
      }
      po32[q] = outw;
    }
  } else {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < PP;
         i += (uint64_t)gridDim.x * blockDim.x) {
      uint32_t v = pr[i];
      if ((fin & kFlagCG) && i > PW) v = (v - cg1(pr[i - PW], pr[i - 1], pr[i - PW - 1])) & 0xffu;
      po[i] = (uint8_t)v;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && !flags_in) {
    flags[f] = (uint8_t)fin;
    if (f == n - 1) counts[3] = fin & 3u;  // guess for the next batch
  }
}

// The same for PW % 16 == 0 (every geometry of the benchmarks), 16 preview pixels at a time.  The 4-pixel kernel
// above is latency bound: its threads walk ~30 dependent iterations.
// One 16-pixel group of a preview plane (PW % 16 == 0): ClampedGradient on the preview's own flat array.
__device__ __forceinline__ uint4 finalize_group(const uint4 c, const uint4 nn, const uint32_t cm, const uint32_t nm,
                                                const bool first_of_row1) {
  uint4 o;
  o.x = finalize_word(c.x, cm, nn.x, nm);
  o.y = finalize_word(c.y, c.x, nn.y, nn.x);
  o.z = finalize_word(c.z, c.y, nn.z, nn.y);
  o.w = finalize_word(c.w, c.z, nn.w, nn.z);
  if (first_of_row1) o.x = (o.x & 0xffffff00u) | (c.x & 0xffu);  // index PW is copied (.cc:578, :584)
  return o;
}

// GROUPS 16-pixel groups per thread, every load issued before the first use (the one-group-per-thread form was
// latency bound: 44 us per 1184 previews of 320x200, 3.4 TB/s of its 152 MB).
constexpr int kFinGroups = 4;
__global__ void __launch_bounds__(256)
k_finalize16(const FrameStat* __restrict__ stats, const uint8_t* __restrict__ preview_raw,
             uint8_t* __restrict__ preview, uint8_t* __restrict__ flags, uint32_t* counts, uint32_t n,
             uint32_t PW, uint64_t PP, int has_low, const uint8_t* __restrict__ flags_in) {
  const uint32_t f = blockIdx.y;
  const uint32_t groups = (uint32_t)(PP / 16), pw16 = PW / 16;
  const FrameStat& st = stats[f];
  const uint32_t fin = flags_in ? flags_in[f]
                                : (st.final_flags & 3u) | (has_low ? (st.low_or == 0 ? kFlagNoLow : 0) : kFlagNoLow);
  if (blockIdx.x == 0 && threadIdx.x == 0 && !flags_in) {
    flags[f] = (uint8_t)fin;
    if (f == n - 1) counts[3] = fin & 3u;  // guess for the next batch
  }
  const uint4* pr = reinterpret_cast<const uint4*>(preview_raw + (uint64_t)f * PP);
  const uint32_t* pr32 = reinterpret_cast<const uint32_t*>(pr);
  uint4* po = reinterpret_cast<uint4*>(preview + (uint64_t)f * PP);
  const uint32_t q0 = blockIdx.x * (blockDim.x * kFinGroups) + threadIdx.x;   // groups q0 + 256 k of this thread
  const bool cg = (fin & kFlagCG) != 0;
  uint4 c[kFinGroups], nn[kFinGroups];
  uint32_t cm[kFinGroups], nm[kFinGroups];
#pragma unroll
  for (int k = 0; k < kFinGroups; k++) {
    const uint32_t q = q0 + 256u * k;
    const bool in = q < groups, pred = in && cg && q >= pw16;
    c[k] = in ? __ldg(pr + q) : make_uint4(0, 0, 0, 0);
    nn[k] = pred ? __ldg(pr + q - pw16) : make_uint4(0, 0, 0, 0);
    cm[k] = pred ? __ldg(pr32 + 4 * q - 1) : 0u;                        // 4 pixels to the west (flat order)
    nm[k] = pred && q > pw16 ? __ldg(pr32 + 4 * (q - pw16) - 1) : 0u;
  }
#pragma unroll
  for (int k = 0; k < kFinGroups; k++) {
    const uint32_t q = q0 + 256u * k;
    if (q >= groups) continue;
    po[q] = (cg && q >= pw16) ? finalize_group(c[k], nn[k], cm[k], nm[k], q == pw16) : c[k];
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

// Frame ctor alone (.cc:388-449): split into byte planes + OR of the low bytes, no prediction.  Serves the
// fpvc::Frame facade (planes of a frame that is never predicted, e.g. one that becomes a delta frame).
template <int MODE>
__global__ void __launch_bounds__(256)
k_split_planes(const uint16_t* __restrict__ frames, uint8_t* __restrict__ high, uint8_t* __restrict__ low,
               uint32_t* __restrict__ low_or, uint64_t P, int s) {
  const uint32_t f = blockIdx.y;
  const uint16_t* img = frames + (uint64_t)f * P;
  uint32_t lor = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t h, l;
    split1<MODE>(img[i], s, h, l);
    high[(uint64_t)f * P + i] = (uint8_t)h;
    if (mode_has_low(MODE)) low[(uint64_t)f * P + i] = (uint8_t)l;
    lor |= l;
  }
  lor = __reduce_or_sync(0xffffffffu, lor);
  if ((threadIdx.x & 31) == 0 && lor) atomicOr(&low_or[f], lor);
}

// =====================================================================================
// Host-side launch logic
// =====================================================================================

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

int enqueue_split(const Geom& g, const uint16_t* frames, uint32_t n, uint8_t* high, uint8_t* low, uint32_t* low_or,
                  cudaStream_t stream, cudaError_t* err) {
  *err = cudaMemsetAsync(low_or, 0, sizeof(uint32_t) * n, stream);
  if (*err != cudaSuccess) return -1;
  unsigned gx = (unsigned)((g.P + 255) / 256);
  if (gx > 1024) gx = 1024;
  FPV_DISPATCH_MODE(g.mode, (k_split_planes<M><<<dim3(gx, n), 256, 0, stream>>>(frames, high, low, low_or, g.P, g.shift)));
  *err = cudaGetLastError();
  return *err == cudaSuccess ? 1 : -1;
}

bool encode_fast_supported(const Geom& g, const EncodeTuning& t) {
  if (g.W % 16 != 0 || g.W < 16) return false;   // 8 pixels per lane; the in-kernel preview pass works on words of 4 preview pixels
  if (g.W > 32 * 31 * 8) return false;  // at most 31 compute warps + 1 producer
  if (g.W > 4096) return false;
  int stages = t.stages < 2 ? 2 : t.stages;
  return fast_smem_bytes(g.W, stages, t.rows_per_stage == 2 ? 2 : 4) <= (size_t)t.max_smem_optin;
}


template <int MODE, bool FULL, int RPS, bool PASS0>
static cudaError_t launch_fast(const FastParams& fp, int grid, int threads, size_t smem,
                               cudaStream_t stream) {
  // per launch: the attribute is per device and is lost by cudaDeviceReset; the call costs about a microsecond
  static const bool wide72 = getenv("FPV_FAST_NO_R72") == nullptr;
  if (threads > 224 && RPS == 4 && wide72) {
    // nine (or more) warps per CTA: the 72-register build, three CTAs per SM
    cudaError_t ea = cudaFuncSetAttribute(k_encode_fast<MODE, FULL, RPS, PASS0, 72>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          232448 - 1024);
    if (ea != cudaSuccess) return ea;
    k_encode_fast<MODE, FULL, RPS, PASS0, 72><<<grid, threads, smem, stream>>>(fp);
    return cudaGetLastError();
  }
  cudaError_t ea = cudaFuncSetAttribute(k_encode_fast<MODE, FULL, RPS, PASS0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        232448 - 1024);
  if (ea != cudaSuccess) return ea;
  k_encode_fast<MODE, FULL, RPS, PASS0><<<grid, threads, smem, stream>>>(fp);
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

int enqueue_encode(const Geom& g, const EncodeTuning& t, EncodeScratch& s,
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

  if (fast) {
    // ---- fast path: one launch does everything per frame (transform, decisions, flags, preview prediction); two
    //      more launches redo the frames whose assumed flags were wrong -- normally none, they exit at once.
    if (s.dirty) {
      k_encode_init<<<n, 256, 0, stream>>>(s.stats, s.lists, s.counts, n, s.cap, has_delta, 0);
      FPV_CHECK_LAUNCH();
      s.dirty = false;
    }
    *err = cudaMemsetAsync(s.counts + 1, 0, 2 * sizeof(uint32_t), stream);
    if (*err != cudaSuccess) return -1;
    FastParams fp;
    fp.frames = frames; fp.delta = delta; fp.stats = s.stats;
    fp.n = n;
    fp.guess_in = s.counts + 3 + (s.calls & 1u);
    fp.guess_out = s.counts + 3 + ((s.calls + 1) & 1u);
    s.calls++;
    if (!s.preview_raw) {
      *err = cudaMalloc(&s.preview_raw, (size_t)s.cap * (g.PP ? g.PP : 1));
      if (*err != cudaSuccess) return -1;
    }
    fp.high = high; fp.low = low; fp.preview_raw = s.preview_raw; fp.flags = flags;
    fp.W = g.W; fp.H = g.H; fp.P = g.P; fp.PP = g.PP; fp.PW = g.PW; fp.shift = g.shift;
    const int rps = t.rows_per_stage == 2 ? 2 : 4;
    fp.stages = t.stages < 2 ? 2 : t.stages;
    fp.rows_per_stage = (uint32_t)rps;
    fp.stage_bytes = ((uint32_t)rps * g.W + kHaloPx) * 2;
    fp.compute_warps = (g.W + kStripPx - 1) / kStripPx;
    fp.qc = make_qconst(g.mode, g.shift);
    // Band height: whole multiples of 4 rows; shrink for small batches so that
    // there are at least ~4 tasks per SM.
    uint32_t band = (uint32_t)t.band_rows;
    band = (band / 4) * 4; if (band < 4) band = 4;
    if (band > 1024) band = 1024;  // per-lane delta-decision counters are 16 bit
    while (band > 8 && (uint64_t)n * ((g.H + band - 1) / band) < (uint64_t)t.num_sms * 4) band = ((band / 2) / 4) * 4;
    if (band > g.H) band = g.H;
    const int threads = (int)(fp.compute_warps + 1) * 32;
    const size_t smem = fast_smem_bytes(g.W, (int)fp.stages, rps);
    int ctas_per_sm = (int)((size_t)(t.max_smem_optin + 1024) / (smem + 1024));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    int by_threads = 2048 / threads; if (by_threads < 1) by_threads = 1;
    if (ctas_per_sm > by_threads) ctas_per_sm = by_threads;
    if (t.max_ctas > 0 && ctas_per_sm > t.max_ctas) ctas_per_sm = t.max_ctas;
    int grid = t.num_sms * ctas_per_sm;
    // A task is a WHOLE FRAME whenever that keeps the CTAs evenly loaded: the frame's statistics then never leave
    // the CTA's shared memory (the decision is taken there, task_done), and no halo row is read twice.  A batch
    // that does not fill its last wave of frame tasks (2500 frames on 296 CTAs) is cut in two launches: whole waves of
    // frame tasks, then the remaining frames as band tasks (band tasks for everything cost 15 % at 2048x2048).
    uint32_t n_whole = 0;             // frames [0, n_whole) go as frame tasks in a launch of their own
    if ((uint64_t)n >= (uint64_t)grid && g.P / 480 < 60000 && !getenv("FPV_NO_FRAME_TASKS")) {
      const uint64_t waves = ((uint64_t)n + grid - 1) / grid;
      if ((double)n / (double)(waves * grid) >= 0.94) band = g.H;
      else if (!getenv("FPV_NO_SPLIT_LAUNCH")) n_whole = (n / (uint32_t)grid) * (uint32_t)grid;
    }
    if (n_whole) {
      // band height of the remaining frames (same rule as above)
      const uint32_t rest = n - n_whole;
      band = ((uint32_t)t.band_rows / 4) * 4; if (band < 4) band = 4;
      if (band > 1024) band = 1024;
      while (band > 8 && (uint64_t)rest * ((g.H + band - 1) / band) < (uint64_t)t.num_sms * 4) band = ((band / 2) / 4) * 4;
      if (band > g.H) band = g.H;
    }
    fp.band_rows = band;
    fp.bands = (g.H + band - 1) / band;
    fp.f0 = n_whole; fp.nf = n - n_whole;
    const int grid_all = grid;
    uint64_t max_tasks = (uint64_t)fp.nf * fp.bands;
    if ((uint64_t)grid > max_tasks) grid = (int)max_tasks;
    const bool full = g.W % kStripPx == 0;
    for (int pass = (n_whole ? -1 : 0); pass < 3; pass++) {
      // pass -1: the whole waves of frame tasks of a batch cut in two (see above)
      FastParams fw = fp;
      if (pass == -1) { fw.band_rows = g.H; fw.bands = 1; fw.f0 = 0; fw.nf = n_whole; }
      const FastParams& fq = pass == -1 ? fw : fp;
      const bool first_launch = pass == (n_whole ? -1 : 0);
      if (pass >= 0) {
        fp.list = pass == 0 ? nullptr : s.lists + (size_t)pass * s.cap;
        fp.count = pass == 0 ? nullptr : s.counts + pass;
        fp.next_list = pass < 2 ? s.lists + (size_t)(pass + 1) * s.cap : nullptr;
        fp.next_count = pass < 2 ? s.counts + pass + 1 : nullptr;
      } else {
        fw.list = nullptr; fw.count = nullptr;
        fw.next_list = s.lists + (size_t)1 * s.cap; fw.next_count = s.counts + 1;
      }
      // redo passes are almost always empty: a small grid is enough
      int gpass = pass == -1 ? grid_all : pass == 0 ? grid : (grid < 2 * t.num_sms ? grid : 2 * t.num_sms);
      if (pass > 0 && gpass < 1) gpass = 1;
      cudaError_t e = cudaSuccess;
      if (hook && first_launch) cudaEventRecord(hook->start, stream);
#define FPV_LAUNCH_FAST(F, R)                                                                                       \
  do {                                                                                                              \
    if (pass <= 0) { FPV_DISPATCH_MODE(g.mode, (e = launch_fast<M, F, R, true>(fq, gpass, threads, smem, stream))); } \
    else { FPV_DISPATCH_MODE(g.mode, (e = launch_fast<M, F, R, false>(fq, gpass, threads, smem, stream))); }         \
  } while (0)
      if (rps == 4) { if (full) FPV_LAUNCH_FAST(true, 4); else FPV_LAUNCH_FAST(false, 4); }
      else { if (full) FPV_LAUNCH_FAST(true, 2); else FPV_LAUNCH_FAST(false, 2); }
#undef FPV_LAUNCH_FAST
      if (hook && pass == 0) cudaEventRecord(hook->stop, stream);   // (covers both pass-0 launches of a batch cut in two)
      launches++;
      if (e != cudaSuccess) { *err = e; return -1; }
    }
    // the preview's own ClampedGradient pass (.cc:575-586), under the flags the passes above made final
    if (g.PW % 16 == 0 && (reinterpret_cast<uintptr_t>(preview) & 15) == 0) {
      dim3 g16((unsigned)((g.PP / 16 + 256 * kFinGroups - 1) / (256 * kFinGroups)), n);
      k_finalize16<<<g16, 256, 0, stream>>>(s.stats, s.preview_raw, preview, flags, s.counts, n, g.PW, g.PP, has_low, flags);
    } else {
      unsigned gfx = (unsigned)((g.PP / 4 + 255) / 256); if (gfx < 1) gfx = 1;
      { unsigned cap = n >= 1024 ? 2u : n >= 256 ? 4u : n >= 64 ? 16u : 64u; if (gfx > cap) gfx = cap; }
      k_finalize<<<dim3(gfx, n), 256, 0, stream>>>(s.stats, s.preview_raw, preview, flags, s.counts, n, g.PW, g.PP, has_low, flags);
    }
    FPV_CHECK_LAUNCH();
    return launches;
  }

  // ---- generic path: statistics first, then transform (plain loads, any xsize % 4 == 0) ----------------------
  s.dirty = true;
  if (!s.preview_raw) {
    *err = cudaMalloc(&s.preview_raw, (size_t)s.cap * (g.PP ? g.PP : 1));
    if (*err != cudaSuccess) return -1;
  }
  k_encode_init<<<n, 256, 0, stream>>>(s.stats, s.lists, s.counts, n, s.cap, has_delta, 0);
  FPV_CHECK_LAUNCH();
  {
    const uint32_t chunk = 8192;
    dim3 gA((unsigned)((g.P + chunk - 1) / chunk), n);
    FPV_DISPATCH_MODE(g.mode, (k_gen_stats_delta<M><<<gA, 256, 0, stream>>>(frames, s.stats, g.P, g.shift, chunk)));
    FPV_CHECK_LAUNCH();
    k_decide<<<n, 256, 0, stream>>>(s.stats, s.lists, s.counts, s.lists + s.cap, s.counts + 1, 1, has_delta, has_low, frames, g.P, g.mode, g.shift);
    FPV_CHECK_LAUNCH();
    uint64_t samples = g.P > (uint64_t)g.W + 1 ? (g.P - g.W - 1 + 30) / 31 : 0;
    unsigned gbx = (unsigned)((samples + 255) / 256); if (gbx < 1) gbx = 1; if (gbx > 1024) gbx = 1024;
    dim3 gB(gbx, n);
    FPV_DISPATCH_MODE(g.mode, (k_gen_stats_cg<M><<<gB, 256, 0, stream>>>(frames, delta, s.stats, g.W, g.P, g.shift)));
    FPV_CHECK_LAUNCH();
    k_decide<<<n, 256, 0, stream>>>(s.stats, s.lists, s.counts, s.lists + s.cap, s.counts + 1, 2, has_delta, has_low, frames, g.P, g.mode, g.shift);
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

  // a few fat blocks per frame: block scheduling, not bandwidth, dominates otherwise
  unsigned gfx = (unsigned)((g.PP / 4 + 255) / 256); if (gfx < 1) gfx = 1;
  { unsigned cap = n >= 1024 ? 2u : n >= 256 ? 4u : n >= 64 ? 16u : 64u; if (gfx > cap) gfx = cap; }
  dim3 gF(gfx, n);
  if (g.PW % 16 == 0 && g.PP % 16 == 0 && (reinterpret_cast<uintptr_t>(preview) & 15) == 0 && !getenv("FPV_FINALIZE4")) {
    dim3 g16((unsigned)((g.PP / 16 + 256 * kFinGroups - 1) / (256 * kFinGroups)), n);
    k_finalize16<<<g16, 256, 0, stream>>>(s.stats, s.preview_raw, preview, flags, s.counts, n, g.PW, g.PP, has_low, nullptr);
  } else {
    k_finalize<<<gF, 256, 0, stream>>>(s.stats, s.preview_raw, preview, flags, s.counts, n, g.PW, g.PP, has_low, nullptr);
  }
  FPV_CHECK_LAUNCH();
#undef FPV_CHECK_LAUNCH
  return launches;
}

}  // namespace fpv

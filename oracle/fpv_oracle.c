/*
 * fpv_oracle.c -- CPU ORACLE for the fusion-power-video pre-entropy transform.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is a plain-C restatement of the
 * reference algorithm (google/fusion-power-video, fusion_power_video.cc).  It
 * exists so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg can check the CUDA path.  Nothing in fusion_power_video_b200/ may
 * include, link or call it: the product path is the CUDA extension and fails
 * loudly without it.
 *
 * Parity status: PINNED.  The reference ships no golden vectors (SURVEY.md
 * section 4), so this restatement is pinned against the reference itself: the
 * unmodified reference sources are compiled in place from /root/reference into
 * oracle/_ref/libfpv_ref.so (oracle/Makefile, oracle/ref_harness.cc) and
 * tests/test_oracle_vs_ref.py compares every function below with it; the
 * vectors that comparison produced are committed under tests/golden/ (made by
 * tests/golden/make_golden.py) so the pin also holds where /root/reference
 * does not exist (the GPU box).
 *
 * All citations are file:line into /root/reference/fusion_power_video.cc.
 * All arithmetic is integer; bytes wrap mod 256.
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define FPVO_USE_DELTA 1u     /* fusion_power_video.h:68-73 FrameFlags */
#define FPVO_USE_CG 2u
#define FPVO_NO_LOW_BYTES 4u

/* floor(log2(v)) for v > 0 (.cc:216-232, approxLog2). */
static uint32_t fpvo_floor_log2(uint64_t v) {
  uint32_t r = 0;
  while (v >>= 1) r++;
  return r;
}

/* .cc:235-244 EstimateEntropy.  The reference accumulates with
 * std::accumulate(..., 0, ...): the accumulator type is `int`, so both sums
 * are truncated to 32 bits after every step.  Bins with v == 0 contribute
 * v * (anything) == 0.  Result is floor(1024 * S / sum) in 64-bit. */
uint64_t fpvo_estimate_entropy(const uint64_t counts[256]) {
  int32_t sum32 = 0;
  for (int i = 0; i < 256; i++) sum32 = (int32_t)((uint64_t)(int64_t)sum32 + counts[i]);
  uint64_t sum = (uint64_t)(int64_t)sum32;
  if (sum == 0) return 0;
  uint64_t log2sum = fpvo_floor_log2(sum);
  int32_t acc = 0;
  for (int i = 0; i < 256; i++) {
    uint64_t v = counts[i];
    uint64_t l = v ? fpvo_floor_log2(v) : 0; /* multiplied by v == 0 anyway */
    uint64_t next = (uint64_t)(int64_t)acc - v * (l - log2sum);
    acc = (int32_t)next;
  }
  uint64_t sum_of_logs = (uint64_t)(int64_t)acc;
  return 1024 * sum_of_logs / sum;
}

/* .cc:247-252 ClampedGradient, restated literally on uint8. */
uint8_t fpvo_clamped_gradient(uint8_t n, uint8_t w, uint8_t nw) {
  uint8_t lo = n < w ? n : w;
  uint8_t hi = n < w ? w : n;
  uint8_t grad = (uint8_t)(n + w - nw);
  uint8_t clamped = (nw < lo) ? hi : grad;
  return (nw > hi) ? lo : clamped;
}

/* .cc:370-451 Frame::Frame(u16): endian-normalise, left-align, split into a
 * high-byte and a low-byte plane.  Six expression variants, restated verbatim
 * (SURVEY.md 9.1).  Host is little-endian so switch_endian == big_endian
 * (.cc:366-368, 379).  Returns the initial flags: NO_LOW_BYTES iff the OR of
 * all low bytes is zero (.cc:447-449; always for shift 8, where no low plane
 * is produced and `low` is not written). */
uint8_t fpvo_split(const uint16_t* img, size_t n, int shift, int big_endian,
                   uint8_t* high, uint8_t* low) {
  uint8_t non_zero_low = 0;
  if (big_endian) {
    if (shift == 0) {                                   /* .cc:391-397 */
      for (size_t i = 0; i < n; i++) {
        uint16_t p = img[i];
        high[i] = (uint8_t)(p & 0xff);
        low[i] = (uint8_t)((p >> 8) & 0xff);
        non_zero_low |= low[i];
      }
    } else if (shift == 8) {                            /* .cc:401-403 */
      for (size_t i = 0; i < n; i++) high[i] = (uint8_t)((img[i] >> 8) & 0xff);
    } else {                                            /* .cc:407-415 */
      int low_shift = 8 - shift, low_shift_high = 16 - shift;
      for (size_t i = 0; i < n; i++) {
        uint16_t p = img[i];
        /* operands promote to int, as in the reference expression */
        high[i] = (uint8_t)((((int)p << shift) | ((int)p >> low_shift_high)) & 0xff);
        low[i] = (uint8_t)(((int)p >> low_shift) & 0xff);
        non_zero_low |= low[i];
      }
    }
  } else if (shift == 0) {                              /* .cc:421-427 */
    for (size_t i = 0; i < n; i++) {
      uint16_t p = img[i];
      high[i] = (uint8_t)((p >> 8) & 0xff);
      low[i] = (uint8_t)(p & 0xff);
      non_zero_low |= low[i];
    }
  } else if (shift == 8) {                              /* .cc:431-433 */
    for (size_t i = 0; i < n; i++) high[i] = (uint8_t)(img[i] & 0xff);
  } else {                                              /* .cc:437-443 */
    for (size_t i = 0; i < n; i++) {
      uint16_t p = (uint16_t)((int)img[i] << shift);
      high[i] = (uint8_t)((p >> 8) & 0xff);
      low[i] = (uint8_t)(p & 0xff);
      non_zero_low |= low[i];
    }
  }
  return non_zero_low ? 0 : FPVO_NO_LOW_BYTES;
}

/* .cc:491-515 GeneratePreview: 4x4 box sum of the RAW high plane,
 * (sum / 16) & 0xfe, (W/4) x (H/4) outputs. */
void fpvo_preview(const uint8_t* high, size_t W, size_t H, uint8_t* preview) {
  size_t pw = W / 4, ph = H / 4;
  for (size_t py = 0; py < ph; py++)
    for (size_t px = 0; px < pw; px++) {
      uint32_t sum = 0;
      for (size_t j = 0; j < 4; j++)
        for (size_t i = 0; i < 4; i++) sum += high[(py * 4 + j) * W + px * 4 + i];
      preview[py * pw + px] = (uint8_t)((sum / 16) & 0xfe);
    }
}

/* .cc:522-533 delta decision.  Reproduces the reference's `d = a - high_[i]`
 * with a == high_[i] (so d == 0 always): countd = {0: N}.  Returns 1 iff
 * EstimateEntropy(countd) < EstimateEntropy(counta). */
int fpvo_delta_decide(const uint8_t* high, size_t n) {
  uint64_t counta[256], countd[256];
  memset(counta, 0, sizeof counta);
  memset(countd, 0, sizeof countd);
  for (size_t i = 0; i < n; i += 15) {
    uint8_t a = high[i];
    uint8_t d = (uint8_t)(a - high[i]);
    counta[a]++;
    countd[d]++;
  }
  return fpvo_estimate_entropy(countd) < fpvo_estimate_entropy(counta);
}

/* .cc:550-564 ClampedGradient decision on the (post-delta) high plane:
 * samples at i = W+1, W+32, ... ; flat 1-D neighbours, no row-start case. */
int fpvo_cg_decide(const uint8_t* high, size_t W, size_t n) {
  uint64_t counta[256], countb[256];
  memset(counta, 0, sizeof counta);
  memset(countb, 0, sizeof countb);
  for (size_t i = W + 1; i < n; i += 31) {
    uint8_t a = high[i];
    uint8_t b = (uint8_t)(a - fpvo_clamped_gradient(high[i - W], high[i - 1], high[i - W - 1]));
    counta[a]++;
    countb[b]++;
  }
  return fpvo_estimate_entropy(countb) < fpvo_estimate_entropy(counta);
}

/* .cc:565-573 (and :577-585 for the preview): forward ClampedGradient over a
 * flat plane of `n` bytes with row pitch W.  Reads only un-predicted values;
 * the first W+1 bytes are copied. */
void fpvo_cg_forward(const uint8_t* in, size_t W, size_t n, uint8_t* out) {
  for (size_t i = 0; i < n && i <= W; i++) out[i] = in[i];
  for (size_t i = W + 1; i < n; i++)
    out[i] = (uint8_t)(in[i] - fpvo_clamped_gradient(in[i - W], in[i - 1], in[i - W - 1]));
}

/* .cc:617-622 / :326-333 inverse ClampedGradient, in place, strictly serial
 * in flat order (each pixel uses the just-reconstructed west neighbour). */
void fpvo_cg_inverse(uint8_t* plane, size_t W, size_t n) {
  for (size_t i = W + 1; i < n; i++)
    plane[i] = (uint8_t)(plane[i] + fpvo_clamped_gradient(plane[i - W], plane[i - 1], plane[i - W - 1]));
}

/*
 * Frame(u16 ctor) followed by Frame::Predict(delta_frame): .cc:370-451 then
 * .cc:777-785 (preview -> delta iff a delta frame is given -> CG).
 *
 *   img          W*H native-endian uint16 as read from the raw file
 *   dhigh/dlow   the delta frame's split planes (9.1 applied to the delta
 *                frame, NOT CG-predicted, .cc:1097); dhigh == NULL means "no
 *                delta frame" (Frame::EMPTY, .cc:780).  dlow is ignored for
 *                shift 8.
 *   high, low    outputs, W*H bytes each (low untouched for shift 8)
 *   preview      output, (W/4)*(H/4) bytes
 * Returns the frame flags byte.  Requires W % 4 == 0 and H % 4 == 0 (the
 * reference reads out of bounds otherwise, .cc:577-578).
 */
uint8_t fpvo_predict(const uint16_t* img, size_t W, size_t H, int shift, int big_endian,
                     const uint8_t* dhigh, const uint8_t* dlow,
                     uint8_t* high, uint8_t* low, uint8_t* preview, uint8_t* scratch) {
  size_t n = W * H;
  uint8_t flags = fpvo_split(img, n, shift, big_endian, high, low);
  fpvo_preview(high, W, H, preview);                     /* .cc:778 */
  if (dhigh) {                                           /* .cc:780-782 */
    if (fpvo_delta_decide(high, n)) {                    /* .cc:533-540 */
      for (size_t i = 0; i < n; i++) high[i] = (uint8_t)(high[i] - dhigh[i]);
      if (shift != 8)
        for (size_t i = 0; i < n; i++) low[i] = (uint8_t)(low[i] - dlow[i]);
      flags |= FPVO_USE_DELTA;
    }
  }
  if (fpvo_cg_decide(high, W, n)) {                      /* .cc:564-589 */
    fpvo_cg_forward(high, W, n, scratch);
    memcpy(high, scratch, n);
    size_t pw = W / 4, pn = n / 16;
    fpvo_cg_forward(preview, pw, pn, scratch);
    memcpy(preview, scratch, pn);
    flags |= FPVO_USE_CG;
  }
  return flags;
}

/*
 * Post-brotli part of DecompressImage (.cc:326-344): inverse CG on the high
 * plane (in place -- `high` is clobbered), then delta add with independent
 * byte wrap and recombination into uint16.  low == NULL means flags & 4
 * (all low bytes zero, .cc:313-314).  delta is the decoded delta frame as
 * uint16 (may be NULL when flags & 1 is clear).
 */
void fpvo_inverse(uint8_t* high, const uint8_t* low, const uint16_t* delta,
                  size_t W, size_t H, uint8_t flags, uint16_t* img) {
  size_t n = W * H;
  if (flags & FPVO_USE_CG) fpvo_cg_inverse(high, W, n);
  if (flags & FPVO_USE_DELTA) {
    for (size_t i = 0; i < n; i++) {
      uint8_t l = low ? low[i] : 0;
      img[i] = (uint16_t)(((high[i] + (delta[i] >> 8)) << 8) | ((l + (delta[i] & 0xff)) & 0xff));
    }
  } else {
    for (size_t i = 0; i < n; i++) img[i] = (uint16_t)((high[i] << 8) | (low ? low[i] : 0));
  }
}

/* .cc:850-862 UnextractFrame. */
void fpvo_unextract(const uint16_t* img, size_t n, int shift, int big_endian, uint8_t* out) {
  for (size_t i = 0; i < n; i++) {
    uint16_t u = (uint16_t)(img[i] >> shift);
    uint8_t a = (uint8_t)(u & 255), b = (uint8_t)(u >> 8);
    if (big_endian) { uint8_t t = a; a = b; b = t; }
    out[i * 2 + 0] = a;
    out[i * 2 + 1] = b;
  }
}

/* Frame::Uncompress on un-compressed planes (.cc:773-774): inverse CG on the
 * high plane and the preview (.cc:612-641), then delta add on both byte
 * planes (.cc:595-610).  Planes are updated in place. */
void fpvo_unpredict_planes(uint8_t* high, uint8_t* low, uint8_t* preview,
                           const uint8_t* dhigh, const uint8_t* dlow,
                           size_t W, size_t H, uint8_t flags) {
  size_t n = W * H;
  if (flags & FPVO_USE_CG) {
    fpvo_cg_inverse(high, W, n);
    if (preview) fpvo_cg_inverse(preview, W / 4, n / 16);
  }
  if ((flags & FPVO_USE_DELTA) && dhigh) {
    for (size_t i = 0; i < n; i++) high[i] = (uint8_t)(high[i] + dhigh[i]);
    if (low && dlow)
      for (size_t i = 0; i < n; i++) low[i] = (uint8_t)(low[i] + dlow[i]);
  }
}

/* Bulk helper for the exhaustive ClampedGradient check: out[(n<<16)|(w<<8)|nw]. */
void fpvo_cg_table(uint8_t* out) {
  for (uint32_t n = 0; n < 256; n++)
    for (uint32_t w = 0; w < 256; w++)
      for (uint32_t nw = 0; nw < 256; nw++)
        out[(n << 16) | (w << 8) | nw] = fpvo_clamped_gradient((uint8_t)n, (uint8_t)w, (uint8_t)nw);
}

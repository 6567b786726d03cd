// frame_parity.cc -- one deterministic walk through the public surface of fpvc::Frame
// (reference fusion_power_video.h:75-139), printing a digest of the frame after every step.
//
// TEST PROGRAM.  It is compiled twice from this one source:
//   * against the UNMODIFIED reference (fusion_power_video.cc + its header)         -> the expected output
//   * against the mirror header (csrc/host/fusion_power_video.h) and the B200 library (or, for the CPU
//     tests, the oracle-backed stand-in of the C ABI)                                -> must print the same
// tests/test_reference_programs.py compares the two outputs byte for byte.
#include <stdint.h>
#include <stdio.h>

#include <vector>

#include "fusion_power_video.h"

using fpvc::Frame;

static uint64_t Hash(const std::vector<uint8_t>& v) {
  uint64_t h = 1469598103934665603ull;
  for (uint8_t b : v) h = (h ^ b) * 1099511628211ull;
  return h;
}

static void Show(const char* what, Frame& f) {
  printf("%-34s state %2u flags %u ts %lld  high %7zu:%016llx  low %7zu:%016llx  preview %6zu:%016llx\n", what,
         (unsigned)f.state(), (unsigned)f.flags(), (long long)f.timestamp(), f.high().size(),
         (unsigned long long)Hash(f.high()), f.low().size(), (unsigned long long)Hash(f.low()), f.preview().size(),
         (unsigned long long)Hash(f.preview()));
}

// smooth blob + deterministic noise, `bits` significant bits
static std::vector<uint16_t> MakeImage(size_t W, size_t H, int bits, uint32_t seed, int kind) {
  std::vector<uint16_t> img(W * H);
  uint32_t s = seed * 2654435761u + 12345u;
  const uint32_t full = (1u << bits) - 1;
  for (size_t y = 0; y < H; y++)
    for (size_t x = 0; x < W; x++) {
      s = s * 1664525u + 1013904223u;
      uint32_t v;
      if (kind == 0) {  // plasma-like
        const double dx = (double)x / W - 0.5, dy = (double)y / H - 0.45;
        const double g = 0.6 / (1.0 + 40.0 * (dx * dx + dy * dy));
        v = (uint32_t)(full * (0.02 + g)) + ((s >> 20) & (full >> 6));
      } else if (kind == 1) {  // pure noise
        v = (s >> 8) & full;
      } else if (kind == 2) {  // constant high byte, busy low byte
        v = (full & ~0xffu & (0x37u << 8)) | ((s >> 12) & 0xff & full);
      } else {  // all low bytes zero after left alignment
        v = ((s >> 16) & (full >> 8)) << 8;
      }
      img[y * W + x] = (uint16_t)(v > full ? full : v);
    }
  return img;
}

static void Scenario(size_t W, size_t H, int bits, int shift, bool big_endian, int kind) {
  printf("== %zux%zu bits %d shift %d big_endian %d kind %d\n", W, H, bits, shift, (int)big_endian, kind);
  std::vector<uint16_t> d = MakeImage(W, H, bits, 1, 0), a = MakeImage(W, H, bits, 2, kind);
  if (big_endian)
    for (auto* v : {&d, &a})
      for (auto& p : *v) p = (uint16_t)((p << 8) | (p >> 8));
  Frame delta(W, H, d.data(), shift, big_endian, 1000);
  Show("delta ctor", delta);
  Frame f(W, H, a.data(), shift, big_endian, 2000);
  Show("frame ctor", f);
  const std::vector<uint8_t> high0 = f.high(), low0 = f.low();

  Frame p = f;
  p.Predict(delta);
  Show("Predict(delta)", p);
  p.Predict(delta);
  Show("Predict(delta) again", p);

  Frame q = f;
  q.Predict();
  Show("Predict()", q);

  // CompressPredicted into caller buffers, serial and parallel (columnar_batch.cc:65-90)
  for (int parallel = 0; parallel < 2; parallel++) {
    std::vector<uint8_t> bh(p.MaxCompressedPlaneSize()), bl(p.MaxCompressedPlaneSize()), bp(p.MaxCompressedPreviewSize());
    size_t sh = bh.size(), sl = bl.size(), sp = bp.size();
    p.CompressPredicted(&sh, bh.data(), &sl, bl.data(), &sp, bp.data(), parallel != 0);
    bh.resize(sh); bl.resize(sl); bp.resize(sp);
    printf("CompressPredicted(parallel=%d)       high %zu:%016llx low %zu:%016llx preview %zu:%016llx\n", parallel, sh,
           (unsigned long long)Hash(bh), sl, (unsigned long long)Hash(bl), sp, (unsigned long long)Hash(bp));
    if (parallel == 1) {
      // the decoder side of the columnar wrapper (columnar_batch.cc:92-110): planes constructor + Uncompress
      uint8_t state = fpvc::FrameState::COMPRESSED | fpvc::FrameState::DELTA_PREDICTED | fpvc::FrameState::CG_PREDICTED |
                      fpvc::FrameState::PREVIEW_GENERATED;
      Frame r(W, H, p.flags(), state, std::move(bh), std::move(bl), std::move(bp), 3000);
      Show("planes ctor (compressed)", r);
      r.Uncompress(delta);
      Show("  Uncompress(delta)", r);
      printf("  planes restored: high %d low %d\n", (int)(r.high() == high0),
             (int)(r.low() == low0 || (r.flags() & fpvc::FrameFlags::NO_LOW_BYTES)));
    }
  }

  Frame c = f;
  c.Compress(delta);
  Show("Compress(delta)", c);
  std::vector<uint8_t> core, full;
  c.OutputCore(&core);
  c.OutputFull(&full);
  printf("OutputCore %zu:%016llx OutputFull %zu:%016llx\n", core.size(), (unsigned long long)Hash(core), full.size(),
         (unsigned long long)Hash(full));
  c.Uncompress(delta);
  Show("  Uncompress(delta)", c);

  Frame e = f;
  e.Compress();
  Show("Compress()", e);
  e.Uncompress();
  Show("  Uncompress()", e);

  // high-byte-only constructor (fusion_power_video.h:105)
  Frame m(W, H, high0.data(), 4000);
  Show("u8 ctor", m);
  m.Predict();
  Show("  Predict()", m);
  m.Compress();
  Show("  Compress()", m);
  m.Uncompress();
  Show("  Uncompress()", m);

  printf("MaxCompressedPlaneSize %zu %zu MaxCompressedPreviewSize %zu %zu\n", Frame::MaxCompressedPlaneSize(W, H),
         f.MaxCompressedPlaneSize(), Frame::MaxCompressedPreviewSize(W, H), f.MaxCompressedPreviewSize());
}

int main() {
  Frame empty;
  Show("default ctor", empty);
  printf("EMPTY state %u flags %u\n", (unsigned)Frame::EMPTY.state(), (unsigned)Frame::EMPTY.flags());
  Scenario(64, 32, 16, 0, false, 0);
  Scenario(128, 64, 12, 4, false, 0);
  Scenario(128, 64, 12, 4, true, 0);
  Scenario(96, 40, 8, 8, false, 0);
  Scenario(96, 40, 16, 0, true, 1);
  Scenario(64, 64, 16, 0, false, 2);
  Scenario(64, 64, 16, 0, false, 3);
  Scenario(256, 128, 10, 6, false, 1);
  Scenario(1280, 160, 12, 4, false, 0);
  return 0;
}

// ref_harness.cc -- C-ABI wrapper around the UNMODIFIED reference sources.
//
// TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/fpv_oracle.c header).
//
// This translation unit textually includes the reference implementation from
// where it lies (/root/reference/fusion_power_video.cc, passed with -I by
// oracle/Makefile) so that the anonymous-namespace functions (ClampedGradient,
// EstimateEntropy, DecompressImage) are reachable too.  No reference source is
// copied into this repository; the build product goes to oracle/_ref/
// (git-ignored, shipped to the GPU box by gpurun).
//
// Everything below is our own glue: plain pointers in, plain pointers out.
#include "fusion_power_video.cc"  // NOLINT: the reference, compiled in place

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

namespace {

double Now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct Sink {
  std::vector<uint8_t>* out;
};

void AppendCallback(const uint8_t* data, size_t size, void* payload) {
  auto* v = static_cast<std::vector<uint8_t>*>(payload);
  v->insert(v->end(), data, data + size);
}

}  // namespace

extern "C" {

// ---- scalar helpers ---------------------------------------------------------
uint64_t ref_estimate_entropy(const uint64_t* counts) {
  std::vector<size_t> v(counts, counts + 256);
  return fpvc::EstimateEntropy(v);
}

uint8_t ref_clamped_gradient(uint8_t n, uint8_t w, uint8_t nw) {
  return fpvc::ClampedGradient(n, w, nw);
}

// out[(n<<16)|(w<<8)|nw] for all 2^24 inputs.
void ref_cg_table(uint8_t* out) {
  for (uint32_t n = 0; n < 256; n++)
    for (uint32_t w = 0; w < 256; w++)
      for (uint32_t nw = 0; nw < 256; nw++)
        out[(n << 16) | (w << 8) | nw] = fpvc::ClampedGradient((uint8_t)n, (uint8_t)w, (uint8_t)nw);
}

// ---- Frame ctor + Predict ---------------------------------------------------
// delta_img == nullptr: predict without a delta frame (Frame::EMPTY).
// Outputs: high/low W*H bytes (low_size receives low().size(), 0 for shift 8),
// preview (W/4)*(H/4) bytes.  Returns flags.
int ref_predict(size_t W, size_t H, const uint16_t* img, int shift, int big_endian,
                const uint16_t* delta_img, uint8_t* high, uint8_t* low, uint8_t* preview,
                size_t* low_size, size_t* preview_size) {
  fpvc::Frame frame(W, H, img, shift, big_endian != 0);
  if (delta_img) {
    fpvc::Frame delta(W, H, delta_img, shift, big_endian != 0);
    frame.Predict(delta);
  } else {
    frame.Predict();
  }
  memcpy(high, frame.high().data(), frame.high().size());
  if (low && !frame.low().empty()) memcpy(low, frame.low().data(), frame.low().size());
  if (preview) memcpy(preview, frame.preview().data(), frame.preview().size());
  if (low_size) *low_size = frame.low().size();
  if (preview_size) *preview_size = frame.preview().size();
  return frame.flags();
}

// Split only (Frame ctor): returns flags after construction.
int ref_split(size_t W, size_t H, const uint16_t* img, int shift, int big_endian,
              uint8_t* high, uint8_t* low, size_t* low_size) {
  fpvc::Frame frame(W, H, img, shift, big_endian != 0);
  memcpy(high, frame.high().data(), frame.high().size());
  if (low && !frame.low().empty()) memcpy(low, frame.low().data(), frame.low().size());
  if (low_size) *low_size = frame.low().size();
  return frame.flags();
}

// Frame::Uncompress on un-compressed planes: inverse CG (+preview) then delta.
// Planes updated in place.  low_size == 0 means "no low plane".
void ref_unpredict_planes(size_t W, size_t H, uint8_t flags, uint8_t* high, uint8_t* low,
                          size_t low_size, uint8_t* preview, const uint16_t* delta_img,
                          int shift, int big_endian) {
  size_t n = W * H;
  std::vector<uint8_t> h(high, high + n), l, p;
  if (low_size) l.assign(low, low + low_size);
  if (preview) p.assign(preview, preview + (W / 4) * (H / 4));
  uint8_t state = fpvc::FrameState::DELTA_PREDICTED | fpvc::FrameState::CG_PREDICTED |
                  (preview ? fpvc::FrameState::PREVIEW_GENERATED : 0);
  fpvc::Frame frame(W, H, flags, state, std::move(h), std::move(l), std::move(p));
  if (delta_img) {
    fpvc::Frame delta(W, H, delta_img, shift, big_endian != 0);
    frame.Uncompress(delta);
  } else {
    frame.Uncompress();
  }
  memcpy(high, frame.high().data(), n);
  if (low_size) memcpy(low, frame.low().data(), low_size);
  if (preview) memcpy(preview, frame.preview().data(), frame.preview().size());
}

// ---- DecompressImage (brotli + inverse) ------------------------------------
int ref_decompress_image(const uint16_t* delta_frame, const uint8_t* in, size_t size,
                         size_t W, size_t H, uint16_t* img) {
  return fpvc::DecompressImage(delta_frame, in, size, W, H, img) ? 1 : 0;
}

void ref_unextract(const uint16_t* img, size_t W, size_t H, int shift, int big_endian,
                   uint8_t* out) {
  fpvc::UnextractFrame(img, W, H, shift, big_endian != 0, out);
}

// ---- whole-stream encode / decode ------------------------------------------
// Encodes nframes frames (delta frame given separately) with the reference
// Encoder.  Returns the stream size, or the required size if cap is too small.
size_t ref_encode_stream(size_t W, size_t H, int shift, int big_endian, size_t threads,
                         const uint16_t* delta_img, const uint16_t* frames, size_t nframes,
                         uint8_t* out, size_t cap) {
  std::vector<uint8_t> stream;
  {
    fpvc::Encoder encoder(threads, shift, big_endian != 0);
    encoder.Init(delta_img, W, H, AppendCallback, &stream);
    for (size_t i = 0; i < nframes; i++)
      encoder.CompressFrame(frames + i * W * H, AppendCallback, &stream);
    encoder.Finish(AppendCallback, &stream);
  }
  if (stream.size() <= cap && out) memcpy(out, stream.data(), stream.size());
  return stream.size();
}

// Decodes a whole stream with StreamingDecoder fed in `block`-byte pieces.
// Writes up to max_frames frames of W*H uint16.  Returns the number of frames
// decoded, or -1 on a decoder failure.
long ref_decode_stream(const uint8_t* bytes, size_t size, size_t block, uint16_t* frames,
                       size_t max_frames, size_t* W_out, size_t* H_out) {
  struct State {
    uint16_t* frames;
    size_t max_frames;
    size_t count = 0;
    bool failed = false;
    size_t W = 0, H = 0;
  } st;
  st.frames = frames;
  st.max_frames = max_frames;
  fpvc::StreamingDecoder decoder;
  if (block == 0) block = size ? size : 1;
  for (size_t pos = 0; pos < size && !st.failed; pos += block) {
    size_t chunk = (pos + block > size) ? size - pos : block;
    decoder.Decode(bytes + pos, chunk,
                   [&st](bool ok, uint16_t* frame, size_t xs, size_t ys, void*) {
                     if (!ok) { st.failed = true; return; }
                     st.W = xs; st.H = ys;
                     if (st.count < st.max_frames && st.frames)
                       memcpy(st.frames + st.count * xs * ys, frame, xs * ys * 2);
                     st.count++;
                   },
                   nullptr);
  }
  if (W_out) *W_out = st.W;
  if (H_out) *H_out = st.H;
  return st.failed ? -1 : (long)st.count;
}

// RandomAccessDecoder: decode frame `index` and its preview.
int ref_random_access_decode(const uint8_t* bytes, size_t size, size_t index, uint16_t* frame,
                             uint8_t* preview, size_t* numframes) {
  fpvc::RandomAccessDecoder dec;
  if (!dec.Init(bytes, size)) return 0;
  if (numframes) *numframes = dec.numframes();
  if (frame && !dec.DecodeFrame(index, frame)) return 0;
  if (preview && !dec.DecodePreview(index, preview)) return 0;
  return 1;
}

// ---- CPU baseline timing (bench.py --impl reference / cpu_baseline) ----------
// Transform stage only: Frame ctor + Predict(delta) for every frame, frames
// distributed over `threads` std::threads (the reference parallelises over
// frames the same way, .cc:1199-1230).  Returns seconds for one pass.
double ref_time_transform(size_t W, size_t H, int shift, int big_endian,
                          const uint16_t* delta_img, const uint16_t* frames, size_t nframes,
                          size_t threads, uint64_t* checksum) {
  if (threads == 0) threads = 1;
  fpvc::Frame delta(W, H, delta_img, shift, big_endian != 0);
  std::atomic<size_t> next(0);
  std::atomic<uint64_t> sum(0);
  double t0 = Now();
  std::vector<std::thread> pool;
  for (size_t t = 0; t < threads; t++) {
    pool.emplace_back([&]() {
      uint64_t local = 0;
      for (;;) {
        size_t i = next.fetch_add(1);
        if (i >= nframes) break;
        fpvc::Frame frame(W, H, frames + i * W * H, shift, big_endian != 0);
        fpvc::Frame d = delta;  // Predict takes a non-const ref; keep threads independent
        frame.Predict(d);
        local += frame.flags() + frame.high()[frame.high().size() / 2] +
                 (frame.low().empty() ? 0 : frame.low()[frame.low().size() / 2]);
      }
      sum += local;
    });
  }
  for (auto& th : pool) th.join();
  double t1 = Now();
  if (checksum) *checksum = sum.load();
  return t1 - t0;
}

// Same, but the per-thread delta copy is hoisted out of the timed loop (the
// copy above costs 2*W*H bytes of memcpy per frame; the reference Encoder
// shares one delta_frame_ between its workers, .cc:1164).
double ref_time_transform_shared(size_t W, size_t H, int shift, int big_endian,
                                 const uint16_t* delta_img, const uint16_t* frames,
                                 size_t nframes, size_t threads, uint64_t* checksum) {
  if (threads == 0) threads = 1;
  std::vector<fpvc::Frame> deltas;
  for (size_t t = 0; t < threads; t++)
    deltas.emplace_back(W, H, delta_img, shift, big_endian != 0);
  std::atomic<size_t> next(0);
  std::atomic<uint64_t> sum(0);
  double t0 = Now();
  std::vector<std::thread> pool;
  for (size_t t = 0; t < threads; t++) {
    pool.emplace_back([&, t]() {
      uint64_t local = 0;
      for (;;) {
        size_t i = next.fetch_add(1);
        if (i >= nframes) break;
        fpvc::Frame frame(W, H, frames + i * W * H, shift, big_endian != 0);
        frame.Predict(deltas[t]);
        local += frame.flags() + frame.high()[frame.high().size() / 2] +
                 (frame.low().empty() ? 0 : frame.low()[frame.low().size() / 2]);
      }
      sum += local;
    });
  }
  for (auto& th : pool) th.join();
  double t1 = Now();
  if (checksum) *checksum = sum.load();
  return t1 - t0;
}

// Full reference encoder (transform + brotli + framing), benchmark.cc:153-180
// timing window without the per-frame stderr prints.  Returns seconds; the
// stream size goes to *stream_size.
double ref_time_encode(size_t W, size_t H, int shift, int big_endian, size_t threads,
                       const uint16_t* delta_img, const uint16_t* frames, size_t nframes,
                       size_t* stream_size) {
  size_t total = 0;
  auto count = [](const uint8_t*, size_t size, void* payload) {
    *static_cast<size_t*>(payload) += size;
  };
  double t0 = Now();
  {
    fpvc::Encoder encoder(threads, shift, big_endian != 0);
    encoder.Init(delta_img, W, H, count, &total);
    for (size_t i = 0; i < nframes; i++)
      encoder.CompressFrame(frames + i * W * H, count, &total);
    encoder.Finish(count, &total);
  }
  double t1 = Now();
  if (stream_size) *stream_size = total;
  return t1 - t0;
}

// Decode-side inverse transform only (the part of DecompressImage after
// brotli, .cc:326-344, exercised through Frame::Uncompress on un-compressed
// planes + recombination is not exposed, so this times Uncompress).
double ref_time_unpredict(size_t W, size_t H, uint8_t flags, const uint8_t* high,
                          const uint8_t* low, const uint16_t* delta_img, int shift,
                          int big_endian, size_t nframes, size_t threads) {
  if (threads == 0) threads = 1;
  size_t n = W * H;
  std::vector<fpvc::Frame> deltas;
  for (size_t t = 0; t < threads; t++)
    deltas.emplace_back(W, H, delta_img, shift, big_endian != 0);
  std::atomic<size_t> next(0);
  std::atomic<uint64_t> sum(0);
  uint8_t state = fpvc::FrameState::DELTA_PREDICTED | fpvc::FrameState::CG_PREDICTED;
  double t0 = Now();
  std::vector<std::thread> pool;
  for (size_t t = 0; t < threads; t++) {
    pool.emplace_back([&, t]() {
      uint64_t local = 0;
      for (;;) {
        size_t i = next.fetch_add(1);
        if (i >= nframes) break;
        std::vector<uint8_t> h(high + i * n, high + (i + 1) * n);
        std::vector<uint8_t> l(low + i * n, low + (i + 1) * n);
        fpvc::Frame frame(W, H, flags, state, std::move(h), std::move(l), std::vector<uint8_t>());
        frame.Uncompress(deltas[t]);
        local += frame.high()[n / 2];
      }
      sum += local;
    });
  }
  for (auto& th : pool) th.join();
  return Now() - t0;
}

unsigned ref_hardware_threads() { return std::thread::hardware_concurrency(); }

}  // extern "C"

// capi.cc -- plain-C entry points over the fpvc:: classes, for language
// bindings (tests/ and bench.py reach the host layer through these).
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <future>
#include <memory>
#include <thread>
#include <vector>

#include "columnar_batch.h"
#include "fusion_power_video.h"
#include "host_internal.h"

namespace {

void Append(const uint8_t* data, size_t size, void* payload) {
  auto* v = static_cast<std::vector<uint8_t>*>(payload);
  v->insert(v->end(), data, data + size);
}

double Now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

extern "C" {

const char* fpvh_last_error(void) { return fpvc::LastError().c_str(); }

// Encodes `nframes` frames with fpvc::Encoder.  Returns the stream size (also
// when it exceeds `cap`, in which case nothing is copied), 0 on failure.
size_t fpvh_encode_stream(size_t xsize, size_t ysize, int shift, int big_endian, size_t threads, uint32_t batch,
                          int device, const uint16_t* delta, const uint16_t* frames, size_t nframes, uint8_t* out,
                          size_t cap) {
  std::vector<uint8_t> stream;
  fpvc::GpuOptions opt;
  opt.device = device;
  if (batch) opt.batch = batch;
  {
    fpvc::Encoder enc(threads, shift, big_endian != 0, opt);
    enc.Init(delta, xsize, ysize, Append, &stream);
    if (!enc.ok()) return 0;
    for (size_t i = 0; i < nframes; i++) enc.CompressFrame(frames + i * xsize * ysize, Append, &stream);
    enc.Finish(Append, &stream);
    if (!enc.ok()) return 0;
  }
  if (out && stream.size() <= cap) memcpy(out, stream.data(), stream.size());
  return stream.size();
}

// The same with one Encoder over several GPUs (GpuOptions::devices); gpu_entropy as in GpuOptions (-1: environment).
size_t fpvh_encode_stream_multi(size_t xsize, size_t ysize, int shift, int big_endian, size_t threads, uint32_t batch,
                                const int* devices, int ndevices, int gpu_entropy, const uint16_t* delta,
                                const uint16_t* frames, size_t nframes, uint8_t* out, size_t cap, double* seconds) {
  std::vector<uint8_t> stream;
  fpvc::GpuOptions opt;
  if (batch) opt.batch = batch;
  opt.gpu_entropy = gpu_entropy;
  for (int i = 0; i < ndevices; i++) opt.devices.push_back(devices[i]);
  const double t0 = Now();
  {
    fpvc::Encoder enc(threads, shift, big_endian != 0, opt);
    enc.Init(delta, xsize, ysize, Append, &stream);
    if (!enc.ok()) return 0;
    for (size_t i = 0; i < nframes; i++) enc.CompressFrame(frames + i * xsize * ysize, Append, &stream);
    enc.Finish(Append, &stream);
    if (!enc.ok()) return 0;
  }
  if (seconds) *seconds = Now() - t0;
  if (out && stream.size() <= cap) memcpy(out, stream.data(), stream.size());
  return stream.size();
}

// Same work, timed like the reference's benchmark (benchmark.cc:153-180: from
// before Encoder construction to after Finish); compressed bytes are counted,
// not kept.  Returns seconds, < 0 on failure.
double fpvh_time_encode(size_t xsize, size_t ysize, int shift, int big_endian, size_t threads, uint32_t batch,
                        int device, const uint16_t* delta, const uint16_t* frames, size_t nframes,
                        size_t* stream_size) {
  size_t total = 0;
  auto count = [](const uint8_t*, size_t size, void* payload) { *static_cast<size_t*>(payload) += size; };
  fpvc::GpuOptions opt;
  opt.device = device;
  if (batch) opt.batch = batch;
  const double t0 = Now();
  {
    fpvc::Encoder enc(threads, shift, big_endian != 0, opt);
    enc.Init(delta, xsize, ysize, count, &total);
    if (!enc.ok()) return -1.0;
    for (size_t i = 0; i < nframes; i++) enc.CompressFrame(frames + i * xsize * ysize, count, &total);
    enc.Finish(count, &total);
    if (!enc.ok()) return -1.0;
  }
  const double t1 = Now();
  if (stream_size) *stream_size = total;
  return t1 - t0;
}

// Decodes a stream with fpvc::StreamingDecoder fed in `block`-byte pieces
// (0 = all at once).  raw_shift < 0: frames are 16-bit images; otherwise raw
// file bytes (UnextractFrame on the GPU) with that shift / endianness.
// seconds: [0] the whole call, [1] the time until the first frame came out.
// Returns the number of frames, -1 on a decoder failure.
long fpvh_decode_stream(const uint8_t* bytes, size_t size, size_t block, uint32_t batch, int device, int raw_shift,
                        int big_endian, uint16_t* frames, size_t max_frames, size_t* xsize_out, size_t* ysize_out,
                        double* seconds) {
  struct State {
    uint16_t* frames;
    size_t max_frames, count = 0, W = 0, H = 0;
    bool failed = false;
    double first = 0;          // when the first frame came out
  } st;
  st.frames = frames;
  st.max_frames = max_frames;
  fpvc::GpuOptions opt;
  opt.device = device;
  if (batch) opt.batch = batch;
  const double t0 = Now();
  {
    fpvc::StreamingDecoder dec(opt);
    if (raw_shift >= 0) dec.SetRawOutput(raw_shift, big_endian != 0);
    if (block == 0) block = size ? size : 1;
    for (size_t pos = 0; pos < size && !st.failed; pos += block) {
      const size_t n = pos + block > size ? size - pos : block;
      dec.Decode(bytes + pos, n,
                 [&st](bool ok, uint16_t* frame, size_t xs, size_t ys, void*) {
                   if (!ok) { st.failed = true; return; }
                   if (st.count == 0) st.first = Now();
                   st.W = xs;
                   st.H = ys;
                   if (st.frames && st.count < st.max_frames) memcpy(st.frames + st.count * xs * ys, frame, xs * ys * 2);
                   st.count++;
                 },
                 nullptr);
    }
  }
  if (seconds) {
    seconds[0] = Now() - t0;                          // decoder construction (GPU context, pinned staging) .. destruction
    seconds[1] = st.count ? st.first - t0 : 0.0;      // until the first frame came out (set-up + the first batch)
  }
  if (xsize_out) *xsize_out = st.W;
  if (ysize_out) *ysize_out = st.H;
  return st.failed ? -1 : (long)st.count;
}

// RandomAccessDecoder: `count` frames from `first` (frames may be NULL) and the
// preview of frame `first` (preview may be NULL).  Returns 1 on success.
int fpvh_random_access(const uint8_t* bytes, size_t size, uint32_t batch, int device, size_t first, size_t count,
                       uint16_t* frames, uint8_t* preview, size_t* numframes, size_t* xsize_out, size_t* ysize_out) {
  fpvc::GpuOptions opt;
  opt.device = device;
  if (batch) opt.batch = batch;
  fpvc::RandomAccessDecoder dec(opt);
  if (!dec.Init(bytes, size)) return 0;
  if (numframes) *numframes = dec.numframes();
  if (xsize_out) *xsize_out = dec.xsize();
  if (ysize_out) *ysize_out = dec.ysize();
  if (frames && count && !dec.DecodeFrames(first, count, frames)) return 0;
  if (preview && !dec.DecodePreview(first, preview)) return 0;
  return 1;
}

// Pushes `n` frames through a ColumnarBatchEncoder whose batch processor feeds a ColumnarBatchDecoder
// (the wiring of the reference's columnar_batch_decoder_test.cc) and collects the decoded images back to
// back in `out` (bytes_per_image each) with their timestamps.  Returns the number of images, -1 on failure.
long fpvh_columnar_roundtrip(size_t xsize, size_t ysize, int shift, int big_endian, int frames_per_batch, int type,
                             int unshift, const uint16_t* frames, const int64_t* timestamps, size_t n, uint8_t* out,
                             int64_t* out_timestamps, size_t capacity, size_t* bytes_per_image, long* batches,
                             int64_t* encoder_close, int64_t* decoder_close, size_t* compressed_bytes) {
  namespace cb = fpvc::columnarbatch;
  struct State {
    uint8_t* out;
    int64_t* ts;
    size_t capacity, used = 0, count = 0, per = 0, compressed = 0;
    long batches = 0;
    bool failed = false;
  } st;
  st.out = out;
  st.ts = out_timestamps;
  st.capacity = capacity;
  std::unique_ptr<cb::ColumnarBatchEncoder> enc;
  std::unique_ptr<cb::ColumnarBatchDecoder> dec;
  dec.reset(new cb::ColumnarBatchDecoder((cb::Image::Type)type, unshift != 0, [&st](cb::Image img) {
    const size_t bytes = img.xsize() * img.ysize() * (img.type() == cb::Image::Type::FULL ? 2 : 1);   // FULL is always 16 bit
    st.per = bytes;
    if (st.used + bytes <= st.capacity) {
      memcpy(st.out + st.used, img.data8(), bytes);
      st.ts[st.count] = img.timestamp();
      st.used += bytes;
    }
    st.count++;
  }));
  enc.reset(new cb::ColumnarBatchEncoder(xsize, ysize, shift, big_endian != 0, [&](cb::BatchPtr batch) {
    if (!batch) return;
    st.batches++;
    st.compressed += batch->preview_column().size() + batch->high_plane_column().size() + batch->low_plane_column().size();
    std::future<cb::BatchPtr> f = dec->PushBatch(batch);
    if (!f.valid()) { st.failed = true; return; }
    enc->ReturnProcessedBatch(f.get());
  }, frames_per_batch));
  for (size_t i = 0; i < n; i++)
    enc->PushFrame((uint64_t)timestamps[i], const_cast<uint16_t*>(frames + i * xsize * ysize), nullptr).wait();
  const int64_t ec = enc->Close().get();
  const int64_t dc = dec->Close().get();
  if (encoder_close) *encoder_close = ec;
  if (decoder_close) *decoder_close = dc;
  if (bytes_per_image) *bytes_per_image = st.per;
  if (batches) *batches = st.batches;
  if (compressed_bytes) *compressed_bytes = st.compressed;
  return (st.failed || !enc->ok()) ? -1 : (long)st.count;
}

// Parity probe for the columnar path: pushes `n` frames through a ColumnarBatchEncoder and brotli-DECODES the plane
// columns of every Batch it emits, frame by frame: flags[i], high[i][P], low[i][P] (zeros if the frame has no low
// plane) and preview[i][P/16] are what Frame::Predict left in the planes (reference columnar_batch.cc:65-90 compresses
// exactly those).  Returns the number of frames, -1 on failure.
long fpvh_columnar_planes(size_t xsize, size_t ysize, int shift, int big_endian, int frames_per_batch,
                          const uint16_t* frames, const int64_t* timestamps, size_t n, uint8_t* flags, uint8_t* high,
                          uint8_t* low, uint8_t* preview) {
  namespace cb = fpvc::columnarbatch;
  const size_t P = xsize * ysize, PP = (xsize / 4) * (ysize / 4);
  size_t count = 0;
  bool failed = false;
  std::unique_ptr<cb::ColumnarBatchEncoder> enc;
  enc.reset(new cb::ColumnarBatchEncoder(xsize, ysize, shift, big_endian != 0, [&](cb::BatchPtr batch) {
    if (!batch) return;
    for (size_t i = 0; i < batch->length() && count < n; i++, count++) {
      const uint8_t fl = batch->flags()[i];
      flags[count] = fl;
      size_t pos = batch->high_plane_offsets()[i];
      if (!fpvc::internal::BrotliUnplane(batch->high_plane_column().data(), batch->high_plane_offsets()[i + 1], &pos, high + count * P, P))
        failed = true;
      pos = batch->preview_offsets()[i];
      if (!fpvc::internal::BrotliUnplane(batch->preview_column().data(), batch->preview_offsets()[i + 1], &pos, preview + count * PP, PP))
        failed = true;
      memset(low + count * P, 0, P);
      if (!(fl & 4)) {
        pos = batch->low_plane_offsets()[i];
        if (!fpvc::internal::BrotliUnplane(batch->low_plane_column().data(), batch->low_plane_offsets()[i + 1], &pos, low + count * P, P))
          failed = true;
      }
    }
    enc->ReturnProcessedBatch(batch);
  }, frames_per_batch));
  for (size_t i = 0; i < n; i++)
    enc->PushFrame((uint64_t)timestamps[i], const_cast<uint16_t*>(frames + i * P), nullptr).wait();
  enc->Close().get();
  return (failed || !enc->ok()) ? -1 : (long)count;
}

// The host decoders' directory walk (ScanCodedPlane) on one plane stream, for tests: chunk offsets into offs[cap],
// their number into *n, the stream's length into *stream_bytes.  Returns 1 if the stream carries the GPU coder's
// directories, 0 if not (a libbrotli stream, a truncated or damaged one).
int fpvh_scan_coded_plane(const uint8_t* stream, size_t avail, size_t plane_bytes, uint64_t* offs, size_t cap, size_t* n,
                          size_t* stream_bytes) {
  std::vector<uint64_t> v;
  size_t len = 0;
  if (!fpvc::internal::ScanCodedPlane(stream, avail, plane_bytes, &v, &len)) return 0;
  for (size_t i = 0; i < v.size() && i < cap; i++) offs[i] = v[i];
  if (n) *n = v.size();
  if (stream_bytes) *stream_bytes = len;
  return 1;
}

// Real-time ingest (BASELINE configs[4]; the loop of reference encode.cc:63-96 driven by a camera clock instead of
// stdin): frames ARRIVE at `fps` for `seconds`, whether or not the encoder keeps up.  The camera side owns a ring of
// `ring_frames` buffers; the feeder thread hands arrived frames to Encoder::CompressFrame in order, and a frame whose
// turn comes more than ring_frames / fps after its arrival has been overwritten by then: it is DROPPED (counted, not
// encoded).  Latency of an encoded frame = time its compressed bytes reach the callback - its arrival time.
// frames: pool of `npool` distinct frames cycled through (frame 0 is also the delta frame).
// out[0..9] = offered, encoded, dropped, p50 ms, p99 ms, max ms, mean ms, achieved frames/s, stream bytes, wall seconds.
// Returns 0, or -1 if the encoder failed.
int fpvh_ingest(size_t xsize, size_t ysize, int shift, int big_endian, size_t threads, uint32_t batch, int device,
                int gpu_entropy, const uint16_t* frames, size_t npool, double fps, double seconds, size_t ring_frames,
                double* out) {
  const size_t P = xsize * ysize;
  const size_t total = (size_t)(fps * seconds);
  if (total == 0 || npool == 0 || !out) return -1;
  struct State {
    std::vector<double> arrival, latency;
    std::atomic<size_t> bytes{0};
    double t0 = 0;
  } st;
  st.arrival.assign(total, 0.0);
  st.latency.assign(total, -1.0);
  fpvc::GpuOptions opt;
  opt.device = device;
  if (batch) opt.batch = batch;
  opt.gpu_entropy = gpu_entropy;
  size_t dropped = 0, encoded = 0;
  double wall = 0;
  {
    fpvc::Encoder enc(threads, shift, big_endian != 0, opt);
    auto sink = [&st](const uint8_t*, size_t size, void* payload) {
      const size_t idx = (size_t)(uintptr_t)payload;
      st.bytes += size;
      if (idx != (size_t)-1) st.latency[idx] = Now() - st.arrival[idx];
    };
    enc.Init(frames, xsize, ysize, sink, (void*)(uintptr_t)(size_t)-1);
    if (!enc.ok()) return -1;
    // warm the pipeline (contexts, EVERY pinned batch buffer -- the first use of one is a page-locked allocation of
    // tens of milliseconds --, worker threads) before the camera starts
    for (size_t i = 0; i < 8 * (size_t)(batch ? batch : 8); i++)
      enc.CompressFrame(frames + (i % npool) * P, sink, (void*)(uintptr_t)(size_t)-1);
    std::this_thread::sleep_for(std::chrono::milliseconds(300));
    const double period = 1.0 / fps, patience = (double)ring_frames / fps;
    st.t0 = Now();
    for (size_t i = 0; i < total; i++) {
      const double due = st.t0 + (double)i * period;
      double now = Now();
      if (now < due) {
        if (due - now > 200e-6) std::this_thread::sleep_for(std::chrono::duration<double>(due - now - 100e-6));
        while ((now = Now()) < due) {}
      }
      st.arrival[i] = due;
      if (now - due > patience) {   // its ring slot has been overwritten by a newer frame
        dropped++;
        continue;
      }
      enc.CompressFrame(frames + (i % npool) * P, sink, (void*)(uintptr_t)i);
      encoded++;
    }
    enc.Finish(sink, (void*)(uintptr_t)(size_t)-1);
    wall = Now() - st.t0;
    if (!enc.ok()) return -1;
  }
  std::vector<double> lat;
  for (double v : st.latency)
    if (v >= 0) lat.push_back(v * 1e3);
  std::sort(lat.begin(), lat.end());
  double mean = 0;
  for (double v : lat) mean += v;
  out[0] = (double)total;
  out[1] = (double)encoded;
  out[2] = (double)dropped;
  out[3] = lat.empty() ? 0 : lat[lat.size() / 2];
  out[4] = lat.empty() ? 0 : lat[std::min(lat.size() - 1, (size_t)(lat.size() * 0.99))];
  out[5] = lat.empty() ? 0 : lat.back();
  out[6] = lat.empty() ? 0 : mean / (double)lat.size();
  out[7] = wall > 0 ? (double)encoded / wall : 0;
  out[8] = (double)st.bytes.load();
  out[9] = wall;
  return 0;
}

void fpvh_unextract(const uint16_t* img, size_t xsize, size_t ysize, int shift, int big_endian, uint8_t* out) {
  fpvc::UnextractFrame(img, xsize, ysize, shift, big_endian != 0, out);
}

}  // extern "C"

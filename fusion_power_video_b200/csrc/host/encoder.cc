// encoder.cc -- fpvc::Encoder on the GPU transform.
//
// Replaces the reference's worker loop (fusion_power_video.cc:1128-1230): where
// the reference runs Frame ctor + Predict + brotli per frame on a pool thread,
// this encoder
//   1. copies each submitted frame into a pinned batch buffer (CompressFrame),
//   2. hands full batches -- or whatever is there when the GPU is idle -- to one
//      GPU thread, which keeps two fpv_encode_submit slots in flight (H2D of
//      batch k+1 overlaps kernels / D2H of batch k),
//   3. fans the landed planes out to the brotli workers, two tasks per frame
//      (low plane; preview + high plane), the second to finish assembles the chunk,
//   4. emits finished frames in submission order under one mutex, recording the
//      frame offsets for the footer exactly as FinishTask does (.cc:1179-1183).
//
// Several GPUs (GpuOptions::devices): where the reference feeds ONE queue to a pool of worker threads
// (.cc:1076-1084, :1199-1230), this encoder feeds one stream of batches to a pool of GPUs: batch k goes to
// device k mod G, each device has its own context, GPU thread and two submit slots, the delta frame is
// uploaded to the first device and reaches the others by peer copy (fpv_copy_delta_peer).  Emission order
// is the submission order whatever device a batch ran on, so the stream is byte-identical to the
// single-GPU one.
//
// Failures (a CUDA error in submit / wait): the batch is dropped, ok() turns false, every later frame is
// dropped as well (fail-stop: nothing that was not produced by the GPU is ever emitted), Finish still
// writes a footer for the frames that did go out, so the stream stays decodable up to the failure.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <iostream>
#include <map>

#include "fusion_power_video.h"
#include "host_internal.h"

namespace fpvc {

using namespace internal;

namespace {
constexpr int kMaxBatches = 16;
}

struct Encoder::Impl {
  size_t threads = 0;
  int shift = 0;
  bool big_endian = false;
  GpuOptions opt;

  fpv_ctx* ctx = nullptr;             // lanes[0]->ctx: header encoding, the synchronous path, size queries
  std::atomic<bool> ok{false};
  size_t W = 0, H = 0, P = 0, PP = 0;
  bool has_low = true;
  uint32_t B = 1;
  bool gpu_entropy = false;           // planes are entropy-coded and framed on the GPU (fpv_encode_stream_submit)
  bool zero_copy = getenv("FPV_NO_ZERO_COPY") == nullptr;   // frames in page-locked memory are uploaded from where they are
  // cudaPointerGetAttributes costs ~15 us on pageable memory: frames that follow each other in memory (the usual
  // case: one big array, a ring of camera buffers) reuse the last answer, which is asked again every 64 frames.
  // Either answer is safe -- a wrong "pinned" only makes the upload a staged copy, a wrong "pageable" only a host copy.
  const uint16_t* pin_last = nullptr;
  bool pin_answer = false;
  uint32_t pin_age = 0;
  bool is_pinned(const uint16_t* img) {
    const ptrdiff_t d = img - pin_last;
    const ptrdiff_t near_by = (ptrdiff_t)(4 * P);
    if (pin_last && d >= -near_by && d <= near_by && ++pin_age < 64) {
      pin_last = img;
      return pin_answer;
    }
    pin_last = img;
    pin_age = 0;
    pin_answer = fpv_host_is_pinned(img) != 0;
    return pin_answer;
  }
  size_t stream_cap = 0;              // fpv_stream_bound(B)

  // Compressed pieces of one frame, filled by its two brotli tasks.
  struct Pieces {
    std::vector<uint8_t> low, preview_high;   // brotli(low) ; brotli(preview) ++ brotli(high)
    size_t preview_bytes = 0;
    std::atomic<int> parts{0};
  };
  struct Batch {
    Pinned frames, high, low, preview, flags;
    Pinned coded, offs;               // gpu_entropy: container chunks back to back, uint64 offsets [B + 1]
    bool allocated = false;
    uint32_t n = 0;
    uint64_t first_id = 0;
    uint64_t seq = 0;                 // batch number in submission order (device = seq % G; emission order of coded batches)
    std::vector<Callback> callbacks;
    std::vector<void*> payloads;
    std::vector<const uint16_t*> src;   // where frame i is uploaded from: its place in `frames`, or the caller's own
                                        // buffer when that is page-locked memory (no host copy at all)
    std::vector<Pieces> pieces;
    std::atomic<uint32_t> pending{0};
    uint32_t slot = 0;
  };
  Batch batches[kMaxBatches];
  // One per GPU: context, GPU thread, its queue of batches that are ready to be submitted.
  struct Lane {
    int device = 0;
    fpv_ctx* ctx = nullptr;
    std::thread thread;
    std::deque<Batch*> ready;
  };
  std::vector<std::unique_ptr<Lane>> lanes;
  int num_batches = 1;   // one filling, up to two on the GPU, the rest feeding brotli

  std::mutex m;                       // batch lists, ids, stop flag
  std::condition_variable cv_free, cv_ready, cv_drained;
  std::deque<Batch*> free_;
  Batch* filling = nullptr;
  size_t gpu_busy = 0;                // batches handed to a GPU thread, not yet landed
  bool stop = false;
  uint64_t next_id = 0;
  uint64_t next_seq = 0;
  std::unique_ptr<Pool> pool;
  // gpu_entropy: no brotli workers; a few helpers share the copy of each submitted frame into the
  // pinned batch, which is otherwise the pipeline's bottleneck (one core copies ~8 GB/s; measured
  // 12 / 18 / 22 / 24 GB/s of frames with 1 / 3 / 5 / 11 helpers).
  std::unique_ptr<Pool> copy_pool;

  std::mutex out_m;                   // ordered emission
  std::condition_variable cv_emit;    // gpu_entropy with several GPUs: batches take turns by seq
  uint64_t emit_seq = 0;
  struct Done {
    std::vector<uint8_t> bytes;
    Callback callback;
    void* payload;
  };
  std::map<uint64_t, Done> done;
  uint64_t next_emit = 0;
  std::vector<uint64_t> offsets;
  uint64_t bytes_written = 0;
  bool finished = false;

  ~Impl() {
    shutdown();
    for (auto& l : lanes)
      if (l->ctx) fpv_destroy(l->ctx);
  }

  void shutdown() {
    {
      std::lock_guard<std::mutex> l(m);
      stop = true;
    }
    cv_ready.notify_all();
    for (auto& l : lanes)
      if (l->thread.joinable()) l->thread.join();
    pool.reset();  // joins the brotli workers after their queue has drained
    copy_pool.reset();
  }

  // Records the failure (message of the calling thread's last C-ABI error) and stops the encoder.
  bool fail(const std::string& what) {
    ok.store(false);
    cv_emit.notify_all();
    return FPV_FAIL(what + std::string(": ") + fpv_last_error(ctx));
  }

  // Queues a filled batch for its GPU (caller holds m).
  void hand_over_locked(Batch* b) {
    b->seq = next_seq++;
    lanes[b->seq % lanes.size()]->ready.push_back(b);
    gpu_busy++;
  }

  // frame -> chunk bytes (frame := total | 0 | 1+|bp| | pflags | bp | core)
  void build_chunk(uint8_t flags, const uint8_t* high, const uint8_t* low, const uint8_t* preview,
                   std::vector<uint8_t>* scratch, std::vector<uint8_t>* out) {
    out->clear();
    out->reserve(P / 2 + 64);
    out->resize(10);
    BrotliPlane(preview, PP, scratch, out);
    const size_t bp = out->size() - 10;
    AppendCore(flags, high, low, P, scratch, out);
    uint8_t* p = out->data();
    StoreU32((uint32_t)out->size(), p);
    p[4] = kChunkFrame;
    StoreU32((uint32_t)(bp + 1), p + 5);
    p[9] = (uint8_t)((flags & FPV_FLAG_USE_CG) | FPV_FLAG_NO_LOW_BYTES);
  }

  bool alloc_batch(Batch* b) {
    if (b->allocated) return true;
    if (gpu_entropy) {
      if (!b->frames.alloc((size_t)B * P * 2) || !b->flags.alloc(B) || !b->coded.alloc(stream_cap) ||
          !b->offs.alloc(((size_t)B + 1) * sizeof(uint64_t)))
        return fail("pinned allocation");
    } else if (!b->frames.alloc((size_t)B * P * 2) || !b->high.alloc((size_t)B * P) || !b->low.alloc((size_t)B * P) ||
               !b->preview.alloc((size_t)B * (PP ? PP : 1)) || !b->flags.alloc(B))
      return fail("pinned allocation");
    b->pieces = std::vector<Pieces>(B);
    b->allocated = true;
    return true;
  }

  // Called by whoever finished frame `id`; emits every frame that is now in order.
  void deliver(uint64_t id, Done&& d) {
    std::lock_guard<std::mutex> l(out_m);
    done.emplace(id, std::move(d));
    for (auto it = done.find(next_emit); it != done.end(); it = done.find(next_emit)) {
      offsets.push_back(bytes_written);
      bytes_written += it->second.bytes.size();
      it->second.callback(it->second.bytes.data(), it->second.bytes.size(), it->second.payload);
      done.erase(it);
      next_emit++;
    }
  }

  // gpu_entropy: the batch's chunks are complete; they are emitted straight out of the pinned buffer, no
  // copy.  One GPU thread lands its batches in submission order; with several GPUs the batches take turns.
  void emit_coded(Batch* b) {
    const uint64_t* off = b->offs.as<uint64_t>();
    const uint8_t* bytes = b->coded.as<uint8_t>();
    // what came back must be a monotonic offset table inside the buffer before any pointer is formed from it
    bool sane = off[0] == 0 && off[b->n] <= stream_cap;
    for (uint32_t i = 0; sane && i < b->n; i++) sane = off[i + 1] >= off[i] + 11;
    {
      std::unique_lock<std::mutex> l(out_m);
      cv_emit.wait(l, [&] { return emit_seq == b->seq || !ok.load(); });
      if (ok.load() && !sane) {
        l.unlock();
        fail("GPU entropy stage returned an inconsistent offset table");
        l.lock();
      }
      if (ok.load()) {
        for (uint32_t i = 0; i < b->n; i++) {
          const size_t size = (size_t)(off[i + 1] - off[i]);
          offsets.push_back(bytes_written);
          bytes_written += size;
          b->callbacks[i](bytes + off[i], size, b->payloads[i]);
          next_emit++;
        }
      }
      emit_seq = b->seq + 1;
    }
    cv_emit.notify_all();
    recycle(b);
  }

  void recycle(Batch* b) {
    {
      std::lock_guard<std::mutex> l(m);
      b->n = 0;
      b->callbacks.clear();
      b->payloads.clear();
      b->src.clear();
      free_.push_back(b);
    }
    cv_free.notify_all();
    cv_drained.notify_all();
  }

  // The second of a frame's two brotli tasks to finish builds the chunk
  // (frame := total | 0 | 1+|bp| | pflags | bp | flags | low? | high) and hands it on.
  void part_done(Batch* b, uint32_t i) {
    Pieces& pc = b->pieces[i];
    if (pc.parts.fetch_add(1) == 0) return;
    const uint8_t flags = b->flags.as<uint8_t>()[i];
    Done d;
    d.callback = b->callbacks[i];
    d.payload = b->payloads[i];
    d.bytes.resize(10);
    d.bytes.reserve(11 + pc.preview_high.size() + pc.low.size());
    d.bytes.insert(d.bytes.end(), pc.preview_high.begin(), pc.preview_high.begin() + (ptrdiff_t)pc.preview_bytes);
    d.bytes.push_back(flags);
    d.bytes.insert(d.bytes.end(), pc.low.begin(), pc.low.end());
    d.bytes.insert(d.bytes.end(), pc.preview_high.begin() + (ptrdiff_t)pc.preview_bytes, pc.preview_high.end());
    uint8_t* p = d.bytes.data();
    StoreU32((uint32_t)d.bytes.size(), p);
    p[4] = kChunkFrame;
    StoreU32((uint32_t)(pc.preview_bytes + 1), p + 5);
    p[9] = (uint8_t)((flags & FPV_FLAG_USE_CG) | FPV_FLAG_NO_LOW_BYTES);
    deliver(b->first_id + i, std::move(d));
    if (b->pending.fetch_sub(1) == 1) recycle(b);
  }

  void compress_batch(Batch* b) {
    // n is read ONCE: the moment the last task of the batch finishes, the batch is recycled and may already be
    // filling again (b->n counting the next frames) while this loop is still on its way out
    const uint32_t n = b->n;
    b->pending.store(n);
    for (uint32_t i = 0; i < n; i++) {
      Pieces& pc = b->pieces[i];
      pc.parts.store(0);
      pc.low.clear();
      pc.preview_high.clear();
      // the low plane carries most of the entropy and takes longest: queue it first
      pool->run([this, b, i, &pc] {
        thread_local std::vector<uint8_t> scratch;
        if (!(b->flags.as<uint8_t>()[i] & FPV_FLAG_NO_LOW_BYTES))
          BrotliPlane(b->low.as<uint8_t>() + (size_t)i * P, P, &scratch, &pc.low);
        part_done(b, i);
      });
      pool->run([this, b, i, &pc] {
        thread_local std::vector<uint8_t> scratch;
        BrotliPlane(b->preview.as<uint8_t>() + (size_t)i * PP, PP, &scratch, &pc.preview_high);
        pc.preview_bytes = pc.preview_high.size();
        BrotliPlane(b->high.as<uint8_t>() + (size_t)i * P, P, &scratch, &pc.preview_high);
        part_done(b, i);
      });
    }
  }

  void gpu_loop(Lane& lane) {
    fpv_bind_thread(lane.ctx);
    std::deque<Batch*> inflight;
    uint32_t next_slot = 0;
    // A batch whose submit or wait failed -- or that lands after another one failed -- is dropped: its
    // buffers hold stale or partial data and must never reach brotli or a callback.
    auto drop = [&](Batch* b) {
      if (gpu_entropy) {
        std::lock_guard<std::mutex> l(out_m);
        if (emit_seq <= b->seq) emit_seq = b->seq + 1;
      }
      cv_emit.notify_all();
      recycle(b);
    };
    auto land = [&] {
      Batch* b = inflight.front();
      inflight.pop_front();
      const bool good = fpv_wait(lane.ctx, b->slot) == FPV_OK;
      if (!good) fail("fpv_wait");
      {
        std::lock_guard<std::mutex> l(m);
        gpu_busy--;
      }
      if (!good || !ok.load()) drop(b);
      else if (gpu_entropy) emit_coded(b);
      else compress_batch(b);
    };
    for (;;) {
      Batch* b = nullptr;
      {
        std::unique_lock<std::mutex> l(m);
        if (inflight.empty()) cv_ready.wait(l, [&] { return stop || !lane.ready.empty(); });
        if (!lane.ready.empty()) {
          b = lane.ready.front();
          lane.ready.pop_front();
        } else if (inflight.empty()) {
          return;  // stop requested and nothing left
        }
      }
      if (b) {
        if (inflight.size() == 2) land();  // its slot is about to be reused
        b->slot = next_slot;
        next_slot ^= 1u;
        int rc = !ok.load() ? FPV_ERR_INVALID_ARG
                 : gpu_entropy
                     ? fpv_encode_stream_submit_v(lane.ctx, b->slot, b->src.data(), b->n, FPV_ENC_DEFAULT,
                                                  b->flags.as<uint8_t>(), b->offs.as<uint64_t>(),
                                                  b->coded.as<uint8_t>(), stream_cap)
                     : fpv_encode_submit_v(lane.ctx, b->slot, b->src.data(), b->n, FPV_ENC_DEFAULT,
                                           b->flags.as<uint8_t>(), b->high.as<uint8_t>(),
                                           has_low ? b->low.as<uint8_t>() : nullptr, b->preview.as<uint8_t>());
        if (rc != FPV_OK) {
          if (ok.load()) fail("fpv_encode_submit");
          {
            std::lock_guard<std::mutex> l(m);
            gpu_busy--;
          }
          drop(b);
        } else {
          inflight.push_back(b);
        }
      } else {
        land();  // nothing new to submit: collect the oldest batch
      }
    }
  }
};

Encoder::Encoder(size_t num_threads, int shift_to_left_align, bool big_endian)
    : Encoder(num_threads, shift_to_left_align, big_endian, GpuOptions()) {}

Encoder::Encoder(size_t num_threads, int shift_to_left_align, bool big_endian, const GpuOptions& options)
    : impl_(new Impl) {
  impl_->threads = num_threads;
  impl_->shift = shift_to_left_align;
  impl_->big_endian = big_endian;
  impl_->opt = options;
}

Encoder::~Encoder() = default;

bool Encoder::ok() const { return impl_->ok; }

size_t Encoder::MaxQueued() const {
  const size_t t = impl_->threads;
  return t == 0 ? 1 : t + (t + 1) / 2;
}

void Encoder::Init(const uint16_t* delta_frame, size_t xsize, size_t ysize, Callback callback, void* payload) {
  Impl& s = *impl_;
  s.W = xsize;
  s.H = ysize;
  s.P = xsize * ysize;
  s.PP = (xsize / 4) * (ysize / 4);
  s.has_low = s.shift != 8;
  s.B = s.threads == 0 ? 1 : (s.opt.batch ? s.opt.batch : 1);
  // one context per GPU; without worker threads everything runs synchronously on the first device
  std::vector<int> devices = s.opt.devices;
  if (devices.empty()) {
    // unmodified callers (the reference's encode.cc / benchmark.cc) pick GPUs through the environment:
    // FPV_DEVICES=0,2,3 or FPV_GPUS=4 (devices 0..3)
    if (const char* v = getenv("FPV_DEVICES")) {
      for (const char* p = v; *p;) {
        char* end = nullptr;
        const long d = strtol(p, &end, 10);
        if (end == p) break;
        devices.push_back((int)d);
        p = *end == ',' ? end + 1 : end;
      }
    } else if (const char* g = getenv("FPV_GPUS")) {
      for (int d = 0; d < atoi(g); d++) devices.push_back(d);
    }
    if (devices.empty()) devices.push_back(s.opt.device);
  }
  if (s.threads == 0) devices.resize(1);
  for (int dev : devices) {
    std::unique_ptr<Impl::Lane> lane(new Impl::Lane);
    lane->device = dev;
    if (fpv_create(&lane->ctx, dev, (uint32_t)xsize, (uint32_t)ysize, s.shift, s.big_endian ? 1 : 0, s.B) != FPV_OK) {
      FPV_FAIL(std::string("fpv_create (device ") + std::to_string(dev) + "): " + fpv_last_error(nullptr));
      return;
    }
    s.lanes.push_back(std::move(lane));
  }
  s.ctx = s.lanes[0]->ctx;
  fpv_bind_thread(s.ctx);   // pinned allocations below belong to the first device's context, not to device 0's
  s.ok = true;
  {
    const char* env = getenv("FPV_GPU_ENTROPY");
    s.gpu_entropy = s.opt.gpu_entropy >= 0 ? s.opt.gpu_entropy != 0 : (env && atoi(env) != 0);
    s.stream_cap = fpv_stream_bound(s.ctx, s.B);
  }
  // enough batches that the brotli workers always have about two frames each queued
  // behind the (up to) three batches that are filling / on the GPU
  // (GPU entropy stage: one filling, two on the GPU, one being emitted)
  // (several GPUs: two more in flight per additional device)
  const size_t extra = 2 * (s.lanes.size() - 1);
  s.num_batches = s.threads == 0 ? 1
                  : s.gpu_entropy ? (int)std::min<size_t>(kMaxBatches, 4 + extra)
                                  : (int)std::min<size_t>(kMaxBatches, 3 + extra + (2 * s.threads + s.B - 1) / s.B);
  for (int i = 0; i < s.num_batches; i++) s.free_.push_back(&s.batches[i]);
  // Batch buffers are allocated on first use; a page-locked allocation takes tens of milliseconds, which a camera
  // feeding CompressFrame at a few thousand frames per second sees as dropped frames (the paced-ingest harness found
  // 20-130 drops at 2000 fps, all in the moments a batch buffer was used for the first time): real-time callers set
  // FPV_EAGER_BATCHES=1 (every buffer allocated here, a longer Init) or push 8 batches of frames before the camera
  // starts, as the harness does.
  const int eager = getenv("FPV_EAGER_BATCHES") ? s.num_batches : 1;
  for (int i = 0; i < eager; i++)
    if (!s.alloc_batch(&s.batches[i])) return;
  if (fpv_set_delta_raw(s.ctx, delta_frame) != FPV_OK) {
    s.fail("fpv_set_delta_raw");
    return;
  }
  // the other GPUs get the split delta planes by peer copy: the only inter-GPU traffic of a stream
  for (size_t k = 1; k < s.lanes.size(); k++)
    if (fpv_copy_delta_peer(s.lanes[k]->ctx, s.ctx) != FPV_OK) {
      s.fail("fpv_copy_delta_peer");
      return;
    }
  // Header: the delta frame itself, predicted without a delta frame
  // (Frame df = delta_frame_; df.Compress(), reference .cc:1099-1101).
  // (once per stream, pageable buffers; the delta chunk's planes always go through libbrotli)
  std::vector<uint8_t> dhigh(s.P), dlow(s.has_low ? s.P : 0), dprev(s.PP ? s.PP : 1);
  uint8_t dflags = 0;
  if (fpv_encode(s.ctx, delta_frame, 1, FPV_ENC_NO_DELTA, &dflags, dhigh.data(), s.has_low ? dlow.data() : nullptr,
                 dprev.data()) != FPV_OK) {
    s.fail("fpv_encode (delta frame)");
    return;
  }
  std::vector<uint8_t> header, scratch;
  AppendU32((uint32_t)xsize, &header);
  AppendU32((uint32_t)ysize, &header);
  AppendU32(0, &header);
  header.push_back(kChunkDelta);
  AppendCore(dflags, dhigh.data(), s.has_low ? dlow.data() : nullptr, s.P, &scratch, &header);
  StoreU32((uint32_t)(header.size() - 8), header.data() + 8);
  s.bytes_written = header.size();
  if (s.threads > 0) {
    if (!s.gpu_entropy) s.pool.reset(new Pool(s.threads));
    else if (s.threads > 1 && s.P * 2 >= (1u << 19)) {
      size_t helpers = 5;
      if (const char* v = getenv("FPV_COPY_THREADS")) helpers = (size_t)std::max(1, atoi(v));
      s.copy_pool.reset(new Pool(std::min<size_t>(s.threads - 1, helpers)));
    }
    for (auto& lane : s.lanes) {
      Impl::Lane* lp = lane.get();
      lp->thread = std::thread([&s, lp] { s.gpu_loop(*lp); });
    }
  }
  callback(header.data(), header.size(), payload);
}

void Encoder::CompressFrame(const uint16_t* img, Callback callback, void* payload) {
  Impl& s = *impl_;
  if (!s.ok) return;
  if (s.threads == 0) {
    // synchronous: transform, brotli and the callback all happen here
    Impl::Batch& b = s.batches[0];
    memcpy(b.frames.as<uint16_t>(), img, s.P * 2);
    if (s.gpu_entropy) {
      if (fpv_encode_stream_submit(s.ctx, 0, b.frames.as<uint16_t>(), 1, FPV_ENC_DEFAULT, b.flags.as<uint8_t>(),
                                   b.offs.as<uint64_t>(), b.coded.as<uint8_t>(), s.stream_cap) != FPV_OK ||
          fpv_wait(s.ctx, 0) != FPV_OK) {
        s.fail("fpv_encode_stream_submit");
        return;
      }
      std::lock_guard<std::mutex> l(s.out_m);
      const size_t size = (size_t)b.offs.as<uint64_t>()[1];
      s.offsets.push_back(s.bytes_written);
      s.bytes_written += size;
      s.next_id++;
      s.next_emit++;
      callback(b.coded.as<uint8_t>(), size, payload);
      return;
    }
    if (fpv_encode(s.ctx, b.frames.as<uint16_t>(), 1, FPV_ENC_DEFAULT, b.flags.as<uint8_t>(),
                   b.high.as<uint8_t>(), s.has_low ? b.low.as<uint8_t>() : nullptr,
                   b.preview.as<uint8_t>()) != FPV_OK) {
      s.fail("fpv_encode");
      return;
    }
    static thread_local std::vector<uint8_t> scratch;
    Impl::Done d;
    d.callback = callback;
    d.payload = payload;
    s.build_chunk(b.flags.as<uint8_t>()[0], b.high.as<uint8_t>(), s.has_low ? b.low.as<uint8_t>() : nullptr,
                  b.preview.as<uint8_t>(), &scratch, &d.bytes);
    s.deliver(s.next_id++, std::move(d));
    return;
  }
  Impl::Batch* b;
  {
    std::unique_lock<std::mutex> l(s.m);
    if (!s.filling) {
      s.cv_free.wait(l, [&] { return !s.free_.empty(); });
      s.filling = s.free_.front();
      s.free_.pop_front();
      s.filling->first_id = s.next_id;
    }
    b = s.filling;
    s.next_id++;
  }
  if (!b->allocated) {
    fpv_bind_thread(s.ctx);
    if (!s.alloc_batch(b)) return;
  }
  // only this (the submitting) thread touches a filling batch
  uint16_t* slot_in_batch = b->frames.as<uint16_t>() + (size_t)b->n * s.P;
  if (s.zero_copy && s.is_pinned(img)) {
    // The caller's buffer is page-locked memory: the GPU reads it where it is.  The reference's contract covers this
    // ("img ... must exist until the callback for this frame is called", fusion_power_video.h:197-199).
    b->src.push_back(img);
  } else if (s.copy_pool) {
    b->src.push_back(slot_in_batch);
    const size_t parts = s.copy_pool->size() + 1, bytes = s.P * 2, step = ((bytes + parts - 1) / parts + 4095) & ~(size_t)4095;
    uint8_t* dst = reinterpret_cast<uint8_t*>(b->frames.as<uint16_t>() + (size_t)b->n * s.P);
    const uint8_t* src = reinterpret_cast<const uint8_t*>(img);
    std::atomic<size_t> left{parts - 1};
    auto copy_part = [&](size_t k) {
      const size_t o = k * step;
      if (o < bytes) memcpy(dst + o, src + o, std::min(step, bytes - o));
    };
    for (size_t k = 1; k < parts; k++)
      s.copy_pool->run([&copy_part, &left, k] {
        copy_part(k);
        left.fetch_sub(1, std::memory_order_release);
      });
    copy_part(0);                                   // the caller copies its share too
    while (left.load(std::memory_order_acquire)) std::this_thread::yield();
  } else {
    b->src.push_back(slot_in_batch);
    memcpy(slot_in_batch, img, s.P * 2);
  }
  b->callbacks.push_back(callback);
  b->payloads.push_back(payload);
  b->n++;
  bool hand_over = b->n == s.B;
  {
    std::lock_guard<std::mutex> l(s.m);
    if (!hand_over && s.gpu_busy == 0) hand_over = true;  // idle pipeline: go now
    if (hand_over) {
      s.hand_over_locked(b);
      s.filling = nullptr;
    }
  }
  if (hand_over) s.cv_ready.notify_all();
}

void Encoder::Finish(Callback callback, void* payload) {
  Impl& s = *impl_;
  if (s.finished) return;
  s.finished = true;
  if (!s.lanes.empty() && s.threads > 0 && (s.lanes[0]->thread.joinable())) {
    {
      std::unique_lock<std::mutex> l(s.m);
      if (s.filling && s.filling->n > 0) {
        s.hand_over_locked(s.filling);
        s.filling = nullptr;
      }
    }
    s.cv_ready.notify_all();
    {
      // every batch back on the free list <=> every frame emitted
      std::unique_lock<std::mutex> l(s.m);
      s.cv_drained.wait(l, [&] { return s.free_.size() + (s.filling ? 1 : 0) == (size_t)s.num_batches; });
    }
    s.shutdown();
  }
  std::vector<uint8_t> footer(5 + 8 * s.offsets.size() + 8);
  StoreU32((uint32_t)footer.size(), footer.data());
  footer[4] = kChunkIndex;
  for (size_t i = 0; i < s.offsets.size(); i++) StoreU64(s.offsets[i], footer.data() + 5 + 8 * i);
  StoreU64(s.offsets.size(), footer.data() + 5 + 8 * s.offsets.size());
  callback(footer.data(), footer.size(), payload);
  if (!s.ok.load()) FPV_FAIL("Encoder: the stream is incomplete, a GPU call failed (see above); the footer indexes the frames that were written");
}

}  // namespace fpvc

// host_common.cc -- error reporting, brotli glue, worker pool, UnextractFrame.
#include <brotli/decode.h>
#include <brotli/encode.h>
#include <stdlib.h>
#include <string.h>
#include <sys/resource.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <iostream>
#include <map>

#include "fusion_power_video.h"
#include "host_internal.h"

namespace fpvc {

namespace {
thread_local std::string g_last_error;
}

const std::string& LastError() { return g_last_error; }

namespace internal {

bool Fail(const char* file, int line, const std::string& message) {
  g_last_error = message;
  std::cerr << "failure at: " << file << ":" << line;
  if (!message.empty()) std::cerr << ": " << message;
  std::cerr << std::endl;
  return false;
}

bool BrotliPlane(const uint8_t* plane, size_t size, std::vector<uint8_t>* scratch, std::vector<uint8_t>* out) {
  // BrotliEncoderMaxCompressedSize is what the reference sizes its buffer with
  // (.cc:355-357); with it the one-shot call cannot fail for lack of room.
  size_t cap = BrotliEncoderMaxCompressedSize(size);
  if (cap == 0) cap = size + 1024;
  if (scratch->size() < cap) scratch->resize(cap);
  size_t n = cap;
  if (!BrotliEncoderCompress(1, BROTLI_DEFAULT_WINDOW, BROTLI_DEFAULT_MODE, size, plane, &n, scratch->data()))
    return FPV_FAIL("brotli compression failed");
  out->insert(out->end(), scratch->data(), scratch->data() + n);
  return true;
}

bool BrotliUnplane(const uint8_t* in, size_t size, size_t* pos, uint8_t* out, size_t expect) {
  if (*pos > size) return FPV_FAIL("out of bounds");
  BrotliDecoderState* st = BrotliDecoderCreateInstance(nullptr, nullptr, nullptr);
  if (!st) return FPV_FAIL("couldn't init brotli decoder");
  size_t avail_in = size - *pos;
  const uint8_t* next_in = in + *pos;
  size_t avail_out = expect;
  uint8_t* next_out = out;
  uint8_t spill[256];
  bool overflow = false;
  BrotliDecoderResult r;
  for (;;) {
    r = BrotliDecoderDecompressStream(st, &avail_in, &next_in, &avail_out, &next_out, nullptr);
    if (r != BROTLI_DECODER_RESULT_NEEDS_MORE_OUTPUT) break;
    // more data than the plane holds: keep consuming to find the stream end, then report the mismatch
    overflow = true;
    avail_out = sizeof spill;
    next_out = spill;
  }
  BrotliDecoderDestroyInstance(st);
  *pos = size - avail_in;
  if (r != BROTLI_DECODER_RESULT_SUCCESS) return FPV_FAIL("brotli decompression failed");
  if (overflow || avail_out != 0) return FPV_FAIL("wrong decompressed plane size");
  return true;
}

void AppendCore(uint8_t flags, const uint8_t* high, const uint8_t* low, size_t plane_bytes,
                std::vector<uint8_t>* scratch, std::vector<uint8_t>* out) {
  out->push_back(flags);
  if (!(flags & FPV_FLAG_NO_LOW_BYTES)) BrotliPlane(low, plane_bytes, scratch, out);
  BrotliPlane(high, plane_bytes, scratch, out);
}

bool ScanCodedPlane(const uint8_t* in, size_t avail, size_t plane_bytes, std::vector<uint64_t>* chunk_offsets,
                    size_t* stream_bytes) {
  constexpr size_t kChunk = 65536, kDirBlock = 236;
  const size_t chunks = (plane_bytes + kChunk - 1) / kChunk;
  const size_t first = chunk_offsets->size();
  size_t pos = 0;
  for (size_t k = 0; k < chunks; k++) {
    // 2 header bytes of the metadata meta-block, then 'F' 'D' version 1 | kind | chunk bytes u24 ...
    if (pos + kDirBlock > avail || in[pos + 2] != 0x46 || in[pos + 3] != 0x44 || in[pos + 4] != 1) {
      chunk_offsets->resize(first);
      return false;
    }
    const size_t cb = (size_t)in[pos + 6] | ((size_t)in[pos + 7] << 8) | ((size_t)in[pos + 8] << 16);
    if (cb < kDirBlock || cb > avail - pos) {
      chunk_offsets->resize(first);
      return false;
    }
    chunk_offsets->push_back(pos);
    pos += cb;
  }
  if (chunks == 0 || pos >= avail || in[pos] != 0x03) {     // ISLAST, ISLASTEMPTY ends the stream
    chunk_offsets->resize(first);
    return false;
  }
  *stream_bytes = pos + 1;
  return true;
}

bool ParseCore(const uint8_t* in, size_t size, size_t plane_bytes, uint8_t* flags, uint8_t* high, uint8_t* low) {
  if (size == 0) return FPV_FAIL("out of bounds");
  size_t pos = 0;
  *flags = in[pos++];
  if (*flags & FPV_FLAG_NO_LOW_BYTES) {
    if (low) memset(low, 0, plane_bytes);
  } else {
    if (!low) return FPV_FAIL("low plane buffer missing");
    if (!BrotliUnplane(in, size, &pos, low, plane_bytes)) return false;
  }
  return BrotliUnplane(in, size, &pos, high, plane_bytes);
}

// ---- pinned block cache -------------------------------------------------------------------
namespace {
std::mutex g_pin_m;
std::multimap<size_t, void*> g_pin_cache;
size_t g_pin_cached_bytes = 0;
// Bound of the cache: FPV_PIN_CACHE_MB (default 2048; 0 disables caching).  Cached blocks are returned to the driver
// by fpvc::TrimPinnedCache() or when the limit is exceeded, never silently kept beyond it.
size_t PinCacheLimit() {
  static const size_t limit = [] {
    const char* v = getenv("FPV_PIN_CACHE_MB");
    return (size_t)(v ? strtoull(v, nullptr, 10) : 2048) << 20;
  }();
  return limit;
}
}  // namespace

void* PinnedAcquire(size_t bytes) {
  if (bytes == 0) bytes = 1;
  {
    std::lock_guard<std::mutex> l(g_pin_m);
    auto it = g_pin_cache.find(bytes);
    if (it != g_pin_cache.end()) {
      void* p = it->second;
      g_pin_cache.erase(it);
      g_pin_cached_bytes -= bytes;
      return p;
    }
  }
  return fpv_host_alloc(bytes);
}

void PinnedRelease(void* p, size_t bytes) {
  if (!p) return;
  if (bytes == 0) bytes = 1;
  {
    std::lock_guard<std::mutex> l(g_pin_m);
    if (g_pin_cached_bytes + bytes <= PinCacheLimit()) {
      g_pin_cache.emplace(bytes, p);
      g_pin_cached_bytes += bytes;
      return;
    }
  }
  fpv_host_free(p);
}

size_t PinnedTrim() {
  std::multimap<size_t, void*> drop;
  size_t bytes = 0;
  {
    std::lock_guard<std::mutex> l(g_pin_m);
    drop.swap(g_pin_cache);
    bytes = g_pin_cached_bytes;
    g_pin_cached_bytes = 0;
  }
  for (auto& e : drop) fpv_host_free(e.second);
  return bytes;
}

// ---- Pool ----------------------------------------------------------------------------
Pool::Pool(size_t threads) {
  for (size_t i = 0; i < threads; i++) threads_.emplace_back([this] { loop(); });
}

Pool::~Pool() {
  {
    std::lock_guard<std::mutex> l(m_);
    stop_ = true;
  }
  cv_work_.notify_all();
  for (auto& t : threads_) t.join();
}

void Pool::run(std::function<void()> task) {
  if (threads_.empty()) {
    task();
    return;
  }
  {
    std::lock_guard<std::mutex> l(m_);
    q_.push_back(std::move(task));
  }
  cv_work_.notify_one();
}

void Pool::wait_idle() {
  std::unique_lock<std::mutex> l(m_);
  cv_idle_.wait(l, [this] { return q_.empty() && busy_ == 0; });
}

void Pool::loop() {
  // The workers only run brotli.  The threads that feed them (the caller copying
  // frames into pinned memory, the GPU thread) need a core the moment they become
  // runnable or the whole pipeline starves, so the workers step back a little.
  setpriority(PRIO_PROCESS, (id_t)syscall(SYS_gettid), 5);
  for (;;) {
    std::function<void()> task;
    {
      std::unique_lock<std::mutex> l(m_);
      cv_work_.wait(l, [this] { return stop_ || !q_.empty(); });
      if (q_.empty()) return;  // stop_ and drained
      task = std::move(q_.front());
      q_.pop_front();
      busy_++;
    }
    task();
    {
      std::lock_guard<std::mutex> l(m_);
      busy_--;
      if (q_.empty() && busy_ == 0) cv_idle_.notify_all();
    }
  }
}

void ParallelFor(Pool* pool, size_t n, const std::function<void(size_t)>& body) {
  if (!pool || pool->size() == 0 || n <= 1) {
    for (size_t i = 0; i < n; i++) body(i);
    return;
  }
  std::mutex m;
  std::condition_variable cv;
  size_t left = n;
  for (size_t i = 0; i < n; i++)
    pool->run([&, i] {
      body(i);
      std::lock_guard<std::mutex> l(m);
      if (--left == 0) cv.notify_all();
    });
  std::unique_lock<std::mutex> l(m);
  cv.wait(l, [&] { return left == 0; });
}

}  // namespace internal

size_t TrimPinnedCache() { return internal::PinnedTrim(); }

void UnextractFrame(const uint16_t* img, size_t xsize, size_t ysize, int shift, bool big_endian, uint8_t* out) {
  const size_t n = xsize * ysize;
  const int lo = big_endian ? 1 : 0, hi = big_endian ? 0 : 1;
  for (size_t i = 0; i < n; i++) {
    const uint16_t v = (uint16_t)(img[i] >> shift);
    out[2 * i + lo] = (uint8_t)(v & 0xff);
    out[2 * i + hi] = (uint8_t)(v >> 8);
  }
}

}  // namespace fpvc

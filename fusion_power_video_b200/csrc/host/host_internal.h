// host_internal.h -- pieces shared by the encoder and the decoders of the host
// layer: little-endian field access, the stream grammar, brotli glue, pinned
// buffers, a small worker pool.  Not installed.
//
// Stream grammar (all little-endian; reference fusion_power_video.cc:56-102,
// :820-846, :1086-1106, :1185-1197):
//   file       := u32 xsize | u32 ysize | deltachunk | frame* | footer
//   deltachunk := u32 (5 + |core|) | u8 1 | core
//   frame      := u32 total | u8 0 | u32 (1 + |bp|) | u8 pflags | bp | core
//                 total = 10 + |bp| + |core|, pflags = (flags & USE_CG) | NO_LOW_BYTES,
//                 bp = brotli(preview plane)
//   core       := u8 flags | brotli(low plane) (absent iff flags & NO_LOW_BYTES) | brotli(high plane)
//   footer     := u32 (13 + 8 N) | u8 2 | u64 offset[N] | u64 N
//                 offset = absolute position of each frame's u32 total
#pragma once

#include <stddef.h>
#include <stdint.h>

#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/fpv_b200.h"

namespace fpvc {
namespace internal {

constexpr uint8_t kChunkFrame = 0, kChunkDelta = 1, kChunkIndex = 2;
constexpr size_t kMaxPixels = 1000000000;  // reference .cc:164

inline uint32_t LoadU32(const uint8_t* p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
inline uint64_t LoadU64(const uint8_t* p) { return (uint64_t)LoadU32(p) | ((uint64_t)LoadU32(p + 4) << 32); }
inline void StoreU32(uint32_t v, uint8_t* p) {
  p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}
inline void StoreU64(uint64_t v, uint8_t* p) { StoreU32((uint32_t)v, p); StoreU32((uint32_t)(v >> 32), p + 4); }
inline void AppendU32(uint32_t v, std::vector<uint8_t>* out) {
  size_t n = out->size();
  out->resize(n + 4);
  StoreU32(v, out->data() + n);
}

// Records the message for LastError() and prints "failure at: ..." to stderr
// like the reference's FAILURE macro (.cc:171-181).  Always returns false.
bool Fail(const char* file, int line, const std::string& message);
#define FPV_FAIL(msg) ::fpvc::internal::Fail(__FILE__, __LINE__, (msg))

// ---- brotli ---------------------------------------------------------------------
// One-shot quality-1 / lgwin-22 / generic-mode compression of a plane, appended
// to *out (reference .cc:653-654).  The scratch vector is reused between calls.
bool BrotliPlane(const uint8_t* plane, size_t size, std::vector<uint8_t>* scratch, std::vector<uint8_t>* out);
// Decodes ONE brotli stream starting at in[*pos] into exactly `expect` bytes at
// `out`; *pos is moved to the end of that stream (streams are concatenated,
// reference .cc:186-214).  Fails if the stream is malformed or its size differs.
bool BrotliUnplane(const uint8_t* in, size_t size, size_t* pos, uint8_t* out, size_t expect);

// Walks the directory meta-blocks of ONE plane stream written by the GPU entropy coder (csrc/fpv_entropy.cu: every
// 64 KiB chunk starts with a metadata meta-block holding, among other things, the chunk's size).  On success the
// offsets of the stream's chunks (relative to `in`) are appended to *chunk_offsets and *stream_bytes is the
// length of the whole stream, final 0x03 included.  false: not such a stream (libbrotli's output carries no
// directory) or a truncated one -- the caller then takes the brotli path.
bool ScanCodedPlane(const uint8_t* in, size_t avail, size_t plane_bytes, std::vector<uint64_t>* chunk_offsets,
                    size_t* stream_bytes);

// core := flags | brotli(low)? | brotli(high)
void AppendCore(uint8_t flags, const uint8_t* high, const uint8_t* low, size_t plane_bytes,
                std::vector<uint8_t>* scratch, std::vector<uint8_t>* out);
// Parses a core chunk into its two planes (low is zero-filled if absent).
bool ParseCore(const uint8_t* in, size_t size, size_t plane_bytes, uint8_t* flags, uint8_t* high, uint8_t* low);

// ---- pinned memory -----------------------------------------------------------------
// Page-locking memory costs on the order of a millisecond per few MB, so freed
// blocks go to a small process-wide cache and are handed out again by size.
void* PinnedAcquire(size_t bytes);
void PinnedRelease(void* p, size_t bytes);
size_t PinnedTrim();   // frees every cached block; returns the bytes released

class Pinned {
 public:
  Pinned() = default;
  ~Pinned() { reset(); }
  Pinned(const Pinned&) = delete;
  Pinned& operator=(const Pinned&) = delete;
  bool alloc(size_t bytes) {
    reset();
    p_ = PinnedAcquire(bytes);
    n_ = p_ ? bytes : 0;
    return p_ != nullptr;
  }
  void reset() {
    if (p_) PinnedRelease(p_, n_);
    p_ = nullptr;
    n_ = 0;
  }
  template <typename T> T* as() const { return static_cast<T*>(p_); }
  size_t bytes() const { return n_; }

 private:
  void* p_ = nullptr;
  size_t n_ = 0;
};

// ---- worker pool ---------------------------------------------------------------------
// Plain FIFO pool.  run() with zero threads executes the task inline.
class Pool {
 public:
  explicit Pool(size_t threads);
  ~Pool();
  void run(std::function<void()> task);
  void wait_idle();
  size_t size() const { return threads_.size(); }

 private:
  void loop();
  std::vector<std::thread> threads_;
  std::mutex m_;
  std::condition_variable cv_work_, cv_idle_;
  std::deque<std::function<void()>> q_;
  size_t busy_ = 0;
  bool stop_ = false;
};

// parallel for over [0, n) on `pool` (inline if the pool has no threads); returns when all are done
void ParallelFor(Pool* pool, size_t n, const std::function<void(size_t)>& body);

}  // namespace internal
}  // namespace fpvc

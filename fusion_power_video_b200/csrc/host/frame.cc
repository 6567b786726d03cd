// frame.cc -- fpvc::Frame on the GPU transform (see the class comment in fusion_power_video.h).
//
// Reference: fusion_power_video.cc:353-846 (class Frame).  State / flag bookkeeping follows the
// reference line by line; the pixel arithmetic is three C-ABI calls:
//   Frame(u16) ctor -> fpv_split            (.cc:370-451)
//   Predict         -> fpv_encode           (.cc:777-785 -> :491-593)
//   Uncompress      -> fpv_unpredict_planes (.cc:773-774 -> :612-641, :595-610)
// and brotli stays on the host (.cc:643-728).
#include <brotli/decode.h>
#include <brotli/encode.h>
#include <string.h>

#include <atomic>
#include <future>
#include <map>
#include <mutex>
#include <tuple>

#include "fusion_power_video.h"
#include "host_internal.h"

namespace fpvc {

using namespace internal;

Frame Frame::EMPTY(0, 0);

namespace {

std::atomic<uint64_t> g_generation{1};
std::atomic<int> g_frame_device{0};

// One brotli stream -> as many bytes as it holds (reference BrotliDecompress, .cc:186-214: the plane size
// is whatever comes out).
bool BrotliToVector(const std::vector<uint8_t>& in, std::vector<uint8_t>* out) {
  BrotliDecoderState* st = BrotliDecoderCreateInstance(nullptr, nullptr, nullptr);
  if (!st) return FPV_FAIL("couldn't init brotli decoder");
  size_t avail_in = in.size();
  const uint8_t* next_in = in.data();
  out->clear();
  std::vector<uint8_t> block(1 << 18);
  BrotliDecoderResult r;
  do {
    size_t avail_out = block.size();
    uint8_t* next_out = block.data();
    r = BrotliDecoderDecompressStream(st, &avail_in, &next_in, &avail_out, &next_out, nullptr);
    out->insert(out->end(), block.data(), next_out);
  } while (r == BROTLI_DECODER_RESULT_NEEDS_MORE_OUTPUT);
  BrotliDecoderDestroyInstance(st);
  if (r != BROTLI_DECODER_RESULT_SUCCESS) return FPV_FAIL("brotli decompression failed");
  return true;
}

}  // namespace

// GPU contexts for Frame methods: a small pool keyed by geometry, one context per concurrent caller (the
// reference's wrappers run Frame methods from std::async threads).  A context remembers which frame's
// planes it holds as its delta frame (generation number), so a stream of Predict(delta) calls uploads the
// delta frame once.  Contexts live until process exit (never destroyed from a static destructor: the CUDA
// runtime may already be gone by then).
struct FrameGpu {
  struct Entry {
    fpv_ctx* ctx = nullptr;
    uint64_t delta_generation = 0;
  };
  typedef std::tuple<size_t, size_t, int, bool, int> Key;   // xsize, ysize, shift, big_endian, device
  std::mutex m;
  std::multimap<Key, Entry> idle;

  static FrameGpu& Get() {
    static FrameGpu* g = new FrameGpu;
    return *g;
  }

  // A context for this geometry, preferring one that already holds `want_generation` as its delta frame.
  bool Acquire(const Key& key, uint64_t want_generation, Entry* out) {
    {
      std::lock_guard<std::mutex> l(m);
      auto range = idle.equal_range(key);
      auto pick = range.first;
      for (auto it = range.first; it != range.second; ++it)
        if (it->second.delta_generation == want_generation) { pick = it; break; }
      if (pick != range.second) {
        *out = pick->second;
        idle.erase(pick);
        return true;
      }
    }
    Entry e;
    if (fpv_create(&e.ctx, std::get<4>(key), (uint32_t)std::get<0>(key), (uint32_t)std::get<1>(key), std::get<2>(key),
                   std::get<3>(key) ? 1 : 0, 1) != FPV_OK)
      return FPV_FAIL(std::string("fpv_create: ") + fpv_last_error(nullptr));
    *out = e;
    return true;
  }
  void Release(const Key& key, const Entry& e) {
    std::lock_guard<std::mutex> l(m);
    idle.emplace(key, e);
  }

  // Makes `delta`'s planes the context's delta frame (image form, reference .cc:337-338), or clears it.
  static bool SetDelta(Entry* e, Frame* delta, size_t size) {
    if (!delta) {
      if (e->delta_generation != 0 && fpv_set_delta_image(e->ctx, nullptr) != FPV_OK) return false;
      e->delta_generation = 0;
      return true;
    }
    if (delta->high_.size() != size) return FPV_FAIL("delta frame planes do not match the frame size");
    if (e->delta_generation == delta->generation_ && delta->generation_ != 0) return true;
    std::vector<uint16_t> image(size);
    const uint8_t* h = delta->high_.data();
    if (delta->low_.size() == size) {
      const uint8_t* l = delta->low_.data();
      for (size_t i = 0; i < size; i++) image[i] = (uint16_t)((h[i] << 8) | l[i]);
    } else {
      for (size_t i = 0; i < size; i++) image[i] = (uint16_t)(h[i] << 8);
    }
    if (fpv_set_delta_image(e->ctx, image.data()) != FPV_OK)
      return FPV_FAIL(std::string("fpv_set_delta_image: ") + fpv_last_error(e->ctx));
    e->delta_generation = delta->generation_;
    return true;
  }
};

void Frame::SetDevice(int device) { g_frame_device.store(device); }

void Frame::Touch() { generation_ = g_generation.fetch_add(1); }

size_t Frame::MaxCompressedPlaneSize(size_t xsize, size_t ysize) { return BrotliEncoderMaxCompressedSize(xsize * ysize); }
size_t Frame::MaxCompressedPreviewSize(size_t xsize, size_t ysize) {
  return BrotliEncoderMaxCompressedSize(xsize * ysize / 16);
}
size_t Frame::MaxCompressedPlaneSize() { return BrotliEncoderMaxCompressedSize(size_); }
size_t Frame::MaxCompressedPreviewSize() { return BrotliEncoderMaxCompressedSize(size_ / 16); }

// reference .cc:370-451: split into byte planes, NO_LOW_BYTES iff every low byte is zero
Frame::Frame(size_t xsize, size_t ysize, const uint16_t* image, int shift_to_left_align, bool big_endian,
             int64_t timestamp)
    : xsize_(xsize), ysize_(ysize), size_(xsize * ysize), timestamp_(timestamp), shift_(shift_to_left_align),
      big_endian_(big_endian) {
  if (!image || size_ == 0) return;
  state_ = FrameState::RAW;
  Touch();
  FrameGpu::Key key(xsize_, ysize_, shift_, big_endian_, g_frame_device.load());
  FrameGpu::Entry e;
  high_.resize(size_);
  if (shift_ != 8) low_.resize(size_);
  uint8_t fl = 0;
  bool good = FrameGpu::Get().Acquire(key, 0, &e);
  if (good) {
    good = fpv_split(e.ctx, image, 1, &fl, high_.data(), shift_ != 8 ? low_.data() : nullptr) == FPV_OK;
    if (!good) FPV_FAIL(std::string("fpv_split: ") + fpv_last_error(e.ctx));
    FrameGpu::Get().Release(key, e);
  }
  if (!good) {
    // no GPU: an empty frame, like a default-constructed one
    high_.clear();
    low_.clear();
    state_ = FrameState::EMPTY;
    return;
  }
  flags_ = fl;
  raw_ = std::make_shared<const std::vector<uint16_t>>(image, image + size_);
}

// reference .cc:453-465
Frame::Frame(size_t xsize, size_t ysize, const uint8_t* image, int64_t timestamp)
    : xsize_(xsize), ysize_(ysize), size_(xsize * ysize), flags_(FrameFlags::NO_LOW_BYTES), timestamp_(timestamp) {
  if (image) {
    state_ = FrameState::RAW;
    high_.assign(image, image + size_);
    Touch();
  }
}

// reference .cc:467-489 (including its two statements that modify the by-value arguments instead of the members)
Frame::Frame(size_t xsize, size_t ysize, uint8_t flags, uint8_t state, std::vector<uint8_t>&& high,
             std::vector<uint8_t>&& low, std::vector<uint8_t>&& preview, int64_t timestamp)
    : xsize_(xsize), ysize_(ysize), size_(xsize * ysize), flags_(flags), state_(state), timestamp_(timestamp) {
  high_ = std::move(high);
  low_ = std::move(low);
  preview_ = std::move(preview);
  if (preview_.empty()) state_ &= ~FrameState::PREVIEW_GENERATED;
  if (!low_.empty()) flags_ &= ~FrameFlags::NO_LOW_BYTES;
  Touch();
}

// reference .cc:777-785: preview, then delta prediction (iff a delta frame is given), then ClampedGradient
void Frame::Predict(Frame& delta_frame) {
  const bool with_delta = delta_frame.state() > FrameState::EMPTY;
  const bool preview_done = state_ & FrameState::PREVIEW_GENERATED;
  const bool delta_done = (state_ & FrameState::DELTA_PREDICTED) || !with_delta;
  const bool cg_done = state_ & FrameState::CG_PREDICTED;
  if (preview_done && delta_done && cg_done) return;   // every step returns early (.cc:492, :518, :547)
  if (size_ == 0 || high_.size() != size_) {
    // nothing to compute on (the reference's loops do not run for an empty frame); only the state moves on
    state_ |= FrameState::PREVIEW_GENERATED;
    if (with_delta) state_ |= FrameState::DELTA_PREDICTED;
    state_ |= FrameState::CG_PREDICTED;
    state_ &= ~FrameState::RAW;
    return;
  }
  if (preview_done || (state_ & FrameState::DELTA_PREDICTED) || cg_done) {
    FPV_FAIL("Frame::Predict: frame is partially predicted; the GPU path applies all steps in one fused call");
    return;
  }
  if (xsize_ % 4 || ysize_ % 4) {
    FPV_FAIL("Frame::Predict requires xsize % 4 == 0 and ysize % 4 == 0 (the reference reads out of bounds otherwise)");
    return;
  }
  // input: the raw image the constructor saw if it is still there, else the planes as a left-aligned image
  std::vector<uint16_t> from_planes;
  const uint16_t* input;
  int shift = 0;
  bool big_endian = false;
  if (raw_) {
    input = raw_->data();
    shift = shift_;
    big_endian = big_endian_;
  } else {
    from_planes.resize(size_);
    if (low_.size() == size_)
      for (size_t i = 0; i < size_; i++) from_planes[i] = (uint16_t)((high_[i] << 8) | low_[i]);
    else
      for (size_t i = 0; i < size_; i++) from_planes[i] = (uint16_t)(high_[i] << 8);
    input = from_planes.data();
  }
  FrameGpu::Key key(xsize_, ysize_, shift, big_endian, g_frame_device.load());
  FrameGpu::Entry e;
  if (!FrameGpu::Get().Acquire(key, with_delta ? delta_frame.generation_ : 0, &e)) return;
  const bool has_low = shift != 8;
  std::vector<uint8_t> high(size_), low(has_low ? size_ : 0), preview(size_ / 16 ? size_ / 16 : 1);
  uint8_t fl = 0;
  bool good = FrameGpu::SetDelta(&e, with_delta ? &delta_frame : nullptr, size_);
  if (good) {
    good = fpv_encode(e.ctx, input, 1, with_delta ? FPV_ENC_DEFAULT : FPV_ENC_NO_DELTA, &fl, high.data(),
                      has_low ? low.data() : nullptr, preview.data()) == FPV_OK;
    if (!good) FPV_FAIL(std::string("fpv_encode: ") + fpv_last_error(e.ctx));
  }
  FrameGpu::Get().Release(key, e);
  if (!good) return;
  preview.resize(size_ / 16);
  // a frame whose low plane the caller never had (u8 constructor / plane constructor without low) keeps none
  const bool keep_low = !low_.empty();
  high_.swap(high);
  if (keep_low && has_low) low_.swap(low);
  preview_.swap(preview);
  // NO_LOW_BYTES is the constructor's business (.cc:447-449); Predict only decides the two predictors
  flags_ = (uint8_t)((flags_ & ~(FrameFlags::USE_DELTA | FrameFlags::USE_CG)) |
                     (fl & (FrameFlags::USE_DELTA | FrameFlags::USE_CG)));
  state_ &= ~FrameState::RAW;
  state_ |= FrameState::PREVIEW_GENERATED | FrameState::CG_PREDICTED;
  if (with_delta) state_ |= FrameState::DELTA_PREDICTED;
  raw_.reset();
  Touch();
}

// reference .cc:643-688
void Frame::ApplyBrotliCompression() {
  if (state_ & FrameState::COMPRESSED) return;
  std::vector<uint8_t> scratch, out;
  BrotliPlane(high_.data(), size_, &scratch, &out);
  high_.swap(out);
  if (flags_ & FrameFlags::NO_LOW_BYTES) {
    low_.clear();
  } else {
    out.clear();
    BrotliPlane(low_.data(), size_, &scratch, &out);
    low_.swap(out);
  }
  if (state_ & FrameState::PREVIEW_GENERATED) {
    out.clear();
    BrotliPlane(preview_.data(), preview_.size(), &scratch, &out);
    preview_.swap(out);
  }
  state_ &= ~FrameState::RAW;
  state_ |= FrameState::COMPRESSED;
  Touch();
}

// reference .cc:738-745
void Frame::Compress(Frame& delta_frame) {
  if (state_ & FrameState::COMPRESSED) return;
  Predict(delta_frame);
  ApplyBrotliCompression();
}

// reference .cc:747-775
void Frame::Uncompress(Frame& delta_frame) {
  if (state_ & FrameState::COMPRESSED) {
    std::vector<uint8_t> plain;
    if (!high_.empty() && BrotliToVector(high_, &plain)) high_.swap(plain);
    if (!(low_.empty() || (flags_ & FrameFlags::NO_LOW_BYTES)) && BrotliToVector(low_, &plain)) low_.swap(plain);
    if ((state_ & FrameState::PREVIEW_GENERATED) && !preview_.empty() && BrotliToVector(preview_, &plain))
      preview_.swap(plain);
    state_ &= ~FrameState::COMPRESSED;
    Touch();
  }
  // .cc:612-641 and :595-610: which of the two undo steps apply
  const bool undo_cg = (state_ & FrameState::CG_PREDICTED) && (flags_ & FrameFlags::USE_CG);
  const bool undo_delta = (state_ & FrameState::DELTA_PREDICTED) && (flags_ & FrameFlags::USE_DELTA) &&
                          delta_frame.state() != FrameState::EMPTY;
  const bool cg_high = undo_cg && high_.size() == size_;
  const bool cg_preview = undo_cg && (state_ & FrameState::PREVIEW_GENERATED) && preview_.size() == size_ / 16 && size_ >= 16;
  if ((cg_high || undo_delta) && high_.size() == size_ && size_ > 0) {
    FrameGpu::Key key(xsize_, ysize_, 0, false, g_frame_device.load());
    FrameGpu::Entry e;
    if (!FrameGpu::Get().Acquire(key, undo_delta ? delta_frame.generation_ : 0, &e)) return;
    bool good = FrameGpu::SetDelta(&e, undo_delta ? &delta_frame : nullptr, size_);
    if (good) {
      const uint8_t fl = (uint8_t)((cg_high ? FrameFlags::USE_CG : 0) | (undo_delta ? FrameFlags::USE_DELTA : 0));
      good = fpv_unpredict_planes(e.ctx, high_.data(), low_.size() == size_ ? low_.data() : nullptr,
                                  cg_preview ? preview_.data() : nullptr, &fl, 1) == FPV_OK;
      if (!good) FPV_FAIL(std::string("fpv_unpredict_planes: ") + fpv_last_error(e.ctx));
    }
    FrameGpu::Get().Release(key, e);
    if (!good) return;
    Touch();
  } else if (cg_preview) {
    FPV_FAIL("Frame::Uncompress: a preview without its high plane cannot be un-predicted on the GPU path");
    return;
  }
  // state / flag bookkeeping exactly as .cc:636-640 and :605-609
  if (undo_cg) {
    flags_ &= ~FrameFlags::USE_CG;
    state_ &= ~FrameState::CG_PREDICTED;
    if (state_ < FrameState::DELTA_PREDICTED) state_ |= FrameState::RAW;
  }
  if (undo_delta) {
    flags_ &= ~FrameFlags::USE_DELTA;
    state_ &= ~FrameState::DELTA_PREDICTED;
    if (state_ < FrameState::DELTA_PREDICTED) state_ |= FrameState::RAW;
  }
}

// reference .cc:787-818 (already compressed: copy out) and :690-728 (compress into the caller's buffers)
void Frame::CompressPredicted(size_t* encoded_high_size, uint8_t* encoded_high_buffer, size_t* encoded_low_size,
                              uint8_t* encoded_low_buffer, size_t* encoded_preview_size,
                              uint8_t* encoded_preview_buffer, bool parallel) {
  if (state_ & FrameState::COMPRESSED) {
    auto copy_out = [](const std::vector<uint8_t>& v, size_t* size, uint8_t* buffer) {
      if (buffer && *size >= v.size()) {
        memcpy(buffer, v.data(), v.size());
        *size = v.size();
      } else {
        *size = 0;
      }
    };
    copy_out(high_, encoded_high_size, encoded_high_buffer);
    copy_out(low_, encoded_low_size, encoded_low_buffer);
    copy_out(preview_, encoded_preview_size, encoded_preview_buffer);
    return;
  }
  // the low plane carries most of the entropy: it is the one compressed in parallel (.cc:697-708)
  std::future<void> low_task;
  auto compress_low = [this, encoded_low_size, encoded_low_buffer] {
    if (!BrotliEncoderCompress(1, BROTLI_DEFAULT_WINDOW, BROTLI_DEFAULT_MODE, size_, low_.data(), encoded_low_size,
                               encoded_low_buffer))
      *encoded_low_size = 0;
  };
  if (!encoded_low_buffer || (flags_ & FrameFlags::NO_LOW_BYTES) || low_.size() != size_) *encoded_low_size = 0;
  else if (parallel) low_task = std::async(std::launch::async, compress_low);
  else compress_low();
  if (encoded_high_buffer && high_.size() == size_) {
    if (!BrotliEncoderCompress(1, BROTLI_DEFAULT_WINDOW, BROTLI_DEFAULT_MODE, size_, high_.data(), encoded_high_size,
                               encoded_high_buffer))
      *encoded_high_size = 0;
  } else {
    *encoded_high_size = 0;
  }
  if (encoded_preview_buffer && (state_ & FrameState::PREVIEW_GENERATED)) {
    if (!BrotliEncoderCompress(1, BROTLI_DEFAULT_WINDOW, BROTLI_DEFAULT_MODE, preview_.size(), preview_.data(),
                               encoded_preview_size, encoded_preview_buffer))
      *encoded_preview_size = 0;
  } else {
    *encoded_preview_size = 0;
  }
  if (low_task.valid()) low_task.wait();
}

// reference .cc:820-828
void Frame::OutputCore(std::vector<uint8_t>* out) {
  if (!(state_ & FrameState::COMPRESSED)) return;
  out->reserve(out->size() + 1 + high_.size() + low_.size());
  out->push_back(flags_);
  out->insert(out->end(), low_.begin(), low_.end());
  out->insert(out->end(), high_.begin(), high_.end());
}

// reference .cc:830-846
void Frame::OutputFull(std::vector<uint8_t>* out) {
  if (!(state_ & FrameState::COMPRESSED)) return;
  const size_t total_size = (9 + 1 + preview_.size()) + (1 + high_.size() + low_.size());
  out->reserve(out->size() + total_size);
  AppendU32((uint32_t)total_size, out);
  out->push_back(0);
  AppendU32((uint32_t)(preview_.size() + 1), out);
  out->push_back((uint8_t)((flags_ & FrameFlags::USE_CG) | FrameFlags::NO_LOW_BYTES));
  out->insert(out->end(), preview_.begin(), preview_.end());
  OutputCore(out);
}

}  // namespace fpvc

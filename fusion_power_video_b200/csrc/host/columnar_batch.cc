// columnar_batch.cc -- see columnar_batch.h.  Reference: columnar_batch/columnar_batch.cc,
// columnar_batch_encoder.cc, columnar_batch_decoder.cc.
#include "columnar_batch.h"

#include <string.h>

#include <algorithm>

#include "host_internal.h"

namespace fpvc {
namespace columnarbatch {

using namespace internal;

namespace {

// A second, lazily created context per schema decodes previews: a preview is a (xsize/4) x (ysize/4)
// high-plane-only image whose flags never carry USE_DELTA (reference .cc:842).
struct PreviewContexts {
  std::mutex m;
  std::vector<std::pair<const BatchSchema*, fpv_ctx*>> list;
  fpv_ctx* get(const BatchSchema* s, int device) {
    std::lock_guard<std::mutex> l(m);
    for (auto& e : list)
      if (e.first == s) return e.second;
    fpv_ctx* c = nullptr;
    if (fpv_create(&c, device, (uint32_t)(s->xsize() / 4), (uint32_t)(s->ysize() / 4), 0, 0,
                   (uint32_t)BatchSchema::kMaxBatch) != FPV_OK)
      return nullptr;
    list.emplace_back(s, c);
    return c;
  }
  void drop(const BatchSchema* s) {
    std::lock_guard<std::mutex> l(m);
    for (size_t i = 0; i < list.size(); i++)
      if (list[i].first == s) {
        fpv_destroy(list[i].second);
        list.erase(list.begin() + (ptrdiff_t)i);
        return;
      }
  }
};
PreviewContexts& previews() {
  static PreviewContexts p;
  return p;
}

}  // namespace

// ---- BatchSchema ---------------------------------------------------------------------------------

// reference columnar_batch.cc:6-24 (which compresses into zero-sized vectors; this does what was meant)
BatchSchema::BatchSchema(size_t xsize, size_t ysize, size_t shifted_left, Frame& uncompressed_delta_frame)
    : xsize_(xsize), ysize_(ysize), shifted_left_(shifted_left), big_endian_(false), device_(0) {
  delta_frame_ = uncompressed_delta_frame;
  const size_t P = xsize * ysize;
  if (delta_frame_.high().size() != P) {
    FPV_FAIL("BatchSchema: the delta frame must be un-compressed and of the schema's size");
    return;
  }
  if (fpv_create(&ctx_, device_, (uint32_t)xsize, (uint32_t)ysize, (int)shifted_left, 0, (uint32_t)kMaxBatch) != FPV_OK) {
    FPV_FAIL(std::string("fpv_create: ") + fpv_last_error(nullptr));
    return;
  }
  // the context's delta frame in image form: (high << 8) | low per pixel (reference .cc:337-338)
  std::vector<uint16_t> image(P);
  const std::vector<uint8_t>& h = delta_frame_.high();
  const std::vector<uint8_t>& l = delta_frame_.low();
  for (size_t i = 0; i < P; i++) image[i] = (uint16_t)((h[i] << 8) | (l.size() == P ? l[i] : 0));
  if (fpv_set_delta_image(ctx_, image.data()) != FPV_OK) {
    FPV_FAIL(std::string("fpv_set_delta_image: ") + fpv_last_error(ctx_));
    return;
  }
  compressed_high_.resize(Frame::MaxCompressedPlaneSize(xsize, ysize));
  compressed_low_.resize(Frame::MaxCompressedPlaneSize(xsize, ysize));
  size_t hs = compressed_high_.size(), ls = compressed_low_.size(), ps = 0;
  uncompressed_delta_frame.CompressPredicted(&hs, compressed_high_.data(), &ls, compressed_low_.data(), &ps, nullptr);
  compressed_high_.resize(hs);
  compressed_low_.resize(ls);
  ok_ = hs > 0;
}

BatchSchema::BatchSchema(size_t xsize, size_t ysize, size_t shifted_left, bool big_endian, const uint16_t* delta_frame,
                         int device)
    : xsize_(xsize), ysize_(ysize), shifted_left_(shifted_left), big_endian_(big_endian), device_(device) {
  if (fpv_create(&ctx_, device, (uint32_t)xsize, (uint32_t)ysize, (int)shifted_left, big_endian ? 1 : 0,
                 (uint32_t)kMaxBatch) != FPV_OK) {
    FPV_FAIL(std::string("fpv_create: ") + fpv_last_error(nullptr));
    return;
  }
  if (fpv_set_delta_raw(ctx_, delta_frame) != FPV_OK) {
    FPV_FAIL(std::string("fpv_set_delta_raw: ") + fpv_last_error(ctx_));
    return;
  }
  // The schema carries the delta frame's own planes, split but neither delta- nor CG-predicted: run the
  // transform without a delta frame and undo whatever ClampedGradient prediction it chose.
  const size_t P = xsize * ysize, PP = (xsize / 4) * (ysize / 4);
  const bool has_low = shifted_left != 8;
  std::vector<uint8_t> high(P), low(has_low ? P : 0), prev(PP ? PP : 1), scratch;
  uint8_t flags = 0;
  if (fpv_encode(ctx_, delta_frame, 1, FPV_ENC_NO_DELTA, &flags, high.data(), has_low ? low.data() : nullptr,
                 prev.data()) != FPV_OK ||
      fpv_unpredict_planes(ctx_, high.data(), has_low ? low.data() : nullptr, nullptr, &flags, 1) != FPV_OK) {
    FPV_FAIL(std::string("delta frame planes: ") + fpv_last_error(ctx_));
    return;
  }
  ok_ = BrotliPlane(high.data(), P, &scratch, &compressed_high_);
  if (ok_ && has_low && !(flags & FPV_FLAG_NO_LOW_BYTES)) ok_ = BrotliPlane(low.data(), P, &scratch, &compressed_low_);
  // the same planes as a Frame, for callers of the reference's delta_frame() accessor
  if (ok_)
    delta_frame_ = Frame(xsize, ysize, (uint8_t)(flags & FPV_FLAG_NO_LOW_BYTES), (uint8_t)FrameState::RAW, std::move(high),
                         std::move(low), std::vector<uint8_t>());
}

BatchSchema::~BatchSchema() {
  previews().drop(this);
  if (ctx_) fpv_destroy(ctx_);
}

// ---- Batch ------------------------------------------------------------------------------------------

Batch::Batch(size_t batch_size, SchemaPtr schema)
    : schema_(schema), batch_size_(batch_size), timestamps_(batch_size), flags_(batch_size),
      preview_offsets_(batch_size + 1, 0), high_offsets_(batch_size + 1, 0), low_offsets_(batch_size + 1, 0) {}

void Batch::Reset() {
  length_ = 0;
  preview_.clear();
  high_.clear();
  low_.clear();
  std::fill(preview_offsets_.begin(), preview_offsets_.end(), 0u);
  std::fill(high_offsets_.begin(), high_offsets_.end(), 0u);
  std::fill(low_offsets_.begin(), low_offsets_.end(), 0u);
}

// reference columnar_batch.cc:65-90
bool Batch::AppendPredicted(Frame predicted_frame) {
  if (length_ >= batch_size_) return false;
  const size_t x = schema_->xsize(), y = schema_->ysize();
  std::vector<uint8_t> high(Frame::MaxCompressedPlaneSize(x, y)), low(Frame::MaxCompressedPlaneSize(x, y)),
      preview(Frame::MaxCompressedPreviewSize(x, y));
  size_t hs = high.size(), ls = low.size(), ps = preview.size();
  predicted_frame.CompressPredicted(&hs, high.data(), &ls, low.data(), &ps, preview.data());
  high.resize(hs);
  low.resize(ls);
  preview.resize(ps);
  return AppendPredicted(predicted_frame.timestamp(), predicted_frame.flags(), preview, high, low);
}

bool Batch::AppendPredicted(int64_t timestamp, uint8_t flags, const std::vector<uint8_t>& preview,
                            const std::vector<uint8_t>& high, const std::vector<uint8_t>& low) {
  if (length_ >= batch_size_) return false;
  timestamps_[length_] = timestamp;
  flags_[length_] = flags;
  preview_.insert(preview_.end(), preview.begin(), preview.end());
  high_.insert(high_.end(), high.begin(), high.end());
  low_.insert(low_.end(), low.begin(), low.end());
  preview_offsets_[length_ + 1] = (uint32_t)preview_.size();
  high_offsets_[length_ + 1] = (uint32_t)high_.size();
  low_offsets_[length_ + 1] = (uint32_t)low_.size();
  length_++;
  return true;
}

// Frames [first, first + count) as images: brotli-decode the planes the type needs, then ONE GPU call for
// the inverse transform (reference: Frame::Uncompress per frame on the CPU, columnar_batch.cc:92-123).
std::vector<Image> Batch::Extract(size_t first, size_t count, Image::Type type) {
  std::vector<Image> out;
  BatchSchema& sc = *schema_;
  const size_t W = sc.xsize(), H = sc.ysize(), P = W * H, PW = W / 4, PH = H / 4, PP = PW * PH;
  if (!sc.ok() || first + count > length_ || count == 0) return out;
  out.reserve(count);
  for (size_t base = first; base < first + count; base += sc.max_batch()) {
    const size_t n = std::min(sc.max_batch(), first + count - base);
    std::vector<uint8_t> fl(n);
    bool good = true;
    if (type == Image::Type::PREVIEW) {
      std::vector<uint8_t> planes(n * PP);
      std::vector<uint16_t> img(n * PP);
      for (size_t i = 0; i < n && good; i++) {
        size_t pos = preview_offsets_[base + i];
        good = BrotliUnplane(preview_.data(), preview_offsets_[base + i + 1], &pos, planes.data() + i * PP, PP);
        fl[i] = (uint8_t)((flags_[base + i] & FPV_FLAG_USE_CG) | FPV_FLAG_NO_LOW_BYTES);
      }
      fpv_ctx* pc = good ? previews().get(&sc, sc.device()) : nullptr;
      {
        std::lock_guard<std::mutex> l(sc.context_mutex());
        good = pc && fpv_decode(pc, planes.data(), nullptr, fl.data(), (uint32_t)n, FPV_DEC_DEFAULT, img.data()) == FPV_OK;
      }
      if (!good) { FPV_FAIL("preview decode failed"); return out; }
      for (size_t i = 0; i < n; i++) {
        std::vector<uint8_t> data(PP);
        for (size_t k = 0; k < PP; k++) data[k] = (uint8_t)(img[i * PP + k] >> 8);
        out.emplace_back(timestamps_[base + i], PW, PH, 8, type, std::move(data));
      }
      continue;
    }
    const bool full = type == Image::Type::FULL;
    std::vector<uint8_t> high(n * P), low(full ? n * P : 0);
    std::vector<uint16_t> img(n * P);
    for (size_t i = 0; i < n && good; i++) {
      size_t pos = high_offsets_[base + i];
      good = BrotliUnplane(high_.data(), high_offsets_[base + i + 1], &pos, high.data() + i * P, P);
      fl[i] = flags_[base + i];
      if (!full) fl[i] |= FPV_FLAG_NO_LOW_BYTES;           // MSB8: the low plane is not looked at (columnar_batch.cc:104)
      if (good && full && !(fl[i] & FPV_FLAG_NO_LOW_BYTES)) {
        pos = low_offsets_[base + i];
        good = BrotliUnplane(low_.data(), low_offsets_[base + i + 1], &pos, low.data() + i * P, P);
      }
    }
    {
      std::lock_guard<std::mutex> l(sc.context_mutex());
      good = good && fpv_decode(sc.context(), high.data(), full ? low.data() : nullptr, fl.data(), (uint32_t)n,
                                FPV_DEC_DEFAULT, img.data()) == FPV_OK;
    }
    if (!good) { FPV_FAIL(std::string("batch decode failed: ") + fpv_last_error(sc.context())); return out; }
    for (size_t i = 0; i < n; i++) {
      if (full) {
        std::vector<uint8_t> data(2 * P);
        memcpy(data.data(), img.data() + i * P, 2 * P);      // low | high << 8 (columnar_batch.cc:119-120)
        out.emplace_back(timestamps_[base + i], W, H, (uint8_t)(16 - sc.shiftedLeft()), type, std::move(data));
      } else {
        std::vector<uint8_t> data(P);
        for (size_t k = 0; k < P; k++) data[k] = (uint8_t)(img[i * P + k] >> 8);
        out.emplace_back(timestamps_[base + i], W, H, 8, type, std::move(data));
      }
    }
  }
  return out;
}

Image Batch::ExtractImage(size_t index, Image::Type type) {
  std::vector<Image> v = Extract(index, 1, type);
  return v.empty() ? Image() : std::move(v[0]);
}

std::vector<Image> Batch::ExtractImages(Image::Type type) { return Extract(0, length_, type); }

// ---- ColumnarBatchEncoder ------------------------------------------------------------------------

ColumnarBatchEncoder::ColumnarBatchEncoder(size_t xsize, size_t ysize, int shift_to_left_align, bool big_endian,
                                           BatchProcessor batch_processor, int frames_per_batch, size_t brotli_threads,
                                           int device)
    : batch_processor_(batch_processor), frames_per_batch_((size_t)std::max(1, frames_per_batch)), xsize_(xsize),
      ysize_(ysize), brotli_threads_(brotli_threads), shift_(shift_to_left_align), device_(device),
      big_endian_(big_endian), closing_timestamp_future_(promised_closing_timestamp_.get_future()),
      encoder_thread_([this] { EncoderTask(); }) {}

ColumnarBatchEncoder::~ColumnarBatchEncoder() {
  Close();
  encoder_thread_.join();
}

std::future<void*> ColumnarBatchEncoder::PushFrame(uint64_t timestamp, uint16_t* frame, void* info) {
  std::promise<void*> done;
  {
    std::unique_lock<std::mutex> lock(queue_mutex_);
    if (closing_) return std::future<void*>();
    if (!schema_) {
      // the first frame doubles as the delta frame (columnar_batch_encoder.cc:37-46) and is encoded like any other
      schema_ = std::make_shared<BatchSchema>(xsize_, ysize_, (size_t)shift_, big_endian_, frame, device_);
      ok_ = schema_->ok();
    }
    const size_t P = xsize_ * ysize_;
    filling_.frames.insert(filling_.frames.end(), frame, frame + P);
    filling_.timestamps.push_back((int64_t)timestamp);
    if (filling_.timestamps.size() == frames_per_batch_) {
      queue_.push_back(std::move(filling_));
      filling_ = Job();
    }
  }
  queue_condition_.notify_one();
  done.set_value(info);
  return done.get_future();
}

std::shared_future<int64_t> ColumnarBatchEncoder::Close() {
  {
    std::unique_lock<std::mutex> lock(queue_mutex_);
    if (!closing_) {
      closing_ = true;
      filling_.close = true;       // the partial batch (possibly empty) ends the stream
      queue_.push_back(std::move(filling_));
      filling_ = Job();
    }
  }
  queue_condition_.notify_one();
  return closing_timestamp_future_;
}

BatchPtr ColumnarBatchEncoder::BatchToFill() {
  std::lock_guard<std::mutex> l(empty_mutex_);
  if (empty_batches_.empty()) return std::make_shared<Batch>(frames_per_batch_, schema_);
  BatchPtr b = empty_batches_.front();
  empty_batches_.pop_front();
  return b;
}

void ColumnarBatchEncoder::ReturnProcessedBatch(BatchPtr processed) {
  if (!processed) return;
  processed->Reset();
  std::lock_guard<std::mutex> l(empty_mutex_);
  empty_batches_.push_back(processed);
}

// One batch: Frame ctor + Predict for all its frames in one GPU call, then brotli per plane on a few threads.
void ColumnarBatchEncoder::EncodeJob(Job& job) {
  const size_t n = job.timestamps.size();
  if (n == 0) {
    if (job.close) batch_processor_(nullptr);          // nothing left to flush (columnar_batch_encoder.cc:86-89)
    return;
  }
  const size_t P = xsize_ * ysize_, PP = (xsize_ / 4) * (ysize_ / 4);
  const bool has_low = shift_ != 8;
  BatchPtr batch = BatchToFill();
  std::vector<uint8_t> flags(n), high(n * P), low(has_low ? n * P : 0), prev(n * (PP ? PP : 1));
  bool good = schema_ && schema_->ok();
  for (size_t off = 0; off < n && good; off += schema_->max_batch()) {
    const size_t m = std::min(schema_->max_batch(), n - off);
    std::lock_guard<std::mutex> l(schema_->context_mutex());
    good = fpv_encode(schema_->context(), job.frames.data() + off * P, (uint32_t)m, FPV_ENC_DEFAULT, flags.data() + off,
                      high.data() + off * P, has_low ? low.data() + off * P : nullptr, prev.data() + off * PP) == FPV_OK;
  }
  if (!good) {
    ok_ = false;
    FPV_FAIL(std::string("columnar batch encode failed: ") + (schema_ ? fpv_last_error(schema_->context()) : "no schema"));
    return;
  }
  std::vector<std::vector<uint8_t>> cp(n), ch(n), cl(n);
  Pool pool(std::min(brotli_threads_, 3 * n));
  ParallelFor(&pool, 3 * n, [&](size_t k) {
    thread_local std::vector<uint8_t> scratch;
    const size_t i = k / 3;
    if (k % 3 == 0) BrotliPlane(high.data() + i * P, P, &scratch, &ch[i]);
    else if (k % 3 == 1) { if (has_low && !(flags[i] & FPV_FLAG_NO_LOW_BYTES)) BrotliPlane(low.data() + i * P, P, &scratch, &cl[i]); }
    else BrotliPlane(prev.data() + i * PP, PP, &scratch, &cp[i]);
  });
  for (size_t i = 0; i < n; i++) batch->AppendPredicted(job.timestamps[i], flags[i], cp[i], ch[i], cl[i]);
  latest_stored_timestamp_ = batch->LatestTimestamp();
  batch_processor_(batch);
}

void ColumnarBatchEncoder::EncoderTask() {
  for (;;) {
    Job job;
    {
      std::unique_lock<std::mutex> lock(queue_mutex_);
      queue_condition_.wait(lock, [this] { return !queue_.empty(); });
      job = std::move(queue_.front());
      queue_.pop_front();
    }
    EncodeJob(job);
    if (job.close) {
      promised_closing_timestamp_.set_value(latest_stored_timestamp_);
      return;
    }
  }
}

// ---- ColumnarBatchDecoder ------------------------------------------------------------------------

ColumnarBatchDecoder::ColumnarBatchDecoder(Image::Type type, bool unshift, ImageProcessor image_processor)
    : image_processor_(image_processor), type_(type), unshift_(unshift),
      closing_timestamp_future_(promised_closing_timestamp_.get_future()), decoder_thread_([this] { DecoderTask(); }) {}

ColumnarBatchDecoder::~ColumnarBatchDecoder() {
  Close();
  decoder_thread_.join();
}

std::future<BatchPtr> ColumnarBatchDecoder::PushBatch(BatchPtr batch) {
  std::future<BatchPtr> f;
  {
    std::unique_lock<std::mutex> lock(queue_mutex_);
    if (!batch) return std::future<BatchPtr>();
    if (!schema_) schema_ = batch->schema();
    if (closing_ || schema_.get() != batch->schema().get()) return std::future<BatchPtr>();
    batch_queue_.emplace_back();
    batch_queue_.back().batch = batch;
    f = batch_queue_.back().promise.get_future();
  }
  queue_condition_.notify_one();
  return f;
}

std::shared_future<int64_t> ColumnarBatchDecoder::Close() {
  {
    std::unique_lock<std::mutex> lock(queue_mutex_);
    if (!closing_) {
      closing_ = true;
      batch_queue_.emplace_back();      // a null batch ends the stream
    }
  }
  queue_condition_.notify_one();
  return closing_timestamp_future_;
}

void ColumnarBatchDecoder::DecoderTask() {
  for (;;) {
    Promised p;
    {
      std::unique_lock<std::mutex> lock(queue_mutex_);
      queue_condition_.wait(lock, [this] { return !batch_queue_.empty(); });
      p = std::move(batch_queue_.front());
      batch_queue_.pop_front();
    }
    if (!p.batch) break;
    std::vector<Image> images = p.batch->ExtractImages(type_);
    const size_t shifted_left = schema_->shiftedLeft();
    for (Image& img : images) {
      if (unshift_ && shifted_left > 0 && img.bpp() > 8) {          // columnar_batch_decoder.cc:82-85
        uint16_t* d = img.data16();
        const size_t px = img.xsize() * img.ysize();
        for (size_t k = 0; k < px; k++) d[k] = (uint16_t)(d[k] >> shifted_left);
      }
      image_processor_(std::move(img));
    }
    latest_provided_timestamp_ = p.batch->LatestTimestamp();
    p.promise.set_value(p.batch);
  }
  promised_closing_timestamp_.set_value(latest_provided_timestamp_);
}

}  // namespace columnarbatch
}  // namespace fpvc

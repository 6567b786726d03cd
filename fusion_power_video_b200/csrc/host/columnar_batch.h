// columnar_batch.h -- the reference's columnar in-memory container (columnar_batch/columnar_batch.h,
// columnar_batch_encoder.h, columnar_batch_decoder.h) on the GPU transform.
//
// Same classes, method names and behaviour: a ColumnarBatchEncoder turns pushed frames into Batches of
// frames_per_batch frames (columns: timestamps, flags, offsets, brotli-compressed preview / high / low
// planes); the first pushed frame doubles as the delta frame and defines the BatchSchema; a
// ColumnarBatchDecoder turns Batches back into Images (PREVIEW, MSB8 or FULL).  Differences: the frame
// transform (Frame ctor + Predict, reference columnar_batch_encoder.cc:61-71) and its inverse
// (Frame::Uncompress, columnar_batch.cc:110) run on the GPU one batch at a time through the C ABI.  The
// reference's Frame-based entry points (BatchSchema from an uncompressed delta Frame, delta_frame(),
// Batch::AppendPredicted(Frame)) are kept next to the batched ones this file's encoder uses (BatchSchema from
// the raw delta frame, AppendPredicted of compressed planes).  Where the reference has defects (compressing into a zero-sized vector,
// columnar_batch.cc:10-22; passing the high plane as the low plane, columnar_batch_decoder.cc:73-74) this
// file implements what was meant.
#ifndef FPV_B200_COLUMNAR_BATCH_H_
#define FPV_B200_COLUMNAR_BATCH_H_

#include <stddef.h>
#include <stdint.h>

#include <condition_variable>
#include <functional>
#include <future>
#include <list>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "fusion_power_video.h"

struct fpv_ctx;

namespace fpvc {
namespace columnarbatch {

class BatchSchema {
 public:
  // As the reference (columnar_batch.h:11): from an un-predicted, un-compressed delta Frame.
  BatchSchema(size_t xsize, size_t ysize, size_t shifted_left, Frame& uncompressed_delta_frame);
  // Extension: from the raw uint16 frame as the camera delivers it (same meaning as the encoders' frames)
  BatchSchema(size_t xsize, size_t ysize, size_t shifted_left, bool big_endian, const uint16_t* delta_frame,
              int device = 0);
  ~BatchSchema();
  BatchSchema(const BatchSchema&) = delete;
  BatchSchema& operator=(const BatchSchema&) = delete;

  size_t xsize() const { return xsize_; }
  size_t ysize() const { return ysize_; }
  size_t shiftedLeft() const { return shifted_left_; }
  bool bigEndian() const { return big_endian_; }
  bool ok() const { return ok_; }
  Frame& delta_frame() { return delta_frame_; }
  // Delta frame is _not_ CG predicted (reference columnar_batch.h:17)
  const std::vector<uint8_t>& compressedDeltaFrameHighPlane() const { return compressed_high_; }
  const std::vector<uint8_t>& compressedDeltaFrameLowPlane() const { return compressed_low_; }

  // the GPU context (geometry + resident delta frame) every Batch of this schema encodes / decodes with
  fpv_ctx* context() const { return ctx_; }
  int device() const { return device_; }
  std::mutex& context_mutex() { return ctx_mutex_; }
  size_t max_batch() const { return kMaxBatch; }
  static constexpr size_t kMaxBatch = 64;

 private:
  size_t xsize_, ysize_, shifted_left_;
  bool big_endian_, ok_ = false;
  int device_ = 0;
  std::vector<uint8_t> compressed_high_, compressed_low_;
  Frame delta_frame_;
  fpv_ctx* ctx_ = nullptr;
  std::mutex ctx_mutex_;
};

typedef std::shared_ptr<BatchSchema> SchemaPtr;

class Image {
 public:
  enum Type { PREVIEW, MSB8, FULL };

  Image(int64_t timestamp = -1, size_t xsize = 0, size_t ysize = 0, uint8_t bpp = 0, Type type = Type::FULL,
        std::vector<uint8_t>&& data = std::vector<uint8_t>())
      : timestamp_(timestamp), xsize_(xsize), ysize_(ysize), bpp_(bpp), data_(std::move(data)), type_(type) {}

  int64_t timestamp() const { return timestamp_; }
  size_t xsize() const { return xsize_; }
  size_t ysize() const { return ysize_; }
  size_t bpp() const { return bpp_; }
  uint8_t* data8() { return data_.data(); }
  uint16_t* data16() { return reinterpret_cast<uint16_t*>(data_.data()); }
  Type type() const { return type_; }

 private:
  int64_t timestamp_;
  size_t xsize_, ysize_;
  uint8_t bpp_;
  std::vector<uint8_t> data_;
  Type type_;
};

typedef std::function<void(Image)> ImageProcessor;

class Batch {
 public:
  Batch(size_t batch_size, SchemaPtr schema);

  void Reset();
  // As the reference (columnar_batch.h:73, columnar_batch.cc:65-90): compresses the predicted frame's planes
  // (Frame::CompressPredicted) into the columns.
  bool AppendPredicted(Frame predicted_frame);
  // Extension: one predicted frame given as flags and the brotli streams of its three planes (low may be empty)
  bool AppendPredicted(int64_t timestamp, uint8_t flags, const std::vector<uint8_t>& preview,
                       const std::vector<uint8_t>& high, const std::vector<uint8_t>& low);

  bool Empty() const { return length_ == 0; }
  bool Full() const { return length_ == batch_size_; }
  int64_t LatestTimestamp() const { return length_ == 0 ? -1 : timestamps_[length_ - 1]; }
  size_t length() const { return length_; }
  Image ExtractImage(size_t index, Image::Type type);
  // every image of the batch in one GPU call (what ColumnarBatchDecoder uses)
  std::vector<Image> ExtractImages(Image::Type type);
  SchemaPtr schema() const { return schema_; }

  // column access (read-only): the compressed bytes of frame i are [offsets[i], offsets[i + 1])
  const int64_t* timestamps() const { return timestamps_.data(); }
  const uint8_t* flags() const { return flags_.data(); }
  const std::vector<uint32_t>& preview_offsets() const { return preview_offsets_; }
  const std::vector<uint32_t>& high_plane_offsets() const { return high_offsets_; }
  const std::vector<uint32_t>& low_plane_offsets() const { return low_offsets_; }
  const std::vector<uint8_t>& preview_column() const { return preview_; }
  const std::vector<uint8_t>& high_plane_column() const { return high_; }
  const std::vector<uint8_t>& low_plane_column() const { return low_; }

 private:
  std::vector<Image> Extract(size_t first, size_t count, Image::Type type);
  SchemaPtr schema_;
  size_t batch_size_, length_ = 0;
  std::vector<int64_t> timestamps_;
  std::vector<uint8_t> flags_;
  std::vector<uint32_t> preview_offsets_, high_offsets_, low_offsets_;
  std::vector<uint8_t> preview_, high_, low_;
};

typedef std::shared_ptr<Batch> BatchPtr;
typedef std::function<void(BatchPtr)> BatchProcessor;

class ColumnarBatchEncoder {
 public:
  ColumnarBatchEncoder(size_t xsize, size_t ysize, int shift_to_left_align, bool big_endian,
                       BatchProcessor batch_processor, int frames_per_batch = 10, size_t brotli_threads = 4,
                       int device = 0);
  ~ColumnarBatchEncoder();

  // The frame is copied before the call returns: the future is ready at once and yields `info`
  // (reference: ready once the frame buffer may be reused, columnar_batch_encoder.cc:27-50).
  std::future<void*> PushFrame(uint64_t timestamp, uint16_t* frame, void* info);
  void ReturnProcessedBatch(BatchPtr processed);
  // Flushes the partial batch (batch_processor(nullptr) if there is none, as the reference does) and
  // resolves to the timestamp of the last frame handed to the batch processor.
  std::shared_future<int64_t> Close();
  bool ok() const { return ok_; }

 private:
  struct Job {
    std::vector<uint16_t> frames;
    std::vector<int64_t> timestamps;
    bool close = false;
  };
  void EncoderTask();
  void EncodeJob(Job& job);
  BatchPtr BatchToFill();

  BatchProcessor batch_processor_;
  size_t frames_per_batch_, xsize_, ysize_, brotli_threads_;
  int shift_, device_;
  bool big_endian_, ok_ = true;
  SchemaPtr schema_;
  Job filling_;
  std::list<Job> queue_;
  std::mutex queue_mutex_;
  std::condition_variable queue_condition_;
  bool closing_ = false, closed_ = false;
  std::promise<int64_t> promised_closing_timestamp_;
  std::shared_future<int64_t> closing_timestamp_future_;
  int64_t latest_stored_timestamp_ = -1;
  std::mutex empty_mutex_;
  std::list<BatchPtr> empty_batches_;
  std::thread encoder_thread_;
};

class ColumnarBatchDecoder {
 public:
  ColumnarBatchDecoder(Image::Type type, bool unshift, ImageProcessor image_processor);
  ~ColumnarBatchDecoder();

  // Resolves to the batch once all its images went through the image processor; an invalid future if
  // the decoder is closing or the batch belongs to another schema (reference columnar_batch_decoder.cc:21-40).
  std::future<BatchPtr> PushBatch(BatchPtr batch);
  std::shared_future<int64_t> Close();

 private:
  struct Promised {
    BatchPtr batch;
    std::promise<BatchPtr> promise;
  };
  void DecoderTask();

  ImageProcessor image_processor_;
  Image::Type type_;
  bool unshift_;
  std::promise<int64_t> promised_closing_timestamp_;
  std::shared_future<int64_t> closing_timestamp_future_;
  std::list<Promised> batch_queue_;
  std::mutex queue_mutex_;
  std::condition_variable queue_condition_;
  bool closing_ = false;
  SchemaPtr schema_;
  int64_t latest_provided_timestamp_ = -1;
  std::thread decoder_thread_;
};

}  // namespace columnarbatch
}  // namespace fpvc

#endif  // FPV_B200_COLUMNAR_BATCH_H_

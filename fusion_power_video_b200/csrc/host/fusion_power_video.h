// fusion_power_video.h -- the fpvc:: stream codec API on top of the B200
// pre-entropy transform (libfpv_b200.so, include/fpv_b200.h).
//
// Drop-in for the reference's fusion_power_video.h: same namespace, class
// names, method signatures, callback types, ordering and error behaviour, so
// that benchmark.cc / encode.cc / decode.cc style callers recompile against
// this header unchanged (reference declarations: fusion_power_video.h:32-57
// UnextractFrame + StreamingDecoder, :143-172 RandomAccessDecoder, :175-211
// Encoder).  What differs is underneath:
//
//   * Frame ctor + Frame::Predict (reference .cc:1162-1164) run on the GPU,
//     batched, through fpv_encode_submit / fpv_wait on pinned double buffers;
//   * brotli (quality 1, lgwin 22, generic mode -- reference .cc:169, :653) stays
//     on host worker threads, one frame per task, fed from the pinned D2H
//     buffers; frames are emitted strictly in submission order;
//   * the post-brotli part of DecompressImage (reference .cc:326-344) runs on
//     the GPU through fpv_decode, batched over all frames that are complete in
//     the bytes handed to one Decode() call.
//
// The byte stream is the reference's (layout in DESIGN.md): streams written
// here decode with the reference decoder and vice versa, and with the same
// libbrotli the files are byte-identical.
//
// There is no CPU implementation of the transform in this library: without a
// CUDA device Encoder::Init / the decoders report failure (see fpvc::LastError).
#ifndef FPV_B200_FUSION_POWER_VIDEO_H_
#define FPV_B200_FUSION_POWER_VIDEO_H_

#include <stddef.h>
#include <stdint.h>

#include <functional>
#include <memory>
#include <string>
#include <vector>

namespace fpvc {

// Converts a decoded 16-bit frame back to the raw file format: value >> shift,
// two bytes per pixel in the file's endianness.  Host convenience kept for API
// compatibility (reference .cc:850-862); the decoders below can also deliver
// frames already in file format, converted on the GPU (SetRawOutput).
void UnextractFrame(const uint16_t* img, size_t xsize, size_t ysize, int shift, bool big_endian,
                    uint8_t* out);

// Description of the most recent failure on the calling thread ("" if none).
// Extension: the reference only prints to stderr (which this library does too).
const std::string& LastError();

// Extension: the encoders / decoders recycle their page-locked staging buffers through a process-wide cache (pinning
// memory costs about a millisecond per few MB), bounded by FPV_PIN_CACHE_MB (default 2048).  Long-lived processes
// that are done with a geometry call this to hand the cached blocks back to the driver; returns the bytes released.
size_t TrimPinnedCache();

// Tuning knobs (extension).  device: CUDA device index; batch: frames per GPU
// submission (the encoder submits earlier when its pipeline is idle).
struct GpuOptions {
  int device = 0;
  uint32_t batch = 8;
  // Entropy stage of the encoder: 0 = libbrotli quality 1 on host threads, byte-identical to the
  // reference's stream; 1 = the chunk-parallel coder on the GPU (valid brotli streams any
  // reference decoder reads, different bytes, no host brotli); -1 = 1 iff the environment
  // variable FPV_GPU_ENTROPY is set to a non-zero value.
  int gpu_entropy = -1;
  // Encoder only: several GPUs of one box behind ONE Encoder, the counterpart of the reference's one pool of
  // worker threads (reference .cc:1076-1084).  Batches go to the devices round-robin, the delta frame crosses
  // GPUs once by peer copy, emission order and stream bytes are those of a single GPU.  Empty = {device}.
  std::vector<int> devices;
};

// ---- fpvc::Frame ---------------------------------------------------------------------------------
// The reference's frame object (reference fusion_power_video.h:59-139): same enums, constructors, accessors and
// methods, so that code written against it -- the reference's columnar_batch/*.cc and arrow/*.cc wrappers --
// recompiles unchanged.  What differs is where the arithmetic runs:
//   * the u16 constructor splits the image into byte planes on the GPU (fpv_split; reference .cc:370-451),
//   * Predict() is ONE fused GPU call (fpv_encode: preview + delta + ClampedGradient with the reference's
//     integer heuristics; reference .cc:777-785),
//   * Uncompress() undoes the predictions on the GPU (fpv_unpredict_planes; reference .cc:595-641, :773-774),
//   * brotli (Compress / CompressPredicted / Uncompress) stays on the host, as in the Encoder.
// This is the single-frame compatibility surface: every call is a synchronous round trip of one frame over
// PCIe.  The batched, overlapped path is fpvc::Encoder / the decoders / columnarbatch::ColumnarBatchEncoder.
// There is no CPU implementation: without a CUDA device the GPU-backed methods report failure (LastError)
// and leave the frame unchanged.
// Restriction: Predict() needs a frame on which none of its three steps has run yet (the state every
// constructor and Uncompress() leave behind when all predictions were undone) or one on which all have
// (no-op, as in the reference); a frame with only some of the steps applied is refused.
enum FrameState {
  EMPTY = 0,
  RAW = 1,
  PREVIEW_GENERATED = 2,
  DELTA_PREDICTED = 4,
  CG_PREDICTED = 8,
  COMPRESSED = 16,
};

enum FrameFlags {
  NONE = 0,
  USE_DELTA = 1,
  USE_CG = 2,
  NO_LOW_BYTES = 4,
};

class Frame {
  size_t xsize_ = 0;
  size_t ysize_ = 0;
  size_t size_ = 0;
  uint8_t flags_ = FrameFlags::NONE;   // FrameFlags
  uint8_t state_ = FrameState::EMPTY;  // FrameState
  int64_t timestamp_ = -1;

 protected:
  std::vector<uint8_t> preview_;
  std::vector<uint8_t> high_;
  std::vector<uint8_t> low_;

 public:
  static Frame EMPTY;

  size_t xsize() const { return xsize_; }
  size_t ysize() const { return ysize_; }
  uint8_t flags() const { return flags_; }
  uint8_t state() const { return state_; }
  int64_t timestamp() const { return timestamp_; }
  const std::vector<uint8_t>& high() { return high_; }
  const std::vector<uint8_t>& low() { return low_; }
  const std::vector<uint8_t>& preview() { return preview_; }
  std::vector<uint8_t>&& MoveOutHigh() { Touch(); return std::move(high_); }
  std::vector<uint8_t>&& MoveOutLow() { Touch(); return std::move(low_); }
  std::vector<uint8_t>&& MoveOutPreview() { return std::move(preview_); }

  Frame(size_t xsize = 0, size_t ysize = 0, const uint16_t* image = nullptr, int shift_to_left_align = 0,
        bool big_endian = false, int64_t timestamp = -1);
  Frame(size_t xsize, size_t ysize, const uint8_t* image, int64_t timestamp = -1);
  Frame(size_t xsize, size_t ysize, uint8_t flags, uint8_t state, std::vector<uint8_t>&& high,
        std::vector<uint8_t>&& low, std::vector<uint8_t>&& preview, int64_t timestamp = -1);

  static size_t MaxCompressedPlaneSize(size_t xsize, size_t ysize);
  static size_t MaxCompressedPreviewSize(size_t xsize, size_t ysize);
  size_t MaxCompressedPlaneSize();
  size_t MaxCompressedPreviewSize();

  void Compress(Frame& delta_frame = EMPTY);
  void Uncompress(Frame& delta_frame = EMPTY);
  void Predict(Frame& delta_frame = EMPTY);
  void CompressPredicted(size_t* encoded_high_size, uint8_t* encoded_high_buffer, size_t* encoded_low_size,
                         uint8_t* encoded_low_buffer, size_t* encoded_preview_size, uint8_t* encoded_preview_buffer,
                         bool parallel = true);
  void OutputCore(std::vector<uint8_t>* out);
  void OutputFull(std::vector<uint8_t>* out);

  // Extension: CUDA device the Frame methods of this process run on (default 0).
  static void SetDevice(int device);

 private:
  void ApplyBrotliCompression();
  void Touch();   // the planes changed: a GPU context holding them as its delta frame must reload them
  // the image the u16 constructor received, kept until Predict() so that Predict is one fused GPU call on
  // the raw pixels (shared between copies of the frame)
  std::shared_ptr<const std::vector<uint16_t>> raw_;
  int shift_ = 0;
  bool big_endian_ = false;
  uint64_t generation_ = 0;   // identity of the current plane contents (see Touch)
  friend struct FrameGpu;
};

class StreamingDecoder {
 public:
  StreamingDecoder();
  explicit StreamingDecoder(const GpuOptions& options);
  ~StreamingDecoder();
  StreamingDecoder(const StreamingDecoder&) = delete;
  StreamingDecoder& operator=(const StreamingDecoder&) = delete;

  // Appends `size` bytes of the stream and calls `callback` once for every
  // frame that has become complete, in stream order, before returning.  `frame`
  // is valid only during the call.  On a malformed stream the callback is
  // invoked once with ok == false, frame == nullptr.
  void Decode(const uint8_t* bytes, size_t size,
              std::function<void(bool ok, uint16_t* frame, size_t xsize, size_t ysize, void* payload)>
                  callback,
              void* payload = nullptr);

  // Extension: deliver frames in raw file format instead (UnextractFrame fused
  // into the GPU kernel); `frame` then points at xsize * ysize * 2 file bytes.
  void SetRawOutput(int shift, bool big_endian);

 private:
  struct Impl;
  std::unique_ptr<Impl> impl_;
};

class RandomAccessDecoder {
 public:
  RandomAccessDecoder();
  explicit RandomAccessDecoder(const GpuOptions& options);
  ~RandomAccessDecoder();
  RandomAccessDecoder(const RandomAccessDecoder&) = delete;
  RandomAccessDecoder& operator=(const RandomAccessDecoder&) = delete;

  // Parses header, delta frame and footer.  `data` must stay valid for the
  // lifetime of the decoder.
  bool Init(const uint8_t* data, size_t size);

  // frame: xsize() * ysize() values.
  bool DecodeFrame(size_t index, uint16_t* frame) const;
  // Extension: `count` consecutive frames in one GPU batch.
  bool DecodeFrames(size_t first, size_t count, uint16_t* frames) const;
  // preview: preview_xsize() * preview_ysize() bytes.
  bool DecodePreview(size_t index, uint8_t* preview) const;

  size_t xsize() const;
  size_t ysize() const;
  size_t preview_xsize() const { return xsize() / 4; }
  size_t preview_ysize() const { return ysize() / 4; }
  size_t numframes() const;

 private:
  struct Impl;
  std::unique_ptr<Impl> impl_;
};

class Encoder {
 public:
  // num_threads host worker threads run brotli; 0 = everything happens inside
  // CompressFrame on the caller's thread.
  Encoder(size_t num_threads = 8, int shift_to_left_align = 0, bool big_endian = false);
  Encoder(size_t num_threads, int shift_to_left_align, bool big_endian, const GpuOptions& options);
  ~Encoder();
  Encoder(const Encoder&) = delete;
  Encoder& operator=(const Encoder&) = delete;

  typedef std::function<void(const uint8_t* compressed, size_t size, void* payload)> Callback;

  // Uploads the delta frame and emits the stream header through `callback`
  // (on the caller's thread).  delta_frame: xsize * ysize pixels.
  void Init(const uint16_t* delta_frame, size_t xsize, size_t ysize, Callback callback, void* payload);

  // Queues one frame.  Its bytes are emitted through `callback` from a worker
  // thread, serialised and in submission order.  The image is copied into
  // pinned staging memory before this returns, which is stricter than the
  // reference's contract (img valid until its callback; up to MaxQueued()
  // frames in flight).  Blocks while the pipeline is full.
  void CompressFrame(const uint16_t* img, Callback callback, void* payload);

  // Drains the pipeline, joins the workers and emits the footer (frame index).
  void Finish(Callback callback, void* payload);

  // As the reference: T + (T + 1) / 2, or 1 without threads.
  size_t MaxQueued() const;

  // Extension: false if Init failed (no CUDA device, unsupported geometry ...).
  bool ok() const;

 private:
  struct Impl;
  std::unique_ptr<Impl> impl_;
};

}  // namespace fpvc

#endif  // FPV_B200_FUSION_POWER_VIDEO_H_

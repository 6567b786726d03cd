/*
 * fpv_b200.h -- C ABI of the B200 (sm_100a) pre-entropy transform for
 * fusion-power-video streams.
 *
 * The reference (google/fusion-power-video) has no plugin/FFI interface: its
 * boundary is the C++ API in fusion_power_video.h.  This header is the thin
 * layer that API is re-plumbed onto.  Every entry point below replaces a
 * specific span of the reference's CPU code; citations are file:line into the
 * reference's fusion_power_video.cc (".cc") and fusion_power_video.h (".h").
 *
 * Conventions: extern "C", opaque handle, plain pointers and sizes, int status
 * (0 = FPV_OK).  No C++ types, no exceptions, no torch types cross this
 * boundary.  A context is bound to one CUDA device and one frame geometry.
 * Thread safety: every entry point takes the context's lock, so any number of
 * threads may share one context; the one blocking call, fpv_wait, releases the
 * lock while it sleeps on the slot's event, so one thread can submit into a
 * slot while another waits for a different slot (the Encoder's pipeline).  A
 * SLOT, however, belongs to one submit/wait sequence at a time: two threads
 * must not submit into the same slot without a wait in between.  Different
 * contexts are independent.  fpv_last_error is per calling thread.
 *
 * Layout of all buffers: frames are dense, frame-major.
 *   raw frames   uint16[n][ysize*xsize]   as read from the raw file (native
 *                                         little-endian load; big_endian says
 *                                         the file data is byte-swapped)
 *   high / low   uint8 [n][ysize*xsize]   byte planes, exactly the bytes
 *                                         Frame::high()/low() hold after
 *                                         Frame::Predict()
 *   preview      uint8 [n][(ysize/4)*(xsize/4)]
 *   flags        uint8 [n]                FrameFlags (.h:68-73): 1 USE_DELTA,
 *                                         2 USE_CG, 4 NO_LOW_BYTES
 *   images       uint16[n][ysize*xsize]   decoded 16-bit frames, as
 *                                         DecompressImage writes them
 */
#ifndef FPV_B200_H_
#define FPV_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FPV_OK 0
#define FPV_ERR_INVALID_ARG 1   /* null pointer, zero size, n > max_batch ...   */
#define FPV_ERR_CUDA 2          /* a CUDA runtime call failed; see last_error   */
#define FPV_ERR_UNSUPPORTED 3   /* geometry the reference itself mishandles     */
#define FPV_ERR_NO_DELTA 4      /* USE_DELTA frame but no delta frame was set   */
#define FPV_ERR_NO_DEVICE 5     /* no CUDA device / driver                      */

/* FrameFlags, .h:68-73 */
#define FPV_FLAG_USE_DELTA 1
#define FPV_FLAG_USE_CG 2
#define FPV_FLAG_NO_LOW_BYTES 4

/* encode options */
#define FPV_ENC_DEFAULT 0u
#define FPV_ENC_NO_DELTA 1u     /* predict with Frame::EMPTY (the delta frame's
                                   own header encoding, .cc:1099-1100)          */
#define FPV_ENC_GENERIC 2u      /* force the generic (non-TMA) kernels          */

/* decode options */
#define FPV_DEC_DEFAULT 0u
#define FPV_DEC_UNEXTRACT 1u    /* also apply UnextractFrame (.cc:850-862):
                                   output is the raw file bytes, not images     */

typedef struct fpv_ctx fpv_ctx;

/* ---- lifetime ---------------------------------------------------------- */

/* Creates a context on CUDA device `device` for xsize*ysize frames.
 * shift_to_left_align / big_endian: as Encoder's constructor (.h:179,
 * .cc:1076-1079) and UnextractFrame (.h:32-34).  max_batch bounds `n` of the
 * host-buffer calls (device staging is sized for it); device-pointer calls
 * accept any n. */
int fpv_create(fpv_ctx** ctx, int device, uint32_t xsize, uint32_t ysize,
               int shift_to_left_align, int big_endian, uint32_t max_batch);
void fpv_destroy(fpv_ctx* ctx);

/* Human-readable description of the last failure of the CALLING THREAD (on any
 * context, or in fpv_create).  The argument is accepted for symmetry and may
 * be NULL.  Never returns NULL; valid until the thread's next failing call. */
const char* fpv_last_error(const fpv_ctx* ctx);

/* Makes the context's device the calling thread's current CUDA device
 * (cudaSetDevice).  Host threads that allocate pinned memory with
 * fpv_host_alloc for a context on a device other than 0 call this first, so
 * that no primary context is created on device 0 behind their back. */
int fpv_bind_thread(const fpv_ctx* ctx);
int fpv_device_of(const fpv_ctx* ctx);

/* Library / device introspection. */
int fpv_device_count(void);
const char* fpv_version(void);
size_t fpv_plane_bytes(const fpv_ctx* ctx);    /* xsize*ysize            */
size_t fpv_preview_bytes(const fpv_ctx* ctx);  /* (xsize/4)*(ysize/4)    */
/* Number of kernels launched by this context so far (bench.py's
 * gpu_launches counts the difference across the timed region). */
uint64_t fpv_kernel_launches(const fpv_ctx* ctx);

/* Live roofline measurement (bench.py): when enabled, every launch of the
 * dominant kernel of a call (the fused encode kernel / the inverse kernel) is
 * bracketed by CUDA events on the launching stream.  fpv_read_kernel_timing
 * synchronises those events, returns the summed device time in milliseconds
 * and the number of launches it covers, and clears the record. */
int fpv_enable_kernel_timing(fpv_ctx* ctx, int on);
int fpv_read_kernel_timing(fpv_ctx* ctx, double* total_ms, uint32_t* launches);

/* Pinned host memory for the overlapped host<->device pipeline. */
void* fpv_host_alloc(size_t bytes);
void fpv_host_free(void* p);

/* ---- delta frame -------------------------------------------------------- */

/* Encoder side.  Replaces `delta_frame_ = Frame(xsize, ysize, delta_frame,
 * shift, big_endian)` in Encoder::Init (.cc:1097): splits the raw delta frame
 * on the device and keeps its high/low planes resident.  raw == NULL clears
 * the delta frame (Frame::EMPTY, .cc:780). */
int fpv_set_delta_raw(fpv_ctx* ctx, const uint16_t* raw_host);
int fpv_set_delta_raw_device(fpv_ctx* ctx, const void* raw_dev, void* stream);

/* Decoder side.  The delta frame as the decoded 16-bit image that
 * DecompressImage receives as `delta_frame` (.cc:296, filled at .cc:902-906 /
 * :983-987). */
int fpv_set_delta_image(fpv_ctx* ctx, const uint16_t* image_host);
int fpv_set_delta_image_device(fpv_ctx* ctx, const void* image_dev, void* stream);

/* Multi-GPU: copies the resident delta planes of `src` (another device) into
 * `dst` by peer copy -- the only inter-GPU traffic of the path (frames are
 * independent given the delta frame, .cc:36-38; the reference shares one
 * delta_frame_ between its pool threads, .cc:1097, :1164). */
int fpv_copy_delta_peer(fpv_ctx* dst, const fpv_ctx* src);

/* The same across PROCESSES (one process per GPU, e.g. torchrun): the owner
 * exports a CUDA IPC handle of its resident delta image, the handle bytes
 * travel over any control plane (a file, a pipe, torch.distributed), and each
 * importer maps the owner's memory and copies it device to device over
 * NVLink / PCIe peer access.  The owner's context must stay alive until every
 * importer has returned.  Fails with FPV_ERR_NO_DELTA if the owner has no
 * delta frame, FPV_ERR_CUDA if the handle is opened in the exporting process
 * (use fpv_copy_delta_peer there). */
#define FPV_IPC_HANDLE_BYTES 64
int fpv_delta_ipc_export(fpv_ctx* ctx, void* handle_out);
int fpv_delta_ipc_import(fpv_ctx* ctx, const void* handle);

/* ---- encode transform --------------------------------------------------- */

/* Replaces, for n frames at once, `Frame(xsize, ysize, img, shift,
 * big_endian)` + `Frame::Predict(delta_frame_)` as called from
 * Encoder::RunTask (.cc:1162-1164 -> :370-451, :777-785): split, preview,
 * optional delta prediction, optional ClampedGradient prediction (high plane
 * and preview), including the reference's integer heuristics (.cc:517-564)
 * bit for bit.  Outputs are exactly the bytes the Frame holds after Predict().
 * `low` may be NULL when shift_to_left_align == 8 (no low plane exists).
 * Requires xsize % 4 == 0 and ysize % 4 == 0 (the reference reads out of
 * bounds otherwise, .cc:577-578): FPV_ERR_UNSUPPORTED.
 *
 * Host-buffer form: synchronous; copies in, transforms, copies out. */
int fpv_encode(fpv_ctx* ctx, const uint16_t* frames_host, uint32_t n, uint32_t options,
               uint8_t* flags_host, uint8_t* high_host, uint8_t* low_host,
               uint8_t* preview_host);

/* Device-pointer form: all pointers are device memory on the context's
 * device; work is enqueued on `stream` (a cudaStream_t; NULL = the legacy
 * default stream) and not synchronised. */
int fpv_encode_device(fpv_ctx* ctx, const void* frames_dev, uint32_t n, uint32_t options,
                      void* flags_dev, void* high_dev, void* low_dev, void* preview_dev,
                      void* stream);

/* Overlapped form for the Encoder pipeline: `slot` (0 .. FPV_NUM_SLOTS - 1) selects one of
 * the staging sets with its own stream; submit returns after enqueueing
 * H2D copy + kernels + D2H copies, wait blocks until that slot's outputs
 * have landed in the (pinned) host buffers. */
#define FPV_NUM_SLOTS 4
int fpv_encode_submit(fpv_ctx* ctx, uint32_t slot, const uint16_t* frames_host, uint32_t n,
                      uint32_t options, uint8_t* flags_host, uint8_t* high_host,
                      uint8_t* low_host, uint8_t* preview_host);
int fpv_wait(fpv_ctx* ctx, uint32_t slot);

/* Frame constructor alone, for n frames: replaces `Frame(xsize, ysize, img, shift, big_endian)`
 * (.cc:370-451) without Predict -- the byte planes as Frame::high() / low() hold them in state RAW,
 * flags[i] = NO_LOW_BYTES iff every low byte of frame i is zero (always for shift 8, where `low` may be
 * NULL).  Host buffers, synchronous.  Any xsize, ysize >= 1. */
int fpv_split(fpv_ctx* ctx, const uint16_t* frames_host, uint32_t n, uint8_t* flags_host, uint8_t* high_host,
              uint8_t* low_host);

/* ---- optional GPU entropy coding ------------------------------------------------
 * Replaces the reference's three BrotliEncoderCompress calls per frame
 * (fusion_power_video.cc:643-688) and its OutputFull framing (.cc:830-846) by
 * a chunk-parallel coder on the device.  Every plane becomes a valid RFC 7932
 * (brotli) stream -- Huffman-coded literals only, independent byte-aligned
 * chunks of 64 KiB -- so the reference's decoder (BrotliDecoderDecompressStream,
 * .cc:186-214) reads the result unchanged; it is NOT byte-identical to what
 * libbrotli's encoder writes.  The output of n frames is their container chunks
 *   u32 total | u8 0 | u32 1+|bp| | u8 pflags | bp | u8 flags | low? | high
 * back to back; frame_off[i] is the byte offset of frame i, frame_off[n] the
 * total.  fpv_stream_bound(n) is the capacity the output buffer must have. */
size_t fpv_stream_bound(const fpv_ctx* ctx, uint32_t n);
/* planes already on the device (as fpv_encode_device leaves them); out_dev and
 * frame_off_dev (uint64[n + 1]) are device buffers; enqueued on `stream`. */
int fpv_entropy_device(fpv_ctx* ctx, const void* flags_dev, const void* high_dev, const void* low_dev,
                       const void* preview_dev, uint32_t n, void* out_dev, size_t capacity,
                       void* frame_off_dev, void* stream);
/* Host-buffer form: raw frames in, container chunks out (pinned buffers).
 * fpv_wait(slot) completes the call: it waits for the sizes, then fetches
 * exactly frame_off_host[n] bytes into out_host. */
int fpv_encode_stream_submit(fpv_ctx* ctx, uint32_t slot, const uint16_t* frames_host, uint32_t n,
                             uint32_t options, uint8_t* flags_host, uint64_t* frame_off_host,
                             uint8_t* out_host, size_t capacity);

/* Scatter forms of the two submit calls: frame_ptrs[i] points at frame i (each
 * xsize * ysize uint16 in host memory).  For callers whose frames already sit in
 * pinned memory of their own (camera DMA buffers registered with
 * cudaHostRegister, or fpv_host_alloc blocks): nothing is copied on the host, each
 * frame goes to the device straight from where it is, and -- as the reference's
 * CompressFrame requires of its callers (fusion_power_video.h:197-199) -- it must
 * stay valid until fpv_wait(slot) returns.  Frames that are neighbours in memory
 * are uploaded with one copy.  fpv_host_is_pinned tells whether a host pointer is
 * page-locked memory CUDA knows about. */
int fpv_encode_submit_v(fpv_ctx* ctx, uint32_t slot, const uint16_t* const* frame_ptrs, uint32_t n,
                        uint32_t options, uint8_t* flags_host, uint8_t* high_host, uint8_t* low_host,
                        uint8_t* preview_host);
int fpv_encode_stream_submit_v(fpv_ctx* ctx, uint32_t slot, const uint16_t* const* frame_ptrs, uint32_t n,
                               uint32_t options, uint8_t* flags_host, uint64_t* frame_off_host,
                               uint8_t* out_host, size_t capacity);
int fpv_host_is_pinned(const void* p);

/* ---- decode (inverse) transform ---------------------------------------- */

/* Replaces the post-brotli part of DecompressImage (.cc:326-344) for n frames
 * at once: inverse ClampedGradient on the high plane (flags & 2), delta add
 * with independent byte wrap (flags & 1), recombination to uint16; low bytes
 * are taken as zero for frames with flags & 4 (their `low` slot is not read;
 * `low` may be NULL if every frame has flags & 4).  With FPV_DEC_UNEXTRACT the
 * output is additionally passed through UnextractFrame (.cc:850-862) with the
 * context's shift/big_endian, i.e. `out` receives the raw file bytes.
 * Any xsize, ysize >= 1. */
int fpv_decode(fpv_ctx* ctx, const uint8_t* high_host, const uint8_t* low_host,
               const uint8_t* flags_host, uint32_t n, uint32_t options, void* out_host);
/* Device-pointer form.  The flags live on the device, so this form cannot
 * refuse a USE_DELTA frame when no delta frame is set (FPV_ERR_NO_DELTA is
 * reported by the host-buffer forms only): such a frame is decoded WITHOUT
 * the delta add.  Callers that build flags themselves must set the delta
 * frame first, as the reference's decoders do (.cc:902-906). */
int fpv_decode_device(fpv_ctx* ctx, const void* high_dev, const void* low_dev,
                      const void* flags_dev, uint32_t n, uint32_t options, void* out_dev,
                      void* stream);
int fpv_decode_submit(fpv_ctx* ctx, uint32_t slot, const uint8_t* high_host,
                      const uint8_t* low_host, const uint8_t* flags_host, uint32_t n,
                      uint32_t options, void* out_host);

/* Frames whose plane streams were written by this library's GPU entropy coder
 * (fpv_encode_stream_submit / fpv_entropy_device): every 64 KiB chunk of such a
 * stream starts with a metadata meta-block -- skipped by brotli decoders, the
 * reference's included -- that carries a directory (chunk size, code lengths,
 * bit positions of 2048-byte spans; layout in csrc/fpv_internal.h).  This call
 * replaces BrotliDecompress (.cc:186-214) + the post-brotli part of
 * DecompressImage (.cc:326-344) for n frames: the coded bytes are uploaded as
 * they are (about half the plane bytes), decoded by one warp per chunk into
 * the context's plane buffers and inverted like fpv_decode.  `blob_host` holds
 * the coded bytes of all frames; `chunks_host` names every chunk of every plane
 * a frame needs (no low-plane chunks for frames with flags & 4): its offset in
 * the blob, the frame, the plane (0 = high, 1 = low) and its index inside the
 * plane (plane bytes [index * 65536, ...)).  The host side finds the offsets by
 * walking the directories (each holds its chunk's size).  Synchronous.
 * Streams libbrotli wrote carry no directory and go through fpv_decode. */
typedef struct fpv_coded_chunk {
  uint64_t offset;
  uint32_t frame;
  uint32_t plane;
  uint32_t index;
  uint32_t reserved;
} fpv_coded_chunk;
int fpv_decode_coded(fpv_ctx* ctx, const uint8_t* blob_host, size_t blob_bytes,
                     const fpv_coded_chunk* chunks_host, uint32_t n_chunks,
                     const uint8_t* flags_host, uint32_t n, uint32_t options, void* out_host);

/* Plane-level inverse, replacing Frame::Uncompress's prediction undo
 * (.cc:773-774 -> :612-641, :595-610): inverse CG on high and preview, delta
 * add on both byte planes; planes are rewritten in place (host buffers).
 * preview may be NULL. */
int fpv_unpredict_planes(fpv_ctx* ctx, uint8_t* high_host, uint8_t* low_host,
                         uint8_t* preview_host, const uint8_t* flags_host, uint32_t n);

#ifdef __cplusplus
}
#endif

#endif  /* FPV_B200_H_ */

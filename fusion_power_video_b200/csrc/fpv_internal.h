// fpv_internal.h -- launcher interface between the C ABI (fpv_cabi.cu) and the
// kernels (fpv_encode.cu, fpv_decode.cu).  Not installed; not part of the ABI.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "fpv_common.cuh"

namespace fpv {

struct Geom {
  uint32_t W = 0, H = 0;
  uint64_t P = 0;        // W * H
  uint32_t PW = 0;       // W / 4
  uint64_t PP = 0;       // (W/4) * (H/4)
  int shift = 0;
  int big_endian = 0;
  int mode = 0;          // SplitMode
};

// Device scratch for one encode pass over up to `cap` frames.
struct EncodeScratch {
  FrameStat* stats = nullptr;     // [cap]
  uint32_t* lists = nullptr;      // [3][cap] frame indices for pass 0,1,2
  uint32_t* counts = nullptr;     // [5]: [1], [2] lengths of the redo lists, [3], [4] flags guess for the next batch
                                  // (two words used alternately, see FastParams::guess_in / guess_out)
  uint8_t* preview_raw = nullptr; // [cap][PP]  (generic path only; allocated on its first use)
  uint32_t cap = 0;
  uint64_t P_PP = 0;              // bytes of one frame's preview plane (for the lazy allocation above)
  // host-side state
  uint32_t calls = 0;             // fast-path calls so far: parity selects the guess word
  bool dirty = false;             // the generic path left statistics behind; the fast path expects them zeroed
};

struct EncodeTuning {
  int num_sms = 148;
  int stages = 2;          // smem ring depth of the fast kernel (2 measured best: occupancy wins)
  int rows_per_stage = 4;  // 4 or 2
  int band_rows = 128;     // rows per task (multiple of 4)
  int max_ctas = 4;        // cap on resident CTAs per SM of the fast kernel (0 = what fits); a fifth CTA costs 8 points at W = 1024
  int max_smem_optin = 0;  // bytes
};

// Optional pair of events recorded around the dominant kernel of a call.
struct TimingHook {
  cudaEvent_t start = nullptr;
  cudaEvent_t stop = nullptr;
};

// True if the TMA fast path supports this geometry.
bool encode_fast_supported(const Geom& g, const EncodeTuning& t);

// Enqueues the whole encode transform for n <= scratch.cap frames on `stream`.
// frames: uint16[n][P] device; delta: uint16[P] image-form device or nullptr.
// Outputs: flags[n], high[n][P], low[n][P] (nullptr iff mode has no low
// plane), preview[n][PP].  Returns the number of kernels launched, or -1 with
// *err set.
int enqueue_encode(const Geom& g, const EncodeTuning& t, EncodeScratch& s,
                   const uint16_t* frames, const uint16_t* delta, uint32_t n, bool force_generic,
                   uint8_t* flags, uint8_t* high, uint8_t* low, uint8_t* preview,
                   cudaStream_t stream, cudaError_t* err, const TimingHook* hook = nullptr);

// ---- GPU entropy coder (fpv_entropy.cu) -----------------------------------------------------
constexpr uint32_t kEntropyChunk = 65536;                 // plane bytes per independently coded chunk
// Every coded chunk starts with a metadata meta-block (skipped by brotli decoders) that carries a directory for
// k_entropy_decode: 2 header bytes, then 'F' 'D' | version | kind | chunk bytes u24 | n - 1 u16 | constant symbol |
// 256 code lengths as nibbles | 32 span bit positions u24 (tests/huffcoder_ref.py states the same layout).
constexpr uint32_t kDirBytes = 234, kDirBlock = 2 + kDirBytes, kDirBits = 8 * kDirBlock;
constexpr uint32_t kDirLengths = 10, kDirSpans = 138;       // payload offsets
constexpr uint32_t kEntropySpan = 2048;                      // plane bytes one decoder lane reconstructs
constexpr uint32_t kKindHuffman = 0, kKindRaw = 1, kKindConstant = 2;
constexpr uint32_t kEntropyChunkCap = kEntropyChunk + kDirBlock + 128;  // scratch bytes per chunk (directory + <= n + 5 bytes)

struct EntropyParams {
  const uint8_t* high;      // [n][P]
  const uint8_t* low;       // [n][P] or nullptr
  const uint8_t* preview;   // [n][PP]
  const uint8_t* flags;     // [n]
  uint64_t P, PP;
  uint32_t n;
  uint32_t cpl, cpp, cpf;   // chunks per full plane, per preview plane, per frame (preview, low, high)
  uint8_t* scratch;         // [n * cpf][kEntropyChunkCap]
  uint32_t* chunk_bytes;    // [n * cpf]
};

// Enqueues chunk coding, layout and gather for n frames; frame_off: uint64[n + 1] (device), out: the
// container chunks of the n frames back to back (device).  Returns kernels launched or -1.
// One chunk of a directory-carrying plane stream (the device-side twin of fpv_coded_chunk).
struct CodedChunk {
  uint64_t offset;   // of the chunk's first byte inside the blob
  uint32_t frame;    // frame index in the batch
  uint32_t plane;    // 0 = high, 1 = low
  uint32_t index;    // chunk index inside the plane: plane bytes [index * kEntropyChunk, ...)
  uint32_t reserved;
};
struct EntropyDecodeParams {
  const uint8_t* blob;
  uint64_t blob_bytes;
  const CodedChunk* chunks;
  uint32_t n_chunks, n_frames;
  uint32_t frame0;   // chunk frame indices are relative to the whole batch: this launch covers [frame0, frame0 + n_frames)
  uint8_t* high;     // [n_frames][P]
  uint8_t* low;      // [n_frames][P] or nullptr
  uint64_t P;
  uint32_t* err;     // set to a non-zero reason by a malformed chunk
};
int enqueue_entropy_decode(const EntropyDecodeParams& p, cudaStream_t stream, cudaError_t* err);
int enqueue_entropy(const EntropyParams& p, uint64_t* frame_off, uint8_t* out, uint64_t capacity, uint32_t* overflow,
                    cudaStream_t stream, cudaError_t* err);

// Frame ctor only: byte planes of n raw frames, low_or[f] = OR of frame f's low bytes (uint32[n], device).
int enqueue_split(const Geom& g, const uint16_t* frames, uint32_t n, uint8_t* high, uint8_t* low, uint32_t* low_or,
                  cudaStream_t stream, cudaError_t* err);

// Splits a raw delta frame into image form ((high << 8) | low per pixel).
int enqueue_delta_from_raw(const Geom& g, const uint16_t* raw, uint16_t* delta_image,
                           cudaStream_t stream, cudaError_t* err);

// Inverse transform for n frames (any n).  out: uint16[n][P] images, or raw
// file bytes when `unextract`.
// `ddup` is the delta image in duplicated form ((d | d << 16) per pixel, see
// enqueue_delta_dup); without it the pair kernel is not used for delta streams.
// Returns the number of kernels launched, -1 on a CUDA error (*err), -2 if no row-pipelined kernel
// takes this geometry (rows too wide for shared memory: the caller uses enqueue_decode_serial).
int enqueue_decode(const Geom& g, int num_sms, const uint8_t* high, const uint8_t* low,
                   const uint8_t* flags, const uint16_t* delta, const uint32_t* ddup, uint32_t n,
                   bool unextract, uint16_t* out, cudaStream_t stream, cudaError_t* err,
                   const TimingHook* hook = nullptr);

// delta image -> pair form for the pair decode kernel (padded, permuted rows; delta_dup_bytes(g) bytes).
size_t delta_dup_bytes(const Geom& g);
int enqueue_delta_dup(const Geom& g, const uint16_t* delta_image, uint32_t* ddup, cudaStream_t stream,
                      cudaError_t* err);

// Serial fallback (one thread per frame runs the chain literally); used for
// rows too wide for the shared-memory row pipeline and as a cross-check.
// `scratch_high` is a WRITABLE copy of the high planes (inverse CG is applied
// in place).
int enqueue_decode_serial(const Geom& g, uint8_t* scratch_high, const uint8_t* low,
                          const uint8_t* flags, const uint16_t* delta, uint32_t n, bool unextract,
                          uint16_t* out, cudaStream_t stream, cudaError_t* err);

// Plane-level inverse (Frame::Uncompress): inverse CG in place on a plane of
// row pitch W and n_px bytes per frame, for frames with flags & 2; then
// (optionally) delta add on high/low planes.
int enqueue_unpredict_planes(const Geom& g, int num_sms, uint8_t* high, uint8_t* low,
                             uint8_t* preview, const uint8_t* flags, const uint16_t* delta,
                             uint32_t n, cudaStream_t stream, cudaError_t* err);

}  // namespace fpv

// fpv_cabi.cu -- implementation of the C ABI declared in include/fpv_b200.h.
//
// Owns: the CUDA context state for one device + one frame geometry, the
// resident delta frame (image form), per-slot device staging and streams for
// the host-buffer entry points, and the scratch the encode kernels need.
// There is deliberately NO CPU fallback here: every entry point either runs
// the CUDA kernels or returns an error code with a message.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/fpv_b200.h"
#include "fpv_internal.h"

using namespace fpv;

namespace {

constexpr int kNumSlots = 4;       // host-pipeline slots (FPV_NUM_SLOTS in the header); staging is allocated on first use
constexpr int kDeviceScratch = kNumSlots;  // scratch index used by the *_device entry points

// Last failure of the calling thread (fpv_last_error): per thread, so that threads sharing a context
// never read each other's half-written message.
thread_local std::string tl_error = "no error";

struct Slot {
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;    // blocking-sync event: fpv_wait sleeps instead of spinning on a host core
  uint16_t* d_frames = nullptr;  // raw frames in / decoded images out
  uint8_t* d_high = nullptr;
  uint8_t* d_low = nullptr;
  uint8_t* d_preview = nullptr;
  uint8_t* d_flags = nullptr;
  // fpv_decode_coded: the coded bytes, their chunk table and the decoder's error word (allocated on first use)
  uint8_t* d_coded = nullptr;
  size_t coded_cap = 0;
  CodedChunk* d_chunks = nullptr;
  size_t chunks_cap = 0;
  uint32_t* d_dec_err = nullptr;
  uint32_t* h_dec_err = nullptr;   // pinned: a pageable destination would make the download synchronous
  bool allocated = false;
  // a pending fpv_encode_stream_submit: fpv_wait fetches the coded bytes once their number is known
  uint8_t* stream_out_host = nullptr;
  uint64_t* stream_off_host = nullptr;
  uint32_t stream_n = 0;
};

// Device buffers of the GPU entropy coder for up to max_batch frames.
struct EntropyBuf {
  uint8_t* scratch = nullptr;      // [cap * cpf][kEntropyChunkCap]
  uint32_t* chunk_bytes = nullptr; // [cap * cpf]
  uint64_t* frame_off = nullptr;   // [cap + 1]
  uint8_t* out = nullptr;          // container chunks, stream_bound(cap) bytes (host-buffer entry points only)
  uint32_t* overflow = nullptr;
};

}  // namespace

struct fpv_ctx {
  int device = 0;
  Geom g;
  EncodeTuning tune;
  uint32_t max_batch = 0;
  bool encode_ok = false;        // W % 4 == 0 && H % 4 == 0
  uint16_t* d_delta = nullptr;   // image form, P pixels
  uint32_t* d_delta_dup = nullptr;  // (d | d << 16) per pixel, for the pair decode kernel
  bool has_delta = false;
  EncodeScratch scratch[kNumSlots + 1];
  Slot slots[kNumSlots];
  EntropyBuf entropy[kNumSlots + 1];
  uint8_t* d_serial_scratch = nullptr;
  size_t serial_scratch_bytes = 0;
  cudaStream_t aux_stream = nullptr;
  uint64_t launches = 0;
  bool force_generic = false;
  bool timing_on = false;
  std::vector<TimingHook> timing;   // event pairs, reused
  size_t timing_used = 0;
  // Every entry point holds this while it touches the context; fpv_wait drops it while it sleeps.
  // Recursive: the synchronous entry points are built from submit + wait.
  std::recursive_mutex mu;
};

#define FPV_LOCK(c) std::lock_guard<std::recursive_mutex> lock__((c)->mu)

namespace {

int fail(fpv_ctx* c, int code, const std::string& msg) {
  (void)c;
  tl_error = msg;
  return code;
}

int cuda_fail(fpv_ctx* c, cudaError_t e, const char* what) {
  return fail(c, FPV_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define FPV_CUDA(call)                                          \
  do {                                                          \
    cudaError_t e__ = (call);                                   \
    if (e__ != cudaSuccess) return cuda_fail(c, e__, #call);    \
  } while (0)

// Next event pair for the dominant kernel, or nullptr when timing is off.
const TimingHook* next_hook(fpv_ctx* c) {
  if (!c->timing_on) return nullptr;
  if (c->timing_used == c->timing.size()) {
    TimingHook h;
    if (cudaEventCreate(&h.start) != cudaSuccess || cudaEventCreate(&h.stop) != cudaSuccess) return nullptr;
    c->timing.push_back(h);
  }
  return &c->timing[c->timing_used++];
}

int ensure_scratch(fpv_ctx* c, int idx) {
  EncodeScratch& s = c->scratch[idx];
  if (s.stats) return FPV_OK;
  uint32_t cap = c->max_batch;
  FPV_CUDA(cudaMalloc(&s.stats, sizeof(FrameStat) * (size_t)cap));
  FPV_CUDA(cudaMemset(s.stats, 0, sizeof(FrameStat) * (size_t)cap));   // the fast path keeps them zero between calls
  FPV_CUDA(cudaMalloc(&s.lists, sizeof(uint32_t) * 3 * (size_t)cap));
  FPV_CUDA(cudaMalloc(&s.counts, sizeof(uint32_t) * 5));
  // (preview_raw: scratch of the generic path only, allocated by enqueue_encode when that path first runs)
  uint32_t init[5] = {0, 0, 0, 3, 3};  // first guess: USE_DELTA | USE_CG
  FPV_CUDA(cudaMemcpy(s.counts, init, sizeof init, cudaMemcpyHostToDevice));
  s.cap = cap;
  return FPV_OK;
}

int ensure_slot(fpv_ctx* c, int idx) {
  Slot& s = c->slots[idx];
  if (s.allocated) return FPV_OK;
  size_t B = c->max_batch, P = c->g.P, PP = c->g.PP ? c->g.PP : 1;
  FPV_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
  FPV_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventBlockingSync | cudaEventDisableTiming));
  FPV_CUDA(cudaMalloc(&s.d_frames, B * P * 2));
  FPV_CUDA(cudaMalloc(&s.d_high, B * P));
  FPV_CUDA(cudaMalloc(&s.d_low, B * P));
  FPV_CUDA(cudaMalloc(&s.d_preview, B * PP));
  FPV_CUDA(cudaMalloc(&s.d_flags, B));
  s.allocated = true;
  return FPV_OK;
}

uint32_t chunks_of(uint64_t bytes) { return (uint32_t)((bytes + kEntropyChunk - 1) / kEntropyChunk); }

// Upper bound of the container chunks of n frames coded by the GPU entropy coder: per frame 11
// container bytes and the three planes, per coded chunk at most 5 bytes of meta-block framing, per
// plane stream the WBITS bit and the final 0x03.
size_t stream_bound(const fpv_ctx* c, uint32_t n) {
  const uint64_t cpf = chunks_of(c->g.PP) + 2ull * chunks_of(c->g.P);
  return (size_t)n * (size_t)(11 + c->g.PP + 2 * c->g.P + (8 + kDirBlock) * cpf + 16);
}

int ensure_entropy(fpv_ctx* c, int idx, bool with_out) {
  EntropyBuf& e = c->entropy[idx];
  const uint64_t cap = c->max_batch, cpf = chunks_of(c->g.PP) + 2ull * chunks_of(c->g.P);
  if (!e.scratch) {
    FPV_CUDA(cudaMalloc(&e.scratch, cap * cpf * kEntropyChunkCap));
    FPV_CUDA(cudaMalloc(&e.chunk_bytes, cap * cpf * sizeof(uint32_t)));
    FPV_CUDA(cudaMalloc(&e.frame_off, (cap + 1) * sizeof(uint64_t)));
    FPV_CUDA(cudaMalloc(&e.overflow, sizeof(uint32_t)));
    FPV_CUDA(cudaMemset(e.overflow, 0, sizeof(uint32_t)));
  }
  if (with_out && !e.out) FPV_CUDA(cudaMalloc(&e.out, stream_bound(c, c->max_batch)));
  return FPV_OK;
}

// Entropy-codes n <= max_batch frames whose planes are on the device.
int entropy_device_impl(fpv_ctx* c, int idx, const uint8_t* flags, const uint8_t* high, const uint8_t* low,
                        const uint8_t* preview, uint32_t n, uint8_t* out, uint64_t capacity, uint64_t* frame_off,
                        cudaStream_t stream) {
  EntropyBuf& e = c->entropy[idx];
  EntropyParams p;
  p.high = high; p.low = mode_has_low(c->g.mode) ? low : nullptr; p.preview = preview; p.flags = flags;
  p.P = c->g.P; p.PP = c->g.PP; p.n = n;
  p.cpl = chunks_of(c->g.P); p.cpp = chunks_of(c->g.PP); p.cpf = p.cpp + 2 * p.cpl;
  p.scratch = e.scratch; p.chunk_bytes = e.chunk_bytes;
  cudaError_t err = cudaSuccess;
  int l = enqueue_entropy(p, frame_off, out, capacity, e.overflow, stream, &err);
  if (l < 0) return cuda_fail(c, err, "entropy kernel launch");
  c->launches += (uint64_t)l;
  return FPV_OK;
}

int encode_device_chunked(fpv_ctx* c, int scratch_idx, const uint16_t* frames, uint32_t n,
                          uint32_t options, uint8_t* flags, uint8_t* high, uint8_t* low,
                          uint8_t* preview, cudaStream_t stream) {
  if (!c->encode_ok)
    return fail(c, FPV_ERR_UNSUPPORTED,
                "encode requires xsize % 4 == 0 and ysize % 4 == 0 (the reference reads out of "
                "bounds otherwise, fusion_power_video.cc:577-578) and shift <= 8 for big-endian data");
  if (mode_has_low(c->g.mode) && !low) return fail(c, FPV_ERR_INVALID_ARG, "low plane buffer is NULL");
  int rc = ensure_scratch(c, scratch_idx);
  if (rc != FPV_OK) return rc;
  const uint16_t* delta = (c->has_delta && !(options & FPV_ENC_NO_DELTA)) ? c->d_delta : nullptr;
  const bool generic = c->force_generic || (options & FPV_ENC_GENERIC);
  const uint64_t P = c->g.P, PP = c->g.PP;
  for (uint32_t off = 0; off < n; off += c->max_batch) {
    uint32_t m = n - off < c->max_batch ? n - off : c->max_batch;
    cudaError_t e = cudaSuccess;
    int l = enqueue_encode(c->g, c->tune, c->scratch[scratch_idx], frames + (uint64_t)off * P, delta, m,
                           generic, flags + off, high + (uint64_t)off * P,
                           low ? low + (uint64_t)off * P : nullptr, preview + (uint64_t)off * PP, stream, &e,
                           next_hook(c));
    if (l < 0) return cuda_fail(c, e, "encode kernel launch");
    c->launches += (uint64_t)l;
  }
  return FPV_OK;
}

int decode_device_impl(fpv_ctx* c, const uint8_t* high, const uint8_t* low, const uint8_t* flags,
                       uint32_t n, uint32_t options, uint16_t* out, cudaStream_t stream,
                       bool high_is_scratch) {
  cudaError_t e = cudaSuccess;
  const uint16_t* delta = c->has_delta ? c->d_delta : nullptr;
  const bool unextract = (options & FPV_DEC_UNEXTRACT) != 0;
  int l = -2;
  if (!getenv("FPV_DECODE_SERIAL"))
    l = enqueue_decode(c->g, c->tune.num_sms, high, low, flags, delta, c->d_delta_dup, n, unextract, out,
                       stream, &e, next_hook(c));
  if (l == -1) return cuda_fail(c, e, "decode kernel launch");
  if (l < 0) {
    // rows too wide for the shared-memory row pipeline (or forced): serial chain
    uint8_t* scratch = const_cast<uint8_t*>(high);
    if (!high_is_scratch) {
      size_t need = (size_t)n * c->g.P;
      if (need > c->serial_scratch_bytes) {
        if (c->d_serial_scratch) cudaFree(c->d_serial_scratch);
        c->d_serial_scratch = nullptr;
        c->serial_scratch_bytes = 0;
        FPV_CUDA(cudaMalloc(&c->d_serial_scratch, need));
        c->serial_scratch_bytes = need;
      }
      FPV_CUDA(cudaMemcpyAsync(c->d_serial_scratch, high, need, cudaMemcpyDeviceToDevice, stream));
      scratch = c->d_serial_scratch;
    }
    l = enqueue_decode_serial(c->g, scratch, low, flags, delta, n, unextract, out, stream, &e);
    if (l < 0) return cuda_fail(c, e, "decode kernel launch");
  }
  c->launches += (uint64_t)l;
  return FPV_OK;
}

}  // namespace

extern "C" {

const char* fpv_version(void) { return "fpv_b200 0.1 (sm_100a)"; }

int fpv_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int fpv_create(fpv_ctx** out, int device, uint32_t xsize, uint32_t ysize, int shift, int big_endian,
               uint32_t max_batch) {
  fpv_ctx* c = nullptr;  // for the FPV_CUDA macro: errors before allocation go to the global slot
  if (!out) return fail(nullptr, FPV_ERR_INVALID_ARG, "ctx out pointer is NULL");
  *out = nullptr;
  if (xsize == 0 || ysize == 0 || xsize > 65536 || ysize > 65536 ||
      (uint64_t)xsize * ysize > 1000000000ull)
    return fail(nullptr, FPV_ERR_INVALID_ARG, "invalid image dimensions");  // .cc:891-895
  int mode = pick_split_mode(shift, big_endian);
  // Big-endian data with shift > 8: the reference's Frame constructor shifts by 8 - shift there (undefined), but its
  // UnextractFrame (.cc:850-862) is fine, so such a context decodes and refuses to encode.
  const bool decode_only = mode < 0 && big_endian && shift > 8 && shift <= 16;
  if (decode_only) mode = kBE0;
  if (mode < 0)
    return fail(nullptr, FPV_ERR_UNSUPPORTED,
                "shift must be 0..16 (0..8 to ENCODE big-endian data: the reference shifts by 8 - shift)");
  if (max_batch == 0 || max_batch > 65535)
    return fail(nullptr, FPV_ERR_INVALID_ARG, "max_batch must be 1..65535 (frames are a grid dimension of the small kernels)");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, FPV_ERR_NO_DEVICE,
                std::string("no CUDA device available: ") + cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(nullptr, FPV_ERR_INVALID_ARG, "device index out of range");
  FPV_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  FPV_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(nullptr, FPV_ERR_NO_DEVICE, "device is not sm_100-class (kernels are built for sm_100a only)");

  c = new (std::nothrow) fpv_ctx();
  if (!c) return fail(nullptr, FPV_ERR_INVALID_ARG, "out of host memory");
  c->device = device;
  c->g.W = xsize; c->g.H = ysize; c->g.P = (uint64_t)xsize * ysize;
  c->g.PW = xsize / 4; c->g.PP = (uint64_t)(xsize / 4) * (ysize / 4);
  c->g.shift = shift; c->g.big_endian = big_endian ? 1 : 0; c->g.mode = mode;
  c->max_batch = max_batch;
  c->encode_ok = (xsize % 4 == 0) && (ysize % 4 == 0) && !decode_only;
  c->tune.num_sms = prop.multiProcessorCount;
  c->tune.max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  if (const char* v = getenv("FPV_STAGES")) c->tune.stages = atoi(v);
  if (const char* v = getenv("FPV_BAND_ROWS")) c->tune.band_rows = atoi(v);
  if (const char* v = getenv("FPV_ROWS_PER_STAGE")) c->tune.rows_per_stage = atoi(v) == 2 ? 2 : 4;
  if (const char* v = getenv("FPV_MAX_CTAS")) c->tune.max_ctas = atoi(v);
  if (c->tune.stages < 2) c->tune.stages = 2;
  if (c->tune.stages > 8) c->tune.stages = 8;
  if (c->tune.band_rows < 4) c->tune.band_rows = 4;
  c->force_generic = getenv("FPV_FORCE_GENERIC") != nullptr;
  cudaError_t e2 = cudaMalloc(&c->d_delta, c->g.P * 2);
  if (e2 == cudaSuccess) e2 = cudaMalloc(&c->d_delta_dup, delta_dup_bytes(c->g));
  if (e2 == cudaSuccess) e2 = cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking);
  if (e2 != cudaSuccess) {
    int rc = cuda_fail(nullptr, e2, "fpv_create allocation");
    fpv_destroy(c);
    return rc;
  }
  *out = c;
  return FPV_OK;
}

void fpv_destroy(fpv_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (auto& s : c->scratch) {
    if (s.stats) cudaFree(s.stats);
    if (s.lists) cudaFree(s.lists);
    if (s.counts) cudaFree(s.counts);
    if (s.preview_raw) cudaFree(s.preview_raw);
  }
  for (auto& s : c->slots) {
    if (s.d_frames) cudaFree(s.d_frames);
    if (s.d_high) cudaFree(s.d_high);
    if (s.d_low) cudaFree(s.d_low);
    if (s.d_preview) cudaFree(s.d_preview);
    if (s.d_flags) cudaFree(s.d_flags);
    if (s.d_coded) cudaFree(s.d_coded);
    if (s.d_chunks) cudaFree(s.d_chunks);
    if (s.d_dec_err) cudaFree(s.d_dec_err);
    if (s.h_dec_err) cudaFreeHost(s.h_dec_err);
    if (s.done) cudaEventDestroy(s.done);
    if (s.stream) cudaStreamDestroy(s.stream);
  }
  for (auto& e : c->entropy) {
    if (e.scratch) cudaFree(e.scratch);
    if (e.chunk_bytes) cudaFree(e.chunk_bytes);
    if (e.frame_off) cudaFree(e.frame_off);
    if (e.out) cudaFree(e.out);
    if (e.overflow) cudaFree(e.overflow);
  }
  for (auto& h : c->timing) {
    if (h.start) cudaEventDestroy(h.start);
    if (h.stop) cudaEventDestroy(h.stop);
  }
  if (c->d_serial_scratch) cudaFree(c->d_serial_scratch);
  if (c->d_delta) cudaFree(c->d_delta);
  if (c->d_delta_dup) cudaFree(c->d_delta_dup);
  if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
  delete c;
}

const char* fpv_last_error(const fpv_ctx* c) {
  (void)c;
  return tl_error.c_str();
}

int fpv_bind_thread(const fpv_ctx* c) {
  if (!c) return FPV_ERR_INVALID_ARG;
  return cudaSetDevice(c->device) == cudaSuccess ? FPV_OK : FPV_ERR_CUDA;
}

int fpv_device_of(const fpv_ctx* c) { return c ? c->device : -1; }

size_t fpv_plane_bytes(const fpv_ctx* c) { return c ? (size_t)c->g.P : 0; }
size_t fpv_preview_bytes(const fpv_ctx* c) { return c ? (size_t)c->g.PP : 0; }
uint64_t fpv_kernel_launches(const fpv_ctx* c) {
  if (!c) return 0;
  std::lock_guard<std::recursive_mutex> l(const_cast<fpv_ctx*>(c)->mu);
  return c->launches;
}

int fpv_enable_kernel_timing(fpv_ctx* c, int on) {
  if (!c) return FPV_ERR_INVALID_ARG;
  FPV_LOCK(c);
  c->timing_on = on != 0;
  c->timing_used = 0;
  return FPV_OK;
}

int fpv_read_kernel_timing(fpv_ctx* c, double* total_ms, uint32_t* launches) {
  if (!c || !total_ms || !launches) return FPV_ERR_INVALID_ARG;
  FPV_LOCK(c);
  FPV_CUDA(cudaSetDevice(c->device));
  double total = 0;
  for (size_t i = 0; i < c->timing_used; i++) {
    FPV_CUDA(cudaEventSynchronize(c->timing[i].stop));
    float ms = 0;
    FPV_CUDA(cudaEventElapsedTime(&ms, c->timing[i].start, c->timing[i].stop));
    total += ms;
  }
  *total_ms = total;
  *launches = (uint32_t)c->timing_used;
  c->timing_used = 0;
  return FPV_OK;
}

void* fpv_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
  return p;
}
void fpv_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

// ---- delta frame ------------------------------------------------------------

// Every way of setting the delta image also refreshes its duplicated form.
static int refresh_delta_dup(fpv_ctx* c, cudaStream_t stream) {
  cudaError_t e = cudaSuccess;
  int l = enqueue_delta_dup(c->g, c->d_delta, c->d_delta_dup, stream, &e);
  if (l < 0) return cuda_fail(c, e, "delta duplication kernel launch");
  c->launches += (uint64_t)l;
  return FPV_OK;
}

int fpv_set_delta_raw_device(fpv_ctx* c, const void* raw_dev, void* stream) {
  if (!c) return FPV_ERR_INVALID_ARG;
  FPV_LOCK(c);
  FPV_CUDA(cudaSetDevice(c->device));
  if (!raw_dev) { c->has_delta = false; return FPV_OK; }
  cudaError_t e = cudaSuccess;
  int l = enqueue_delta_from_raw(c->g, static_cast<const uint16_t*>(raw_dev), c->d_delta,
                                 static_cast<cudaStream_t>(stream), &e);
  if (l < 0) return cuda_fail(c, e, "delta split kernel launch");
  c->launches += (uint64_t)l;
  int rc = refresh_delta_dup(c, static_cast<cudaStream_t>(stream));
  if (rc != FPV_OK) return rc;
  c->has_delta = true;
  return FPV_OK;
}

int fpv_set_delta_raw(fpv_ctx* c, const uint16_t* raw_host) {
  if (!c) return FPV_ERR_INVALID_ARG;
  FPV_LOCK(c);
  FPV_CUDA(cudaSetDevice(c->device));
  if (!raw_host) { c->has_delta = false; return FPV_OK; }
  uint16_t* tmp = nullptr;
  FPV_CUDA(cudaMalloc(&tmp, c->g.P * 2));
  cudaError_t e = cudaMemcpyAsync(tmp, raw_host, c->g.P * 2, cudaMemcpyHostToDevice, c->aux_stream);
  int rc = FPV_OK;
  if (e != cudaSuccess) rc = cuda_fail(c, e, "delta H2D copy");
  if (rc == FPV_OK) rc = fpv_set_delta_raw_device(c, tmp, c->aux_stream);
  cudaError_t es = cudaStreamSynchronize(c->aux_stream);
  cudaFree(tmp);
  if (rc == FPV_OK && es != cudaSuccess) rc = cuda_fail(c, es, "delta split");
  return rc;
}

int fpv_set_delta_image_device(fpv_ctx* c, const void* image_dev, void* stream) {
  if (!c) return FPV_ERR_INVALID_ARG;
  FPV_LOCK(c);
  FPV_CUDA(cudaSetDevice(c->device));
  if (!image_dev) { c->has_delta = false; return FPV_OK; }
  FPV_CUDA(cudaMemcpyAsync(c->d_delta, image_dev, c->g.P * 2, cudaMemcpyDeviceToDevice,
                           static_cast<cudaStream_t>(stream)));
  int rc = refresh_delta_dup(c, static_cast<cudaStream_t>(stream));
  if (rc != FPV_OK) return rc;
  c->has_delta = true;
  return FPV_OK;
}

int fpv_set_delta_image(fpv_ctx* c, const uint16_t* image_host) {
  if (!c) return FPV_ERR_INVALID_ARG;
  FPV_LOCK(c);
  FPV_CUDA(cudaSetDevice(c->device));
  if (!image_host) { c->has_delta = false; return FPV_OK; }
  FPV_CUDA(cudaMemcpy(c->d_delta, image_host, c->g.P * 2, cudaMemcpyHostToDevice));
  int rc = refresh_delta_dup(c, nullptr);
  if (rc != FPV_OK) return rc;
  FPV_CUDA(cudaStreamSynchronize(nullptr));
  c->has_delta = true;
  return FPV_OK;
}

int fpv_copy_delta_peer(fpv_ctx* dst, const fpv_ctx* src) {
  fpv_ctx* c = dst;
  if (!dst || !src) return FPV_ERR_INVALID_ARG;
  if (dst == src) return FPV_OK;
  // both contexts are touched; lock in address order so that two opposite copies cannot deadlock
  fpv_ctx* first = dst < src ? dst : const_cast<fpv_ctx*>(src);
  fpv_ctx* second = dst < src ? const_cast<fpv_ctx*>(src) : dst;
  std::lock_guard<std::recursive_mutex> l1(first->mu);
  std::lock_guard<std::recursive_mutex> l2(second->mu);
  if (dst->g.P != src->g.P) return fail(dst, FPV_ERR_INVALID_ARG, "geometry mismatch between contexts");
  if (!src->has_delta) { dst->has_delta = false; return FPV_OK; }
  FPV_CUDA(cudaSetDevice(dst->device));
  FPV_CUDA(cudaMemcpyPeerAsync(dst->d_delta, dst->device, src->d_delta, src->device, dst->g.P * 2,
                               dst->aux_stream));
  int rc = refresh_delta_dup(dst, dst->aux_stream);
  if (rc != FPV_OK) return rc;
  FPV_CUDA(cudaStreamSynchronize(dst->aux_stream));
  dst->has_delta = true;
  return FPV_OK;
}

int fpv_delta_ipc_export(fpv_ctx* c, void* handle_out) {
  if (!c || !handle_out) return FPV_ERR_INVALID_ARG;
  FPV_LOCK(c);
  static_assert(sizeof(cudaIpcMemHandle_t) == FPV_IPC_HANDLE_BYTES, "IPC handle size");
  if (!c->has_delta) return fail(c, FPV_ERR_NO_DELTA, "no delta frame to export");
  FPV_CUDA(cudaSetDevice(c->device));
  FPV_CUDA(cudaStreamSynchronize(c->aux_stream));   // the image is complete before anyone maps it
  cudaIpcMemHandle_t h;
  FPV_CUDA(cudaIpcGetMemHandle(&h, c->d_delta));
  memcpy(handle_out, &h, sizeof h);
  return FPV_OK;
}

int fpv_delta_ipc_import(fpv_ctx* c, const void* handle) {
  if (!c || !handle) return FPV_ERR_INVALID_ARG;
  FPV_LOCK(c);
  FPV_CUDA(cudaSetDevice(c->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof h);
  void* peer = nullptr;
  FPV_CUDA(cudaIpcOpenMemHandle(&peer, h, cudaIpcMemLazyEnablePeerAccess));
  // device-to-device over NVLink / PCIe peer access; the owner keeps its copy
  cudaError_t e = cudaMemcpyAsync(c->d_delta, peer, c->g.P * 2, cudaMemcpyDefault, c->aux_stream);
  int rc = FPV_OK;
  if (e != cudaSuccess) rc = cuda_fail(c, e, "delta peer copy (IPC)");
  if (rc == FPV_OK) rc = refresh_delta_dup(c, c->aux_stream);
  cudaError_t es = cudaStreamSynchronize(c->aux_stream);
  cudaIpcCloseMemHandle(peer);
  if (rc == FPV_OK && es != cudaSuccess) rc = cuda_fail(c, es, "delta peer copy (IPC)");
  if (rc == FPV_OK) c->has_delta = true;
  return rc;
}

// ---- encode -------------------------------------------------------------------

int fpv_encode_device(fpv_ctx* c, const void* frames_dev, uint32_t n, uint32_t options, void* flags_dev,
                      void* high_dev, void* low_dev, void* preview_dev, void* stream) {
  if (!c) return FPV_ERR_INVALID_ARG;
  FPV_LOCK(c);
  if (n == 0) return FPV_OK;
  if (!frames_dev || !flags_dev || !high_dev || !preview_dev)
    return fail(c, FPV_ERR_INVALID_ARG, "NULL device buffer");
  if ((reinterpret_cast<uintptr_t>(frames_dev) & 15) || (reinterpret_cast<uintptr_t>(high_dev) & 15) ||
      (reinterpret_cast<uintptr_t>(low_dev) & 15) || (reinterpret_cast<uintptr_t>(preview_dev) & 15))
    return fail(c, FPV_ERR_INVALID_ARG, "device buffers must be 16-byte aligned");
  FPV_CUDA(cudaSetDevice(c->device));
  return encode_device_chunked(c, kDeviceScratch, static_cast<const uint16_t*>(frames_dev), n, options,
                               static_cast<uint8_t*>(flags_dev), static_cast<uint8_t*>(high_dev),
                               static_cast<uint8_t*>(low_dev), static_cast<uint8_t*>(preview_dev),
                               static_cast<cudaStream_t>(stream));
}

// frames_host (n frames back to back) or frame_ptrs (n pointers to one frame each) -> the slot's device frames
static int upload_frames(fpv_ctx* c, Slot& s, const uint16_t* frames_host, const uint16_t* const* frame_ptrs, uint32_t n) {
  const size_t P = c->g.P;
  if (frames_host) {
    FPV_CUDA(cudaMemcpyAsync(s.d_frames, frames_host, (size_t)n * P * 2, cudaMemcpyHostToDevice, s.stream));
    return FPV_OK;
  }
  // neighbouring frames that are also neighbours in host memory go as one copy
  for (uint32_t i = 0; i < n;) {
    uint32_t j = i + 1;
    while (j < n && frame_ptrs[j] == frame_ptrs[j - 1] + P) j++;
    if (!frame_ptrs[i]) return fail(c, FPV_ERR_INVALID_ARG, "NULL frame pointer");
    FPV_CUDA(cudaMemcpyAsync(s.d_frames + (size_t)i * P, frame_ptrs[i], (size_t)(j - i) * P * 2, cudaMemcpyHostToDevice, s.stream));
    i = j;
  }
  return FPV_OK;
}

static int encode_submit_impl(fpv_ctx* c, uint32_t slot, const uint16_t* frames_host, const uint16_t* const* frame_ptrs,
                              uint32_t n, uint32_t options, uint8_t* flags_host, uint8_t* high_host, uint8_t* low_host,
                              uint8_t* preview_host) {
  if (!c) return FPV_ERR_INVALID_ARG;
  FPV_LOCK(c);
  if (slot >= kNumSlots) return fail(c, FPV_ERR_INVALID_ARG, "slot out of range");
  if (n == 0) return FPV_OK;
  if (n > c->max_batch) return fail(c, FPV_ERR_INVALID_ARG, "n exceeds max_batch");
  if ((!frames_host && !frame_ptrs) || !flags_host || !high_host || !preview_host)
    return fail(c, FPV_ERR_INVALID_ARG, "NULL host buffer");
  const bool has_low = mode_has_low(c->g.mode);
  if (has_low && !low_host) return fail(c, FPV_ERR_INVALID_ARG, "low plane buffer is NULL");
  FPV_CUDA(cudaSetDevice(c->device));
  int rc = ensure_slot(c, (int)slot);
  if (rc != FPV_OK) return rc;
  Slot& s = c->slots[slot];
  const size_t P = c->g.P, PP = c->g.PP;
  rc = upload_frames(c, s, frames_host, frame_ptrs, n);
  if (rc != FPV_OK) return rc;
  rc = encode_device_chunked(c, (int)slot, s.d_frames, n, options, s.d_flags, s.d_high, s.d_low,
                             s.d_preview, s.stream);
  if (rc != FPV_OK) return rc;
  FPV_CUDA(cudaMemcpyAsync(high_host, s.d_high, (size_t)n * P, cudaMemcpyDeviceToHost, s.stream));
  if (has_low)
    FPV_CUDA(cudaMemcpyAsync(low_host, s.d_low, (size_t)n * P, cudaMemcpyDeviceToHost, s.stream));
  FPV_CUDA(cudaMemcpyAsync(preview_host, s.d_preview, (size_t)n * PP, cudaMemcpyDeviceToHost, s.stream));
  FPV_CUDA(cudaMemcpyAsync(flags_host, s.d_flags, (size_t)n, cudaMemcpyDeviceToHost, s.stream));
  FPV_CUDA(cudaEventRecord(s.done, s.stream));
  return FPV_OK;
}

int fpv_encode_submit(fpv_ctx* c, uint32_t slot, const uint16_t* frames_host, uint32_t n, uint32_t options,
                      uint8_t* flags_host, uint8_t* high_host, uint8_t* low_host, uint8_t* preview_host) {
  return encode_submit_impl(c, slot, frames_host, nullptr, n, options, flags_host, high_host, low_host, preview_host);
}

int fpv_encode_submit_v(fpv_ctx* c, uint32_t slot, const uint16_t* const* frame_ptrs, uint32_t n, uint32_t options,
                        uint8_t* flags_host, uint8_t* high_host, uint8_t* low_host, uint8_t* preview_host) {
  return encode_submit_impl(c, slot, nullptr, frame_ptrs, n, options, flags_host, high_host, low_host, preview_host);
}

int fpv_host_is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (!p || cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return a.type == cudaMemoryTypeHost ? 1 : 0;
}

int fpv_wait(fpv_ctx* c, uint32_t slot) {
  if (!c) return FPV_ERR_INVALID_ARG;
  if (slot >= kNumSlots) return fail(c, FPV_ERR_INVALID_ARG, "slot out of range");
  std::unique_lock<std::recursive_mutex> lk(c->mu);
  Slot& s = c->slots[slot];
  if (!s.allocated) return FPV_OK;
  FPV_CUDA(cudaSetDevice(c->device));
  // Sleep WITHOUT the context lock: another thread may submit into a different slot meanwhile.  (When the
  // caller is one of the synchronous entry points the lock is held once more further up; that is fine,
  // those calls are blocking by contract.)
  lk.unlock();
  cudaError_t e = cudaEventSynchronize(s.done);                    // everything submitted on the slot so far
  if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);       // (returns at once; surfaces stream errors)
  lk.lock();
  if (e != cudaSuccess) return cuda_fail(c, e, "fpv_wait");
  if (s.stream_n) {
    // second half of fpv_encode_stream_submit: the coded size is known now, fetch exactly that many bytes
    const uint64_t total = s.stream_off_host[s.stream_n];
    const uint32_t n = s.stream_n;
    s.stream_n = 0;
    if (total > stream_bound(c, n)) return fail(c, FPV_ERR_CUDA, "entropy coder overflowed its output bound");
    FPV_CUDA(cudaMemcpyAsync(s.stream_out_host, c->entropy[slot].out, (size_t)total, cudaMemcpyDeviceToHost, s.stream));
    FPV_CUDA(cudaEventRecord(s.done, s.stream));
    lk.unlock();
    e = cudaEventSynchronize(s.done);
    lk.lock();
    if (e != cudaSuccess) return cuda_fail(c, e, "fpv_wait (coded bytes)");
  }
  return FPV_OK;
}

int fpv_encode(fpv_ctx* c, const uint16_t* frames_host, uint32_t n, uint32_t options, uint8_t* flags_host,
               uint8_t* high_host, uint8_t* low_host, uint8_t* preview_host) {
  if (!c) return FPV_ERR_INVALID_ARG;
  FPV_LOCK(c);
  const size_t P = c->g.P, PP = c->g.PP;
  for (uint32_t off = 0; off < n; off += c->max_batch) {
    uint32_t m = n - off < c->max_batch ? n - off : c->max_batch;
    int rc = fpv_encode_submit(c, 0, frames_host + (size_t)off * P, m, options, flags_host + off,
                               high_host + (size_t)off * P, low_host ? low_host + (size_t)off * P : nullptr,
                               preview_host + (size_t)off * PP);
    if (rc != FPV_OK) return rc;
    rc = fpv_wait(c, 0);
    if (rc != FPV_OK) return rc;
  }
  return FPV_OK;
}

int fpv_split(fpv_ctx* c, const uint16_t* frames_host, uint32_t n, uint8_t* flags_host, uint8_t* high_host,
              uint8_t* low_host) {
  if (!c) return FPV_ERR_INVALID_ARG;
  FPV_LOCK(c);
  if (n == 0) return FPV_OK;
  if (!frames_host || !flags_host || !high_host) return fail(c, FPV_ERR_INVALID_ARG, "NULL host buffer");
  const bool has_low = mode_has_low(c->g.mode);
  if (has_low && !low_host) return fail(c, FPV_ERR_INVALID_ARG, "low plane buffer is NULL");
  FPV_CUDA(cudaSetDevice(c->device));
  int rc = ensure_slot(c, 0);
  if (rc != FPV_OK) return rc;
  Slot& s = c->slots[0];
  const size_t P = c->g.P;
  std::vector<uint32_t> lor(c->max_batch);
  for (uint32_t off = 0; off < n; off += c->max_batch) {
    const uint32_t m = n - off < c->max_batch ? n - off : c->max_batch;
    FPV_CUDA(cudaMemcpyAsync(s.d_frames, frames_host + (size_t)off * P, (size_t)m * P * 2, cudaMemcpyHostToDevice, s.stream));
    cudaError_t e = cudaSuccess;
    // the preview staging buffer doubles as the per-frame OR words (4 bytes per frame; it holds >= 1 byte per frame only
    // for tiny geometries, so use the flags + a dedicated allocation instead)
    uint32_t* d_or = nullptr;
    FPV_CUDA(cudaMallocAsync(&d_or, sizeof(uint32_t) * m, s.stream));
    int l = enqueue_split(c->g, s.d_frames, m, s.d_high, has_low ? s.d_low : nullptr, d_or, s.stream, &e);
    if (l < 0) { cudaFreeAsync(d_or, s.stream); return cuda_fail(c, e, "split kernel launch"); }
    c->launches += (uint64_t)l;
    FPV_CUDA(cudaMemcpyAsync(high_host + (size_t)off * P, s.d_high, (size_t)m * P, cudaMemcpyDeviceToHost, s.stream));
    if (has_low)
      FPV_CUDA(cudaMemcpyAsync(low_host + (size_t)off * P, s.d_low, (size_t)m * P, cudaMemcpyDeviceToHost, s.stream));
    FPV_CUDA(cudaMemcpyAsync(lor.data(), d_or, sizeof(uint32_t) * m, cudaMemcpyDeviceToHost, s.stream));
    FPV_CUDA(cudaFreeAsync(d_or, s.stream));
    FPV_CUDA(cudaStreamSynchronize(s.stream));
    for (uint32_t i = 0; i < m; i++)
      flags_host[off + i] = (!has_low || (lor[i] & 0xffu) == 0) ? FPV_FLAG_NO_LOW_BYTES : 0;
  }
  return FPV_OK;
}

// ---- GPU entropy coding ---------------------------------------------------------

size_t fpv_stream_bound(const fpv_ctx* c, uint32_t n) { return c ? stream_bound(c, n) : 0; }

int fpv_entropy_device(fpv_ctx* c, const void* flags_dev, const void* high_dev, const void* low_dev,
                       const void* preview_dev, uint32_t n, void* out_dev, size_t capacity, void* frame_off_dev,
                       void* stream) {
  if (!c) return FPV_ERR_INVALID_ARG;
  FPV_LOCK(c);
  if (n == 0) return FPV_OK;
  if (n > c->max_batch) return fail(c, FPV_ERR_INVALID_ARG, "n exceeds max_batch");
  if (!flags_dev || !high_dev || !preview_dev || !out_dev || !frame_off_dev)
    return fail(c, FPV_ERR_INVALID_ARG, "NULL device buffer");
  if (mode_has_low(c->g.mode) && !low_dev) return fail(c, FPV_ERR_INVALID_ARG, "low plane buffer is NULL");
  if (capacity < stream_bound(c, n)) return fail(c, FPV_ERR_INVALID_ARG, "capacity below fpv_stream_bound(n)");
  FPV_CUDA(cudaSetDevice(c->device));
  int rc = ensure_entropy(c, kDeviceScratch, false);
  if (rc != FPV_OK) return rc;
  return entropy_device_impl(c, kDeviceScratch, static_cast<const uint8_t*>(flags_dev),
                             static_cast<const uint8_t*>(high_dev), static_cast<const uint8_t*>(low_dev),
                             static_cast<const uint8_t*>(preview_dev), n, static_cast<uint8_t*>(out_dev), capacity,
                             static_cast<uint64_t*>(frame_off_dev), static_cast<cudaStream_t>(stream));
}

static int encode_stream_submit_impl(fpv_ctx* c, uint32_t slot, const uint16_t* frames_host, const uint16_t* const* frame_ptrs,
                                     uint32_t n, uint32_t options, uint8_t* flags_host, uint64_t* frame_off_host,
                                     uint8_t* out_host, size_t capacity) {
  if (!c) return FPV_ERR_INVALID_ARG;
  FPV_LOCK(c);
  if (slot >= kNumSlots) return fail(c, FPV_ERR_INVALID_ARG, "slot out of range");
  if (n == 0) return FPV_OK;
  if (n > c->max_batch) return fail(c, FPV_ERR_INVALID_ARG, "n exceeds max_batch");
  if ((!frames_host && !frame_ptrs) || !flags_host || !frame_off_host || !out_host)
    return fail(c, FPV_ERR_INVALID_ARG, "NULL host buffer");
  if (capacity < stream_bound(c, n)) return fail(c, FPV_ERR_INVALID_ARG, "capacity below fpv_stream_bound(n)");
  FPV_CUDA(cudaSetDevice(c->device));
  int rc = ensure_slot(c, (int)slot);
  if (rc == FPV_OK) rc = ensure_entropy(c, (int)slot, true);
  if (rc != FPV_OK) return rc;
  Slot& s = c->slots[slot];
  if (s.stream_n) return fail(c, FPV_ERR_INVALID_ARG, "slot has an unfinished stream submit: call fpv_wait first");
  EntropyBuf& e = c->entropy[slot];
  rc = upload_frames(c, s, frames_host, frame_ptrs, n);
  if (rc != FPV_OK) return rc;
  rc = encode_device_chunked(c, (int)slot, s.d_frames, n, options, s.d_flags, s.d_high, s.d_low, s.d_preview,
                             s.stream);
  if (rc != FPV_OK) return rc;
  rc = entropy_device_impl(c, (int)slot, s.d_flags, s.d_high, s.d_low, s.d_preview, n, e.out,
                           stream_bound(c, c->max_batch), e.frame_off, s.stream);
  if (rc != FPV_OK) return rc;
  FPV_CUDA(cudaMemcpyAsync(flags_host, s.d_flags, (size_t)n, cudaMemcpyDeviceToHost, s.stream));
  FPV_CUDA(cudaMemcpyAsync(frame_off_host, e.frame_off, (size_t)(n + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost,
                           s.stream));
  FPV_CUDA(cudaEventRecord(s.done, s.stream));
  s.stream_out_host = out_host;
  s.stream_off_host = frame_off_host;
  s.stream_n = n;
  return FPV_OK;
}

int fpv_encode_stream_submit(fpv_ctx* c, uint32_t slot, const uint16_t* frames_host, uint32_t n, uint32_t options,
                             uint8_t* flags_host, uint64_t* frame_off_host, uint8_t* out_host, size_t capacity) {
  return encode_stream_submit_impl(c, slot, frames_host, nullptr, n, options, flags_host, frame_off_host, out_host, capacity);
}

int fpv_encode_stream_submit_v(fpv_ctx* c, uint32_t slot, const uint16_t* const* frame_ptrs, uint32_t n, uint32_t options,
                               uint8_t* flags_host, uint64_t* frame_off_host, uint8_t* out_host, size_t capacity) {
  return encode_stream_submit_impl(c, slot, nullptr, frame_ptrs, n, options, flags_host, frame_off_host, out_host, capacity);
}

// ---- decode -------------------------------------------------------------------

int fpv_decode_device(fpv_ctx* c, const void* high_dev, const void* low_dev, const void* flags_dev,
                      uint32_t n, uint32_t options, void* out_dev, void* stream) {
  if (!c) return FPV_ERR_INVALID_ARG;
  FPV_LOCK(c);
  if (n == 0) return FPV_OK;
  if (!high_dev || !flags_dev || !out_dev) return fail(c, FPV_ERR_INVALID_ARG, "NULL device buffer");
  if ((reinterpret_cast<uintptr_t>(high_dev) & 15) || (reinterpret_cast<uintptr_t>(low_dev) & 15) ||
      (reinterpret_cast<uintptr_t>(out_dev) & 15))
    return fail(c, FPV_ERR_INVALID_ARG, "device buffers must be 16-byte aligned");
  FPV_CUDA(cudaSetDevice(c->device));
  return decode_device_impl(c, static_cast<const uint8_t*>(high_dev), static_cast<const uint8_t*>(low_dev),
                            static_cast<const uint8_t*>(flags_dev), n, options,
                            static_cast<uint16_t*>(out_dev), static_cast<cudaStream_t>(stream), false);
}

int fpv_decode_submit(fpv_ctx* c, uint32_t slot, const uint8_t* high_host, const uint8_t* low_host,
                      const uint8_t* flags_host, uint32_t n, uint32_t options, void* out_host) {
  if (!c) return FPV_ERR_INVALID_ARG;
  FPV_LOCK(c);
  if (slot >= kNumSlots) return fail(c, FPV_ERR_INVALID_ARG, "slot out of range");
  if (n == 0) return FPV_OK;
  if (n > c->max_batch) return fail(c, FPV_ERR_INVALID_ARG, "n exceeds max_batch");
  if (!high_host || !flags_host || !out_host) return fail(c, FPV_ERR_INVALID_ARG, "NULL host buffer");
  bool any_delta = false, any_low = false;
  for (uint32_t i = 0; i < n; i++) {
    if (flags_host[i] & FPV_FLAG_USE_DELTA) any_delta = true;
    if (!(flags_host[i] & FPV_FLAG_NO_LOW_BYTES)) any_low = true;
  }
  if (any_delta && !c->has_delta)
    return fail(c, FPV_ERR_NO_DELTA, "delta frame not given");  // .cc:310
  if (any_low && !low_host) return fail(c, FPV_ERR_INVALID_ARG, "low plane buffer is NULL");
  FPV_CUDA(cudaSetDevice(c->device));
  int rc = ensure_slot(c, (int)slot);
  if (rc != FPV_OK) return rc;
  Slot& s = c->slots[slot];
  const size_t P = c->g.P;
  FPV_CUDA(cudaMemcpyAsync(s.d_high, high_host, (size_t)n * P, cudaMemcpyHostToDevice, s.stream));
  if (low_host)
    FPV_CUDA(cudaMemcpyAsync(s.d_low, low_host, (size_t)n * P, cudaMemcpyHostToDevice, s.stream));
  FPV_CUDA(cudaMemcpyAsync(s.d_flags, flags_host, (size_t)n, cudaMemcpyHostToDevice, s.stream));
  rc = decode_device_impl(c, s.d_high, low_host ? s.d_low : nullptr, s.d_flags, n, options, s.d_frames,
                          s.stream, true);
  if (rc != FPV_OK) return rc;
  FPV_CUDA(cudaMemcpyAsync(out_host, s.d_frames, (size_t)n * P * 2, cudaMemcpyDeviceToHost, s.stream));
  FPV_CUDA(cudaEventRecord(s.done, s.stream));
  return FPV_OK;
}

int fpv_decode(fpv_ctx* c, const uint8_t* high_host, const uint8_t* low_host, const uint8_t* flags_host,
               uint32_t n, uint32_t options, void* out_host) {
  if (!c) return FPV_ERR_INVALID_ARG;
  FPV_LOCK(c);
  const size_t P = c->g.P;
  for (uint32_t off = 0; off < n; off += c->max_batch) {
    uint32_t m = n - off < c->max_batch ? n - off : c->max_batch;
    int rc = fpv_decode_submit(c, 0, high_host + (size_t)off * P, low_host ? low_host + (size_t)off * P : nullptr,
                               flags_host + off, m, options, static_cast<uint8_t*>(out_host) + (size_t)off * P * 2);
    if (rc != FPV_OK) return rc;
    rc = fpv_wait(c, 0);
    if (rc != FPV_OK) return rc;
  }
  return FPV_OK;
}

int fpv_decode_coded(fpv_ctx* c, const uint8_t* blob_host, size_t blob_bytes, const fpv_coded_chunk* chunks_host,
                     uint32_t n_chunks, const uint8_t* flags_host, uint32_t n, uint32_t options, void* out_host) {
  if (!c) return FPV_ERR_INVALID_ARG;
  FPV_LOCK(c);
  if (n == 0) return FPV_OK;
  if (n > c->max_batch) return fail(c, FPV_ERR_INVALID_ARG, "n exceeds max_batch");
  if (!blob_host || !chunks_host || !flags_host || !out_host) return fail(c, FPV_ERR_INVALID_ARG, "NULL host buffer");
  static_assert(sizeof(fpv_coded_chunk) == sizeof(CodedChunk), "fpv_coded_chunk layout");
  const size_t P = c->g.P;
  const uint32_t cpl = chunks_of(P);
  // every plane a frame needs must be covered by exactly its chunks (a missing chunk would leave stale bytes)
  std::vector<uint32_t> seen((size_t)n * 2, 0);
  bool any_delta = false;
  for (uint32_t i = 0; i < n_chunks; i++) {
    const fpv_coded_chunk& k = chunks_host[i];
    if (k.frame >= n || k.plane > 1 || k.index >= cpl || k.offset >= blob_bytes)
      return fail(c, FPV_ERR_INVALID_ARG, "chunk table entry out of range");
    seen[(size_t)k.frame * 2 + k.plane]++;
  }
  for (uint32_t f = 0; f < n; f++) {
    if (flags_host[f] & FPV_FLAG_USE_DELTA) any_delta = true;
    const bool has_low = !(flags_host[f] & FPV_FLAG_NO_LOW_BYTES);
    if (seen[(size_t)f * 2] != cpl || seen[(size_t)f * 2 + 1] != (has_low ? cpl : 0u))
      return fail(c, FPV_ERR_INVALID_ARG, "chunk table does not cover the planes of every frame exactly once");
  }
  if (any_delta && !c->has_delta) return fail(c, FPV_ERR_NO_DELTA, "delta frame not given");  // .cc:310
  FPV_CUDA(cudaSetDevice(c->device));
  int rc = ensure_slot(c, 0);
  if (rc != FPV_OK) return rc;
  Slot& s0 = c->slots[0];
  if (s0.coded_cap < blob_bytes + 16) {
    if (s0.d_coded) cudaFree(s0.d_coded);
    s0.d_coded = nullptr;
    s0.coded_cap = 0;
    const size_t want = std::max(blob_bytes + 16, stream_bound(c, c->max_batch));
    FPV_CUDA(cudaMalloc(&s0.d_coded, want));
    s0.coded_cap = want;
  }
  if (s0.chunks_cap < n_chunks) {
    if (s0.d_chunks) cudaFree(s0.d_chunks);
    s0.d_chunks = nullptr;
    s0.chunks_cap = 0;
    const size_t want = std::max<size_t>(n_chunks, (size_t)c->max_batch * 2 * cpl);
    FPV_CUDA(cudaMalloc(&s0.d_chunks, want * sizeof(CodedChunk)));
    s0.chunks_cap = want;
  }
  if (!s0.d_dec_err) {
    FPV_CUDA(cudaMalloc(&s0.d_dec_err, kNumSlots * sizeof(uint32_t)));
    FPV_CUDA(cudaMemset(s0.d_dec_err, 0, kNumSlots * sizeof(uint32_t)));
    FPV_CUDA(cudaMallocHost(&s0.h_dec_err, kNumSlots * sizeof(uint32_t)));
  }
  // The batch is cut into up to kNumSlots pieces of whole frames, each on its own slot (stream + plane buffers):
  // the upload of one piece overlaps the kernels of another and the download of a third.  That needs the chunk
  // table grouped by frame (the host decoders build it that way); otherwise the batch goes as one piece.
  bool grouped = true;
  for (uint32_t i = 1; i < n_chunks && grouped; i++) grouped = chunks_host[i].frame >= chunks_host[i - 1].frame;
  uint32_t pieces = grouped ? std::min<uint32_t>(kNumSlots, (n + 15) / 16) : 1;
  if (pieces < 1) pieces = 1;
  uint32_t* dec_err = s0.h_dec_err;
  for (int i = 0; i < kNumSlots; i++) dec_err[i] = 0;
  uint32_t c0 = 0;
  for (uint32_t pc = 0; pc < pieces; pc++) {
    const uint32_t f0 = (uint32_t)((uint64_t)n * pc / pieces), f1 = (uint32_t)((uint64_t)n * (pc + 1) / pieces);
    if (f1 == f0) continue;
    rc = ensure_slot(c, (int)pc);
    if (rc != FPV_OK) return rc;
    Slot& s = c->slots[pc];
    uint32_t c1 = c0;
    uint64_t lo = blob_bytes, hi = 0;
    if (grouped) {
      while (c1 < n_chunks && chunks_host[c1].frame < f1) c1++;
    } else {
      c1 = n_chunks;
    }
    for (uint32_t i = c0; i < c1; i++) {
      // the chunk's size sits in its directory (bytes 6..8); a chunk that claims more than the blob holds is
      // cut off here and refused by the kernel
      const uint64_t o = chunks_host[i].offset;
      uint64_t len = blob_bytes - o;
      if (len >= 9) {
        const uint64_t cb = (uint64_t)blob_host[o + 6] | ((uint64_t)blob_host[o + 7] << 8) | ((uint64_t)blob_host[o + 8] << 16);
        if (cb + 1 < len) len = cb + 1;            // + the stream's final 0x03 after a plane's last chunk
      }
      lo = std::min(lo, o);
      hi = std::max(hi, o + len);
    }
    if (hi > lo) {
      lo &= ~(uint64_t)15;
      FPV_CUDA(cudaMemcpyAsync(s0.d_coded + lo, blob_host + lo, hi - lo, cudaMemcpyHostToDevice, s.stream));
    }
    if (c1 > c0)
      FPV_CUDA(cudaMemcpyAsync(s0.d_chunks + c0, chunks_host + c0, (size_t)(c1 - c0) * sizeof(CodedChunk), cudaMemcpyHostToDevice,
                               s.stream));
    FPV_CUDA(cudaMemcpyAsync(s.d_flags, flags_host + f0, (size_t)(f1 - f0), cudaMemcpyHostToDevice, s.stream));
    FPV_CUDA(cudaMemsetAsync(s0.d_dec_err + pc, 0, sizeof(uint32_t), s.stream));
    EntropyDecodeParams p;
    p.blob = s0.d_coded; p.blob_bytes = blob_bytes; p.chunks = s0.d_chunks + c0; p.n_chunks = c1 - c0;
    p.n_frames = f1 - f0; p.frame0 = f0;
    p.high = s.d_high; p.low = s.d_low; p.P = P; p.err = s0.d_dec_err + pc;
    cudaError_t e = cudaSuccess;
    const int l = enqueue_entropy_decode(p, s.stream, &e);
    if (l < 0) return cuda_fail(c, e, "entropy decode kernel launch");
    c->launches += (uint64_t)l;
    rc = decode_device_impl(c, s.d_high, s.d_low, s.d_flags, f1 - f0, options, s.d_frames, s.stream, true);
    if (rc != FPV_OK) return rc;
    FPV_CUDA(cudaMemcpyAsync(&dec_err[pc], s0.d_dec_err + pc, sizeof(uint32_t), cudaMemcpyDeviceToHost, s.stream));
    FPV_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(out_host) + (size_t)f0 * P * 2, s.d_frames, (size_t)(f1 - f0) * P * 2,
                             cudaMemcpyDeviceToHost, s.stream));
    c0 = c1;
  }
  for (uint32_t pc = 0; pc < pieces; pc++)
    if (c->slots[pc].allocated) FPV_CUDA(cudaStreamSynchronize(c->slots[pc].stream));
  for (uint32_t pc = 0; pc < pieces; pc++)
    if (dec_err[pc])
      return fail(c, FPV_ERR_INVALID_ARG, "malformed coded chunk (entropy decoder reason " + std::to_string(dec_err[pc]) + ")");
  return FPV_OK;
}

int fpv_unpredict_planes(fpv_ctx* c, uint8_t* high_host, uint8_t* low_host, uint8_t* preview_host,
                         const uint8_t* flags_host, uint32_t n) {
  if (!c) return FPV_ERR_INVALID_ARG;
  FPV_LOCK(c);
  if (n == 0) return FPV_OK;
  if (!high_host || !flags_host) return fail(c, FPV_ERR_INVALID_ARG, "NULL host buffer");
  FPV_CUDA(cudaSetDevice(c->device));
  int rc = ensure_slot(c, 0);
  if (rc != FPV_OK) return rc;
  Slot& s = c->slots[0];
  const size_t P = c->g.P, PP = c->g.PP;
  for (uint32_t off = 0; off < n; off += c->max_batch) {
    uint32_t m = n - off < c->max_batch ? n - off : c->max_batch;
    FPV_CUDA(cudaMemcpyAsync(s.d_high, high_host + (size_t)off * P, (size_t)m * P, cudaMemcpyHostToDevice, s.stream));
    if (low_host)
      FPV_CUDA(cudaMemcpyAsync(s.d_low, low_host + (size_t)off * P, (size_t)m * P, cudaMemcpyHostToDevice, s.stream));
    if (preview_host)
      FPV_CUDA(cudaMemcpyAsync(s.d_preview, preview_host + (size_t)off * PP, (size_t)m * PP, cudaMemcpyHostToDevice, s.stream));
    FPV_CUDA(cudaMemcpyAsync(s.d_flags, flags_host + off, (size_t)m, cudaMemcpyHostToDevice, s.stream));
    cudaError_t e = cudaSuccess;
    int l = enqueue_unpredict_planes(c->g, c->tune.num_sms, s.d_high, low_host ? s.d_low : nullptr,
                                     preview_host ? s.d_preview : nullptr, s.d_flags,
                                     c->has_delta ? c->d_delta : nullptr, m, s.stream, &e);
    if (l < 0) return cuda_fail(c, e, "unpredict kernel launch");
    c->launches += (uint64_t)l;
    FPV_CUDA(cudaMemcpyAsync(high_host + (size_t)off * P, s.d_high, (size_t)m * P, cudaMemcpyDeviceToHost, s.stream));
    if (low_host)
      FPV_CUDA(cudaMemcpyAsync(low_host + (size_t)off * P, s.d_low, (size_t)m * P, cudaMemcpyDeviceToHost, s.stream));
    if (preview_host)
      FPV_CUDA(cudaMemcpyAsync(preview_host + (size_t)off * PP, s.d_preview, (size_t)m * PP, cudaMemcpyDeviceToHost, s.stream));
    FPV_CUDA(cudaStreamSynchronize(s.stream));
  }
  return FPV_OK;
}

}  // extern "C"

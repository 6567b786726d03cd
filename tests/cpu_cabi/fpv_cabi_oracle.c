/*
 * fpv_cabi_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A CPU stand-in for libfpv_b200.so: the C ABI of include/fpv_b200.h implemented on the oracle
 * (oracle/fpv_oracle.c).  It exists so that the HOST layer (csrc/host: fpvc::Encoder, the decoders, the
 * fpvc::Frame facade, the columnar batch classes, the unmodified reference CLIs built against the mirror
 * header) can be exercised by `pytest -m "not gpu"` in a container without a GPU.  tests/cpu_host.py
 * compiles it into tests/_build/ together with a second copy of the host sources; nothing under
 * fusion_power_video_b200/ links it, ships it or falls back to it -- the product library has no CPU path
 * and fpv_create there fails with FPV_ERR_NO_DEVICE when there is no GPU.
 *
 * Not implemented here: the GPU entropy coder (FPV_ERR_UNSUPPORTED), kernel timing, IPC.
 */
#include <stdlib.h>
#include <string.h>

#include "../../include/fpv_b200.h"
#include "../../oracle/fpv_oracle.c"

struct fpv_ctx {
  uint32_t W, H;
  size_t P, PP;
  int shift, big_endian, has_low;
  uint32_t max_batch;
  int has_delta;
  uint8_t *dhigh, *dlow;   /* split delta planes */
  uint16_t* dimage;        /* (dhigh << 8) | dlow */
  uint8_t* scratch;
  uint64_t launches;
};

static __thread const char* tl_err = "no error";
static int fail(int code, const char* msg) { tl_err = msg; return code; }

const char* fpv_version(void) { return "fpv_b200 CPU TEST STAND-IN (oracle) -- not the product"; }
int fpv_device_count(void) { return 1; }
const char* fpv_last_error(const fpv_ctx* c) { (void)c; return tl_err; }
int fpv_bind_thread(const fpv_ctx* c) { return c ? FPV_OK : FPV_ERR_INVALID_ARG; }
int fpv_device_of(const fpv_ctx* c) { (void)c; return 0; }
size_t fpv_plane_bytes(const fpv_ctx* c) { return c ? c->P : 0; }
size_t fpv_preview_bytes(const fpv_ctx* c) { return c ? c->PP : 0; }
uint64_t fpv_kernel_launches(const fpv_ctx* c) { return c ? c->launches : 0; }
int fpv_enable_kernel_timing(fpv_ctx* c, int on) { (void)c; (void)on; return FPV_OK; }
int fpv_read_kernel_timing(fpv_ctx* c, double* ms, uint32_t* n) { (void)c; *ms = 0; *n = 0; return FPV_OK; }
void* fpv_host_alloc(size_t bytes) { void* p = NULL; return posix_memalign(&p, 4096, bytes ? bytes : 1) ? NULL : p; }
void fpv_host_free(void* p) { free(p); }

int fpv_create(fpv_ctx** out, int device, uint32_t xsize, uint32_t ysize, int shift, int big_endian, uint32_t max_batch) {
  if (!out) return fail(FPV_ERR_INVALID_ARG, "ctx out pointer is NULL");
  *out = NULL;
  if (device < 0 || device >= 8) return fail(FPV_ERR_INVALID_ARG, "device index out of range");   /* pretend 8 devices: multi-GPU host logic */
  if (xsize == 0 || ysize == 0 || xsize > 65536 || ysize > 65536 || (uint64_t)xsize * ysize > 1000000000ull)
    return fail(FPV_ERR_INVALID_ARG, "invalid image dimensions");
  if (shift < 0 || shift > 16 || (big_endian && shift > 8)) return fail(FPV_ERR_UNSUPPORTED, "shift");
  if (max_batch == 0 || max_batch > 65535) return fail(FPV_ERR_INVALID_ARG, "max_batch");
  fpv_ctx* c = (fpv_ctx*)calloc(1, sizeof *c);
  c->W = xsize; c->H = ysize; c->P = (size_t)xsize * ysize; c->PP = (size_t)(xsize / 4) * (ysize / 4);
  c->shift = shift; c->big_endian = big_endian != 0; c->has_low = shift != 8; c->max_batch = max_batch;
  c->dhigh = (uint8_t*)malloc(c->P); c->dlow = (uint8_t*)calloc(c->P, 1);
  c->dimage = (uint16_t*)malloc(c->P * 2); c->scratch = (uint8_t*)malloc(c->P);
  *out = c;
  return FPV_OK;
}
void fpv_destroy(fpv_ctx* c) {
  if (!c) return;
  free(c->dhigh); free(c->dlow); free(c->dimage); free(c->scratch); free(c);
}

static void refresh_image(fpv_ctx* c) {
  for (size_t i = 0; i < c->P; i++) c->dimage[i] = (uint16_t)((c->dhigh[i] << 8) | c->dlow[i]);
}
int fpv_set_delta_raw(fpv_ctx* c, const uint16_t* raw) {
  if (!c) return FPV_ERR_INVALID_ARG;
  if (!raw) { c->has_delta = 0; return FPV_OK; }
  memset(c->dlow, 0, c->P);
  fpvo_split(raw, c->P, c->shift, c->big_endian, c->dhigh, c->dlow);
  refresh_image(c);
  c->has_delta = 1; c->launches++;
  return FPV_OK;
}
int fpv_set_delta_raw_device(fpv_ctx* c, const void* raw, void* stream) { (void)stream; return fpv_set_delta_raw(c, (const uint16_t*)raw); }
int fpv_set_delta_image(fpv_ctx* c, const uint16_t* img) {
  if (!c) return FPV_ERR_INVALID_ARG;
  if (!img) { c->has_delta = 0; return FPV_OK; }
  for (size_t i = 0; i < c->P; i++) { c->dimage[i] = img[i]; c->dhigh[i] = (uint8_t)(img[i] >> 8); c->dlow[i] = (uint8_t)(img[i] & 0xff); }
  c->has_delta = 1; c->launches++;
  return FPV_OK;
}
int fpv_set_delta_image_device(fpv_ctx* c, const void* img, void* stream) { (void)stream; return fpv_set_delta_image(c, (const uint16_t*)img); }
int fpv_copy_delta_peer(fpv_ctx* dst, const fpv_ctx* src) {
  if (!dst || !src || dst->P != src->P) return fail(FPV_ERR_INVALID_ARG, "geometry mismatch between contexts");
  if (!src->has_delta) { dst->has_delta = 0; return FPV_OK; }
  return fpv_set_delta_image(dst, src->dimage);
}
int fpv_delta_ipc_export(fpv_ctx* c, void* h) { (void)c; (void)h; return fail(FPV_ERR_UNSUPPORTED, "no IPC in the CPU stand-in"); }
int fpv_delta_ipc_import(fpv_ctx* c, const void* h) { (void)c; (void)h; return fail(FPV_ERR_UNSUPPORTED, "no IPC in the CPU stand-in"); }

int fpv_encode(fpv_ctx* c, const uint16_t* frames, uint32_t n, uint32_t options, uint8_t* flags, uint8_t* high,
               uint8_t* low, uint8_t* preview) {
  if (!c) return FPV_ERR_INVALID_ARG;
  if (n == 0) return FPV_OK;
  if (!frames || !flags || !high || !preview) return fail(FPV_ERR_INVALID_ARG, "NULL host buffer");
  if (c->W % 4 || c->H % 4) return fail(FPV_ERR_UNSUPPORTED, "encode requires xsize % 4 == 0 and ysize % 4 == 0");
  if (c->has_low && !low) return fail(FPV_ERR_INVALID_ARG, "low plane buffer is NULL");
  const int use_delta = c->has_delta && !(options & FPV_ENC_NO_DELTA);
  for (uint32_t i = 0; i < n; i++) {
    flags[i] = fpvo_predict(frames + (size_t)i * c->P, c->W, c->H, c->shift, c->big_endian, use_delta ? c->dhigh : NULL,
                            use_delta ? c->dlow : NULL, high + (size_t)i * c->P, low ? low + (size_t)i * c->P : NULL,
                            preview + (size_t)i * c->PP, c->scratch);
    c->launches++;
  }
  return FPV_OK;
}
int fpv_split(fpv_ctx* c, const uint16_t* frames, uint32_t n, uint8_t* flags, uint8_t* high, uint8_t* low) {
  if (!c) return FPV_ERR_INVALID_ARG;
  if (n == 0) return FPV_OK;
  if (!frames || !flags || !high) return fail(FPV_ERR_INVALID_ARG, "NULL host buffer");
  if (c->has_low && !low) return fail(FPV_ERR_INVALID_ARG, "low plane buffer is NULL");
  for (uint32_t i = 0; i < n; i++)
    flags[i] = fpvo_split(frames + (size_t)i * c->P, c->P, c->shift, c->big_endian, high + (size_t)i * c->P,
                          low ? low + (size_t)i * c->P : NULL);
  return FPV_OK;
}
int fpv_encode_device(fpv_ctx* c, const void* frames, uint32_t n, uint32_t options, void* flags, void* high, void* low,
                      void* preview, void* stream) {
  (void)stream;
  return fpv_encode(c, (const uint16_t*)frames, n, options, (uint8_t*)flags, (uint8_t*)high, (uint8_t*)low, (uint8_t*)preview);
}
int fpv_encode_submit(fpv_ctx* c, uint32_t slot, const uint16_t* frames, uint32_t n, uint32_t options, uint8_t* flags,
                      uint8_t* high, uint8_t* low, uint8_t* preview) {
  if (slot >= FPV_NUM_SLOTS) return fail(FPV_ERR_INVALID_ARG, "slot out of range");
  if (c && n > c->max_batch) return fail(FPV_ERR_INVALID_ARG, "n exceeds max_batch");
  return fpv_encode(c, frames, n, options, flags, high, low, preview);
}
int fpv_encode_submit_v(fpv_ctx* c, uint32_t slot, const uint16_t* const* frame_ptrs, uint32_t n, uint32_t options,
                        uint8_t* flags, uint8_t* high, uint8_t* low, uint8_t* preview) {
  if (slot >= FPV_NUM_SLOTS) return fail(FPV_ERR_INVALID_ARG, "slot out of range");
  if (!c || !frame_ptrs) return fail(FPV_ERR_INVALID_ARG, "NULL");
  if (n > c->max_batch) return fail(FPV_ERR_INVALID_ARG, "n exceeds max_batch");
  for (uint32_t i = 0; i < n; i++) {
    int rc = fpv_encode(c, frame_ptrs[i], 1, options, flags + i, high + (size_t)i * c->P, low ? low + (size_t)i * c->P : NULL,
                        preview + (size_t)i * ((c->W / 4) * (size_t)(c->H / 4)));
    if (rc != FPV_OK) return rc;
  }
  return FPV_OK;
}
int fpv_host_is_pinned(const void* p) { (void)p; return 0; }
int fpv_wait(fpv_ctx* c, uint32_t slot) { return c && slot < FPV_NUM_SLOTS ? FPV_OK : FPV_ERR_INVALID_ARG; }

size_t fpv_stream_bound(const fpv_ctx* c, uint32_t n) { return c ? (size_t)n * (3 * c->P + 4096) : 0; }
int fpv_entropy_device(fpv_ctx* c, const void* a, const void* b, const void* d, const void* e, uint32_t n, void* o,
                       size_t cap, void* off, void* s) {
  (void)c; (void)a; (void)b; (void)d; (void)e; (void)n; (void)o; (void)cap; (void)off; (void)s;
  return fail(FPV_ERR_UNSUPPORTED, "the GPU entropy coder has no CPU stand-in");
}
int fpv_encode_stream_submit(fpv_ctx* c, uint32_t slot, const uint16_t* f, uint32_t n, uint32_t o, uint8_t* fl,
                             uint64_t* off, uint8_t* out, size_t cap) {
  (void)c; (void)slot; (void)f; (void)n; (void)o; (void)fl; (void)off; (void)out; (void)cap;
  return fail(FPV_ERR_UNSUPPORTED, "the GPU entropy coder has no CPU stand-in");
}
int fpv_encode_stream_submit_v(fpv_ctx* c, uint32_t slot, const uint16_t* const* f, uint32_t n, uint32_t o, uint8_t* fl,
                             uint64_t* off, uint8_t* out, size_t cap) {
  (void)c; (void)slot; (void)f; (void)n; (void)o; (void)fl; (void)off; (void)out; (void)cap;
  return fail(FPV_ERR_UNSUPPORTED, "the GPU entropy coder has no CPU stand-in");
}

int fpv_decode(fpv_ctx* c, const uint8_t* high, const uint8_t* low, const uint8_t* flags, uint32_t n, uint32_t options,
               void* out) {
  if (!c) return FPV_ERR_INVALID_ARG;
  if (n == 0) return FPV_OK;
  if (!high || !flags || !out) return fail(FPV_ERR_INVALID_ARG, "NULL host buffer");
  uint16_t* img = (uint16_t*)malloc(c->P * 2);
  for (uint32_t i = 0; i < n; i++) {
    const uint8_t f = flags[i];
    if ((f & FPV_FLAG_USE_DELTA) && !c->has_delta) { free(img); return fail(FPV_ERR_NO_DELTA, "delta frame not given"); }
    const int lowp = !(f & FPV_FLAG_NO_LOW_BYTES);
    if (lowp && !low) { free(img); return fail(FPV_ERR_INVALID_ARG, "low plane buffer is NULL"); }
    memcpy(c->scratch, high + (size_t)i * c->P, c->P);
    fpvo_inverse(c->scratch, lowp ? low + (size_t)i * c->P : NULL, c->dimage, c->W, c->H, f, img);
    if (options & FPV_DEC_UNEXTRACT) fpvo_unextract(img, c->P, c->shift, c->big_endian, (uint8_t*)out + (size_t)i * c->P * 2);
    else memcpy((uint8_t*)out + (size_t)i * c->P * 2, img, c->P * 2);
    c->launches++;
  }
  free(img);
  return FPV_OK;
}
int fpv_decode_device(fpv_ctx* c, const void* high, const void* low, const void* flags, uint32_t n, uint32_t options,
                      void* out, void* stream) {
  (void)stream;
  return fpv_decode(c, (const uint8_t*)high, (const uint8_t*)low, (const uint8_t*)flags, n, options, out);
}
int fpv_decode_coded(fpv_ctx* c, const uint8_t* blob, size_t blob_bytes, const fpv_coded_chunk* chunks, uint32_t n_chunks,
                     const uint8_t* flags, uint32_t n, uint32_t options, void* out) {
  (void)c; (void)blob; (void)blob_bytes; (void)chunks; (void)n_chunks; (void)flags; (void)n; (void)options; (void)out;
  return fail(FPV_ERR_UNSUPPORTED, "the GPU entropy decoder has no CPU stand-in");
}
int fpv_decode_submit(fpv_ctx* c, uint32_t slot, const uint8_t* high, const uint8_t* low, const uint8_t* flags, uint32_t n,
                      uint32_t options, void* out) {
  if (slot >= FPV_NUM_SLOTS) return fail(FPV_ERR_INVALID_ARG, "slot out of range");
  if (c && n > c->max_batch) return fail(FPV_ERR_INVALID_ARG, "n exceeds max_batch");
  return fpv_decode(c, high, low, flags, n, options, out);
}
int fpv_unpredict_planes(fpv_ctx* c, uint8_t* high, uint8_t* low, uint8_t* preview, const uint8_t* flags, uint32_t n) {
  if (!c) return FPV_ERR_INVALID_ARG;
  if (n == 0) return FPV_OK;
  if (!high || !flags) return fail(FPV_ERR_INVALID_ARG, "NULL host buffer");
  for (uint32_t i = 0; i < n; i++)
    fpvo_unpredict_planes(high + (size_t)i * c->P, low ? low + (size_t)i * c->P : NULL,
                          preview ? preview + (size_t)i * c->PP : NULL, c->has_delta ? c->dhigh : NULL,
                          c->has_delta ? c->dlow : NULL, c->W, c->H, flags[i]);
  return FPV_OK;
}

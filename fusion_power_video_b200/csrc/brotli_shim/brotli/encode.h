/* Prototype shim: see types.h in this directory. */
#ifndef FPV_BROTLI_SHIM_ENCODE_H_
#define FPV_BROTLI_SHIM_ENCODE_H_

#include "types.h"

#ifdef __cplusplus
extern "C" {
#endif

#define BROTLI_MIN_QUALITY 0
#define BROTLI_MAX_QUALITY 11
#define BROTLI_DEFAULT_QUALITY 11
#define BROTLI_DEFAULT_WINDOW 22

typedef enum BrotliEncoderMode {
  BROTLI_MODE_GENERIC = 0,
  BROTLI_MODE_TEXT = 1,
  BROTLI_MODE_FONT = 2
} BrotliEncoderMode;

#define BROTLI_DEFAULT_MODE BROTLI_MODE_GENERIC

size_t BrotliEncoderMaxCompressedSize(size_t input_size);

BROTLI_BOOL BrotliEncoderCompress(int quality, int lgwin, BrotliEncoderMode mode,
                                  size_t input_size, const uint8_t* input_buffer,
                                  size_t* encoded_size, uint8_t* encoded_buffer);

uint32_t BrotliEncoderVersion(void);

#ifdef __cplusplus
}
#endif

#endif  /* FPV_BROTLI_SHIM_ENCODE_H_ */

/* Prototype shim: see types.h in this directory. */
#ifndef FPV_BROTLI_SHIM_DECODE_H_
#define FPV_BROTLI_SHIM_DECODE_H_

#include "types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct BrotliDecoderStateStruct BrotliDecoderState;

typedef enum {
  BROTLI_DECODER_RESULT_ERROR = 0,
  BROTLI_DECODER_RESULT_SUCCESS = 1,
  BROTLI_DECODER_RESULT_NEEDS_MORE_INPUT = 2,
  BROTLI_DECODER_RESULT_NEEDS_MORE_OUTPUT = 3
} BrotliDecoderResult;

BrotliDecoderState* BrotliDecoderCreateInstance(brotli_alloc_func alloc_func,
                                                brotli_free_func free_func,
                                                void* opaque);
void BrotliDecoderDestroyInstance(BrotliDecoderState* state);

BrotliDecoderResult BrotliDecoderDecompress(size_t encoded_size,
                                            const uint8_t* encoded_buffer,
                                            size_t* decoded_size,
                                            uint8_t* decoded_buffer);

BrotliDecoderResult BrotliDecoderDecompressStream(BrotliDecoderState* state,
                                                  size_t* available_in,
                                                  const uint8_t** next_in,
                                                  size_t* available_out,
                                                  uint8_t** next_out,
                                                  size_t* total_out);

const uint8_t* BrotliDecoderTakeOutput(BrotliDecoderState* state, size_t* size);
BROTLI_BOOL BrotliDecoderHasMoreOutput(const BrotliDecoderState* state);
BROTLI_BOOL BrotliDecoderIsFinished(const BrotliDecoderState* state);
uint32_t BrotliDecoderVersion(void);

#ifdef __cplusplus
}
#endif

#endif  /* FPV_BROTLI_SHIM_DECODE_H_ */

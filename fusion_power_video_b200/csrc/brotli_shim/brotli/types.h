/* Minimal prototype shim for libbrotli 1.1.0.
 *
 * This image ships the brotli runtime libraries
 * (/usr/lib/x86_64-linux-gnu/libbrotli{enc,dec,common}.so.1) but neither the
 * development headers nor the pkg-config files.  These three files declare
 * only the handful of public entry points the fusion-power-video path uses
 * (RFC 7932 one-shot encode, streaming decode).  They are declarations of a
 * third-party ABI, written from the documented public API; link against the
 * .so.1 files by full path. */
#ifndef FPV_BROTLI_SHIM_TYPES_H_
#define FPV_BROTLI_SHIM_TYPES_H_

#include <stddef.h>
#include <stdint.h>

#define BROTLI_BOOL int
#define BROTLI_TRUE 1
#define BROTLI_FALSE 0

typedef void* (*brotli_alloc_func)(void* opaque, size_t size);
typedef void (*brotli_free_func)(void* opaque, void* address);

#endif  /* FPV_BROTLI_SHIM_TYPES_H_ */

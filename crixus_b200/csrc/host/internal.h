// crixus_b200 -- library-internal entry points used by the host pipeline (not part of the public C-ABI).
#pragma once
#include <stddef.h>
#include <stdint.h>
#include "../../../include/crixus_b200.h"
extern "C" {
// grow-only arenas owned by the context: >= bytes of device memory / pinned host memory, contents undefined
void *cx_internal_dev_arena(cx_ctx *ctx, size_t bytes);
void *cx_internal_host_arena(cx_ctx *ctx, size_t bytes);
// dst[i] = (src[3i], src[3i+1], src[3i+2], 0) on the device, i < n
int cx_internal_expand3(cx_ctx *ctx, const float *src3_d, float *dst4_d, int64_t n);
}

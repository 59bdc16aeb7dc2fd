// crixus_b200 -- library-internal entry points used by the host pipeline (not part of the public C-ABI).
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <nvtx3/nvToolsExt.h>
#include "../../../include/crixus_b200.h"

// NVTX range of the running stage (header-only NVTX 3: a no-op unless a profiler is attached); closed on every exit path
struct CxStage {
  bool open = false;
  void next(const char *name) { end(); nvtxRangePushA(name); open = true; }
  void end() { if (open) { nvtxRangePop(); open = false; } }
  ~CxStage() { end(); }
};
extern "C" {
// grow-only arenas owned by the context: >= bytes of device memory / pinned host memory, contents undefined
void *cx_internal_dev_arena(cx_ctx *ctx, size_t bytes);
void *cx_internal_norm_stage(cx_ctx *ctx, size_t bytes);   /* grow-only device staging of the raw normals */
void *cx_internal_host_arena(cx_ctx *ctx, size_t bytes);
// dst[i] = (src[3i], src[3i+1], src[3i+2], 0) on the device, i < n
int cx_internal_expand3(cx_ctx *ctx, const float *src3_d, float *dst4_d, int64_t n);
// dst[i] = (src[i].x + offset, src[i].y + offset, src[i].z + offset, 0) on the device, i < n (segments appended behind a mesh)
int cx_internal_offset_ep4(cx_ctx *ctx, const int32_t *src_ep4_d, int32_t *dst_ep4_d, int64_t n, int32_t offset);
// in-place all-gather over the context's communicator: rank r owns bytes [r*bytes, (r+1)*bytes) of buf_d
int cx_internal_allgather(cx_ctx *ctx, void *buf_d, int64_t bytes);
}

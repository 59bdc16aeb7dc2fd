// crixus_b200 -- shared device/host helpers.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "../../include/crixus_b200.h"

#define CX_NUM_SMS_FALLBACK 148

struct cx_ctx {
  int device = 0;
  int num_sms = CX_NUM_SMS_FALLBACK;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string last_error;
  int64_t launches = 0;
  // scratch arena (device) -- grows on demand, reused between calls
  void *scratch = nullptr;
  size_t scratch_bytes = 0;
  // small pinned host mailbox for counters read back per call / per round
  int64_t *h_mail = nullptr;  // 64 x int64 pinned
  int64_t *d_mail = nullptr;  // 64 x int64 device
  // persistent fill state for the slab variant
  void *fill_state = nullptr;
  int fill_max_rounds = 0;     // 0: cx_fill_geometry_slab runs until its slab is converged; n: at most ~n rounds per call
  int fill_unfinished = 0;     // the last slab call stopped with active tiles left
  int fill_resolve_mode = 0;   // cx_fill_set_resolve_mode: 0 = by dr (record-driven kernels at fine dr), 1 = cell-driven kernel always
  // grow-only arenas of cx_preprocess_host: device working set and pinned host result buffers (cudaMalloc and
  // page-locking cost milliseconds per call; the arenas make repeated calls allocation-free)
  void *dev_arena = nullptr, *host_arena = nullptr;
  size_t dev_arena_bytes = 0, host_arena_bytes = 0;
  // pinned staging of the streamed file writers (records.cu)
  void *pin_stage = nullptr;
  size_t pin_stage_bytes = 0;
  // device staging of the raw STL normals of cx_preprocess_host (grow-only; they travel on the copy stream while the weld runs)
  void *norm_stage = nullptr;
  size_t norm_stage_bytes = 0;
  // multi-GPU: collective back-end of this rank (dist.h: NCCL communicator or in-process group); nullptr = single GPU
  void *dist = nullptr;
  int dist_fill_mode = 0;       // cx_dist_set_fill_mode: 0 = by lattice size, 1 = replicated flood, 2 = z-slab flood
};

int cx_fail(cx_ctx *ctx, int code, const char *what, cudaError_t e = cudaSuccess);
// returns a device pointer to >= bytes of scratch (contents undefined); nullptr on failure
void *cx_scratch(cx_ctx *ctx, size_t bytes);
void *cx_scratch_pinned(cx_ctx *ctx, size_t bytes);  // grow-only pinned host staging; nullptr on failure
void cx_fill_state_free(cx_ctx *ctx);  // fill.cu
void cx_dist_release(cx_ctx *ctx);     // dist.cu

#define CX_CUDA(ctx, call)                                              \
  do {                                                                  \
    cudaError_t e__ = (call);                                           \
    if (e__ != cudaSuccess) return cx_fail(ctx, CX_CUDA_ERROR, #call, e__); \
  } while (0)

#define CX_LAUNCH_CHECK(ctx, name)                                      \
  do {                                                                  \
    (ctx)->launches++;                                                  \
    cudaError_t e__ = cudaGetLastError();                               \
    if (e__ != cudaSuccess) return cx_fail(ctx, CX_CUDA_ERROR, name, e__); \
  } while (0)

static inline int cx_blocks(int64_t n, int threads, int64_t cap = 1 << 30)
{
  int64_t b = (n + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > cap) b = cap;
  return (int)b;
}

// ------------------------------------------------------------------------------------------------ numerics
// The reference's float expression trees with the exact multiply-add fusion nvcc 12.9 emits for them on sm_100
// (read from the SASS of the reference build, DESIGN.md "Numerics").  Every operation is pinned with an
// intrinsic so that this translation unit's own -fmad setting cannot change results.
//   R1  sp = 0; sp += a[o]*b[o]                -> fma(a2,b2, fma(a1,b1, fma(a0,b0, 0)))
//   R2  a*b - c*d                              -> fma(a, b, -(c*d))
//   R3  a0*b0 + a1*b1 + a2*b2                  -> fma(a2,b2, fma(a0,b0, a1*b1))
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float dot_r1(float a0, float a1, float a2, float b0, float b1, float b2)
{
  return ffma(a2, b2, ffma(a1, b1, ffma(a0, b0, 0.0f)));
}
__device__ __forceinline__ float dot_r3(float a0, float a1, float a2, float b0, float b1, float b2)
{
  return ffma(a2, b2, ffma(a0, b0, fmul(a1, b1)));
}
__device__ __forceinline__ float det_r2(float a, float b, float c, float d) { return ffma(a, b, -fmul(c, d)); }

struct f3 { float x, y, z; };
__device__ __forceinline__ f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ f3 mk3(float4 a) { return mk3(a.x, a.y, a.z); }
__device__ __forceinline__ f3 sub3(f3 a, f3 b) { return mk3(fsub(a.x, b.x), fsub(a.y, b.y), fsub(a.z, b.z)); }
__device__ __forceinline__ f3 add3(f3 a, f3 b) { return mk3(fadd(a.x, b.x), fadd(a.y, b.y), fadd(a.z, b.z)); }
__device__ __forceinline__ f3 neg3(f3 a) { return mk3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot3_r1(f3 a, f3 b) { return dot_r1(a.x, a.y, a.z, b.x, b.y, b.z); }
__device__ __forceinline__ float dot3_r3(f3 a, f3 b) { return dot_r3(a.x, a.y, a.z, b.x, b.y, b.z); }
// cross(a,b) of src/vector_math.h:48-55 under R2
__device__ __forceinline__ f3 cross_r2(f3 a, f3 b)
{
  return mk3(det_r2(a.y, b.z, a.z, b.y), det_r2(a.z, b.x, a.x, b.z), det_r2(a.x, b.y, a.y, b.x));
}
__device__ __forceinline__ float comp(f3 a, int n) { return n == 0 ? a.x : (n == 1 ? a.y : a.z); }

// 3dr-cell coordinate of a point: max(min((int)((x-dmin)/griddr), gridn-1), 0)   (src/crixus_d.cu:55)
__device__ __forceinline__ int cell_coord(float x, float dmin, float griddr, int gridn)
{
  int g = (int)__fdiv_rn(fsub(x, dmin), griddr);
  return max(min(g, gridn - 1), 0);
}
// neighbour cell with periodic wrap / skip (src/crixus_d.cu:327-347); returns -1 when skipped
__device__ __forceinline__ int64_t nb_cell(const cx_params &p, int gx, int gy, int gz)
{
  int g[3] = {gx, gy, gz};
#pragma unroll
  for (int j = 0; j < 3; j++) {
    if (g[j] == -1 || g[j] == p.gridn[j]) {
      if (p.per[j]) g[j] = (g[j] == -1) ? p.gridn[j] - 1 : 0;
      else return -1;
    }
  }
  return (int64_t)g[0] * p.gridn[1] * p.gridn[2] + (int64_t)g[1] * p.gridn[2] + g[2];
}

__device__ __forceinline__ float4 ldg4(const float *base, int64_t i) { return __ldg(reinterpret_cast<const float4 *>(base) + i); }
__device__ __forceinline__ int4 ldg4i(const int32_t *base, int64_t i) { return __ldg(reinterpret_cast<const int4 *>(base) + i); }

// ------------------------------------------------------------------------------------------------ device scan
// exclusive scan of int32 -> int32 (totals < 2^31), n up to ~2^31; implemented in scan.cu
int cx_exclusive_scan_i32(cx_ctx *ctx, const int32_t *in_d, int32_t *out_d, int64_t n);

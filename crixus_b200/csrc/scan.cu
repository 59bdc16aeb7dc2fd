// crixus_b200 -- exclusive prefix sum of int32 (cell histogram -> cell_idx, ordered compaction of face vertices).
// Replaces the reference's serial add_up_indices<<<1,1>>> (/root/reference/src/crixus_d.cu:62-78).
// Reduce-then-scan, 3 passes over the data (12 B / element), recursive on the per-block sums.
#include "common.cuh"

#define SCAN_THREADS 256
#define SCAN_ITEMS 16
#define SCAN_CHUNK (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ int warp_incl_scan(int v)
{
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// block-wide exclusive scan of one value per thread; returns the exclusive prefix, *total = block sum
__device__ __forceinline__ int block_excl_scan(int v, int *total)
{
  __shared__ int wsum[SCAN_THREADS / 32];
  __shared__ int wtot;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = warp_incl_scan(v);
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  if (w == 0) {
    int s = lane < SCAN_THREADS / 32 ? wsum[lane] : 0;
    int si = warp_incl_scan(s);
    if (lane < SCAN_THREADS / 32) wsum[lane] = si - s;
    if (lane == SCAN_THREADS / 32 - 1) wtot = si;
  }
  __syncthreads();
  int r = inc - v + wsum[w];
  *total = wtot;
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const int32_t *__restrict__ in, int32_t *__restrict__ sums, int64_t n)
{
  const int64_t base = (int64_t)blockIdx.x * SCAN_CHUNK;
  int s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    int64_t i = base + k * SCAN_THREADS + threadIdx.x;  // coalesced
    if (i < n) s += in[i];
  }
  int tot;
  block_excl_scan(s, &tot);
  if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

// vec: in and out are 16-byte aligned (int4 loads / stores); the scalar path serves caller-owned pointers that are not (the
// C-ABI passes e.g. cell_idx_d straight through)
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const int32_t *__restrict__ in, int32_t *__restrict__ out,
                                                             const int32_t *__restrict__ offs, int64_t n, int vec)
{
  const int64_t base = (int64_t)blockIdx.x * SCAN_CHUNK + (int64_t)threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  int s = 0;
  const bool whole = vec && base + SCAN_ITEMS <= n;
  if (whole) {
    const int4 *p = reinterpret_cast<const int4 *>(in + base);  // base is a multiple of 16 ints
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS / 4; k++) {
      int4 q = p[k];
      v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) v[k] = (base + k < n) ? in[base + k] : 0;
  }
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) s += v[k];
  int tot;
  int ex = block_excl_scan(s, &tot) + (offs ? offs[blockIdx.x] : 0);
  if (whole) {
    int4 *p = reinterpret_cast<int4 *>(out + base);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS / 4; k++) {
      int4 q;
      q.x = ex; ex += v[4 * k];
      q.y = ex; ex += v[4 * k + 1];
      q.z = ex; ex += v[4 * k + 2];
      q.w = ex; ex += v[4 * k + 3];
      p[k] = q;
    }
  } else {
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++)
      if (base + k < n) { out[base + k] = ex; ex += v[k]; }
  }
}

int cx_exclusive_scan_i32(cx_ctx *ctx, const int32_t *in_d, int32_t *out_d, int64_t n)
{
  if (n <= 0) return CX_OK;
  const int64_t nb = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
  const int vec = (((uintptr_t)in_d | (uintptr_t)out_d) & 15u) == 0 ? 1 : 0;
  if (nb == 1) {
    k_scan_apply<<<1, SCAN_THREADS, 0, ctx->stream>>>(in_d, out_d, nullptr, n, vec);
    CX_LAUNCH_CHECK(ctx, "k_scan_apply");
    return CX_OK;
  }
  int32_t *sums = nullptr;
  CX_CUDA(ctx, cudaMallocAsync(&sums, sizeof(int32_t) * (size_t)nb, ctx->stream));
  k_scan_reduce<<<(unsigned)nb, SCAN_THREADS, 0, ctx->stream>>>(in_d, sums, n);
  CX_LAUNCH_CHECK(ctx, "k_scan_reduce");
  int rc = cx_exclusive_scan_i32(ctx, sums, sums, nb);  // in-place is safe: each block reads its chunk before writing
  if (rc != CX_OK) return rc;
  k_scan_apply<<<(unsigned)nb, SCAN_THREADS, 0, ctx->stream>>>(in_d, out_d, sums, n, vec);
  CX_LAUNCH_CHECK(ctx, "k_scan_apply");
  CX_CUDA(ctx, cudaFreeAsync(sums, ctx->stream));
  return CX_OK;
}

// crixus_b200 -- vertex weld and bounding box on the device.
// Reference: remove_duplicate_vertices, /root/reference/src/crixus.cpp:343-455 (host, std::map<mint3, list<intb>>), and
// the bounding box accumulation of src/crixus.cu:185-208.  Same outcome as csrc/host/weld.cpp (outcome-exact, see the
// comment there and SURVEY.md App. B-1):
//   * corners are bucketed by cell ((int)(x/dr), (int)(y/dr), (int)(z/dr));
//   * cells in lexicographic (x,y,z) order, corners of a cell in corner-index order;
//   * an unclaimed corner becomes the next vertex and claims every LATER corner of its cell closer than
//     sqrt(1e-10)*dr (also corners that were claimed before);
//   * the neighbour-cell searches of the reference have no lasting effect (by-value map copy, crixus.cpp:331).
// Device formulation: cell key = lexicographic rank of the cell inside the bounding box (<= 2^63), stable LSD radix
// sort of (key, corner index) -> cells are contiguous runs, run order = cell order, order inside a run = corner order;
// one thread per run replays the claim loop; representatives are numbered by a prefix sum over the sorted sequence,
// which is exactly the reference's vertex numbering.
#include <math.h>
#include "common.cuh"

// ------------------------------------------------------------------------------------------------ bounding box
__device__ __forceinline__ unsigned f2ord(float f)
{
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned o)
{
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

__global__ void k_bbox_init(unsigned *mm)
{
  if (threadIdx.x < 3) { mm[threadIdx.x] = 0xffffffffu; mm[3 + threadIdx.x] = 0u; }
}

__global__ void __launch_bounds__(256) k_bbox(const float *__restrict__ c, int64_t ncorner, unsigned *mm)
{
  float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < ncorner; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const float x = c[3 * i + a];
      lo[a] = fminf(lo[a], x);
      hi[a] = fmaxf(hi[a], x);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; a++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(&mm[a], f2ord(lo[a]));
      atomicMax(&mm[3 + a], f2ord(hi[a]));
    }
  }
}

__global__ void k_bbox_finish(const unsigned *mm, float *out)
{
  if (threadIdx.x < 6) out[threadIdx.x] = ord2f(mm[threadIdx.x]);
}

extern "C" int cx_bbox_device(cx_ctx *ctx, const float *corners_d, int64_t ncorner, float bbmin[3], float bbmax[3])
{
  if (!ctx || !corners_d || ncorner <= 0) return CX_WRONG_INPUT;
  unsigned *mm = (unsigned *)(ctx->d_mail + 32);
  float *out = (float *)(ctx->d_mail + 36);
  k_bbox_init<<<1, 32, 0, ctx->stream>>>(mm);
  CX_LAUNCH_CHECK(ctx, "k_bbox_init");
  k_bbox<<<cx_blocks(ncorner, 256, ctx->num_sms * 8), 256, 0, ctx->stream>>>(corners_d, ncorner, mm);
  CX_LAUNCH_CHECK(ctx, "k_bbox");
  k_bbox_finish<<<1, 32, 0, ctx->stream>>>(mm, out);
  CX_LAUNCH_CHECK(ctx, "k_bbox_finish");
  CX_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail + 36, out, 6 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  CX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const float *h = (const float *)(ctx->h_mail + 36);
  for (int a = 0; a < 3; a++) { bbmin[a] = h[a]; bbmax[a] = h[3 + a]; }
  return CX_OK;
}

// ------------------------------------------------------------------------------------------------ mesh precondition
// resources/test-triangle-size.c: every vertex-barycentre distance must be <= dr - 1e-4*dr (README.md: triangles larger
// than that break the 27-cell neighbourhood assumption of calc_vert_volume and checkCollision).  Same arithmetic:
// barycentre accumulated in float from double quotients (:43-46), distance in double, narrowed to float (:48).
__global__ void __launch_bounds__(256) k_triangle_size(const float *__restrict__ c, int64_t ntri, float dr, float eps, unsigned *mm,
                                                       unsigned long long *failed)
{
  float mind = 1e10f, maxd = 0.f;
  unsigned long long nf = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < ntri; i += (int64_t)gridDim.x * blockDim.x) {
    float v[9], s[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 9; q++) v[q] = c[9 * i + q];
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
      for (int k = 0; k < 3; k++) s[k] = (float)((double)s[k] + (double)v[3 * j + k] / 3.0);
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const double dx = (double)fsub(s[0], v[3 * j]), dy = (double)fsub(s[1], v[3 * j + 1]), dz = (double)fsub(s[2], v[3 * j + 2]);
      const float dist = (float)sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
      maxd = fmaxf(maxd, dist);
      mind = fminf(mind, dist);
      if (dist > fsub(dr, eps)) nf++;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mind = fminf(mind, __shfl_xor_sync(0xffffffffu, mind, o));
    maxd = fmaxf(maxd, __shfl_xor_sync(0xffffffffu, maxd, o));
    nf += __shfl_xor_sync(0xffffffffu, nf, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&mm[0], f2ord(mind));
    atomicMax(&mm[3], f2ord(maxd));
    if (nf) atomicAdd(failed, nf);
  }
}

extern "C" int cx_check_triangle_size(cx_ctx *ctx, const float *corners_d, int64_t ntri, float dr, int64_t *failed_host, float *mind_host,
                                      float *maxd_host)
{
  if (!ctx || !corners_d || ntri <= 0 || !(dr > 0)) return CX_WRONG_INPUT;
  unsigned *mm = (unsigned *)(ctx->d_mail + 32);
  float *out = (float *)(ctx->d_mail + 36);
  unsigned long long *failed = (unsigned long long *)(ctx->d_mail + 46);
  k_bbox_init<<<1, 32, 0, ctx->stream>>>(mm);
  CX_LAUNCH_CHECK(ctx, "k_bbox_init");
  CX_CUDA(ctx, cudaMemsetAsync(failed, 0, sizeof(unsigned long long), ctx->stream));
  const float eps = (float)(dr * 1e-4);     // test-triangle-size.c:24
  k_triangle_size<<<cx_blocks(ntri, 256, ctx->num_sms * 8), 256, 0, ctx->stream>>>(corners_d, ntri, dr, eps, mm, failed);
  CX_LAUNCH_CHECK(ctx, "k_triangle_size");
  k_bbox_finish<<<1, 32, 0, ctx->stream>>>(mm, out);
  CX_LAUNCH_CHECK(ctx, "k_bbox_finish");
  CX_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail + 36, out, 6 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  CX_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail + 46, failed, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
  CX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const float *h = (const float *)(ctx->h_mail + 36);
  if (failed_host) *failed_host = ctx->h_mail[46];
  if (mind_host) *mind_host = h[0];
  if (maxd_host) *maxd_host = h[3];
  return CX_OK;
}

// ------------------------------------------------------------------------------------------------ radix sort
// Stable LSD radix sort of (uint64 key, uint32 value) pairs, 8 bits per pass.  Per pass: per-block digit histograms
// (bin-major) -> exclusive scan -> stable scatter (rank inside the block = warp match + per-warp digit counts).
#define RS_THREADS 256
#define RS_ROUNDS 16
#define RS_TILE (RS_THREADS * RS_ROUNDS)

__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const uint64_t *__restrict__ keys, int64_t n, int shift, int32_t *__restrict__ hist,
                                                        int nblocks)
{
  __shared__ int h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll 4
  for (int r = 0; r < RS_ROUNDS; r++) {
    const int64_t i = base + (int64_t)r * RS_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&h[(int)((keys[i] >> shift) & 255u)], 1);
  }
  __syncthreads();
  hist[(int64_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals, int64_t n,
                                                           int shift, const int32_t *__restrict__ offs, int nblocks,
                                                           uint64_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out)
{
  __shared__ int off[256];          // next output slot of each digit for this block
  __shared__ int wc[RS_THREADS / 32][256];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  off[threadIdx.x] = offs[(int64_t)threadIdx.x * nblocks + blockIdx.x];
  const int64_t base = (int64_t)blockIdx.x * RS_TILE;
  for (int r = 0; r < RS_ROUNDS; r++) {
    const int64_t i = base + (int64_t)r * RS_THREADS + threadIdx.x;
    if (base + (int64_t)r * RS_THREADS >= n) break;   // block-uniform
    const bool valid = i < n;
#pragma unroll
    for (int q = 0; q < RS_THREADS / 32; q++) wc[q][threadIdx.x] = 0;
    __syncthreads();
    uint64_t k = 0;
    uint32_t v = 0;
    int d = 256 + lane;   // invalid lanes: a digit nobody shares
    if (valid) { k = keys[i]; v = vals[i]; d = (int)((k >> shift) & 255u); }
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const int wrank = __popc(peers & ((1u << lane) - 1u));
    if (valid && wrank == 0) wc[w][d] = __popc(peers);
    __syncthreads();
    {   // thread t owns digit t: prefix over the warps, advance the block offset
      int run = off[threadIdx.x];
#pragma unroll
      for (int q = 0; q < RS_THREADS / 32; q++) {
        const int c = wc[q][threadIdx.x];
        wc[q][threadIdx.x] = run;
        run += c;
      }
      off[threadIdx.x] = run;
    }
    __syncthreads();
    if (valid) {
      const int p = wc[w][d] + wrank;
      keys_out[p] = k;
      vals_out[p] = v;
    }
    __syncthreads();
  }
}

// sorts in place semantics: returns through *keys_res / *vals_res which of the two buffer pairs holds the result
static int radix_sort_pairs(cx_ctx *ctx, uint64_t *keys, uint32_t *vals, uint64_t *keys_tmp, uint32_t *vals_tmp, int64_t n, int bits,
                            int32_t *hist, uint64_t **keys_res, uint32_t **vals_res)
{
  const int nblocks = (int)((n + RS_TILE - 1) / RS_TILE);
  for (int shift = 0; shift < bits; shift += 8) {
    k_rs_hist<<<nblocks, RS_THREADS, 0, ctx->stream>>>(keys, n, shift, hist, nblocks);
    CX_LAUNCH_CHECK(ctx, "k_rs_hist");
    int rc = cx_exclusive_scan_i32(ctx, hist, hist, (int64_t)256 * nblocks);
    if (rc != CX_OK) return rc;
    k_rs_scatter<<<nblocks, RS_THREADS, 0, ctx->stream>>>(keys, vals, n, shift, hist, nblocks, keys_tmp, vals_tmp);
    CX_LAUNCH_CHECK(ctx, "k_rs_scatter");
    uint64_t *tk = keys; keys = keys_tmp; keys_tmp = tk;
    uint32_t *tv = vals; vals = vals_tmp; vals_tmp = tv;
  }
  *keys_res = keys;
  *vals_res = vals;
  return CX_OK;
}

// ------------------------------------------------------------------------------------------------ weld
#define CX_WELD_BIG_RUN 256   // corners per dr-cell above which the claim loop of a cell is run by a whole block

struct WeldGrid {
  int32_t cmin[3];
  int64_t ny, nz;
  float dr;
};

__global__ void __launch_bounds__(256) k_weld_keys(const float *__restrict__ c, int64_t ncorner, WeldGrid g, uint64_t *__restrict__ keys,
                                                   uint32_t *__restrict__ vals)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < ncorner; i += (int64_t)gridDim.x * blockDim.x) {
    const int cx = (int)__fdiv_rn(c[3 * i + 0], g.dr), cy = (int)__fdiv_rn(c[3 * i + 1], g.dr), cz = (int)__fdiv_rn(c[3 * i + 2], g.dr);  // crixus.cpp:354
    keys[i] = ((uint64_t)(cx - g.cmin[0]) * (uint64_t)g.ny + (uint64_t)(cy - g.cmin[1])) * (uint64_t)g.nz + (uint64_t)(cz - g.cmin[2]);
    vals[i] = (uint32_t)i;
  }
}

__global__ void __launch_bounds__(256) k_weld_heads(const uint64_t *__restrict__ keys, int64_t n, int32_t *__restrict__ head)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i <= n; i += (int64_t)gridDim.x * blockDim.x)
    head[i] = (i < n && (i == 0 || keys[i] != keys[i - 1])) ? 1 : 0;
}

// run starts: start[r] = first sorted position of run r (ordered compaction through the scanned head flags)
__global__ void __launch_bounds__(256) k_weld_starts(const uint64_t *__restrict__ keys, const int32_t *__restrict__ headscan, int64_t n,
                                                     int32_t *__restrict__ start)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (i == 0 || keys[i] != keys[i - 1]) start[headscan[i]] = (int32_t)i;
}

// longest run (corners per cell): runs above CX_WELD_BIG_RUN are claimed by a whole block (k_weld_cells_big)
__global__ void __launch_bounds__(256) k_weld_maxrun(const int32_t *__restrict__ start, int64_t nruns, int64_t n, int32_t *__restrict__ mx)
{
  int m = 0;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < nruns; r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = (r + 1 < nruns) ? (int64_t)start[r + 1] : n;
    m = max(m, (int)(e - start[r]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(mx, m);
}

// one thread per cell: the claim loop of crixus.cpp:363-380 + find_in_cell (:317-329)
__global__ void __launch_bounds__(128) k_weld_cells(const float *__restrict__ c, const uint32_t *__restrict__ vals, const int32_t *__restrict__ start,
                                                    int64_t nruns, int64_t n, double thr, int32_t *__restrict__ rep, int32_t *__restrict__ isrep)
{
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < nruns; r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = start[r], e = (r + 1 < nruns) ? (int64_t)start[r + 1] : n;
    if (e - s > CX_WELD_BIG_RUN) continue;   // k_weld_cells_big
    for (int64_t i = s; i < e; i++) { rep[i] = -1; isrep[i] = 0; }
    for (int64_t i = s; i < e; i++) {
      if (rep[i] >= 0) continue;   // found
      rep[i] = (int32_t)i;
      isrep[i] = 1;
      const float *pi = c + 3 * (int64_t)vals[i];
      const float xi = pi[0], yi = pi[1], zi = pi[2];
      for (int64_t k = i + 1; k < e; k++) {
        const float *pk = c + 3 * (int64_t)vals[k];
        const float xij = __fsub_rn(pk[0], xi), yij = __fsub_rn(pk[1], yi), zij = __fsub_rn(pk[2], zi);
        const float dist = __fadd_rn(__fadd_rn(__fmul_rn(xij, xij), __fmul_rn(yij, yij)), __fmul_rn(zij, zij));   // crixus.cpp:320-323
        if ((double)dist < thr) rep[k] = (int32_t)i;
      }
    }
  }
}

// the same claim loop for over-full cells (dr far larger than the mesh spacing: thousands of corners in one dr-cell), one BLOCK
// per cell: the outer loop over unclaimed corners stays sequential (the claim order is the outcome), the inner loop over the
// later corners of the cell is spread over the threads.  Same result as the one-thread loop for every run length.
__global__ void __launch_bounds__(256) k_weld_cells_big(const float *__restrict__ c, const uint32_t *__restrict__ vals,
                                                        const int32_t *__restrict__ start, int64_t nruns, int64_t n, double thr,
                                                        volatile int32_t *rep, int32_t *__restrict__ isrep)
{
  for (int64_t r = blockIdx.x; r < nruns; r += gridDim.x) {
    const int64_t s = start[r], e = (r + 1 < nruns) ? (int64_t)start[r + 1] : n;
    if (e - s <= CX_WELD_BIG_RUN) continue;
    for (int64_t i = s + threadIdx.x; i < e; i += blockDim.x) { rep[i] = -1; isrep[i] = 0; }
    __syncthreads();
    for (int64_t i = s; i < e; i++) {
      if (rep[i] >= 0) continue;   // block-uniform: every write to rep[] is followed by a barrier before the next read
      if (threadIdx.x == 0) { rep[i] = (int32_t)i; isrep[i] = 1; }
      const float *pi = c + 3 * (int64_t)vals[i];
      const float xi = pi[0], yi = pi[1], zi = pi[2];
      for (int64_t k = i + 1 + threadIdx.x; k < e; k += blockDim.x) {
        const float *pk = c + 3 * (int64_t)vals[k];
        const float xij = __fsub_rn(pk[0], xi), yij = __fsub_rn(pk[1], yi), zij = __fsub_rn(pk[2], zi);
        const float dist = __fadd_rn(__fadd_rn(__fmul_rn(xij, xij), __fmul_rn(yij, yij)), __fmul_rn(zij, zij));   // crixus.cpp:320-323
        if ((double)dist < thr) rep[k] = (int32_t)i;
      }
      __syncthreads();
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) k_weld_out(const float *__restrict__ c, const uint32_t *__restrict__ vals, const int32_t *__restrict__ rep,
                                                  const int32_t *__restrict__ isrep, const int32_t *__restrict__ vid, int64_t n,
                                                  float4 *__restrict__ pos4, int32_t *__restrict__ ep4)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t o = vals[i];
    const int32_t v = vid[rep[i]];
    ep4[4 * (o / 3) + (o % 3)] = v;
    if (isrep[i]) pos4[v] = make_float4(c[3 * o], c[3 * o + 1], c[3 * o + 2], 0.f);
  }
}

__global__ void __launch_bounds__(256) k_ep_w0(int32_t *__restrict__ ep4, int64_t nbe)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nbe; i += (int64_t)gridDim.x * blockDim.x) ep4[4 * i + 3] = 0;
}

// corners_d: ncorner*3 floats (ncorner = 3*nbe, STL order).  pos4_d: room for ncorner float4 (the welded vertices come
// first; only nvert of them are written).  ep4_d: nbe int4.  bbmin/bbmax: bounding box of the corners (cx_bbox_device).
extern "C" int cx_weld_device(cx_ctx *ctx, const float *corners_d, int64_t ncorner, float dr, const float bbmin[3], const float bbmax[3],
                              float *pos4_d, int32_t *ep4_d, int64_t *nvert_host)
{
  if (!ctx || !corners_d || !pos4_d || !ep4_d || !nvert_host || ncorner <= 0 || ncorner % 3 || !(dr > 0)) return CX_WRONG_INPUT;
  if (ncorner > 0x7fffffffLL) return cx_fail(ctx, CX_WRONG_INPUT, "weld: more than 2^31 corners");
  WeldGrid g;
  g.dr = dr;
  int64_t dim[3];
  double cells = 1;
  for (int a = 0; a < 3; a++) {
    g.cmin[a] = (int32_t)(bbmin[a] / dr);            // the cell map (int)(x/dr) is monotone: every corner lies in [cmin, cmax]
    const int32_t cmax = (int32_t)(bbmax[a] / dr);
    dim[a] = (int64_t)cmax - g.cmin[a] + 1;
    cells *= (double)dim[a];
  }
  if (cells >= 9.0e18) return cx_fail(ctx, CX_WRONG_INPUT, "weld: cell grid too large for 64-bit keys");
  g.ny = dim[1]; g.nz = dim[2];
  int bits = 1;
  while (bits < 64 && ((double)(1ull << bits)) < cells) bits++;
  const int64_t n = ncorner;
  const int nblocks = (int)((n + RS_TILE - 1) / RS_TILE);
  // scratch layout
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t o_k0 = take(8 * (size_t)n), o_k1 = take(8 * (size_t)n), o_v0 = take(4 * (size_t)n), o_v1 = take(4 * (size_t)n),
               o_hist = take(4 * (size_t)256 * nblocks), o_head = take(4 * (size_t)(n + 1)), o_start = take(4 * (size_t)(n + 1)),
               o_rep = take(4 * (size_t)n), o_isrep = take(4 * (size_t)(n + 1));
  char *base = (char *)cx_scratch(ctx, off);
  if (!base) return CX_CUDA_ERROR;
  uint64_t *k0 = (uint64_t *)(base + o_k0), *k1 = (uint64_t *)(base + o_k1);
  uint32_t *v0 = (uint32_t *)(base + o_v0), *v1 = (uint32_t *)(base + o_v1);
  int32_t *hist = (int32_t *)(base + o_hist), *head = (int32_t *)(base + o_head), *start = (int32_t *)(base + o_start),
          *rep = (int32_t *)(base + o_rep), *isrep = (int32_t *)(base + o_isrep);
  cudaStream_t st = ctx->stream;
  const int gb = cx_blocks(n, 256, ctx->num_sms * 16);
  k_weld_keys<<<gb, 256, 0, st>>>(corners_d, n, g, k0, v0);
  CX_LAUNCH_CHECK(ctx, "k_weld_keys");
  uint64_t *ks;
  uint32_t *vs;
  int rc = radix_sort_pairs(ctx, k0, v0, k1, v1, n, bits, hist, &ks, &vs);
  if (rc != CX_OK) return rc;
  k_weld_heads<<<gb, 256, 0, st>>>(ks, n, head);
  CX_LAUNCH_CHECK(ctx, "k_weld_heads");
  rc = cx_exclusive_scan_i32(ctx, head, head, n + 1);   // head[i] = run index of position i; head[n] = number of runs
  if (rc != CX_OK) return rc;
  k_weld_starts<<<gb, 256, 0, st>>>(ks, head, n, start);
  CX_LAUNCH_CHECK(ctx, "k_weld_starts");
  CX_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail + 44, head + n, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CX_CUDA(ctx, cudaStreamSynchronize(st));
  const int64_t nruns = *(int32_t *)(ctx->h_mail + 44);
  int32_t maxrun = 0;
  {
    int32_t *mx = (int32_t *)(ctx->d_mail + 45);
    CX_CUDA(ctx, cudaMemsetAsync(mx, 0, sizeof(int32_t), st));
    k_weld_maxrun<<<cx_blocks(nruns, 256, ctx->num_sms * 8), 256, 0, st>>>(start, nruns, n, mx);
    CX_LAUNCH_CHECK(ctx, "k_weld_maxrun");
    CX_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail + 45, mx, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CX_CUDA(ctx, cudaStreamSynchronize(st));
    maxrun = *(int32_t *)(ctx->h_mail + 45);
  }
  const double thr = 1e-10 * dr * dr;                   // crixus.cpp:324 (double)
  k_weld_cells<<<cx_blocks(nruns, 128, ctx->num_sms * 32), 128, 0, st>>>(corners_d, vs, start, nruns, n, thr, rep, isrep);
  CX_LAUNCH_CHECK(ctx, "k_weld_cells");
  if (maxrun > CX_WELD_BIG_RUN) {   // over-full cells: one block each (no host fallback)
    k_weld_cells_big<<<cx_blocks(nruns, 1, ctx->num_sms * 8), 256, 0, st>>>(corners_d, vs, start, nruns, n, thr, rep, isrep);
    CX_LAUNCH_CHECK(ctx, "k_weld_cells_big");
  }
  CX_CUDA(ctx, cudaMemsetAsync(isrep + n, 0, sizeof(int32_t), st));
  int32_t *vid = head;                                  // reuse: exclusive scan of the representative flags = vertex ids
  rc = cx_exclusive_scan_i32(ctx, isrep, vid, n + 1);
  if (rc != CX_OK) return rc;
  k_weld_out<<<gb, 256, 0, st>>>(corners_d, vs, rep, isrep, vid, n, (float4 *)pos4_d, ep4_d);
  CX_LAUNCH_CHECK(ctx, "k_weld_out");
  k_ep_w0<<<cx_blocks(n / 3, 256, ctx->num_sms * 16), 256, 0, st>>>(ep4_d, n / 3);
  CX_LAUNCH_CHECK(ctx, "k_ep_w0");
  CX_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail + 44, vid + n, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CX_CUDA(ctx, cudaStreamSynchronize(st));
  *nvert_host = *(int32_t *)(ctx->h_mail + 44);
  return CX_OK;
}

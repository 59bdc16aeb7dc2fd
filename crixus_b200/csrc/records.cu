// crixus_b200 -- particle-record assembly and file bodies on the device.
// Reference: /root/reference/src/crixus.cu:973-1097 (OutBuf assembly: fluid particles in bit order, alive vertices in index
// order, segments stably grouped by KENT), /root/reference/src/crixus.cpp:12-60,62-259 (h5sph / vtu writers).
// The reference walks the bitfield bit by bit and the arrays element by element on one host thread, then converts the
// 64-byte records to the VTU's per-field arrays with one fwrite per value.  Here: popcount / alive-flag prefix sums give
// every particle its final index, one thread per bitfield word / vertex / segment writes its 64-byte records, a second
// kernel lays out the VTU appended-data body (12 arrays, each prefixed by its int32 byte count), and the bytes leave the
// GPU through two pinned staging buffers while the previous chunk is being written to the file.
// HBM-bound: 64 B written per particle (+84 B for the VTU body).
#include <algorithm>
#include <string>
#include <vector>
#include "common.cuh"
#include "host/output.h"

struct __align__(16) Rec {   // OutBuf, src/crixus.h:33-36
  float x, y, z, nx, ny, nz, vol, surf;
  int32_t kpar, kfluid, kent, kparmob, iref, ep1, ep2, ep3;
};
static_assert(sizeof(Rec) == 64, "OutBuf is 64 bytes");

__device__ __forceinline__ void store_rec(Rec *dst, const Rec &r)
{
  const float4 *s = reinterpret_cast<const float4 *>(&r);
  float4 *d = reinterpret_cast<float4 *>(dst);
  d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = s[3];
}

__global__ void __launch_bounds__(256) k_rec_flags(const uint32_t *__restrict__ fpos, int64_t nwords, int32_t *__restrict__ wcount,
                                                   const float4 *__restrict__ pos, int64_t nvert, int32_t *__restrict__ alive)
{
  const int64_t n = nwords > nvert ? nwords : nvert;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i <= n; i += (int64_t)gridDim.x * blockDim.x) {
    if (i <= nwords) wcount[i] = i < nwords ? __popc(fpos[i]) : 0;
    if (i <= nvert) alive[i] = (i < nvert && !(pos[i].x < -1e9f)) ? 1 : 0;   // crixus.cu:1027
  }
}

// fluid particles, crixus.cu:991-1019: one warp per bitfield word, one lane per bit -- the set lanes of a warp write a contiguous
// run of records (rank inside the word = popcount of the lower set bits)
__global__ void __launch_bounds__(256) k_rec_fluid(const uint32_t *__restrict__ fpos, const int32_t *__restrict__ wbase, int64_t nwords,
                                                   int64_t d0, int64_t d1, float3 fcmin, float dr, float fluid_vol, Rec *__restrict__ rec)
{
  const int lane = threadIdx.x & 31;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; w < nwords; w += nwarp) {
    const uint32_t bits = __ldg(fpos + w);
    if (!(bits >> lane & 1u)) continue;
    const int32_t k = __ldg(wbase + w) + __popc(bits & ((1u << lane) - 1u));
    const int64_t j = w * 32 + lane;
    const int64_t m = j / (d1 * d0), n = j % (d1 * d0);
    Rec r = {};
    r.z = fadd(fcmin.z, fmul(dr, (float)m));          // host code of the reference: no fused multiply-add
    r.y = fadd(fcmin.y, fmul(dr, (float)(n / d0)));
    r.x = fadd(fcmin.x, fmul(dr, (float)(n % d0)));
    r.vol = fluid_vol;
    r.kpar = 1; r.kfluid = 1;
    r.iref = k;
    store_rec(rec + k, r);
  }
}

// vertex particles, crixus.cu:1026-1053
__global__ void __launch_bounds__(256) k_rec_vertex(const float4 *__restrict__ pos, const float *__restrict__ vol,
                                                    const int32_t *__restrict__ sbid, const int32_t *__restrict__ abase, int64_t nvert,
                                                    int32_t nfluid_rec, Rec *__restrict__ rec)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvert; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 p = pos[i];
    if (p.x < -1e9f) continue;
    const int32_t k = nfluid_rec + abase[i];
    Rec r = {};
    r.x = p.x; r.y = p.y; r.z = p.z;
    r.vol = vol[i];
    r.kpar = 2; r.kfluid = 1;
    r.kent = sbid ? sbid[i] : 0;
    r.iref = k;
    store_rec(rec + k, r);
  }
}

// boundary segments, crixus.cu:1056-1097: slot[i] = position among the segments after the stable grouping by KENT
__global__ void __launch_bounds__(256) k_rec_segment(const float4 *__restrict__ pos, const float4 *__restrict__ norm,
                                                     const int4 *__restrict__ ep, const float *__restrict__ surf,
                                                     const int32_t *__restrict__ sbid, const int32_t *__restrict__ slot,
                                                     const int32_t *__restrict__ abase, int64_t nvert, int64_t nbe, int32_t ncur,
                                                     int32_t nfluid_idx, Rec *__restrict__ rec)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nbe; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 p = pos[nvert + i], n = norm[i];
    const int4 e = ep[i];
    const int32_t l = ncur + (slot ? slot[i] : (int32_t)i);
    Rec r = {};
    r.x = p.x; r.y = p.y; r.z = p.z;
    r.nx = n.x; r.ny = n.y; r.nz = n.z;
    r.surf = surf[i];
    r.kpar = 3; r.kfluid = 1;
    r.kent = sbid ? sbid[nvert + i] : 0;
    r.iref = l;
    // ep - nvshift[ep] (crixus.cu:1082-1084): the number of alive vertices before ep for an alive vertex; nvshift of a linked
    // (deleted) vertex is 0, should one ever be referenced
    const int ax = abase[e.x], ay = abase[e.y], az = abase[e.z];
    r.ep1 = nfluid_idx + (abase[e.x + 1] != ax ? ax : e.x);
    r.ep2 = nfluid_idx + (abase[e.y + 1] != ay ? ay : e.y);
    r.ep3 = nfluid_idx + (abase[e.z + 1] != az ? az : e.z);
    store_rec(rec + l, r);
  }
}

// chunked variants for the streamed writer (cx_stream_h5sph_device): records are written relative to the chunk's first record
__global__ void __launch_bounds__(256) k_rec_fluid_chunk(const uint32_t *__restrict__ fpos, const int32_t *__restrict__ wbase, int64_t w0, int64_t w1,
                                                         int64_t d0, int64_t d1, float3 fcmin, float dr, float fluid_vol, int32_t rec0,
                                                         Rec *__restrict__ rec)
{
  const int lane = threadIdx.x & 31;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t w = w0 + ((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5); w < w1; w += nwarp) {
    const uint32_t bits = __ldg(fpos + w);
    if (!(bits >> lane & 1u)) continue;
    const int32_t k = __ldg(wbase + w) + __popc(bits & ((1u << lane) - 1u));
    const int64_t j = w * 32 + lane;
    const int64_t m = j / (d1 * d0), n = j % (d1 * d0);
    Rec r = {};
    r.z = fadd(fcmin.z, fmul(dr, (float)m));
    r.y = fadd(fcmin.y, fmul(dr, (float)(n / d0)));
    r.x = fadd(fcmin.x, fmul(dr, (float)(n % d0)));
    r.vol = fluid_vol;
    r.kpar = 1; r.kfluid = 1;
    r.iref = k;
    store_rec(rec + (k - rec0), r);
  }
}

__global__ void __launch_bounds__(256) k_rec_segment_chunk(const float4 *__restrict__ pos, const float4 *__restrict__ norm,
                                                           const int4 *__restrict__ ep, const float *__restrict__ surf,
                                                           const int32_t *__restrict__ abase, int64_t nvert, int64_t s0, int64_t s1, int32_t ncur,
                                                           int32_t nfluid_idx, Rec *__restrict__ rec)
{
  for (int64_t i = s0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < s1; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 p = pos[nvert + i], n = norm[i];
    const int4 e = ep[i];
    Rec r = {};
    r.x = p.x; r.y = p.y; r.z = p.z;
    r.nx = n.x; r.ny = n.y; r.nz = n.z;
    r.surf = surf[i];
    r.kpar = 3; r.kfluid = 1;
    r.iref = ncur + (int32_t)i;
    const int ax = abase[e.x], ay = abase[e.y], az = abase[e.z];
    r.ep1 = nfluid_idx + (abase[e.x + 1] != ax ? ax : e.x);
    r.ep2 = nfluid_idx + (abase[e.y + 1] != ay ? ay : e.y);
    r.ep3 = nfluid_idx + (abase[e.z + 1] != az ? az : e.z);
    store_rec(rec + (i - s0), r);
  }
}

__global__ void __launch_bounds__(256) k_kent_max(const int32_t *__restrict__ kent, int64_t n, int32_t *__restrict__ out)
{
  int m = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = max(m, kent[i]);
  m = __reduce_max_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}
__global__ void __launch_bounds__(256) k_kent_flag(const int32_t *__restrict__ kent, int64_t n, int32_t k, int32_t *__restrict__ flag)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i <= n; i += (int64_t)gridDim.x * blockDim.x)
    flag[i] = (i < n && kent[i] == k) ? 1 : 0;
}
__global__ void __launch_bounds__(256) k_kent_place(const int32_t *__restrict__ kent, const int32_t *__restrict__ rank, int64_t n, int32_t k,
                                                    int32_t base, int32_t *__restrict__ slot)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (kent[i] == k) slot[i] = base + rank[i];
}

extern "C" int cx_assemble_records_device(cx_ctx *ctx, const cx_params *pfill, const int64_t dimg[3], const uint32_t *fpos_d,
                                          const float *pos_d, const float *norm_d, const int32_t *ep_d, const float *surf_d,
                                          const float *vol_d, const int32_t *sbid_d, int64_t nvert, int64_t nbe,
                                          int64_t nfluid_for_indices, void *records_d, int64_t capacity, int64_t *n_host)
{
  if (!ctx || !pfill || !dimg || !fpos_d || !pos_d || !norm_d || !ep_d || !surf_d || !vol_d || !n_host || nvert < 0 || nbe < 0)
    return CX_WRONG_INPUT;
  const int64_t G = dimg[0] * dimg[1] * dimg[2], nwords = (G + 31) / 32;
  if (nwords + 1 >= (int64_t)1 << 31 || nvert + 1 >= (int64_t)1 << 31 || nbe + 1 >= (int64_t)1 << 31) return CX_WRONG_INPUT;
  cudaStream_t st = ctx->stream;
  int32_t *wbase = nullptr, *abase = nullptr, *slot = nullptr, *flag = nullptr;
  CX_CUDA(ctx, cudaMallocAsync(&wbase, 4 * (size_t)(nwords + 1), st));
  CX_CUDA(ctx, cudaMallocAsync(&abase, 4 * (size_t)(nvert + 1), st));
  const int grid = ctx->num_sms * 8;
  k_rec_flags<<<grid, 256, 0, st>>>(fpos_d, nwords, wbase, (const float4 *)pos_d, nvert, abase);
  CX_LAUNCH_CHECK(ctx, "k_rec_flags");
  int rc = cx_exclusive_scan_i32(ctx, wbase, wbase, nwords + 1);
  if (rc == CX_OK) rc = cx_exclusive_scan_i32(ctx, abase, abase, nvert + 1);
  if (rc != CX_OK) return rc;
  int32_t tot[2] = {0, 0}, kmax = 0;
  CX_CUDA(ctx, cudaMemcpyAsync(&tot[0], wbase + nwords, 4, cudaMemcpyDeviceToHost, st));
  CX_CUDA(ctx, cudaMemcpyAsync(&tot[1], abase + nvert, 4, cudaMemcpyDeviceToHost, st));
  if (sbid_d) {
    CX_CUDA(ctx, cudaMemsetAsync(ctx->d_mail, 0, 4, st));
    k_kent_max<<<grid, 256, 0, st>>>(sbid_d + nvert, nbe, (int32_t *)ctx->d_mail);
    CX_LAUNCH_CHECK(ctx, "k_kent_max");
    CX_CUDA(ctx, cudaMemcpyAsync(&kmax, ctx->d_mail, 4, cudaMemcpyDeviceToHost, st));
  }
  CX_CUDA(ctx, cudaStreamSynchronize(st));
  const int64_t nfl = (uint32_t)tot[0], nalive = tot[1], n = nfl + nalive + nbe;
  *n_host = n;
  if (n >= (int64_t)1 << 31 || nfluid_for_indices + nvert >= (int64_t)1 << 31) {   // 32-bit particle indices (AbsoluteIndex,
    cudaFreeAsync(wbase, st); cudaFreeAsync(abase, st);                              // VertexParticle are ints in both formats)
    return cx_fail(ctx, CX_WRONG_INPUT, "records: more than 2^31 particles");
  }
  if (!records_d) { cudaFreeAsync(wbase, st); cudaFreeAsync(abase, st); return CX_OK; }   // sizing call
  if (capacity < n) { cudaFreeAsync(wbase, st); cudaFreeAsync(abase, st); return CX_WRONG_INPUT; }
  if (sbid_d && kmax > 0) {   // stable grouping by KENT: one flag-scan per id (ids are few: the number of special-boundary grids)
    CX_CUDA(ctx, cudaMallocAsync(&slot, 4 * (size_t)(nbe + 1), st));
    CX_CUDA(ctx, cudaMallocAsync(&flag, 4 * (size_t)(nbe + 1), st));
    int32_t base = 0;
    for (int32_t k = 0; k <= kmax; k++) {
      k_kent_flag<<<grid, 256, 0, st>>>(sbid_d + nvert, nbe, k, flag);
      CX_LAUNCH_CHECK(ctx, "k_kent_flag");
      rc = cx_exclusive_scan_i32(ctx, flag, flag, nbe + 1);
      if (rc != CX_OK) return rc;
      k_kent_place<<<grid, 256, 0, st>>>(sbid_d + nvert, flag, nbe, k, base, slot);
      CX_LAUNCH_CHECK(ctx, "k_kent_place");
      int32_t cnt = 0;
      CX_CUDA(ctx, cudaMemcpyAsync(&cnt, flag + nbe, 4, cudaMemcpyDeviceToHost, st));
      CX_CUDA(ctx, cudaStreamSynchronize(st));
      base += cnt;
    }
  }
  Rec *rec = (Rec *)records_d;
  const float dr = pfill->dr;
  const float fluid_vol = (float)pow(dr, 3);   // crixus.cu:986 (double pow, rounded to float)
  if (nwords > 0) {
    k_rec_fluid<<<grid, 256, 0, st>>>(fpos_d, wbase, nwords, dimg[0], dimg[1], make_float3(pfill->fcmin[0], pfill->fcmin[1], pfill->fcmin[2]),
                                      dr, fluid_vol, rec);
    CX_LAUNCH_CHECK(ctx, "k_rec_fluid");
  }
  if (nvert > 0) {
    k_rec_vertex<<<grid, 256, 0, st>>>((const float4 *)pos_d, vol_d, sbid_d, abase, nvert, (int32_t)nfl, rec);
    CX_LAUNCH_CHECK(ctx, "k_rec_vertex");
  }
  if (nbe > 0) {
    k_rec_segment<<<grid, 256, 0, st>>>((const float4 *)pos_d, (const float4 *)norm_d, (const int4 *)ep_d, surf_d, sbid_d, slot, abase, nvert,
                                        nbe, (int32_t)(nfl + nalive), (int32_t)nfluid_for_indices, rec);
    CX_LAUNCH_CHECK(ctx, "k_rec_segment");
  }
  CX_CUDA(ctx, cudaFreeAsync(wbase, st));
  CX_CUDA(ctx, cudaFreeAsync(abase, st));
  if (slot) CX_CUDA(ctx, cudaFreeAsync(slot, st));
  if (flag) CX_CUDA(ctx, cudaFreeAsync(flag, st));
  return CX_OK;
}

// ------------------------------------------------------------------------------------------------ VTU appended-data body
// crixus.cpp:149-252: Volume, Surface, ParticleType, FluidType, KENT, MovingBoundary, AbsoluteIndex (4 B each), Normal,
// VertexParticle (12 B), Points (3 doubles), connectivity, offsets (4 B); each array preceded by its int32 byte count.
// Every offset is a multiple of 4, so the body is written as 32-bit words (doubles as two words).
struct VtuOff { int64_t o[12]; };   // word offsets of the 12 arrays' first value

__global__ void __launch_bounds__(256) k_vtu_body(const Rec *__restrict__ rec, int64_t n, VtuOff off, uint32_t *__restrict__ body)
{
  if (blockIdx.x == 0 && threadIdx.x < 12) {
    const int elem[12] = {4, 4, 4, 4, 4, 4, 4, 12, 12, 24, 4, 4};
    body[off.o[threadIdx.x] - 1] = (uint32_t)(int)((int64_t)elem[threadIdx.x] * n);   // int numbytes, as in the reference
  }
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 *s = reinterpret_cast<const float4 *>(rec + i);
    const float4 a = s[0], b = s[1];
    const int4 c = reinterpret_cast<const int4 *>(s)[2], d = reinterpret_cast<const int4 *>(s)[3];
    body[off.o[0] + i] = __float_as_uint(b.z);   // vol
    body[off.o[1] + i] = __float_as_uint(b.w);   // surf
    body[off.o[2] + i] = (uint32_t)c.x;          // kpar
    body[off.o[3] + i] = (uint32_t)c.y;          // kfluid
    body[off.o[4] + i] = (uint32_t)c.z;          // kent
    body[off.o[5] + i] = (uint32_t)c.w;          // kparmob
    body[off.o[6] + i] = (uint32_t)d.x;          // iref
    uint32_t *nn = body + off.o[7] + 3 * i;
    nn[0] = __float_as_uint(a.w); nn[1] = __float_as_uint(b.x); nn[2] = __float_as_uint(b.y);
    uint32_t *ee = body + off.o[8] + 3 * i;
    ee[0] = (uint32_t)d.y; ee[1] = (uint32_t)d.z; ee[2] = (uint32_t)d.w;
    uint32_t *pp = body + off.o[9] + 6 * i;
    const double px = (double)a.x, py = (double)a.y, pz = (double)a.z;
    pp[0] = (uint32_t)__double2loint(px); pp[1] = (uint32_t)__double2hiint(px);
    pp[2] = (uint32_t)__double2loint(py); pp[3] = (uint32_t)__double2hiint(py);
    pp[4] = (uint32_t)__double2loint(pz); pp[5] = (uint32_t)__double2hiint(pz);
    body[off.o[10] + i] = (uint32_t)i;           // connectivity
    body[off.o[11] + i] = (uint32_t)(i + 1);     // offsets
  }
}

// device bytes -> FILE through two pinned staging buffers: the copy of chunk c+1 runs while chunk c is written
static int stream_to_file(cx_ctx *ctx, FILE *f, const void *dev, size_t bytes)
{
  const size_t CH = (size_t)16 << 20;
  char *stage = (char *)cx_scratch_pinned(ctx, 2 * CH);
  if (!stage) return cx_fail(ctx, CX_CUDA_ERROR, "pinned staging");
  cudaEvent_t ev[2];
  for (int k = 0; k < 2; k++) CX_CUDA(ctx, cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming));
  const size_t nch = (bytes + CH - 1) / CH;
  bool ok = true;
  for (size_t c = 0; c <= nch; c++) {
    if (c < nch) {
      const size_t o = c * CH, m = std::min(CH, bytes - o);
      if (cudaMemcpyAsync(stage + (c & 1) * CH, (const char *)dev + o, m, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) ok = false;
      cudaEventRecord(ev[c & 1], ctx->stream);
    }
    if (c > 0) {
      const size_t o = (c - 1) * CH, m = std::min(CH, bytes - o);
      if (cudaEventSynchronize(ev[(c - 1) & 1]) != cudaSuccess) ok = false;
      if (ok && fwrite(stage + ((c - 1) & 1) * CH, 1, m, f) != m) ok = false;
    }
  }
  for (int k = 0; k < 2; k++) cudaEventDestroy(ev[k]);
  return ok ? CX_OK : CX_WRITE_FAIL;
}

extern "C" int cx_write_vtu_records_device(cx_ctx *ctx, const void *records_d, int64_t n, const char *filename)
{
  if (!ctx || (!records_d && n > 0) || !filename || n < 0 || n >= (int64_t)1 << 31) return CX_WRONG_INPUT;
  FILE *f = fopen(filename, "w");
  if (!f) return CX_NO_WRITE_FILE;
  setvbuf(f, nullptr, _IOFBF, 8 << 20);
  const std::string head = cx_vtu_header((int)n);
  bool ok = fwrite(head.data(), 1, head.size(), f) == head.size();
  const int elem[12] = {1, 1, 1, 1, 1, 1, 1, 3, 3, 6, 1, 1};   // words per particle
  VtuOff off;
  int64_t w = 0;
  for (int k = 0; k < 12; k++) { off.o[k] = w + 1; w += 1 + elem[k] * n; }
  uint32_t *body = nullptr;
  int rc = CX_OK;
  if (cudaMallocAsync(&body, 4 * (size_t)w, ctx->stream) != cudaSuccess) { fclose(f); return cx_fail(ctx, CX_CUDA_ERROR, "vtu body"); }
  k_vtu_body<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>((const Rec *)records_d, n, off, body);
  ctx->launches++;
  if (cudaGetLastError() != cudaSuccess) rc = cx_fail(ctx, CX_CUDA_ERROR, "k_vtu_body");
  if (rc == CX_OK) rc = stream_to_file(ctx, f, body, 4 * (size_t)w);
  cudaFreeAsync(body, ctx->stream);
  if (rc == CX_OK && ok) fputs(cx_vtu_trailer(), f);
  fclose(f);
  return rc != CX_OK ? rc : (ok ? CX_OK : CX_WRITE_FAIL);
}

extern "C" int cx_write_h5sph_records_device(cx_ctx *ctx, const void *records_d, int64_t n, const char *filename)
{
  if (!ctx || (!records_d && n > 0) || !filename || n < 0) return CX_WRONG_INPUT;
  std::vector<uint8_t> head;
  int rc = cx_h5sph_header((uint64_t)n, head);
  if (rc != CX_OK) return rc;
  FILE *f = fopen(filename, "wb");
  if (!f) return CX_NO_WRITE_FILE;
  setvbuf(f, nullptr, _IOFBF, 8 << 20);
  const bool ok = fwrite(head.data(), 1, head.size(), f) == head.size();
  if (n > 0) rc = stream_to_file(ctx, f, records_d, 64 * (size_t)n);
  fclose(f);
  return rc != CX_OK ? rc : (ok ? CX_OK : CX_WRITE_FAIL);
}


// ------------------------------------------------------------------------------------------------ streamed .h5sph (large runs)
// cx_assemble_records_device needs all n records resident (64 B each: 64 GB at BASELINE configs[3]).  The .h5sph body is just the
// records in order -- fluid particles in bit order, vertices, segments -- so it can be produced and written in CHUNKS: the prefix
// sums over the bitfield popcounts / alive flags give every chunk its first record index, a chunk of bitfield words (or of
// vertices / segments) is assembled into a bounded device buffer and leaves through the pinned staging while the next one is
// being assembled.  Same bytes as cx_assemble_records_device + cx_write_h5sph_records_device (tests/test_gpu_pipeline.py).
// Not available with special boundaries (the stable grouping by KENT permutes the segment records): CX_WRONG_INPUT.
__global__ void __launch_bounds__(256) k_gather_stride(const int32_t *__restrict__ src, int64_t stride, int64_t n, int64_t limit,
                                                       int32_t *__restrict__ dst)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = i * stride;
    dst[i] = src[j < limit ? j : limit];
  }
}

extern "C" int cx_stream_h5sph_device(cx_ctx *ctx, const cx_params *pfill, const int64_t dimg[3], const uint32_t *fpos_d,
                                      const float *pos_d, const float *norm_d, const int32_t *ep_d, const float *surf_d,
                                      const float *vol_d, int64_t nvert, int64_t nbe, int64_t nfluid_for_indices, const char *filename,
                                      int64_t chunk_items, int64_t *n_host)
{
  if (!ctx || !pfill || !dimg || !fpos_d || !pos_d || !norm_d || !ep_d || !surf_d || !vol_d || !filename || nvert < 0 || nbe < 0)
    return CX_WRONG_INPUT;
  const int64_t G = dimg[0] * dimg[1] * dimg[2], nwords = (G + 31) / 32;
  if (nwords + 1 >= (int64_t)1 << 31 || nvert + 1 >= (int64_t)1 << 31 || nbe + 1 >= (int64_t)1 << 31) return CX_WRONG_INPUT;
  if (chunk_items <= 0) chunk_items = (int64_t)1 << 20;       // bitfield words / vertices / segments per chunk
  cudaStream_t st = ctx->stream;
  const int grid = ctx->num_sms * 8;
  int32_t *wbase = nullptr, *abase = nullptr, *cstart = nullptr;
  Rec *buf = nullptr;
  const int64_t nch_w = (nwords + chunk_items - 1) / chunk_items;
  CX_CUDA(ctx, cudaMallocAsync(&wbase, 4 * (size_t)(nwords + 1), st));
  CX_CUDA(ctx, cudaMallocAsync(&abase, 4 * (size_t)(nvert + 1), st));
  CX_CUDA(ctx, cudaMallocAsync(&cstart, 4 * (size_t)(nch_w + 2), st));
  CX_CUDA(ctx, cudaMallocAsync(&buf, sizeof(Rec) * (size_t)(32 * chunk_items), st));   // a chunk of words holds <= 32 records per word
  struct Cleanup {
    cudaStream_t st; void *a, *b, *c, *d; FILE *f = nullptr;
    ~Cleanup() { cudaFreeAsync(a, st); cudaFreeAsync(b, st); cudaFreeAsync(c, st); cudaFreeAsync(d, st); if (f) fclose(f); }
  } cl{st, wbase, abase, cstart, buf};
  k_rec_flags<<<grid, 256, 0, st>>>(fpos_d, nwords, wbase, (const float4 *)pos_d, nvert, abase);
  CX_LAUNCH_CHECK(ctx, "k_rec_flags");
  int rc = cx_exclusive_scan_i32(ctx, wbase, wbase, nwords + 1);
  if (rc == CX_OK) rc = cx_exclusive_scan_i32(ctx, abase, abase, nvert + 1);
  if (rc != CX_OK) return rc;
  k_gather_stride<<<cx_blocks(nch_w + 1, 256), 256, 0, st>>>(wbase, chunk_items, nch_w + 1, nwords, cstart);   // first record of every chunk
  CX_LAUNCH_CHECK(ctx, "k_gather_stride");
  std::vector<int32_t> hstart((size_t)nch_w + 1);
  int32_t nalive32 = 0;
  CX_CUDA(ctx, cudaMemcpyAsync(hstart.data(), cstart, 4 * (size_t)(nch_w + 1), cudaMemcpyDeviceToHost, st));
  CX_CUDA(ctx, cudaMemcpyAsync(&nalive32, abase + nvert, 4, cudaMemcpyDeviceToHost, st));
  CX_CUDA(ctx, cudaStreamSynchronize(st));
  const int64_t nfl = (uint32_t)hstart[(size_t)nch_w], nalive = nalive32, n = nfl + nalive + nbe;
  if (n_host) *n_host = n;
  if (n >= (int64_t)1 << 31 || nfluid_for_indices + nvert >= (int64_t)1 << 31)
    return cx_fail(ctx, CX_WRONG_INPUT, "records: more than 2^31 particles (the file formats index particles with 32-bit integers)");
  std::vector<uint8_t> head;
  rc = cx_h5sph_header((uint64_t)n, head);
  if (rc != CX_OK) return rc;
  cl.f = fopen(filename, "wb");
  if (!cl.f) return CX_NO_WRITE_FILE;
  setvbuf(cl.f, nullptr, _IOFBF, 8 << 20);
  if (fwrite(head.data(), 1, head.size(), cl.f) != head.size()) return CX_WRITE_FAIL;
  const float dr = pfill->dr;
  const float fluid_vol = (float)pow(dr, 3);   // crixus.cu:986
  // ---- fluid particles, a chunk of bitfield words at a time (records rebased to the chunk's first record)
  for (int64_t c = 0; c < nch_w; c++) {
    const int64_t w0 = c * chunk_items, w1 = std::min(nwords, w0 + chunk_items);
    const int64_t r0 = (uint32_t)hstart[(size_t)c], r1 = (uint32_t)hstart[(size_t)c + 1];
    if (r1 == r0) continue;
    k_rec_fluid_chunk<<<grid, 256, 0, st>>>(fpos_d, wbase, w0, w1, dimg[0], dimg[1], make_float3(pfill->fcmin[0], pfill->fcmin[1], pfill->fcmin[2]),
                                            dr, fluid_vol, (int32_t)r0, buf);
    CX_LAUNCH_CHECK(ctx, "k_rec_fluid_chunk");
    rc = stream_to_file(ctx, cl.f, buf, sizeof(Rec) * (size_t)(r1 - r0));
    if (rc != CX_OK) return rc;
  }
  // ---- vertex particles: chunks of vertex indices (alive vertices of a chunk are consecutive records)
  std::vector<int32_t> two(2);
  for (int64_t v0 = 0; v0 < nvert; v0 += 32 * chunk_items) {
    const int64_t v1 = std::min(nvert, v0 + 32 * chunk_items);
    CX_CUDA(ctx, cudaMemcpyAsync(&two[0], abase + v0, 4, cudaMemcpyDeviceToHost, st));
    CX_CUDA(ctx, cudaMemcpyAsync(&two[1], abase + v1, 4, cudaMemcpyDeviceToHost, st));
    CX_CUDA(ctx, cudaStreamSynchronize(st));
    if (two[1] == two[0]) continue;
    k_rec_vertex<<<grid, 256, 0, st>>>((const float4 *)pos_d + v0, vol_d + v0, nullptr, abase + v0, v1 - v0, (int32_t)nfl, buf - (nfl + two[0]));
    CX_LAUNCH_CHECK(ctx, "k_rec_vertex");
    rc = stream_to_file(ctx, cl.f, buf, sizeof(Rec) * (size_t)(two[1] - two[0]));
    if (rc != CX_OK) return rc;
  }
  // ---- boundary segments: chunks of segment indices
  for (int64_t s0 = 0; s0 < nbe; s0 += 32 * chunk_items) {
    const int64_t s1 = std::min(nbe, s0 + 32 * chunk_items);
    k_rec_segment_chunk<<<grid, 256, 0, st>>>((const float4 *)pos_d, (const float4 *)norm_d, (const int4 *)ep_d, surf_d, abase, nvert, s0, s1,
                                              (int32_t)(nfl + nalive), (int32_t)nfluid_for_indices, buf);
    CX_LAUNCH_CHECK(ctx, "k_rec_segment_chunk");
    rc = stream_to_file(ctx, cl.f, buf, sizeof(Rec) * (size_t)(s1 - s0));
    if (rc != CX_OK) return rc;
  }
  return CX_OK;
}

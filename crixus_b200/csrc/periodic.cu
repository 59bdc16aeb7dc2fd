// crixus_b200 -- periodic vertex pairing (find_links) and segment relinking (periodicity_links), plus calc_trisize.
// Reference: /root/reference/src/crixus_d.cu:213-267, driver loop /root/reference/src/crixus.cu:387-435.
// The reference scans ALL vertices for every max-face vertex (O(n_face * nvert)); here both faces are compacted first
// (flag -> scan -> scatter), the min face is bucketed on a 2-D grid, and every max-face vertex tests the 3 x 3 buckets around
// it -- O(n_face) -- keeping the reference's rule "lowest-index partner wins".
#include "common.cuh"

// face membership.  max face: fabs(x_d - dmax_d) < 1e-4*dr  (double compare, crixus_d.cu:217).
// min face candidates: a partner must satisfy sqrt(dy^2+dz^2+(x_d-dmin_d)^2) < 1e-4*dr, hence |x_d-dmin_d| <= 2e-4*dr
// is a safe superset; the exact test is applied in k_pair.
__global__ void __launch_bounds__(256) k_face_flags(const float4 *__restrict__ pos, int64_t nvert, int idim, float dmin, float dmax,
                                                    double thr, int32_t *__restrict__ fmax, int32_t *__restrict__ fmin)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvert; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 q = pos[i];
    const float x = idim == 0 ? q.x : (idim == 1 ? q.y : q.z);
    fmax[i] = ((double)fabsf(fsub(x, dmax)) < thr) ? 1 : 0;
    fmin[i] = ((double)fabsf(fsub(x, dmin)) <= 2.0 * thr) ? 1 : 0;
  }
}

__global__ void __launch_bounds__(256) k_compact(const int32_t *__restrict__ flag, const int32_t *__restrict__ offs, int64_t n,
                                                 int32_t *__restrict__ list)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (flag[i]) list[offs[i]] = (int32_t)i;
}

// The min-face vertices are bucketed on a uniform 2-D grid over the two in-face coordinates (cell edge >= dr/2, i.e. thousands
// of pairing radii): a partner lies within 1e-4 dr of the max-face vertex in both coordinates, hence in the same or an
// adjacent bucket.  count -> scan -> scatter; the order inside a bucket is irrelevant because the pairing keeps the LOWEST
// index among the vertices that pass the reference's test (crixus_d.cu:218-228 scans in index order and stops at the first).
struct PairGrid {
  float o1, o2, inv;
  int n1, n2;
};
__device__ __forceinline__ int pg_coord(float a, float o, float inv, int n)
{
  const int c = (int)((a - o) * inv);
  return max(min(c, n - 1), 0);
}
__device__ __forceinline__ float comp4(float4 q, int d) { return d == 0 ? q.x : (d == 1 ? q.y : q.z); }

__global__ void __launch_bounds__(256) k_pg_count(const float4 *__restrict__ pos, const int32_t *__restrict__ lmin, int nmin, int idim,
                                                  PairGrid g, int32_t *__restrict__ cnt)
{
  const int d1 = (idim + 1) % 3, d2 = (idim + 2) % 3;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < nmin; s += gridDim.x * blockDim.x) {
    const float4 q = __ldg(&pos[lmin[s]]);
    atomicAdd(&cnt[pg_coord(comp4(q, d1), g.o1, g.inv, g.n1) * g.n2 + pg_coord(comp4(q, d2), g.o2, g.inv, g.n2)], 1);
  }
}

__global__ void __launch_bounds__(256) k_pg_scatter(const float4 *__restrict__ pos, const int32_t *__restrict__ lmin, int nmin, int idim,
                                                    PairGrid g, const int32_t *__restrict__ start, int32_t *__restrict__ cursor,
                                                    int32_t *__restrict__ bucket)
{
  const int d1 = (idim + 1) % 3, d2 = (idim + 2) % 3;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < nmin; s += gridDim.x * blockDim.x) {
    const int j = lmin[s];
    const float4 q = __ldg(&pos[j]);
    const int c = pg_coord(comp4(q, d1), g.o1, g.inv, g.n1) * g.n2 + pg_coord(comp4(q, d2), g.o2, g.inv, g.n2);
    bucket[start[c] + atomicAdd(&cursor[c], 1)] = j;
  }
}

// one thread per max-face vertex: the reference's test (crixus_d.cu:218-228) against the min-face vertices of the 3 x 3
// buckets around it; lowest index wins
__global__ void __launch_bounds__(128) k_pair(const float4 *__restrict__ pos, const int32_t *__restrict__ lmax, int nmax, PairGrid g,
                                              const int32_t *__restrict__ start, const int32_t *__restrict__ bucket, int idim, float dmin,
                                              double thr, int32_t *__restrict__ newlink, int64_t *__restrict__ missing)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nmax) return;
  const int i = lmax[t];
  const float4 pi = pos[i];
  const int d1 = (idim + 1) % 3, d2 = (idim + 2) % 3;
  const float pi1 = comp4(pi, d1), pi2 = comp4(pi, d2);
  const int c1 = pg_coord(pi1, g.o1, g.inv, g.n1), c2 = pg_coord(pi2, g.o2, g.inv, g.n2);
  int hit = 0x7fffffff;
  for (int a1 = max(c1 - 1, 0); a1 <= min(c1 + 1, g.n1 - 1); a1++)
    for (int a2 = max(c2 - 1, 0); a2 <= min(c2 + 1, g.n2 - 1); a2++) {
      const int c = a1 * g.n2 + a2;
      for (int s = start[c]; s < start[c + 1]; s++) {
        const int j = bucket[s];
        if (j == i || j >= hit) continue;
        const float4 pj = __ldg(&pos[j]);
        const float a = fsub(pi1, comp4(pj, d1)), b = fsub(pi2, comp4(pj, d2)), c0 = fsub(comp4(pj, idim), dmin);
        // pow(x,(float)2.) of crixus_d.cu:220-222 is an inlined powf; restated as x*x (DESIGN.md "Numerics")
        const float d = __fsqrt_rn(fadd(fadd(fmul(a, a), fmul(b, b)), fmul(c0, c0)));
        if ((double)d < thr) hit = j;
      }
    }
  if (hit != 0x7fffffff) newlink[i] = hit;
  else atomicAdd((unsigned long long *)missing, 1ull);
}

// "delete" paired vertices: pos = -1e10 (crixus_d.cu:225-226); separate pass so that k_pair reads unmodified positions
__global__ void __launch_bounds__(128) k_delete_linked(float4 *__restrict__ pos, const int32_t *__restrict__ lmax, int nmax,
                                                       const int32_t *__restrict__ newlink)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nmax) return;
  const int i = lmax[t];
  if (newlink[i] != -1) {
    float4 q = pos[i];
    q.x = -1e10f; q.y = -1e10f; q.z = -1e10f;
    pos[i] = q;
  }
}

extern "C" int cx_find_links(cx_ctx *ctx, const cx_params *p, float *pos_d, int64_t nvert, int32_t *newlink_d, int idim)
{
  if (!ctx || !p || !pos_d || !newlink_d || idim < 0 || idim > 2 || nvert < 0) return CX_WRONG_INPUT;
  if (nvert == 0) return CX_OK;
  cudaStream_t st = ctx->stream;
  const double thr = 1e-4 * p->dr;  // double product of a double literal and float dr (crixus_d.cu:217)
  int32_t *fmax, *fmin, *omax, *omin, *lmax, *lmin;
  const int64_t stride = (nvert + 1 + 63) & ~(int64_t)63;  // keep every sub-array 16-byte aligned (vectorised scan)
  CX_CUDA(ctx, cudaMallocAsync(&fmax, sizeof(int32_t) * (size_t)stride * 6, st));
  fmin = fmax + stride; omax = fmin + stride; omin = omax + stride; lmax = omin + stride; lmin = lmax + stride;
  CX_CUDA(ctx, cudaMemsetAsync(fmax, 0, sizeof(int32_t) * (size_t)stride * 2, st));
  const int grid = cx_blocks(nvert, 256, ctx->num_sms * 16);
  k_face_flags<<<grid, 256, 0, st>>>((const float4 *)pos_d, nvert, idim, p->dmin[idim], p->dmax[idim], thr, fmax, fmin);
  CX_LAUNCH_CHECK(ctx, "k_face_flags");
  int rc = cx_exclusive_scan_i32(ctx, fmax, omax, nvert + 1);
  if (rc != CX_OK) return rc;
  rc = cx_exclusive_scan_i32(ctx, fmin, omin, nvert + 1);
  if (rc != CX_OK) return rc;
  k_compact<<<grid, 256, 0, st>>>(fmax, omax, nvert, lmax);
  CX_LAUNCH_CHECK(ctx, "k_compact");
  k_compact<<<grid, 256, 0, st>>>(fmin, omin, nvert, lmin);
  CX_LAUNCH_CHECK(ctx, "k_compact");
  int32_t *hn = (int32_t *)ctx->h_mail;
  CX_CUDA(ctx, cudaMemcpyAsync(hn, omax + nvert, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CX_CUDA(ctx, cudaMemcpyAsync(hn + 1, omin + nvert, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CX_CUDA(ctx, cudaMemsetAsync(ctx->d_mail + 8, 0, sizeof(int64_t), st));
  CX_CUDA(ctx, cudaStreamSynchronize(st));
  const int nmax = hn[0], nmin = hn[1];
  int64_t missing = 0;
  if (nmax > 0) {
    // bucket grid over the two in-face coordinates: cell edge >= dr/2 and at most ~16 M cells
    const int d1 = (idim + 1) % 3, d2 = (idim + 2) % 3;
    const double e1 = (double)p->dmax[d1] - p->dmin[d1], e2 = (double)p->dmax[d2] - p->dmin[d2];
    double cedge = 0.5 * p->dr;
    if (e1 * e2 / (cedge * cedge) > 1.6e7) cedge = sqrt(e1 * e2 / 1.6e7);
    PairGrid g;
    g.o1 = p->dmin[d1]; g.o2 = p->dmin[d2]; g.inv = (float)(1.0 / cedge);
    g.n1 = (int)(e1 / cedge) + 2; g.n2 = (int)(e2 / cedge) + 2;
    const int64_t ncell = (int64_t)g.n1 * g.n2;
    const int64_t cstride = (ncell + 1 + 63) & ~(int64_t)63;
    int32_t *cnt, *cstart, *bucket;
    CX_CUDA(ctx, cudaMallocAsync(&cnt, sizeof(int32_t) * (size_t)(2 * cstride + nmin + 64), st));
    cstart = cnt + cstride; bucket = cstart + cstride;
    CX_CUDA(ctx, cudaMemsetAsync(cnt, 0, sizeof(int32_t) * (size_t)cstride, st));
    if (nmin > 0) {
      k_pg_count<<<cx_blocks(nmin, 256, ctx->num_sms * 16), 256, 0, st>>>((const float4 *)pos_d, lmin, nmin, idim, g, cnt);
      CX_LAUNCH_CHECK(ctx, "k_pg_count");
    }
    rc = cx_exclusive_scan_i32(ctx, cnt, cstart, ncell + 1);
    if (rc != CX_OK) return rc;
    CX_CUDA(ctx, cudaMemsetAsync(cnt, 0, sizeof(int32_t) * (size_t)cstride, st));
    if (nmin > 0) {
      k_pg_scatter<<<cx_blocks(nmin, 256, ctx->num_sms * 16), 256, 0, st>>>((const float4 *)pos_d, lmin, nmin, idim, g, cstart, cnt, bucket);
      CX_LAUNCH_CHECK(ctx, "k_pg_scatter");
    }
    k_pair<<<cx_blocks(nmax, 128), 128, 0, st>>>((const float4 *)pos_d, lmax, nmax, g, cstart, bucket, idim, p->dmin[idim], thr, newlink_d,
                                                 ctx->d_mail + 8);
    CX_LAUNCH_CHECK(ctx, "k_pair");
    CX_CUDA(ctx, cudaFreeAsync(cnt, st));
    k_delete_linked<<<cx_blocks(nmax, 128), 128, 0, st>>>((float4 *)pos_d, lmax, nmax, newlink_d);
    CX_LAUNCH_CHECK(ctx, "k_delete_linked");
    CX_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail + 8, ctx->d_mail + 8, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CX_CUDA(ctx, cudaStreamSynchronize(st));
    missing = ctx->h_mail[8];
  }
  CX_CUDA(ctx, cudaFreeAsync(fmax, st));
  if (missing > 0) return cx_fail(ctx, CX_NO_PER_VERT, "find_links: max-face vertex without a periodic partner");
  return CX_OK;
}

// crixus_d.cu:240-256
__global__ void __launch_bounds__(256) k_relink(int4 *__restrict__ ep, int64_t nbe, const int32_t *__restrict__ newlink)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nbe; i += (int64_t)gridDim.x * blockDim.x) {
    int4 e = ep[i];
    const int a = __ldg(&newlink[e.x]), b = __ldg(&newlink[e.y]), c = __ldg(&newlink[e.z]);
    if (a != -1 || b != -1 || c != -1) {
      if (a != -1) e.x = a;
      if (b != -1) e.y = b;
      if (c != -1) e.z = c;
      ep[i] = e;
    }
  }
}

extern "C" int cx_periodicity_links(cx_ctx *ctx, int32_t *ep_d, int64_t nbe, const int32_t *newlink_d)
{
  if (!ctx || !ep_d || !newlink_d || nbe < 0) return CX_WRONG_INPUT;
  if (nbe == 0) return CX_OK;
  k_relink<<<cx_blocks(nbe, 256, ctx->num_sms * 16), 256, 0, ctx->stream>>>((int4 *)ep_d, nbe, newlink_d);
  CX_LAUNCH_CHECK(ctx, "k_relink");
  return CX_OK;
}

// crixus_d.cu:258-267: triangles per vertex
__global__ void __launch_bounds__(256) k_trisize(const int4 *__restrict__ ep, int32_t *__restrict__ trisize, int64_t nbe)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nbe; i += (int64_t)gridDim.x * blockDim.x) {
    const int4 e = ep[i];
    atomicAdd(&trisize[e.x], 1);
    atomicAdd(&trisize[e.y], 1);
    atomicAdd(&trisize[e.z], 1);
  }
}

extern "C" int cx_calc_trisize(cx_ctx *ctx, const int32_t *ep_d, int32_t *trisize_d, int64_t nbe)
{
  if (!ctx || !ep_d || !trisize_d || nbe < 0) return CX_WRONG_INPUT;
  if (nbe == 0) return CX_OK;
  k_trisize<<<cx_blocks(nbe, 256, ctx->num_sms * 16), 256, 0, ctx->stream>>>((const int4 *)ep_d, trisize_d, nbe);
  CX_LAUNCH_CHECK(ctx, "k_trisize");
  return CX_OK;
}

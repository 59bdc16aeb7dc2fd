// crixus_b200 -- periodic vertex pairing (find_links) and segment relinking (periodicity_links), plus calc_trisize.
// Reference: /root/reference/src/crixus_d.cu:213-267, driver loop /root/reference/src/crixus.cu:387-435.
// The reference scans ALL vertices for every max-face vertex (O(n_face * nvert)); here both faces are compacted in
// index order first (flag -> scan -> scatter) and only face x face pairs are tested, keeping the reference's rule
// "lowest-index partner wins".
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

// one thread per max-face vertex; scans the (index-ordered) min-face list, first hit = lowest index (crixus_d.cu:218-228)
__global__ void __launch_bounds__(128) k_pair(const float4 *__restrict__ pos, const int32_t *__restrict__ lmax, int nmax,
                                              const int32_t *__restrict__ lmin, int nmin, int idim, float dmin, double thr,
                                              int32_t *__restrict__ newlink, int64_t *__restrict__ missing)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nmax) return;
  const int i = lmax[t];
  const float4 pi = pos[i];
  const int d1 = (idim + 1) % 3, d2 = (idim + 2) % 3;
  const float pi1 = d1 == 0 ? pi.x : (d1 == 1 ? pi.y : pi.z), pi2 = d2 == 0 ? pi.x : (d2 == 1 ? pi.y : pi.z);
  int hit = -1;
  for (int s = 0; s < nmin; s++) {
    const int j = lmin[s];
    if (j == i) continue;
    const float4 pj = __ldg(&pos[j]);
    const float pj0 = idim == 0 ? pj.x : (idim == 1 ? pj.y : pj.z);
    const float pj1 = d1 == 0 ? pj.x : (d1 == 1 ? pj.y : pj.z), pj2 = d2 == 0 ? pj.x : (d2 == 1 ? pj.y : pj.z);
    const float a = fsub(pi1, pj1), b = fsub(pi2, pj2), c = fsub(pj0, dmin);
    // pow(x,(float)2.) of crixus_d.cu:220-222 is an inlined powf; restated as x*x (DESIGN.md "Numerics")
    const float d = __fsqrt_rn(fadd(fadd(fmul(a, a), fmul(b, b)), fmul(c, c)));
    if ((double)d < thr) { hit = j; break; }
  }
  if (hit >= 0) newlink[i] = hit;
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
    k_pair<<<cx_blocks(nmax, 128), 128, 0, st>>>((const float4 *)pos_d, lmax, nmax, lmin, nmin, idim, p->dmin[idim], thr, newlink_d,
                                                 ctx->d_mail + 8);
    CX_LAUNCH_CHECK(ctx, "k_pair");
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

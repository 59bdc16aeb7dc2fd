// crixus_b200 -- special-boundary (KENT) tagging: which segments have their barycentre inside a triangle of a
// special-boundary grid, and which vertices belong to such a segment.
// Reference: /root/reference/src/crixus_d.cu:1060-1112 (segInTri, identifySpecialBoundarySegments), driver loop
// /root/reference/src/crixus.cu:469-617.  The reference tests every segment against every grid triangle, O(nbe*sbnbe),
// recomputing the triangle's normal and edge dot products per pair.  Here the per-triangle part is hoisted into a record
// (same operations, same order: bit-identical per-pair results) and each grid triangle is only tested against the
// segments of the 3dr-cells its conservatively grown bounding box touches -- the segments are already cell-sorted, so a
// z-run of cells is one contiguous index range.  "First hit wins" (crixus_d.cu:1102-1107) writes the same value for every
// hit, so the outcome is the set {segments with a hit} and any order reaches it.
// HBM-bound by construction (16 B barycentre + 16 B ep per candidate segment), in practice a few microseconds.
#include "common.cuh"

struct SbRec {            // 6 x float4
  float4 v0_thr;          // vb[0].xyz, eps*|nx+ny+nz|      (crixus_d.cu:1065,1096-1101: norm.a[3] is the SUM of the components)
  float4 n_inv;           // un-normalised normal (R2 cross product), 1/det
  float4 ba_d00;          // vb[1]-vb[0], ba.ba
  float4 ca_d11;          // vb[2]-vb[0], ca.ca
  float4 d01_lo;          // ba.ca, cell range lo (3 ints)
  float4 hi_pad;          // cell range hi (3 ints)
};

__global__ void __launch_bounds__(128) k_sb_prepare(const float4 *__restrict__ sbpos, const int4 *__restrict__ sbep, int64_t sbnbe,
                                                    cx_params p, SbRec *__restrict__ rec)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= sbnbe) return;
  const int4 e = sbep[i];
  const f3 a = mk3(sbpos[e.x]), b = mk3(sbpos[e.y]), c = mk3(sbpos[e.z]);
  const f3 ba = sub3(b, a), ca = sub3(c, a);
  const f3 n = cross_r2(ba, ca);                                   // :1097-1098 under R2
  const float n3 = fabsf(fadd(fadd(fadd(0.0f, n.x), n.y), n.z));   // :1095,1099,1101
  const float d00 = dot3_r1(ba, ba), d01 = dot3_r1(ba, ca), d11 = dot3_r1(ca, ca);   // :1069-1076 (R1 chains from 0)
  const float det = det_r2(d00, d11, d01, d01);
  const float inv = __frcp_rn(det);                                // :1078 1.0/det in double, rounded to float == float reciprocal
  SbRec r;
  r.v0_thr = make_float4(a.x, a.y, a.z, fmul(n3, p.eps));
  r.n_inv = make_float4(n.x, n.y, n.z, inv);
  r.ba_d00 = make_float4(ba.x, ba.y, ba.z, d00);
  r.ca_d11 = make_float4(ca.x, ca.y, ca.z, d11);
  // conservative bounding box of the accept region {u >= -eps, v >= -eps, u+v <= 1+eps, |plane distance| <= sqrt(3) eps}:
  // a triangle with corners v0 - e(ba+ca), v1 + 2e ba - e ca, v2 - e ba + 2e ca.  e = eps + 0.1 covers the rounding of u, v
  // for well-conditioned triangles (sin^2 >= 1e-2, edge ratio <= 30: error of u, v < 5e-3); anything else is tested
  // against every cell.
  const double lb = sqrt((double)d00), lc = sqrt((double)d11), L = lb + lc;
  const double s2 = (double)det / ((double)d00 * (double)d11);
  const bool well = (s2 >= 1e-2) && (lb <= 30.0 * lc) && (lc <= 30.0 * lb) && isfinite(L) && isfinite((double)inv) && L > 0.0;
  int lo[3] = {0, 0, 0}, hi[3] = {p.gridn[0] - 1, p.gridn[1] - 1, p.gridn[2] - 1};
  if (well) {
    const double m = 2.0 * (0.1 + (double)p.eps) * L + 4.0 * (double)p.eps + 1e-5 * L;
    const float va[3] = {a.x, a.y, a.z}, vbb[3] = {b.x, b.y, b.z}, vc[3] = {c.x, c.y, c.z};
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const float l = __double2float_rd((double)fminf(va[j], fminf(vbb[j], vc[j])) - m);
      const float h = __double2float_ru((double)fmaxf(va[j], fmaxf(vbb[j], vc[j])) + m);
      lo[j] = cell_coord(l, p.dmin[j], p.griddr[j], p.gridn[j]);    // same monotone mapping as the binning
      hi[j] = cell_coord(h, p.dmin[j], p.griddr[j], p.gridn[j]);
    }
  }
  r.d01_lo = make_float4(d01, __int_as_float(lo[0]), __int_as_float(lo[1]), __int_as_float(lo[2]));
  r.hi_pad = make_float4(__int_as_float(hi[0]), __int_as_float(hi[1]), __int_as_float(hi[2]), 0.0f);
  rec[i] = r;
}

// segInTri, crixus_d.cu:1060-1083, on the hoisted record
__device__ __forceinline__ bool sb_seg_in_tri(const SbRec &r, float4 s, float eps)
{
  const f3 pa = mk3(fsub(s.x, r.v0_thr.x), fsub(s.y, r.v0_thr.y), fsub(s.z, r.v0_thr.z));
  const float d = dot_r1(r.n_inv.x, r.n_inv.y, r.n_inv.z, pa.x, pa.y, pa.z);
  if (fabsf(d) > r.v0_thr.w) return false;                          // :1065
  const float d02 = dot_r1(r.ba_d00.x, r.ba_d00.y, r.ba_d00.z, pa.x, pa.y, pa.z);
  const float d12 = dot_r1(r.ca_d11.x, r.ca_d11.y, r.ca_d11.z, pa.x, pa.y, pa.z);
  const float d01 = r.d01_lo.x;
  const float u = fmul(det_r2(d02, r.ca_d11.w, d01, d12), r.n_inv.w);   // :1079
  const float v = fmul(det_r2(r.ba_d00.w, d12, d01, d02), r.n_inv.w);   // :1080
  return !(u < -eps || v < -eps || fadd(u, v) > fadd(eps, 1.0f));      // :1081
}

// block (x = grid triangle, y = column chunk); one warp per (gx, gy) column of the triangle's cell range; lanes stride the
// contiguous segment range of the column's z-run
__global__ void __launch_bounds__(256) k_sb_test(const SbRec *__restrict__ rec, int64_t sbnbe, const float4 *__restrict__ pos,
                                                 const int4 *__restrict__ ep, const int32_t *__restrict__ cell_idx, int64_t nvert,
                                                 cx_params p, int32_t *__restrict__ sbid, int32_t sbi)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int64_t i = blockIdx.x; i < sbnbe; i += gridDim.x) {
    const SbRec r = rec[i];
    const int lx = __float_as_int(r.d01_lo.y), ly = __float_as_int(r.d01_lo.z), lz = __float_as_int(r.d01_lo.w);
    const int hx = __float_as_int(r.hi_pad.x), hy = __float_as_int(r.hi_pad.y), hz = __float_as_int(r.hi_pad.z);
    const int ny = hy - ly + 1;
    const int64_t ncol = (int64_t)(hx - lx + 1) * ny;
    for (int64_t col = (int64_t)blockIdx.y * nwarp + warp; col < ncol; col += (int64_t)gridDim.y * nwarp) {
      const int gx = lx + (int)(col / ny), gy = ly + (int)(col % ny);
      const int64_t c0 = ((int64_t)gx * p.gridn[1] + gy) * p.gridn[2];
      const int b = __ldg(cell_idx + c0 + lz), e = __ldg(cell_idx + c0 + hz + 1);
      for (int id = b + lane; id < e; id += 32) {
        if (sb_seg_in_tri(r, __ldg(pos + nvert + id), p.eps)) {
          const int4 v = __ldg(ep + id);
          sbid[nvert + id] = sbi;                                       // :1103-1105 (every writer stores the same value)
          sbid[v.x] = sbi;
          sbid[v.y] = sbi;
          sbid[v.z] = sbi;
        }
      }
    }
  }
}

extern "C" int cx_identify_special_boundary(cx_ctx *ctx, const cx_params *p, const float *pos_d, const int32_t *ep_d,
                                            const int32_t *cell_idx_d, int64_t nvert, int64_t nbe, const float *sbpos_d,
                                            const int32_t *sbep_d, int64_t sbnbe, int32_t *sbid_d, int32_t sbi)
{
  if (!ctx || !p || nbe < 0 || nvert < 0 || sbnbe < 0 || p->gridn[3] <= 0) return CX_WRONG_INPUT;
  if (nbe == 0 || sbnbe == 0) return CX_OK;
  if (!pos_d || !ep_d || !cell_idx_d || !sbpos_d || !sbep_d || !sbid_d) return CX_WRONG_INPUT;
  cudaStream_t st = ctx->stream;
  SbRec *rec = nullptr;
  CX_CUDA(ctx, cudaMallocAsync(&rec, sizeof(SbRec) * (size_t)sbnbe, st));
  k_sb_prepare<<<cx_blocks(sbnbe, 128), 128, 0, st>>>((const float4 *)sbpos_d, (const int4 *)sbep_d, sbnbe, *p, rec);
  CX_LAUNCH_CHECK(ctx, "k_sb_prepare");
  const int gx = (int)(sbnbe < 32768 ? sbnbe : 32768);
  int gy = (ctx->num_sms * 8 + gx - 1) / gx;
  gy = gy < 1 ? 1 : (gy > 64 ? 64 : gy);
  k_sb_test<<<dim3(gx, gy), 256, 0, st>>>(rec, sbnbe, (const float4 *)pos_d, (const int4 *)ep_d, cell_idx_d, nvert, *p, sbid_d, sbi);
  CX_LAUNCH_CHECK(ctx, "k_sb_test");
  CX_CUDA(ctx, cudaFreeAsync(rec, st));
  return CX_OK;
}

// crixus_b200 -- fill_fluid (box) and fill_fluid_complex (geometry) on the dr lattice.
// Reference: /root/reference/src/crixus_d.cu:680-744 (box), :753-963 (checkTriangleCollision, checkCollision),
// :965-1058 (fill_fluid_complex, one full-grid sweep per launch, looped up to 10 000 times by crixus.cu:934-947).
//
// The fill result is the least fixed point of "s unset, a 6-neighbour e set, !checkCollision(s,e) => set s"
// (SURVEY.md A.6), so any schedule reaches the same bits.  B200 mapping (DESIGN.md "fill"):
//   * the packed reference bitfield is unpacked to a row-aligned layout (x-rows padded to 32-bit words);
//   * checkCollision(s, .) can only be true for a lattice cell s whose 3dr-cell has segments in its 27-neighbourhood, or that
//     lies near a free-surface (fshape) triangle.  All of these directed tests are pure functions of the geometry and are
//     resolved ONCE, up front, into six "blocked" bit planes B[d] (bit set = cell s must not be entered from direction d):
//       - k_resolve_blocks: one warp per non-empty 3dr-cell neighbourhood.  The lattice cells of a 3dr-cell (3-4 per axis)
//         share their 27-cell neighbourhood, so the candidate segment records are streamed once per 3dr-cell, rejected
//         against the whole block of lattice cells by their plane distance, staged in shared memory, and then every lane
//         tests ITS lattice cell against the staged records (broadcast reads) with exact, division-free pre-filters in front
//         of the reference's arithmetic;
//       - k_resolve_fs: the free-surface triangles against the lattice rows they can touch;
//   * propagation is then pure bit arithmetic on 32-cell words: tiles of 1 word x 16 x 16 rows are iterated to local
//     convergence in shared memory (k_fill_tiles); a tile that changed activates its 6 face neighbours; rounds continue until
//     no tile is active.  Only reached tiles are touched: work ~ filled volume + wall area.
//   * multi-GPU (cx_dist_fill_geometry): the directed tests are dealt to the ranks by 3dr-cell chunk, the blocked planes are
//     reduced to the z-slab owners, each rank floods its slab and exchanges boundary planes with its z-neighbours.
// Bound: HBM, 0.25 B/cell + 8 B/filled cell -- in practice the directed tests (ALU) and the round latency.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <algorithm>
#include <vector>
#include "common.cuh"
#include "fill_enum.h"
#include "dist.h"

#define FT_Y 16
#define FT_Z 16
#define FT_THREADS (FT_Y * FT_Z)

// ================================================================================================ box fill
// crixus_d.cu:686-717.  One thread per packed 32-bit word; nfi counts every in-box cell (even if already set).
struct BoxArgs {
  int64_t dimg[3];
  uint32_t mn[3], dim[3];
};

__global__ void __launch_bounds__(256) k_fill_box(uint32_t *__restrict__ fpos, BoxArgs b, int64_t nwords, unsigned long long *nfi)
{
  unsigned long long cnt = 0;
  const int64_t G = b.dimg[0] * b.dimg[1] * b.dimg[2], plane = b.dimg[0] * b.dimg[1];
  for (int64_t wi = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; wi < nwords; wi += (int64_t)gridDim.x * blockDim.x) {
    int64_t bit0 = wi * 32;
    int64_t k = bit0 / plane, rem = bit0 - k * plane;
    int64_t j = rem / b.dimg[0], i = rem - j * b.dimg[0];
    uint32_t m = 0;
    for (int ii = 0; ii < 32 && bit0 + ii < G; ii++) {
      const uint32_t ui = (uint32_t)i, uj = (uint32_t)j, uk = (uint32_t)k;
      if (ui >= b.mn[0] && uj >= b.mn[1] && uk >= b.mn[2] && ui - b.mn[0] < b.dim[0] && uj - b.mn[1] < b.dim[1] &&
          uk - b.mn[2] < b.dim[2]) {
        m |= 1u << ii;
        cnt++;
      }
      if (++i == b.dimg[0]) { i = 0; if (++j == b.dimg[1]) { j = 0; k++; } }
    }
    if (m) fpos[wi] |= m;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(nfi, cnt);
}

extern "C" int cx_fill_box(cx_ctx *ctx, const cx_params *p, uint32_t *fpos_d, const float lo[3], const float hi[3], int64_t *nfi_host)
{
  if (!ctx || !p || !fpos_d || !lo || !hi) return CX_WRONG_INPUT;
  BoxArgs b;
  cx_fill_dims(p, b.dimg);
  const float eps = p->eps, dr = p->dr;
  for (int j = 0; j < 3; j++) {
    b.dim[j] = (uint32_t)(floorf((hi[j] + eps - lo[j]) / dr) + 1);  // crixus_d.cu:686-688
    const double r = round(((float)(uint32_t)b.dimg[j] - 1.) * (lo[j] - p->fcmin[j]) / (p->fcmax[j] - p->fcmin[j]));  // :694-696
    b.mn[j] = r < 0 ? 0xffffffffu : (uint32_t)r;
  }
  const int64_t G = b.dimg[0] * b.dimg[1] * b.dimg[2], nwords = (G + 31) / 32;
  unsigned long long *cnt = (unsigned long long *)(ctx->d_mail + 20);
  CX_CUDA(ctx, cudaMemsetAsync(cnt, 0, sizeof(int64_t), ctx->stream));
  k_fill_box<<<cx_blocks(nwords, 256, ctx->num_sms * 16), 256, 0, ctx->stream>>>(fpos_d, b, nwords, cnt);
  CX_LAUNCH_CHECK(ctx, "k_fill_box");
  if (nfi_host) {
    CX_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail + 20, cnt, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *nfi_host = ctx->h_mail[20];
  }
  return CX_OK;
}

// ================================================================================================ geometry fill
// conservative description of a free-surface (fshape) triangle: where checkTriangleCollision(s, ., triangle) can be true
struct FsDesc {
  float lo[3], hi[3];  // bounding box grown by the margin outside of which the segment test cannot hit the triangle
  float n[3], v0[3];   // plane, for the "ray in plane => collision" rule (crixus_d.cu:766-769)
  int quirk;           // some axis direction can be parallel to the plane within eps: cells within 4 eps of the plane matter too
  int cull;            // 0: the triangle is ill-conditioned / its normal disagrees with its plane -> every cell is tested
};

struct FillArgs {
  cx_params p;
  const float4 *pos;
  const float4 *norm;
  const int4 *ep;
  const int32_t *cell_idx;
  int64_t nbe, cnbe;
  int dimg[3];       // lattice size
  int wr;            // words per x-row
  int ty, tz;        // tile counts in y and z
  int64_t ntiles;    // wr * ty * tz
  int k_lo, k_hi;    // owned planes [k_lo, k_hi): only these are written (z-slab sharding)
  uint32_t *F;       // row-aligned filled bits, wr*dimg[1]*dimg[2] words
  uint32_t *B;       // 6 blocked planes, same shape each (direction d at B + d*nw)
  int64_t nw;        // words per plane set
  const uint8_t *nbflag;  // per 3dr-cell: 27-neighbourhood contains a segment
  const FsDesc *fs;       // fshape descriptors (cnbe of them)
  uint8_t *act_cur, *act_next;  // tile activation flags
  int32_t *list;                // active tile list
  int32_t *count;               // [0] = list length, [1] = cursor
  int32_t *count_next;          // the pair of the next propagation round (zeroed by k_fill_tiles of this round)
  unsigned long long *count64;  // work counter of the up-front resolve
  unsigned long long *changed;  // newly set cells
  int res_nparts, res_part;     // up-front resolve: 3dr-cell chunks dealt to nparts ranks, this rank resolves its share (0/1: all)
  const int32_t *blist;         // work list of the up-front resolve (3dr-cells of this rank's share), nblist entries
  long long nblist;
  int force_activate;           // first round of a call: processed tiles wake their neighbours unconditionally
  int merge;                    // k_unpack: OR into the existing row-aligned state (later slab calls)
  int64_t w_lo, w_hi;           // word range k_unpack / k_pack walk (slab calls: the owned planes, +-1 halo plane for k_unpack)
  const float4 *rec;            // per-segment invariants of checkCollision (k_seg_prep), 8 x float4 per segment
  float rrej;                   // plane distance beyond which a well-formed triangle cannot matter
  float drw2;                   // dr_wall*dr_wall
  double one_eps;               // 1.0 + eps (double, crixus_d.cu:774)
  int dist_dead;                // eps >= dr_wall^2: the distance test `dist + eps < dr_wall^2` (crixus_d.cu:940) can never hold
  int nofilter;                 // debugging: no pre-filters, every pair inside rrej runs the reference's arithmetic
  float sph2;                   // (1.01 (1 + 3 eps))^2: hit points further than sqrt(sph2 * max(uu, vv)) from v0 miss the triangle
  float lstep;                  // reach of one lattice step with the ray parameter's tolerance (fe_lstep)
  int skip3;                    // k_resolve_blocks: records of flag 3 are resolved by k_resolve_cross / k_resolve_inplane
  int32_t *nrest;               // k_seg_prep: number of wall records whose flag is not 3
  const float4 *plane;          // per wall segment (n, n.(v0 - fcmin)) if it can take part in the in-plane rule, (0, 0, 0, inf) otherwise
  const int32_t *cellid;        // per wall segment: its 3dr-cell (k_seg_cells)
};

// ---- the reference's directed test -------------------------------------------------------------------------------
// Everything in checkCollision that depends on the triangle alone is hoisted into a per-segment record (k_seg_prep):
// the same operations in the same order, so the results are bit-identical to evaluating them per (cell, segment) pair.
//   R0 = (n, wf)   R1 = (v0, uu)   R2 = (v1, uv)   R3 = (v2, vv)   R4 = (segpos, D)   R5..R7 = (nr[0..2], -)
// wf: 0 = no assumption, 1 = well-formed normal (plane-distance reject is valid), 2 = and well-conditioned edges (the
// bounding-sphere pre-filter of the segment test is valid too), 3 = and every lattice cell the segment can block is a
// 27-neighbour of its 3dr-cell (set by k_seg_cells; the record-driven kernels take these).  R5.w = the segment's 3dr-cell.
#define FREC 8

// checkTriangleCollision, crixus_d.cu:753-799 (u, v, uu, uv, vv, D from the record)
__device__ __forceinline__ bool tri_collision(const FillArgs &A, f3 s, f3 dir, f3 n, f3 v0, f3 u, f3 v, float uu, float uv, float vv,
                                              float D)
{
  const float eps = A.p.eps;
  const f3 w0 = sub3(s, v0);
  const float a = -dot3_r3(n, w0), b = dot3_r3(n, dir);
  if (fabsf(b) < eps) return fabsf(a) < eps;
  const float r = __fdiv_rn(a, b);
  if (r < -eps || (double)r > A.one_eps) return false;
  const f3 w = mk3(fsub(ffma(r, dir.x, s.x), v0.x), fsub(ffma(r, dir.y, s.y), v0.y), fsub(ffma(r, dir.z, s.z), v0.z));
  const float wu = dot3_r3(w, u), wv = dot3_r3(w, v);
  const float t1 = __fdiv_rn(det_r2(uv, wv, vv, wu), D);
  if (t1 < -eps || (double)t1 > A.one_eps) return false;
  const float t2 = __fdiv_rn(det_r2(uv, wu, uu, wv), D);
  if (t2 < -eps || (double)fadd(t1, t2) > A.one_eps) return false;
  return true;
}

// squared distance from s to triangle (v0,v1,v2) exactly as crixus_d.cu:859-939 computes it (segpos, nr from the record)
__device__ __forceinline__ float tri_dist2(const FillArgs &A, f3 s, f3 n, f3 v0, f3 v1, f3 v2, f3 segpos, f3 nr0, f3 nr1, f3 nr2)
{
  const float eps = A.p.eps;
  const f3 sm = sub3(s, segpos);
  const float pd = dot3_r3(sm, n);
  const f3 pp = mk3(ffma(-n.x, pd, sm.x), ffma(-n.y, pd, sm.y), ffma(-n.z, pd, sm.z));
  const f3 q = add3(pp, segpos);
  const float o0 = dot3_r3(sub3(q, v0), nr0), o1 = dot3_r3(sub3(q, v1), nr1), o2 = dot3_r3(sub3(q, v2), nr2);
  f3 t;
  if (o0 > -eps && o1 > -eps && o2 > -eps) t = sub3(pp, sm);
  else if (o0 < eps && o1 < eps) t = sub3(s, v2);
  else if (o1 < eps && o2 < eps) t = sub3(s, v0);
  else if (o2 < eps && o0 < eps) t = sub3(s, v1);
  else {
    f3 x1, x2;
    if (o0 < eps) { x1 = v0; x2 = v1; }
    else if (o1 < eps) { x1 = v1; x2 = v2; }
    else { x1 = v2; x2 = v0; }
    const f3 x21 = sub3(x2, x1), sx1 = sub3(s, x1);
    const float td = dot3_r3(x21, sx1), l2 = dot3_r3(x21, x21);
    t = mk3(ffma(-__fdiv_rn(x21.x, l2), td, sx1.x), ffma(-__fdiv_rn(x21.y, l2), td, sx1.y), ffma(-__fdiv_rn(x21.z, l2), td, sx1.z));
  }
  return dot3_r3(t, t);
}

// one thread per segment: the triangle-only parts of crixus_d.cu:859-896 (segpos, edge normals nr) and :755-783 (uu, uv,
// vv, D), plus the flags that license the pre-filters of the resolve kernels.
__global__ void __launch_bounds__(256) k_seg_prep(FillArgs A, float4 *__restrict__ rec, float4 *__restrict__ plane)
{
  const int64_t nwall = A.nbe - A.cnbe;
  int rest = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < A.nbe; i += (int64_t)gridDim.x * blockDim.x) {
    const f3 n = mk3(__ldg(&A.norm[i]));
    const int4 E = __ldg(&A.ep[i]);
    const f3 v0 = mk3(__ldg(&A.pos[E.x])), v1 = mk3(__ldg(&A.pos[E.y])), v2 = mk3(__ldg(&A.pos[E.z]));
    const f3 segpos = mk3(__fdiv_rn(fadd(fadd(v0.x, v1.x), v2.x), 3.0f), __fdiv_rn(fadd(fadd(v0.y, v1.y), v2.y), 3.0f),
                          __fdiv_rn(fadd(fadd(v0.z, v1.z), v2.z), 3.0f));
    f3 nr[3];
    const f3 vv3[3] = {v0, v1, v2};
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const f3 a = vv3[k], b = vv3[(k + 1) % 3];
      f3 ed = sub3(b, a);
      const float len = __fsqrt_rn(dot3_r3(ed, ed));
      ed = mk3(__fdiv_rn(ed.x, len), __fdiv_rn(ed.y, len), __fdiv_rn(ed.z, len));
      const f3 sv = sub3(segpos, a);
      const float d = dot3_r3(ed, sv);
      nr[k] = mk3(ffma(-ed.x, d, sv.x), ffma(-ed.y, d, sv.y), ffma(-ed.z, d, sv.z));
    }
    const f3 u = sub3(v1, v0), v = sub3(v2, v0);
    const float uu = dot3_r3(u, u), uv = dot3_r3(u, v), vv = dot3_r3(v, v);
    const float D = det_r2(uv, uv, uu, vv);
    // well-formed / well-conditioned: fe_classify (fill_enum.h, shared with the CPU check of the culling rules)
    const float na[3] = {n.x, n.y, n.z}, ua[3] = {u.x, u.y, u.z}, va[3] = {v.x, v.y, v.z};
    int cls = fe_classify(na, ua, va, D);
    // wall segments: flag 3 if every lattice cell the segment can block is a 27-neighbour of its 3dr-cell (fe_rec_setup: safe) --
    // the record-driven kernels take these; the plane (n, n.v0) of those that have an axis parallel to it (in-plane rule)
    int cell = 0;
    if (i < nwall) {
      cell = __ldg(&A.cellid[i]);
      float4 pl = make_float4(0.f, 0.f, 0.f, __int_as_float(0x7f800000));
      if (cls == 2) {
        int g[3];
        fe_cell_coords(&A.p, cell, g);
        const float a0[3] = {v0.x, v0.y, v0.z}, a1[3] = {v1.x, v1.y, v1.z}, a2[3] = {v2.x, v2.y, v2.z};
        FeRec T;
        fe_rec_setup(&A.p, A.lstep, na, a0, a1, a2, uu, vv, g, &T);
        if (T.safe) {
          cls = 3;
          if (fe_capable(&A.p, A.lstep, na)) pl = make_float4(n.x, n.y, n.z, fe_plane_offset(&A.p, na, a0));
        }
      }
      plane[i] = pl;
      if (cls != 3) rest++;
    }
    float4 *r = rec + i * FREC;
    r[0] = make_float4(n.x, n.y, n.z, (float)cls);
    r[1] = make_float4(v0.x, v0.y, v0.z, uu);
    r[2] = make_float4(v1.x, v1.y, v1.z, uv);
    r[3] = make_float4(v2.x, v2.y, v2.z, vv);
    r[4] = make_float4(segpos.x, segpos.y, segpos.z, D);
    r[5] = make_float4(nr[0].x, nr[0].y, nr[0].z, __int_as_float(cell));
    r[6] = make_float4(nr[1].x, nr[1].y, nr[1].z, 0.f);
    r[7] = make_float4(nr[2].x, nr[2].y, nr[2].z, 0.f);
  }
  rest = __reduce_add_sync(0xffffffffu, rest);
  if ((threadIdx.x & 31) == 0 && rest) atomicAdd(A.nrest, rest);
}

// the 3dr-cell of every wall segment (the cell the BINNING put it in: cell_idx is the authority, not a recomputed barycentre).
// A warp takes 32 consecutive 3dr-cells: their runs are one contiguous range of records, lane l holds the end of cell c0 + l's
// run and a record finds its cell by a 5-step search over the lanes.
__global__ void __launch_bounds__(256) k_seg_cells(int64_t n, const int32_t *__restrict__ cell_idx, int64_t nwall, int32_t *__restrict__ cellid)
{
  const int lane = threadIdx.x & 31;
  for (int64_t c0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) & ~31ll; c0 < n; c0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t cl = min(c0 + lane, n - 1);
    const int me = __ldg(&cell_idx[cl + 1]);
    const int64_t B = __ldg(&cell_idx[c0]), E = min((int64_t)__shfl_sync(0xffffffffu, me, 31), nwall);
    for (int64_t i0 = B; i0 < E; i0 += 32) {
      const int64_t i = i0 + lane;
      int k = 0;   // number of lanes whose run ends at or before i = the lane of i's cell
#pragma unroll
      for (int step = 16; step > 0; step >>= 1) {
        const int e = __shfl_sync(0xffffffffu, me, k + step - 1);
        if (e <= i) k += step;
      }
      if (i < E) cellid[i] = (int)min(c0 + k, n - 1);
    }
  }
}

// 3dr-cells whose segments do not all share one plane for the in-plane rule: bit 1 of nbflag.  A cell without the bit either
// holds no segment that takes part (the plane of its first segment is not finite) or only segments with the normal of the first
// one, bit for bit, and plane offsets within eps / 4 of its offset -- k_resolve_inplane then decides about the whole cell from
// its first segment.
__global__ void __launch_bounds__(256) k_cell_planes(FillArgs A, uint8_t *__restrict__ nbflag)
{
  const int64_t nwall = A.nbe - A.cnbe;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nwall; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 P = __ldg(&A.plane[i]);
    if (!(P.w < 3.0e38f && P.w > -3.0e38f)) continue;
    const int c = __ldg(&A.cellid[i]);
    const float4 P0 = __ldg(&A.plane[__ldg(&A.cell_idx[c])]);
    if (!(P0.x == P.x && P0.y == P.y && P0.z == P.z && fabsf(P0.w - P.w) < 0.25f * A.p.eps)) nbflag[c] = (uint8_t)(nbflag[c] | 2u);
  }
}

// lattice position along one axis, crixus_d.cu:805-807 (R5)
__device__ __forceinline__ float lat_pos(const cx_params &p, int a, int idx) { return ffma((float)idx, p.dr, p.fcmin[a]); }
__device__ __forceinline__ int lat_cell(const cx_params &p, int a, int idx)
{
  return cell_coord(lat_pos(p, a, idx), p.dmin[a], p.griddr[a], p.gridn[a]);
}

// the body of checkCollision's segment loop (crixus_d.cu:846-948) for ONE (lattice cell, segment record) pair and the
// directions in `need` (bit d: 0:-x 1:+x 2:-y 3:+y 4:-z 5:+z, the reference's neighbour order): influence-radius test,
// distance test, segment-triangle test.  Returns the directions that collide.  Called by diverged lanes (rare path).
__device__ __noinline__ uint32_t pair_full(const FillArgs &A, int si, int sj, int sk, int64_t ri, uint32_t need)
{
  const cx_params &p = A.p;
  const float dr = p.dr, eps = p.eps;
  const int sidx[3] = {si, sj, sk};
  float sc[3];
#pragma unroll
  for (int a = 0; a < 3; a++) sc[a] = lat_pos(p, a, sidx[a]);
  const f3 s = mk3(sc[0], sc[1], sc[2]);
  const float4 *r = A.rec + ri * FREC;
  const float4 R0 = __ldg(r), R1 = __ldg(r + 1), R2 = __ldg(r + 2), R3 = __ldg(r + 3), R4 = __ldg(r + 4);
  const f3 n = mk3(R0), v0 = mk3(R1), v1 = mk3(R2), v2 = mk3(R3);
  uint32_t pass = 0;
  float dirc[6];
#pragma unroll
  for (int d = 0; d < 6; d++) {
    dirc[d] = 0.f;
    if (!(need >> d & 1)) continue;
    const int a = d >> 1;
    const float ec = lat_pos(p, a, sidx[a] + ((d & 1) ? 1 : -1));
    dirc[d] = fsub(ec, sc[a]);
    const float hd = fmul(fsub(sc[a], ec), 0.5f);  // (s-e)/2.0f; the two other components are exactly +0
    const f3 hv = mk3(a == 0 ? hd : 0.f, a == 1 ? hd : 0.f, a == 2 ? hd : 0.f);
    const float l = __fsqrt_rn(dot3_r3(hv, hv));
    const float rr = fadd(ffma(dr, 3.0f, l), p.dr_wall);  // crixus_d.cu:815 (R6)
    const float infl = fmul(rr, rr);
    const float midc = fadd(sc[a], ec);
    const f3 tt = mk3(ffma(a == 0 ? midc : fadd(sc[0], sc[0]), 0.5f, -v0.x), ffma(a == 1 ? midc : fadd(sc[1], sc[1]), 0.5f, -v0.y),
                      ffma(a == 2 ? midc : fadd(sc[2], sc[2]), 0.5f, -v0.z));
    if (!(dot3_r3(tt, tt) > infl)) pass |= 1u << d;  // crixus_d.cu:855-857
  }
  if (!pass) return 0;
  if (!A.dist_dead) {
    const float4 R5 = __ldg(r + 5), R6 = __ldg(r + 6), R7 = __ldg(r + 7);
    const float dist = tri_dist2(A, s, n, v0, v1, v2, mk3(R4), mk3(R5), mk3(R6), mk3(R7));
    if (fadd(dist, eps) < A.drw2) return pass;  // crixus_d.cu:940
  }
  uint32_t mine = 0;
  const f3 u = sub3(v1, v0), v = sub3(v2, v0);
#pragma unroll
  for (int d = 0; d < 6; d++) {
    if (!(pass >> d & 1)) continue;
    const int a = d >> 1;
    const f3 dir = mk3(a == 0 ? dirc[d] : 0.f, a == 1 ? dirc[d] : 0.f, a == 2 ? dirc[d] : 0.f);
    if (tri_collision(A, s, dir, n, v0, u, v, R1.w, R2.w, R3.w, R4.w)) mine |= 1u << d;
  }
  return mine;
}

// ---- preparation kernels ---------------------------------------------------------------------------------------
// nbflag bit 0: the 27 neighbour cells of a 3dr-cell (periodic wrap / skip per axis as in the reference's lookup,
// crixus_d.cu:327-347) hold a segment.  The neighbourhood is a product of three per-axis neighbourhoods: three passes, each an
// OR over the cell and its two neighbours along one axis (axis 2 = z reads the runs themselves).
__global__ void __launch_bounds__(256) k_cell3_flags(cx_params p, int axis, const int32_t *__restrict__ cell_idx, const uint8_t *__restrict__ in,
                                                     uint8_t *__restrict__ out)
{
  const int64_t n = p.gridn[3];
  const int64_t stride = axis == 2 ? 1 : (axis == 1 ? p.gridn[2] : (int64_t)p.gridn[2] * p.gridn[1]);
  const int gn = p.gridn[axis];
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < n; c += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)((c / stride) % gn);
    int any = 0;
#pragma unroll
    for (int d = -1; d <= 1; d++) {
      int gg = g + d;
      if (gg == -1 || gg == gn) {
        if (!p.per[axis]) continue;
        gg = gg == -1 ? gn - 1 : 0;
      }
      const int64_t q = c + (int64_t)(gg - g) * stride;
      any |= axis == 2 ? (__ldg(&cell_idx[q + 1]) > __ldg(&cell_idx[q]) ? 1 : 0) : (in[q] & 1);
    }
    out[c] = (uint8_t)any;
  }
}

// packed (reference) layout -> row-aligned words; marks tiles that contain set bits as active
__global__ void __launch_bounds__(256) k_unpack(FillArgs A, const uint32_t *__restrict__ fpos)
{
  const int64_t rows = (int64_t)A.dimg[1] * A.dimg[2];
  for (int64_t wi = A.w_lo + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; wi < A.w_hi; wi += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = wi / A.wr;
    const int wx = (int)(wi - row * A.wr);
    if (row >= rows) continue;
    const int64_t b0 = row * A.dimg[0] + (int64_t)wx * 32;
    const int nvalid = min(32, A.dimg[0] - wx * 32);
    const int64_t w0 = b0 >> 5;
    const int sh = (int)(b0 & 31);
    uint32_t v = fpos[w0] >> sh;
    if (sh && sh + nvalid > 32) v |= fpos[w0 + 1] << (32 - sh);
    if (nvalid < 32) v &= (1u << nvalid) - 1u;
    uint32_t fresh_bits = v;
    if (A.merge) {
      const uint32_t old = A.F[wi];
      fresh_bits = v & ~old;
      v |= old;
    }
    A.F[wi] = v;
    if (fresh_bits) {
      const int j = (int)(row % A.dimg[1]), k = (int)(row / A.dimg[1]);
      A.act_cur[((int64_t)(k / FT_Z) * A.ty + j / FT_Y) * A.wr + wx] = 1;
    }
  }
}

__global__ void __launch_bounds__(256) k_pack(FillArgs A, uint32_t *__restrict__ fpos)
{
  for (int64_t wi = A.w_lo + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; wi < A.w_hi; wi += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t v = A.F[wi];
    if (!v) continue;
    const int64_t row = wi / A.wr;
    const int wx = (int)(wi - row * A.wr);
    const int k = (int)(row / A.dimg[1]);
    if (k < A.k_lo || k >= A.k_hi) continue;
    const int64_t b0 = row * A.dimg[0] + (int64_t)wx * 32;
    const int64_t w0 = b0 >> 5;
    const int sh = (int)(b0 & 31);
    atomicOr(&fpos[w0], v << sh);
    if (sh && (v >> (32 - sh))) atomicOr(&fpos[w0 + 1], v >> (32 - sh));
  }
}

// a boundary plane received from a z-neighbour (row-aligned words of lattice plane k) is OR-ed into F; tiles that got new
// bits are activated and *fresh is raised
__global__ void __launch_bounds__(256) k_merge_plane(FillArgs A, const uint32_t *__restrict__ plane, int k, int *fresh)
{
  const int64_t wplane = (int64_t)A.wr * A.dimg[1];
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < wplane; q += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t v = plane[q];
    if (!v) continue;
    const int64_t wi = (int64_t)k * wplane + q;
    const uint32_t old = A.F[wi];
    if (v & ~old) {
      A.F[wi] = old | v;
      const int j = (int)(q / A.wr), wx = (int)(q - (int64_t)j * A.wr);
      A.act_cur[((int64_t)(k / FT_Z) * A.ty + j / FT_Y) * A.wr + wx] = 1;
      *fresh = 1;
    }
  }
}

__global__ void k_seed(FillArgs A, int64_t seed)
{
  if (seed < 0) return;
  const int64_t plane = (int64_t)A.dimg[0] * A.dimg[1];
  const int k = (int)(seed / plane), j = (int)((seed % plane) / A.dimg[0]), i = (int)(seed % A.dimg[0]);
  const int64_t wi = ((int64_t)k * A.dimg[1] + j) * A.wr + (i >> 5);
  const uint32_t bit = 1u << (i & 31);
  if (k >= A.k_lo && k < A.k_hi && !(A.F[wi] & bit)) {  // crixus_d.cu:972-980: counted only if it was unset
    A.F[wi] |= bit;
    atomicAdd(A.changed, 1ull);
  }
  A.act_cur[((int64_t)(k / FT_Z) * A.ty + j / FT_Y) * A.wr + (i >> 5)] = 1;
}

// ---- up-front resolution of the directed tests ---------------------------------------------------------------------
// smallest lattice index i in [0, dim] whose 3dr-cell coordinate along `axis` is >= g (dim if there is none).  The cell
// coordinate is a monotone function of i (every operation of lat_cell is correctly rounded and monotone): binary search.
__device__ int idx_lo(const cx_params &p, int axis, int g, int dim)
{
  if (g <= 0) return 0;
  if (g >= p.gridn[axis]) return dim;
  int lo = 0, hi = dim;   // invariant: cells < lo have coordinate < g, cells >= hi have coordinate >= g
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (lat_cell(p, axis, mid) >= g) hi = mid; else lo = mid + 1;
  }
  return lo;
}

// which rank resolves the 3dr-cells of chunk `chunk` (32 consecutive cells): every group of nparts chunks is dealt one chunk
// per rank, the order rotated by a hash of the group (walls are planes of the cell grid: plain round-robin would be periodic
// in them).  The same function on every rank, so the shares are disjoint and complete.
__device__ __forceinline__ int chunk_owner(long long chunk, int nparts)
{
  const long long g = chunk / nparts;
  const int slot = (int)(chunk - g * nparts);
  const int rot = (int)((((unsigned)g * 2654435761u) >> 16) % (unsigned)nparts);
  return (slot + nparts - rot) % nparts;
}

#define RB_WARPS 8
#define RB_CAP 96    // staged records per warp (phase 1 appends up to 32 per step and flushes above RB_CAP - 32)
#define RB_QUEUE 64  // ring of candidate (cell, record) pairs per warp

// per-warp working set of k_resolve_blocks
struct RBWarp {
  float4 R0[RB_CAP], R1[RB_CAP];   // (n, flags) and (v0, uu) of the staged records
  int idx[RB_CAP];                 // record index
  float rad[RB_CAP];               // bounding radius of the accepted barycentric region around v0 (well-conditioned records)
  unsigned short queue[RB_QUEUE];  // candidate pairs: (cell << 8) | staged slot
  uint32_t coll[64];               // blocked directions per lattice cell of the sub-block
  unsigned char cxyz[64];          // local coordinates of the cells: x | y << 2 | z << 4
};

// geometry of a sub-block of lattice cells: cell q = (i0 + q % nx, j0 + (q / nx) % ny, k0 + q / (nx ny)), q < nx ny nz <= 64
struct RBSub {
  int i0, j0, k0, nx, ny, nz;
};

__device__ __forceinline__ uint32_t cell_dirs(const FillArgs &A, int si, int sj, int sk)
{
  return (si > 0 ? 1u : 0u) | (si + 1 < A.dimg[0] ? 2u : 0u) | (sj > 0 ? 4u : 0u) | (sj + 1 < A.dimg[1] ? 8u : 0u) | (sk > 0 ? 16u : 0u) |
         (sk + 1 < A.dimg[2] ? 32u : 0u);
}

// One candidate (lattice cell, staged record) pair, exactly: the directions that can collide are narrowed with exact, division-
// free tests, the reference's arithmetic (pair_full) decides the rest.
//   With a = -n.(s - v0) and b = n.dir exactly as checkTriangleCollision computes them (crixus_d.cu:757-765), a direction along
//   axis x can only collide if |b| < eps and |a| < eps (ray in the plane), or r = a / b lies in [-eps, 1 + eps], which needs
//   |a| <= |n_x| |dir| (1 + eps) up to rounding; the hit point s + r dir of an accepted (t1, t2) lies within rad of v0.
// Only for well-formed triangles (R0.w != 0: finite, unit normal perpendicular to the edges) whose distance test is provably
// false; everything else takes the reference's arithmetic unfiltered.
__device__ __forceinline__ uint32_t pair_exact(const FillArgs &A, float4 R0, float4 R1, float rad, int64_t ridx, int si, int sj, int sk,
                                               uint32_t need)
{
  const cx_params &p = A.p;
  const float eps = p.eps;
  const f3 s = mk3(lat_pos(p, 0, si), lat_pos(p, 1, sj), lat_pos(p, 2, sk));
  const f3 n = mk3(R0), v0 = mk3(R1);
  const f3 w0 = sub3(s, v0);
  const float a = -dot3_r3(n, w0);
  const float aa = fabsf(a);
  const bool wf = R0.w != 0.f;
  if (wf && aa > A.rrej) return 0;   // plane-distance reject: neither test can hold
  uint32_t cand = need;
  // the distance test is alive unless eps >= dr_wall^2 or (well-formed triangle) the plane distance alone exceeds dr_wall:
  // every distance the reference computes is >= |a| / 1.2
  if (wf && (A.dist_dead || a * a >= 1.44f * A.drw2) && !A.nofilter) {
    cand = 0;
    uint32_t m = need;
    const float grow = 1.0001f * (1.0f + eps);
    while (m) {
      const int d = __ffs(m) - 1;
      m &= m - 1;
      const int ax = d >> 1;
      const int sidx = ax == 0 ? si : ax == 1 ? sj : sk;
      const float sc = ax == 0 ? s.x : ax == 1 ? s.y : s.z;
      const float dc = fsub(lat_pos(p, ax, sidx + ((d & 1) ? 1 : -1)), sc);
      const f3 dir = mk3(ax == 0 ? dc : 0.f, ax == 1 ? dc : 0.f, ax == 2 ? dc : 0.f);
      const float b = dot3_r3(n, dir);
      const float ab = fabsf(b);
      if (ab < eps) {
        if (aa < eps) cand |= 1u << d;   // ray in the plane: collision whatever the triangle's extent (crixus_d.cu:766-769)
        continue;
      }
      const float qn = b > 0.f ? a : -a;                            // r = qn / |b|
      if (qn > ab * grow || qn < -ab * eps * 1.0001f - 1e-37f) continue;   // r outside [-eps, 1+eps] for sure
      if (R0.w >= 2.f) {
        // hit point P = s + r dir (r from a fast division: the 1 % margin of rad covers its error many times over)
        const float r = __fdividef(qn, ab);
        const float px = ax == 0 ? w0.x + r * dc : w0.x, py = ax == 1 ? w0.y + r * dc : w0.y, pz = ax == 2 ? w0.z + r * dc : w0.z;
        if (px * px + py * py + pz * pz > rad * rad) continue;
      }
      cand |= 1u << d;
    }
    if (!cand) return 0;
  }
  return pair_full(A, si, sj, sk, ridx, cand);
}

// the lattice cells of one sub-block against the staged records.
//   filter: every lane holds a cell (two passes for sub-blocks of more than 32 cells) and walks the staged records with a
//           cheap, CONSERVATIVE, warp-uniform test (approximate plane distance; for well-conditioned triangles a sphere that
//           contains every hit point a lattice step from this cell can produce; cells in the plane keep every record while an
//           in-plane direction is open: the "ray in plane" rule ignores the triangle's extent).  Survivors are appended to a
//           ring of (cell, record) pairs -- ballot + popc compaction;
//   pairs : as soon as 32 pairs wait, every lane takes one and runs the exact test (pair_exact) -- full lanes for the
//           expensive part instead of a warp that diverges on every record.
__device__ __forceinline__ void sub_vs_staged(const FillArgs &A, RBWarp &W, const RBSub &S, int cnt, int lane)
{
  const cx_params &p = A.p;
  const int ncl = S.nx * S.ny * S.nz;
  int qh = 0, qn = 0;   // ring head / fill (warp-uniform)
  auto drain = [&](int take) {
    const bool have = lane < take;
    if (have) {
      const unsigned e = W.queue[(qh + lane) & (RB_QUEUE - 1)];
      const int cell = (int)(e >> 8), q = (int)(e & 255u);
      const unsigned xyz = W.cxyz[cell];
      const int si = S.i0 + (int)(xyz & 3u), sj = S.j0 + (int)((xyz >> 2) & 3u), sk = S.k0 + (int)(xyz >> 4);
      const uint32_t need = cell_dirs(A, si, sj, sk) & ~W.coll[cell];
      if (need) {
        const uint32_t got = pair_exact(A, W.R0[q], W.R1[q], W.rad[q], W.idx[q], si, sj, sk, need);
        if (got) atomicOr(&W.coll[cell], got);
      }
    }
    qh = (qh + take) & (RB_QUEUE - 1);
    qn -= take;
    __syncwarp();
  };
  for (int sl = 0; sl * 32 < ncl; sl++) {
    const int cell = lane + 32 * sl;
    const bool valid = cell < ncl;
    const unsigned xyz = valid ? W.cxyz[cell] : 0u;
    const int si = S.i0 + (int)(xyz & 3u), sj = S.j0 + (int)((xyz >> 2) & 3u), sk = S.k0 + (int)(xyz >> 4);
    const uint32_t dm = valid ? cell_dirs(A, si, sj, sk) : 0u;
    const float sx = lat_pos(p, 0, si), sy = lat_pos(p, 1, sj), sz = lat_pos(p, 2, sk);
    // reach of a lattice step from this cell, with the tolerances of the ray parameter: (1 + eps) |dir|, 2 % on top
    const float L = 1.02f * (1.0f + p.eps) * fmaxf(fmaxf(fabsf(lat_pos(p, 0, si + 1) - sx), fabsf(lat_pos(p, 0, si - 1) - sx)),
                                                  fmaxf(fmaxf(fabsf(lat_pos(p, 1, sj + 1) - sy), fabsf(lat_pos(p, 1, sj - 1) - sy)),
                                                        fmaxf(fabsf(lat_pos(p, 2, sk + 1) - sz), fabsf(lat_pos(p, 2, sk - 1) - sz))));
    const float eps2 = 2.0f * p.eps;
    const float lim_rej = A.rrej * 1.001f, lim_alive = A.dist_dead ? 0.f : 1.45f * A.drw2;
    uint32_t need = valid ? (dm & ~W.coll[cell]) : 0u;   // refreshed after every drain (the only place coll changes)
    for (int q = 0; q < cnt; q++) {
      if (!__any_sync(0xffffffffu, need != 0u)) break;   // every cell of this pass has all its directions blocked
      bool c = false;
      if (need) {
        const float4 R0 = W.R0[q], R1 = W.R1[q];
        const float wx = sx - R1.x, wy = sy - R1.y, wz = sz - R1.z;
        const float aa = fabsf(wx * R0.x + wy * R0.y + wz * R0.z);   // approximate: every threshold below has slack for its rounding
        if (R0.w == 0.f || A.nofilter) c = true;                     // no assumption about this triangle: exact path
        else if (aa <= lim_rej) {
          if (aa * aa < lim_alive) c = true;                         // the distance test may hold: exact path
          else if (aa <= L || aa < eps2) {
            const float rr = W.rad[q] + L;
            c = R0.w < 2.f || wx * wx + wy * wy + wz * wz <= rr * rr;   // a lattice step from here can hit inside the triangle's sphere
            if (!c && aa < eps2)   // in the plane: while a direction parallel to the plane is open every record counts
              c = (need & ((fabsf(R0.x) * L < eps2 ? 3u : 0u) | (fabsf(R0.y) * L < eps2 ? 12u : 0u) | (fabsf(R0.z) * L < eps2 ? 48u : 0u))) != 0u;
          }
        }
      }
      const unsigned bal = __ballot_sync(0xffffffffu, c);
      if (bal) {
        if (c) W.queue[(qh + qn + __popc(bal & ((1u << lane) - 1u))) & (RB_QUEUE - 1)] = (unsigned short)((cell << 8) | q);
        qn += __popc(bal);
        __syncwarp();
        if (qn >= 32) {
          drain(32);
          need = valid ? (dm & ~W.coll[cell]) : 0u;
        }
      }
    }
    if (qn > 0) drain(qn);
  }
}

#ifndef FILL_LB
#define FILL_LB 2
#endif
__global__ void __launch_bounds__(RB_WARPS * 32, FILL_LB) k_resolve_blocks(const __grid_constant__ FillArgs A)
{
  __shared__ RBWarp sW[RB_WARPS];
  const int lane = threadIdx.x & 31;
  const cx_params &p = A.p;
  RBWarp &W = sW[threadIdx.x >> 5];
  const float radf = 1.01f * (1.0f + 3.0f * p.eps);
  for (;;) {
    long long item = 0;
    if (lane == 0) item = (long long)atomicAdd(A.count64, 1ull);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= A.nblist) break;
    const long long c = A.blist[item];
    const int gz = (int)(c % p.gridn[2]), gy = (int)((c / p.gridn[2]) % p.gridn[1]), gx = (int)(c / ((long long)p.gridn[2] * p.gridn[1]));
    // lattice index ranges of the cells that map to this 3dr-cell: lanes 0..5 search one bound each
    int bound = 0;
    if (lane < 6) {
      const int ax = lane >> 1, g = (ax == 0 ? gx : ax == 1 ? gy : gz) + (lane & 1);
      bound = idx_lo(p, ax, g, A.dimg[ax]);
    }
    const int bi0 = __shfl_sync(0xffffffffu, bound, 0), bi1 = __shfl_sync(0xffffffffu, bound, 1);
    const int bj0 = __shfl_sync(0xffffffffu, bound, 2), bj1 = __shfl_sync(0xffffffffu, bound, 3);
    const int bk0 = max(__shfl_sync(0xffffffffu, bound, 4), A.k_lo), bk1 = min(__shfl_sync(0xffffffffu, bound, 5), A.k_hi);
    if (bi1 <= bi0 || bj1 <= bj0 || bk1 <= bk0) continue;
    // the neighbourhood as 9 (x,y) columns; the z-neighbours of a column are adjacent in memory, so one column is one run of
    // records (plus, at a periodic z border, the wrapped cell as a run of its own): lane l < 27 holds run l (9 columns x 3 parts)
    int rb = 0, re = 0;
    if (lane < 27) {
      const int c9 = lane / 3, part = lane % 3;
      const int cx = gx + c9 / 3 - 1, cy = gy + c9 % 3 - 1;
      if (part == 0) {
        const int z0 = max(gz - 1, 0), z1 = min(gz + 1, p.gridn[2] - 1);
        const int64_t q0 = nb_cell(p, cx, cy, z0);
        if (q0 >= 0) { rb = __ldg(&A.cell_idx[q0]); re = __ldg(&A.cell_idx[q0 + (z1 - z0) + 1]); }
      } else {
        const int zz = part == 1 ? gz - 1 : gz + 1;
        if (zz < 0 || zz >= p.gridn[2]) {   // out of range: wrapped (periodic z) or skipped
          const int64_t q0 = nb_cell(p, cx, cy, zz);
          if (q0 >= 0) { rb = __ldg(&A.cell_idx[q0]); re = __ldg(&A.cell_idx[q0 + 1]); }
        }
      }
    }
    // sub-blocks of at most 4 x 4 x 4 lattice cells (a 3dr-cell holds 3-4 cells per axis; border cells of a container that is
    // larger than the domain hold more)
    for (int k0 = bk0; k0 < bk1; k0 += 4)
      for (int j0 = bj0; j0 < bj1; j0 += 4)
        for (int i0 = bi0; i0 < bi1; i0 += 4) {
          RBSub S;
          S.i0 = i0; S.j0 = j0; S.k0 = k0;
          S.nx = min(4, bi1 - i0); S.ny = min(4, bj1 - j0); S.nz = min(4, bk1 - k0);
          const int ncl = S.nx * S.ny * S.nz;
          W.coll[lane] = 0u; W.coll[lane + 32] = 0u;
          for (int q = lane; q < ncl; q += 32) W.cxyz[q] = (unsigned char)((q % S.nx) | (((q / S.nx) % S.ny) << 2) | ((q / (S.nx * S.ny)) << 4));
          __syncwarp();
          // centre and half extents of the sub-block, for the block-level plane-distance reject
          const float xlo = lat_pos(p, 0, i0), xhi = lat_pos(p, 0, i0 + S.nx - 1), ylo = lat_pos(p, 1, j0), yhi = lat_pos(p, 1, j0 + S.ny - 1);
          const float zlo = lat_pos(p, 2, k0), zhi = lat_pos(p, 2, k0 + S.nz - 1);
          const float bcx = 0.5f * (xlo + xhi), bcy = 0.5f * (ylo + yhi), bcz = 0.5f * (zlo + zhi);
          const float bhx = 0.5f * (xhi - xlo), bhy = 0.5f * (yhi - ylo), bhz = 0.5f * (zhi - zlo);
          int cnt = 0;
          for (int run = 0; run < 27; run++) {
            const int b = __shfl_sync(0xffffffffu, rb, run), e = __shfl_sync(0xffffffffu, re, run);
            for (int base = b; base < e; base += 32) {
              const int i = base + lane;
              bool keep = false;
              float4 R0 = make_float4(0.f, 0.f, 0.f, 0.f), R1 = R0;
              float rad = 0.f;
              if (i < e) {
                const float4 *r = A.rec + (int64_t)i * FREC;
                R0 = __ldg(r); R1 = __ldg(r + 1);
                const float tc = (bcx - R1.x) * R0.x + (bcy - R1.y) * R0.y + (bcz - R1.z) * R0.z;
                const float hr = fabsf(R0.x) * bhx + fabsf(R0.y) * bhy + fabsf(R0.z) * bhz;
                keep = !(R0.w != 0.f && fabsf(tc) - hr > A.rrej * 1.001f) && !(A.skip3 && R0.w == 3.f);
                // an accepted (t1, t2) puts the hit point within (1 + 3 eps) max(|u|, |v|) of v0 (well-conditioned triangles), 1 % on top
                if (keep && R0.w >= 2.f) rad = radf * sqrtf(fmaxf(R1.w, __ldg(r + 3).w));
              }
              const unsigned bal = __ballot_sync(0xffffffffu, keep);
              if (!bal) continue;
              if (keep) {
                const int at = cnt + __popc(bal & ((1u << lane) - 1u));
                W.R0[at] = R0; W.R1[at] = R1; W.idx[at] = i; W.rad[at] = rad;
              }
              cnt += __popc(bal);
              __syncwarp();
              if (cnt > RB_CAP - 32) {
                sub_vs_staged(A, W, S, cnt, lane);
                cnt = 0;
                __syncwarp();
              }
            }
          }
          if (cnt) {
            sub_vs_staged(A, W, S, cnt, lane);
            __syncwarp();
          }
          for (int cell = lane; cell < ncl; cell += 32) {
            const int si = i0 + cell % S.nx, sj = j0 + (cell / S.nx) % S.ny, sk = k0 + cell / (S.nx * S.ny);
            const uint32_t cl = W.coll[cell] & cell_dirs(A, si, sj, sk);
            if (!cl) continue;
            const int64_t wi = ((int64_t)sk * A.dimg[1] + sj) * A.wr + (si >> 5);
            const uint32_t bit = 1u << (si & 31);
#pragma unroll
            for (int d = 0; d < 6; d++)
              if (cl >> d & 1) atomicOr(&A.B[d * A.nw + wi], bit);
          }
          __syncwarp();
        }
  }
}

// work list of k_resolve_blocks: the 3dr-cells whose neighbourhood holds segments and that this rank resolves.  One list entry
// per cell (a warp takes one at a time): with coarser items a warp that drew a chunk of 32 wall cells ran for milliseconds
// after all others had finished.
__global__ void __launch_bounds__(256) k_block_list(cx_params p, const uint8_t *__restrict__ nbflag, int nparts, int part, int32_t *list,
                                                    unsigned long long *count)
{
  const long long n = p.gridn[3];
  for (long long c0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) & ~31ll; c0 < n; c0 += (long long)gridDim.x * blockDim.x) {
    const long long c = c0 + (threadIdx.x & 31);
    const bool mine = c < n && (nbflag[c] & 1) && (nparts <= 1 || chunk_owner(c >> 5, nparts) == part);
    const unsigned bal = __ballot_sync(0xffffffffu, mine);
    if (!bal) continue;
    unsigned long long base = 0;
    if ((threadIdx.x & 31) == 0) base = atomicAdd(count, (unsigned long long)__popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (mine && list) list[base + __popc(bal & ((1u << (threadIdx.x & 31)) - 1u))] = (int32_t)c;
  }
}

// ---- record-driven resolution (fine dr: dead distance test; records of flag 3) ----------------------------------------------
// With the distance test dead (eps >= dr_wall^2, crixus_d.cu:940 can never hold) a wall segment blocks a lattice cell in two ways
// only (fill_enum.h): a lattice step CROSSES the triangle -- then the cell lies in the triangle's bounding box across the step and
// within one step of it along the step: a handful of cells per segment, enumerated by the segment (k_resolve_cross) -- or the cell
// lies IN THE PLANE of a triangle that is parallel to the step, whatever the triangle's extent, as far as the influence radius
// reaches -- found per 3dr-cell from the segments of its 27 neighbours (k_resolve_inplane).  Both only produce candidates; the
// reference's arithmetic (pair_exact / pair_full) decides.  tests/test_resolve_enum.py checks the enumeration against the oracle
// on the CPU, tests/test_gpu_parity.py the blocked planes against the cell-driven kernel.
__device__ __forceinline__ void block_dirs(const FillArgs &A, int si, int sj, int sk, uint32_t got)
{
  const int64_t wi = ((int64_t)sk * A.dimg[1] + sj) * A.wr + (si >> 5);
  const uint32_t bit = 1u << (si & 31);
#pragma unroll
  for (int d = 0; d < 6; d++)
    if (got >> d & 1) atomicOr(&A.B[d * A.nw + wi], bit);
}

#ifndef RX_LB
#define RX_LB 6
#endif
// one lane per wall segment; the lanes of a warp walk their candidate boxes in lock-step (neighbouring segments of a wall have
// congruent boxes)
__global__ void __launch_bounds__(128, RX_LB) k_resolve_cross(const __grid_constant__ FillArgs A)
{
  const cx_params &p = A.p;
  const int64_t nwall = A.nbe - A.cnbe;
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  // z-slab ownership: only cells of the planes [k_lo, k_hi) are written
  const float zlo = lat_pos(p, 2, A.k_lo) - A.lstep, zhi = lat_pos(p, 2, A.k_hi - 1) + A.lstep;
  for (int64_t base = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32; base < nwall; base += nwarps * 32) {
    const int64_t i = base + lane;
    bool live = false;
    FeRec T;
    float4 R0 = make_float4(0.f, 0.f, 0.f, 0.f), R1 = R0;
    if (i < nwall) {
      const float4 *r = A.rec + i * FREC;
      R0 = __ldg(r);
      if (R0.w == 3.f) {
        R1 = __ldg(r + 1);
        const float4 R2 = __ldg(r + 2), R3 = __ldg(r + 3);
        const int g[3] = {1, 1, 1};   // flag 3 = "safe" was established by k_seg_prep: the cell is not needed again
        const float na[3] = {R0.x, R0.y, R0.z}, a0[3] = {R1.x, R1.y, R1.z}, a1[3] = {R2.x, R2.y, R2.z}, a2[3] = {R3.x, R3.y, R3.z};
        fe_rec_setup(&p, A.lstep, na, a0, a1, a2, R1.w, R3.w, g, &T);
        live = !(T.bhi[2] < zlo || T.blo[2] > zhi);
      }
    }
    if (!__any_sync(0xffffffffu, live)) continue;
#pragma unroll
    for (int ax = 0; ax < 3; ax++) {
      int lo0 = 0, hi0 = -1, lo1 = 0, hi1 = -1, lo2 = 0, hi2 = -1;
      unsigned cnt = 0;
      if (live && fe_axis_crossing(&p, A.lstep, &T, ax)) {
        fe_axis_range(&p, &T, 0, ax == 0 ? A.lstep : 0.f, 0, A.dimg[0], &lo0, &hi0);
        fe_axis_range(&p, &T, 1, ax == 1 ? A.lstep : 0.f, 0, A.dimg[1], &lo1, &hi1);
        fe_axis_range(&p, &T, 2, ax == 2 ? A.lstep : 0.f, A.k_lo, A.k_hi, &lo2, &hi2);
        if (lo0 <= hi0 && lo1 <= hi1 && lo2 <= hi2) {
          const unsigned long long v = (unsigned long long)(hi0 - lo0 + 1) * (unsigned long long)(hi1 - lo1 + 1) * (unsigned long long)(hi2 - lo2 + 1);
          cnt = v > 0x7fffffffull ? 0x7fffffffu : (unsigned)v;   // (flag 3: the box lies within the 27 neighbour cells)
        }
      }
      const unsigned mx = __reduce_max_sync(0xffffffffu, cnt);
      int ci = lo0, cj = lo1, ck = lo2;
#pragma unroll 1
      for (unsigned t = 0; t < mx; t++) {
        if (t < cnt) {
          bool cand = fe_cross_candidate(&p, A.lstep, &T, ax, ci, cj, ck) != 0;
          if (cand && A.res_nparts > 1) {   // directed tests dealt to the ranks by 3dr-cell chunk (k_block_list)
            const long long c3 = ((long long)lat_cell(p, 0, ci) * p.gridn[1] + lat_cell(p, 1, cj)) * p.gridn[2] + lat_cell(p, 2, ck);
            cand = chunk_owner(c3 >> 5, A.res_nparts) == A.res_part;
          }
          if (cand) {
            const uint32_t need = cell_dirs(A, ci, cj, ck) & (3u << (2 * ax));
            if (need) {
              const uint32_t got = pair_exact(A, R0, R1, T.rad, i, ci, cj, ck, need);
              if (got) block_dirs(A, ci, cj, ck, got);
            }
          }
          if (++ci > hi0) {
            ci = lo0;
            if (++cj > hi1) { cj = lo1; ++ck; }
          }
        }
      }
    }
  }
}

// One warp per 3dr-cell of the work list (as k_resolve_blocks), sub-blocks of at most 4 x 4 x 4 lattice cells with the fixed
// numbering q = x | y << 2 | z << 4.  Lanes = segments of the 27 neighbour cells, 32 at a time: a lane whose segment has an axis
// parallel to its plane ("capable", fe_capable) and whose plane passes through the sub-block marks the cells within 2 eps of the
// plane in a 64-bit mask; cells that still have an open direction along a capable axis are tested with the reference's
// arithmetic (pair_full) -- one lane per distinct cell and round, each lane its own segment.  At a wall in a lattice plane the
// first 32 segments block every in-plane cell and the rest of the neighbourhood costs the mask only.
__global__ void __launch_bounds__(RB_WARPS * 32) k_resolve_inplane(const __grid_constant__ FillArgs A)
{
  __shared__ uint32_t sColl[RB_WARPS][64];
  const int lane = threadIdx.x & 31;
  const cx_params &p = A.p;
  uint32_t *coll = sColl[threadIdx.x >> 5];
  const float eps2 = 2.0f * p.eps;
  const float reach2 = fe_infl_reach(&p, A.lstep) * fe_infl_reach(&p, A.lstep);   // beyond it the influence-radius test fails
  for (;;) {
    long long item = 0;
    if (lane == 0) item = (long long)atomicAdd(A.count64, 1ull);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= A.nblist) break;
    const long long c = A.blist[item];
    const int gz = (int)(c % p.gridn[2]), gy = (int)((c / p.gridn[2]) % p.gridn[1]), gx = (int)(c / ((long long)p.gridn[2] * p.gridn[1]));
    int bound = 0;
    if (lane < 6) {
      const int ax = lane >> 1, g = (ax == 0 ? gx : ax == 1 ? gy : gz) + (lane & 1);
      bound = idx_lo(p, ax, g, A.dimg[ax]);
    }
    const int bi0 = __shfl_sync(0xffffffffu, bound, 0), bi1 = __shfl_sync(0xffffffffu, bound, 1);
    const int bj0 = __shfl_sync(0xffffffffu, bound, 2), bj1 = __shfl_sync(0xffffffffu, bound, 3);
    const int bk0 = max(__shfl_sync(0xffffffffu, bound, 4), A.k_lo), bk1 = min(__shfl_sync(0xffffffffu, bound, 5), A.k_hi);
    if (bi1 <= bi0 || bj1 <= bj0 || bk1 <= bk0) continue;
    // lane l < 27 holds neighbour cell l (periodic wrap as in the reference's lookup): its run of records, whether its segments
    // share one plane (k_cell_planes) and the plane of its first segment
    int rb = 0, re = 0, mixed = 0;
    float4 P0 = make_float4(0.f, 0.f, 0.f, __int_as_float(0x7f800000));
    if (lane < 27) {
      const int64_t q0 = nb_cell(p, gx + lane / 9 - 1, gy + (lane / 3) % 3 - 1, gz + lane % 3 - 1);
      if (q0 >= 0) {
        rb = __ldg(&A.cell_idx[q0]); re = __ldg(&A.cell_idx[q0 + 1]);
        if (re > rb) {
          mixed = (__ldg(&A.nbflag[q0]) & 2) ? 1 : 0;
          P0 = __ldg(&A.plane[rb]);
        }
      }
    }
    const bool p0fin = P0.w < 3.0e38f && P0.w > -3.0e38f;
    uint32_t cap0 = 0u;
    if (p0fin) {
      const float na[3] = {P0.x, P0.y, P0.z};
      cap0 = fe_capable(&p, A.lstep, na);
    }
    for (int k0 = bk0; k0 < bk1; k0 += 4)
      for (int j0 = bj0; j0 < bj1; j0 += 4)
        for (int i0 = bi0; i0 < bi1; i0 += 4) {
          const int nx = min(4, bi1 - i0), ny = min(4, bj1 - j0), nz = min(4, bk1 - k0);
          coll[lane] = 0u; coll[lane + 32] = 0u;
          __syncwarp();
          // this lane's two cells (q = lane, lane + 32) and the directions that exist for them
          const int lx = lane & 3, ly = (lane >> 2) & 3, lz = lane >> 4;
          const uint32_t dmA = (lx < nx && ly < ny && lz < nz) ? cell_dirs(A, i0 + lx, j0 + ly, k0 + lz) : 0u;
          const uint32_t dmB = (lx < nx && ly < ny && lz + 2 < nz) ? cell_dirs(A, i0 + lx, j0 + ly, k0 + lz + 2) : 0u;
          unsigned long long openx, openy, openz;   // cells with an open direction along x / y / z
          auto refresh = [&]() {
            const uint32_t na = dmA & ~coll[lane], nb = dmB & ~coll[lane + 32];
            openx = (unsigned long long)__ballot_sync(0xffffffffu, (na & 3u) != 0u) | ((unsigned long long)__ballot_sync(0xffffffffu, (nb & 3u) != 0u) << 32);
            openy = (unsigned long long)__ballot_sync(0xffffffffu, (na & 12u) != 0u) | ((unsigned long long)__ballot_sync(0xffffffffu, (nb & 12u) != 0u) << 32);
            openz = (unsigned long long)__ballot_sync(0xffffffffu, (na & 48u) != 0u) | ((unsigned long long)__ballot_sync(0xffffffffu, (nb & 48u) != 0u) << 32);
          };
          refresh();
          // lattice positions of the sub-block relative to the lattice origin (fe_plane_offset), its centre and half extents
          // (plane-through-the-block test)
          float px[4], py[4], pz[4];
#pragma unroll
          for (int t = 0; t < 4; t++) {
            px[t] = lat_pos(p, 0, i0 + min(t, nx - 1)) - p.fcmin[0];
            py[t] = lat_pos(p, 1, j0 + min(t, ny - 1)) - p.fcmin[1];
            pz[t] = lat_pos(p, 2, k0 + min(t, nz - 1)) - p.fcmin[2];
          }
          const float bcx = 0.5f * (px[0] + px[3]), bcy = 0.5f * (py[0] + py[3]), bcz = 0.5f * (pz[0] + pz[3]);
          const float bhx = 0.5f * (px[3] - px[0]), bhy = 0.5f * (py[3] - py[0]), bhz = 0.5f * (pz[3] - pz[0]);
          float cnx = 0.f, cny = 0.f, cnz = 0.f, cd = 0.f;   // the remembered plane (warp-uniform)
          bool cvalid = false;
          // cells to look at: mixed ones, and single-plane cells whose plane passes through the sub-block
          bool look = re > rb && mixed;
          if (re > rb && !mixed && cap0) {
            const float tc = fmaf(bcx, P0.x, fmaf(bcy, P0.y, bcz * P0.z)) - P0.w;
            const float hr = fabsf(P0.x) * bhx + fabsf(P0.y) * bhy + fabsf(P0.z) * bhz;
            look = fabsf(tc) - hr < eps2 * 1.01f;
          }
          unsigned todo = __ballot_sync(0xffffffffu, look);
          while (todo) {
            const int src = (todo >> 13 & 1u) ? 13 : __ffs(todo) - 1;   // the 3dr-cell itself first: its segments are the nearest
            todo &= ~(1u << src);
            const int b = __shfl_sync(0xffffffffu, rb, src), e = __shfl_sync(0xffffffffu, re, src);
            const int single = !__shfl_sync(0xffffffffu, mixed, src);
            const float snx = __shfl_sync(0xffffffffu, P0.x, src), sny = __shfl_sync(0xffffffffu, P0.y, src);
            const float snz = __shfl_sync(0xffffffffu, P0.z, src), sd = __shfl_sync(0xffffffffu, P0.w, src);
            for (int base = b; base < e; base += 32) {
              // a single-plane cell whose plane is (within eps / 2) the remembered finished one has nothing left to give
              if (single && cvalid && snx == cnx && sny == cny && snz == cnz && fabsf(sd - cd) < 0.5f * p.eps) break;
              if (!(openx | openy | openz)) break;   // every direction of every cell is blocked
              const int i = base + lane;
              bool live = false;
              uint32_t cap = 0u;
              float4 P = make_float4(0.f, 0.f, 0.f, 0.f);
              if (i < e) {
                P = __ldg(&A.plane[i]);
                if (P.w < 3.0e38f && P.w > -3.0e38f) {   // (n, d) of a segment that takes part in the in-plane rule
                  const float na[3] = {P.x, P.y, P.z};
                  cap = fe_capable(&p, A.lstep, na);
                  const float tc = fmaf(bcx, P.x, fmaf(bcy, P.y, bcz * P.z)) - P.w;
                  const float hr = fabsf(P.x) * bhx + fabsf(P.y) * bhy + fabsf(P.z) * bhz;
                  // a plane that is (within eps / 2) one whose in-plane cells are all blocked along its capable axes has none left
                  const bool known = cvalid && P.x == cnx && P.y == cny && P.z == cnz && fabsf(P.w - cd) < 0.5f * p.eps;
                  live = cap != 0u && !known && fabsf(tc) - hr < eps2 * 1.01f;
                }
              }
              if (!__any_sync(0xffffffffu, live)) continue;
              unsigned long long cm = 0ull, cmp = 0ull;   // cells within 2 eps of this lane's plane: cmp; and within the influence radius: cm
              if (live) {
                float4 R1 = __ldg(A.rec + (int64_t)i * FREC + 1);   // v0, for the influence radius
                R1.x -= p.fcmin[0]; R1.y -= p.fcmin[1]; R1.z -= p.fcmin[2];
#pragma unroll
                for (int z = 0; z < 4; z++) {
                  if (z < nz) {
                    const float wz = pz[z] - R1.z;
                    const float tz = fmaf(pz[z], P.z, -P.w), dz2 = wz * wz;
#pragma unroll
                    for (int y = 0; y < 4; y++) {
                      if (y < ny) {
                        const float wy = py[y] - R1.y;
                        const float tyz = fmaf(py[y], P.y, tz), dyz2 = fmaf(wy, wy, dz2);
#pragma unroll
                        for (int x = 0; x < 4; x++) {
                          const float wx = px[x] - R1.x;
                          if (x < nx && fabsf(fmaf(px[x], P.x, tyz)) < eps2) {
                            cmp |= 1ull << (x | (y << 2) | (z << 4));
                            if (fmaf(wx, wx, dyz2) <= reach2) cm |= 1ull << (x | (y << 2) | (z << 4));
                          }
                        }
                      }
                    }
                  }
                }
              }
              for (;;) {
                const unsigned long long cand =
                    live ? (cm & (((cap & 3u) ? openx : 0ull) | ((cap & 12u) ? openy : 0ull) | ((cap & 48u) ? openz : 0ull))) : 0ull;
                if (!__any_sync(0xffffffffu, cand != 0ull)) break;
                int q = 64 + lane;
                if (cand) {   // spread the lanes over the set bits, so that a round tests as many distinct cells as possible
                  unsigned long long tmp = cand;
                  for (int k = (lane * __popcll(cand)) >> 5; k > 0; k--) tmp &= tmp - 1ull;
                  q = __ffsll((long long)tmp) - 1;
                }
                const unsigned grp = __match_any_sync(0xffffffffu, q);
                if (cand && (__ffs(grp) - 1) == lane) {   // one lane per cell
                  cm &= ~(1ull << q);
                  const int si = i0 + (q & 3), sj = j0 + ((q >> 2) & 3), sk = k0 + (q >> 4);
                  const uint32_t need = cell_dirs(A, si, sj, sk) & ~coll[q] & cap;
                  if (need) {
                    const uint32_t got = pair_full(A, si, sj, sk, i, need);
                    if (got) atomicOr(&coll[q], got);
                  }
                }
                __syncwarp();
                refresh();
              }
              // remember one plane whose in-plane cells (all of them, also those beyond this segment's influence radius) are done
              const bool done = live && (cmp & (((cap & 3u) ? openx : 0ull) | ((cap & 12u) ? openy : 0ull) | ((cap & 48u) ? openz : 0ull))) == 0ull;
              const unsigned dn = __ballot_sync(0xffffffffu, done);
              if (dn) {
                const int src = __ffs(dn) - 1;
                cnx = __shfl_sync(0xffffffffu, P.x, src); cny = __shfl_sync(0xffffffffu, P.y, src);
                cnz = __shfl_sync(0xffffffffu, P.z, src); cd = __shfl_sync(0xffffffffu, P.w, src);
                cvalid = true;
              }
            }
          }
          __syncwarp();
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const int q = lane + 32 * h;
            const uint32_t cl = coll[q] & (h ? dmB : dmA);
            if (cl) block_dirs(A, i0 + (q & 3), j0 + ((q >> 2) & 3), k0 + (q >> 4), cl);
          }
          __syncwarp();
        }
  }
}

// free-surface triangles: all of them against every lattice cell, segment test only, no cell lookup (crixus_d.cu:951-960).
// One thread per (triangle, lattice row); the cells of the row that can be hit are an index interval (FsDesc).
__global__ void __launch_bounds__(128) k_resolve_fs(FillArgs A)
{
  const cx_params &p = A.p;
  const int64_t rows = (int64_t)A.dimg[1] * (A.k_hi - A.k_lo);
  const int64_t total = rows * A.cnbe;
  const int nparts = A.res_nparts > 1 ? A.res_nparts : 1;
  for (int64_t it = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; it < total; it += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(it / rows);
    const int64_t row = it - (int64_t)t * rows;
    const int j = (int)(row % A.dimg[1]), k = A.k_lo + (int)(row / A.dimg[1]);
    const FsDesc f = A.fs[t];
    const float sy = lat_pos(p, 1, j), sz = lat_pos(p, 2, k);
    int ia = 0, ib = A.dimg[0] - 1;
    if (f.cull && !A.nofilter) {
      ia = A.dimg[0]; ib = -1;
      if (sy >= f.lo[1] && sy <= f.hi[1] && sz >= f.lo[2] && sz <= f.hi[2]) {   // the bounding box part
        const double a = floor(((double)f.lo[0] - p.fcmin[0]) / p.dr) - 1.0, b = ceil(((double)f.hi[0] - p.fcmin[0]) / p.dr) + 1.0;
        ia = (int)fmax(a, 0.0); ib = (int)fmin(b, (double)A.dimg[0] - 1.0);
      }
      if (f.quirk) {   // cells within 4 eps of the plane: |a0 + i c| < 4 eps
        const double a0 = (double)f.n[0] * ((double)p.fcmin[0] - f.v0[0]) + (double)f.n[1] * ((double)sy - f.v0[1]) +
                          (double)f.n[2] * ((double)sz - f.v0[2]);
        const double c = (double)f.n[0] * p.dr, band = 4.0 * p.eps + 1e-5 * (fabs(a0) + fabs(c) * A.dimg[0]);
        int qa, qb;
        if (fabs(c) * A.dimg[0] < 0.25 * band) { qa = fabs(a0) < 2.0 * band ? 0 : A.dimg[0]; qb = fabs(a0) < 2.0 * band ? A.dimg[0] - 1 : -1; }
        else {
          const double i1 = (-a0 - band) / c, i2 = (-a0 + band) / c;
          qa = (int)fmax(floor(fmin(i1, i2)) - 1.0, 0.0); qb = (int)fmin(ceil(fmax(i1, i2)) + 1.0, (double)A.dimg[0] - 1.0);
        }
        if (qa <= qb) { ia = min(ia, qa); ib = max(ib, qb); }
      }
    }
    if (ia > ib) continue;
    const int64_t ri = A.nbe - A.cnbe + t;
    const float4 *r = A.rec + ri * FREC;
    const float4 R0 = __ldg(r), R1 = __ldg(r + 1), R2 = __ldg(r + 2), R3 = __ldg(r + 3), R4 = __ldg(r + 4);
    const f3 n = mk3(R0), v0 = mk3(R1), u = sub3(mk3(R2), v0), v = sub3(mk3(R3), v0);
    const int gy = nparts > 1 ? cell_coord(sy, p.dmin[1], p.griddr[1], p.gridn[1]) : 0;
    const int gz = nparts > 1 ? cell_coord(sz, p.dmin[2], p.griddr[2], p.gridn[2]) : 0;
    for (int i = ia; i <= ib; i++) {
      const float sx = lat_pos(p, 0, i);
      if (nparts > 1) {   // the rank that resolves the cell's 3dr-cell owns the cell (disjoint bits: a SUM merge is an OR)
        const int gx = cell_coord(sx, p.dmin[0], p.griddr[0], p.gridn[0]);
        const long long c = ((long long)gx * p.gridn[1] + gy) * p.gridn[2] + gz;
        if (chunk_owner(c >> 5, nparts) != A.res_part) continue;
      }
      const f3 s = mk3(sx, sy, sz);
      const int idx[3] = {i, j, k};
      uint32_t cl = 0;
#pragma unroll
      for (int d = 0; d < 6; d++) {
        const int ax = d >> 1, e = idx[ax] + ((d & 1) ? 1 : -1);
        if (e < 0 || e >= A.dimg[ax]) continue;
        const float dc = fsub(lat_pos(p, ax, e), ax == 0 ? sx : ax == 1 ? sy : sz);
        const f3 dir = mk3(ax == 0 ? dc : 0.f, ax == 1 ? dc : 0.f, ax == 2 ? dc : 0.f);
        if (tri_collision(A, s, dir, n, v0, u, v, R1.w, R2.w, R3.w, R4.w)) cl |= 1u << d;
      }
      if (!cl) continue;
      const int64_t wi = ((int64_t)k * A.dimg[1] + j) * A.wr + (i >> 5);
#pragma unroll
      for (int d = 0; d < 6; d++)
        if (cl >> d & 1) atomicOr(&A.B[d * A.nw + wi], 1u << (i & 31));
    }
  }
}

// conservative fshape descriptors, one thread per free-surface triangle (DESIGN.md "fill / free surface").
// A triangle is "cullable" when its stored normal is consistent with its plane and it is well conditioned; then
// checkTriangleCollision(s, dir, .) can only be true for s inside the triangle's bounding box grown by the margin below, or
// (ray-in-plane rule, crixus_d.cu:766-769) within 4 eps of its plane when some axis direction is parallel to the plane within
// eps.  eps is ABSOLUTE (1e-5 x the domain extent) but enters the segment test as a RELATIVE tolerance (r <= 1 + eps,
// t1, t2 >= -eps in barycentric units), so every margin scales with it: the ray may start (1 + eps) dr away, and the accepted
// barycentric region overhangs the triangle by eps (|u| + |v|).
__global__ void k_fs_desc(FillArgs A, FsDesc *out)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= A.cnbe) return;
  const int64_t i = A.nbe - A.cnbe + t;
  const float4 n4 = A.norm[i];
  const int4 E = A.ep[i];
  const float4 a4 = A.pos[E.x], b4 = A.pos[E.y], c4 = A.pos[E.z];
  const float n[3] = {n4.x, n4.y, n4.z}, a[3] = {a4.x, a4.y, a4.z}, b[3] = {b4.x, b4.y, b4.z}, c[3] = {c4.x, c4.y, c4.z};
  double lu = 0, lv = 0, ln = 0, nu = 0, nv = 0, uv = 0;
  for (int k = 0; k < 3; k++) {
    const double u = (double)b[k] - a[k], v = (double)c[k] - a[k];
    lu += u * u; lv += v * v; ln += (double)n[k] * n[k];
    nu += (double)n[k] * u; nv += (double)n[k] * v; uv += u * v;
  }
  lu = sqrt(lu); lv = sqrt(lv); ln = sqrt(ln);
  const double eps = A.p.eps;
  FsDesc d;
  d.cull = ln > 0 && lu > 0 && lv > 0 && fabs(nu) <= 1e-3 * ln * lu && fabs(nv) <= 1e-3 * ln * lv &&
           (lu * lu * lv * lv - uv * uv) >= 1e-4 * lu * lu * lv * lv && eps < 0.25;
  const double margin = (1.5 + eps) * A.p.dr + (0.01 + 2.0 * eps) * (lu + lv) + 10.0 * eps;
  d.quirk = 0;
  for (int k = 0; k < 3; k++) {
    const float mn = fminf(a[k], fminf(b[k], c[k])), mx = fmaxf(a[k], fmaxf(b[k], c[k]));
    d.lo[k] = (float)(mn - margin); d.hi[k] = (float)(mx + margin);
    d.n[k] = n[k]; d.v0[k] = a[k];
    if (fabs((double)n[k]) * A.p.dr * 0.5 < eps) d.quirk = 1;
  }
  out[t] = d;
}

// ---- one round -------------------------------------------------------------------------------------------------
__global__ void k_round_prep(int32_t *count) { count[0] = 0; count[1] = 0; }

__global__ void __launch_bounds__(256) k_build_list(FillArgs A)
{
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < A.ntiles; t += (int64_t)gridDim.x * blockDim.x)
    if (A.act_cur[t]) {
      A.act_cur[t] = 0;
      A.list[atomicAdd(&A.count[0], 1)] = (int32_t)t;
    }
}

struct TileSM {
  uint32_t F[FT_Z + 2][FT_Y + 2];  // filled, with y/z halo
  int changed, newbits;
};

// propagation inside active tiles: pure bit arithmetic (all directed tests are already in the B planes)
__global__ void __launch_bounds__(FT_THREADS) k_fill_tiles(FillArgs A)
{
  __shared__ TileSM S;
  __shared__ int s_t;
  __shared__ int s_face[6];   // new cells on the face towards -x, +x, -y, +y, -z, +z: only those neighbours see a change
  const int tid = threadIdx.x;
  const int ly = tid % FT_Y, lz = tid / FT_Y;
  if (blockIdx.x == 0 && tid < 2) A.count_next[tid] = 0;   // list length and cursor of the NEXT round (the two pairs alternate)
  for (;;) {
    if (tid == 0) {
      const int q = atomicAdd(&A.count[1], 1);
      s_t = q < A.count[0] ? A.list[q] : -1;
      S.newbits = 0;
    }
    if (tid < 6) s_face[tid] = 0;
    __syncthreads();
    const int tile = s_t;
    if (tile < 0) break;
    const int wx = tile % A.wr, tyi = (tile / A.wr) % A.ty, tzi = tile / (A.wr * A.ty);
    const int j0 = tyi * FT_Y, k0 = tzi * FT_Z;
    const int nvalid = min(32, A.dimg[0] - wx * 32);
    const uint32_t vmask = nvalid < 32 ? (1u << nvalid) - 1u : 0xffffffffu;
    // ---- load (halo rows outside the lattice read as 0)
    for (int h = tid; h < (FT_Z + 2) * (FT_Y + 2); h += FT_THREADS) {
      const int hz = h / (FT_Y + 2), hy = h % (FT_Y + 2);
      const int j = j0 + hy - 1, k = k0 + hz - 1;
      uint32_t v = 0;
      if (j >= 0 && j < A.dimg[1] && k >= 0 && k < A.dimg[2]) v = A.F[((int64_t)k * A.dimg[1] + j) * A.wr + wx];
      S.F[hz][hy] = v;
    }
    const int j = j0 + ly, k = k0 + lz;
    const bool inrow = j < A.dimg[1] && k < A.dimg[2];
    const bool owned = inrow && k >= A.k_lo && k < A.k_hi;
    const int64_t wi = ((int64_t)k * A.dimg[1] + j) * A.wr + wx;
    uint32_t xl = 0, xr = 0;
    // per direction: cells that may be entered from that neighbour (every cell unless its directed test blocks it)
    uint32_t o0 = vmask, o1 = vmask, o2 = vmask, o3 = vmask, o4 = vmask, o5 = vmask;
    if (owned) {
      if (wx > 0) xl = A.F[wi - 1] >> 31;
      if (wx + 1 < A.wr) xr = A.F[wi + 1] << 31;
      o0 &= ~A.B[0 * A.nw + wi]; o1 &= ~A.B[1 * A.nw + wi]; o2 &= ~A.B[2 * A.nw + wi];
      o3 &= ~A.B[3 * A.nw + wi]; o4 &= ~A.B[4 * A.nw + wi]; o5 &= ~A.B[5 * A.nw + wi];
    }
    __syncthreads();
    const uint32_t f_start = S.F[lz + 1][ly + 1];
    // ---- iterate to local convergence (every productive iteration sets at least one bit)
    for (int iter = 0; iter < 200000; iter++) {
      if (tid == 0) S.changed = 0;
      __syncthreads();
      if (owned) {
        const uint32_t f = S.F[lz + 1][ly + 1];
        uint32_t nf = f | (S.F[lz + 1][ly] & o2) | (S.F[lz + 1][ly + 2] & o3) | (S.F[lz][ly + 1] & o4) | (S.F[lz + 2][ly + 1] & o5);
        for (;;) {   // x direction inside the word (and from the two neighbouring words)
          const uint32_t g = ((((nf << 1) | xl) & o0) | (((nf >> 1) | xr) & o1)) & ~nf;
          if (!g) break;
          nf |= g;
        }
        if (nf != f) {
          S.F[lz + 1][ly + 1] = nf;
          S.changed = 1;
        }
      }
      __syncthreads();
      if (!S.changed) break;
      __syncthreads();
    }
    // ---- write back + activate neighbours
    const uint32_t f_end = S.F[lz + 1][ly + 1];
    if (owned && f_end != f_start) {
      const uint32_t nb = f_end & ~f_start;
      A.F[wi] = f_end;
      atomicAdd(&S.newbits, __popc(nb));
      // a neighbouring tile reads this tile's face only (6-connectivity): bit 0 / bit 31 of the word, first / last row in y and z
      if (nb & 1u) s_face[0] = 1;
      if (nb >> 31) s_face[1] = 1;
      if (ly == 0) s_face[2] = 1;
      if (ly == FT_Y - 1) s_face[3] = 1;
      if (lz == 0) s_face[4] = 1;
      if (lz == FT_Z - 1) s_face[5] = 1;
    }
    __syncthreads();
    if (tid == 0 && S.newbits) atomicAdd(A.changed, (unsigned long long)S.newbits);
    if (tid == 0 && (S.newbits || A.force_activate)) {
      const bool all = A.force_activate != 0;
      if (wx > 0 && (all || s_face[0])) A.act_next[tile - 1] = 1;
      if (wx + 1 < A.wr && (all || s_face[1])) A.act_next[tile + 1] = 1;
      if (tyi > 0 && (all || s_face[2])) A.act_next[tile - A.wr] = 1;
      if (tyi + 1 < A.ty && (all || s_face[3])) A.act_next[tile + A.wr] = 1;
      if (tzi > 0 && (all || s_face[4])) A.act_next[tile - A.wr * A.ty] = 1;
      if (tzi + 1 < A.tz && (all || s_face[5])) A.act_next[tile + A.wr * A.ty] = 1;
    }
    __syncthreads();
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------
struct FillState {
  const uint32_t *fpos_key = nullptr;
  int64_t dimg[3] = {0, 0, 0};
  int64_t nbe = 0, cnbe = 0;
  uint32_t *F = nullptr, *B = nullptr;   // B: 6 blocked planes
  uint32_t *halo = nullptr;              // two received boundary planes (multi-GPU)
  int32_t *blist = nullptr;              // work list of the up-front resolve
  int64_t cap_blist = 0;
  float4 *rec = nullptr, *plane = nullptr;
  int32_t *cellid = nullptr;
  int64_t cap_nw = 0, cap_ntiles = 0, cap_ncell = 0, cap_rec = 0, cap_fs = 0, cap_halo = 0;
  uint8_t *nbflag = nullptr, *nbtmp = nullptr, *act = nullptr;
  int32_t *list = nullptr, *count = nullptr;
  FsDesc *fs = nullptr;
  int64_t nw = 0, ntiles = 0;
  int resolved_lo = 0, resolved_hi = 0;   // planes whose directed tests are in B
  int act_swapped = 0;                    // which half of `act` is the current one (kept across slab calls)
};

void cx_fill_state_free(cx_ctx *ctx)
{
  FillState *s = (FillState *)ctx->fill_state;
  if (!s) return;
  cudaFree(s->F); cudaFree(s->B); cudaFree(s->nbflag); cudaFree(s->act); cudaFree(s->list);
  cudaFree(s->count); cudaFree(s->fs); cudaFree(s->rec); cudaFree(s->halo); cudaFree(s->blist); cudaFree(s->plane); cudaFree(s->cellid); cudaFree(s->nbtmp);
  delete s;
  ctx->fill_state = nullptr;
}

static int fill_prepare(cx_ctx *ctx, const cx_params *p, const uint32_t *fpos_d, const float *norm_d, const int32_t *ep_d,
                        const float *pos_d, int64_t nbe, int64_t cnbe, const int32_t *cell_idx_d, int first, FillArgs *A, FillState **out)
{
  cudaStream_t st = ctx->stream;
  FillState *s = (FillState *)ctx->fill_state;
  int64_t dimg[3];
  cx_fill_dims(p, dimg);
  for (int k = 0; k < 3; k++)
    if (dimg[k] <= 0 || dimg[k] > 0x7fffffff) return cx_fail(ctx, CX_WRONG_INPUT, "fill: bad lattice dimensions");
  const int wr = (int)((dimg[0] + 31) / 32);
  const int64_t nw = (int64_t)wr * dimg[1] * dimg[2];
  const int ty = (int)((dimg[1] + FT_Y - 1) / FT_Y), tz = (int)((dimg[2] + FT_Z - 1) / FT_Z);
  const int64_t ntiles = (int64_t)wr * ty * tz;
  if (ntiles > 0x7fffffff) return cx_fail(ctx, CX_WRONG_INPUT, "fill: too many tiles");
  const bool fresh = first || !s || s->fpos_key != fpos_d || s->dimg[0] != dimg[0] || s->dimg[1] != dimg[1] || s->dimg[2] != dimg[2] ||
                     s->nbe != nbe || s->cnbe != cnbe;
  if (fresh) {
    if (!s) {
      s = new FillState();
      ctx->fill_state = s;
    }
    // buffers are kept between calls and only grow (cudaMalloc/cudaFree cost milliseconds and synchronise)
    if (nw > s->cap_nw) {
      cudaFree(s->F); cudaFree(s->B);
      s->F = s->B = nullptr; s->cap_nw = 0;
      CX_CUDA(ctx, cudaMalloc(&s->F, sizeof(uint32_t) * (size_t)nw));
      CX_CUDA(ctx, cudaMalloc(&s->B, sizeof(uint32_t) * (size_t)nw * 6));
      s->cap_nw = nw;
    }
    if (ntiles > s->cap_ntiles) {
      cudaFree(s->act); cudaFree(s->list);
      s->act = nullptr; s->list = nullptr; s->cap_ntiles = 0;
      CX_CUDA(ctx, cudaMalloc(&s->act, (size_t)ntiles * 2));
      CX_CUDA(ctx, cudaMalloc(&s->list, sizeof(int32_t) * (size_t)ntiles));
      s->cap_ntiles = ntiles;
    }
    if (p->gridn[3] > s->cap_ncell) {
      cudaFree(s->nbflag); cudaFree(s->nbtmp); s->nbflag = s->nbtmp = nullptr; s->cap_ncell = 0;
      CX_CUDA(ctx, cudaMalloc(&s->nbflag, (size_t)p->gridn[3]));
      CX_CUDA(ctx, cudaMalloc(&s->nbtmp, (size_t)p->gridn[3]));
      s->cap_ncell = p->gridn[3];
    }
    if (nbe > s->cap_rec) {
      cudaFree(s->rec); cudaFree(s->plane); cudaFree(s->cellid);
      s->rec = s->plane = nullptr; s->cellid = nullptr; s->cap_rec = 0;
      CX_CUDA(ctx, cudaMalloc(&s->rec, sizeof(float4) * FREC * (size_t)nbe));
      CX_CUDA(ctx, cudaMalloc(&s->plane, sizeof(float4) * (size_t)nbe));
      CX_CUDA(ctx, cudaMalloc(&s->cellid, sizeof(int32_t) * (size_t)nbe));
      s->cap_rec = nbe;
    }
    if (cnbe > s->cap_fs) {
      cudaFree(s->fs); s->fs = nullptr; s->cap_fs = 0;
      CX_CUDA(ctx, cudaMalloc(&s->fs, sizeof(FsDesc) * (size_t)cnbe));
      s->cap_fs = cnbe;
    }
    if (!s->count) CX_CUDA(ctx, cudaMalloc(&s->count, sizeof(int32_t) * 16));   // [0..1] tile list, [4..5] work counter, [6] halo flag, [8] nrest
    s->fpos_key = fpos_d; s->nbe = nbe; s->cnbe = cnbe; s->nw = nw; s->ntiles = ntiles;
    for (int k = 0; k < 3; k++) s->dimg[k] = dimg[k];
    CX_CUDA(ctx, cudaMemsetAsync(s->B, 0, sizeof(uint32_t) * (size_t)nw * 6, st));
    CX_CUDA(ctx, cudaMemsetAsync(s->act, 0, (size_t)ntiles * 2, st));
    s->resolved_lo = s->resolved_hi = 0;
    s->act_swapped = 0;
  }
  memset(A, 0, sizeof(*A));
  A->p = *p;
  A->pos = (const float4 *)pos_d; A->norm = (const float4 *)norm_d; A->ep = (const int4 *)ep_d; A->cell_idx = cell_idx_d;
  A->nbe = nbe; A->cnbe = cnbe;
  for (int k = 0; k < 3; k++) A->dimg[k] = (int)dimg[k];
  A->wr = wr; A->ty = ty; A->tz = tz; A->ntiles = ntiles; A->nw = nw;
  A->k_lo = 0; A->k_hi = (int)dimg[2];
  A->w_lo = 0; A->w_hi = nw;
  A->F = s->F; A->B = s->B; A->nbflag = s->nbflag; A->fs = s->fs;
  A->act_cur = s->act + (s->act_swapped ? ntiles : 0); A->act_next = s->act + (s->act_swapped ? 0 : ntiles);
  A->list = s->list; A->count = s->count;
  A->count64 = (unsigned long long *)(s->count + 4);
  A->changed = (unsigned long long *)(ctx->d_mail + 24);
  A->rec = s->rec; A->plane = s->plane; A->cellid = s->cellid;
  // plane-distance reject: a lattice step reaches (1 + eps) dr in barycentric units of the ray parameter, and the distance
  // test reaches dr_wall; 20 % on top for everything the stored normal and the float arithmetic can add
  A->rrej = (1.2f + 2.0f * p->eps) * fmaxf(p->dr_wall, p->dr);
  A->drw2 = p->dr_wall * p->dr_wall;
  A->one_eps = 1.0 + (double)p->eps;
  A->dist_dead = (p->eps >= A->drw2) ? 1 : 0;
  static const bool nofilter = getenv("CX_FILL_NOFILTER") != nullptr;
  A->nofilter = nofilter ? 1 : 0;
  {
    const float f = 1.01f * (1.0f + 3.0f * p->eps);
    A->sph2 = f * f;
  }
  A->lstep = fe_lstep(p, A->dimg);
  A->nrest = s->count + 8;
  if (fresh) {
    if (nbe - cnbe > 0) {
      k_seg_cells<<<cx_blocks(p->gridn[3], 256, ctx->num_sms * 16), 256, 0, st>>>(p->gridn[3], cell_idx_d, nbe - cnbe, s->cellid);
      CX_LAUNCH_CHECK(ctx, "k_seg_cells");
    }
    CX_CUDA(ctx, cudaMemsetAsync(A->nrest, 0, sizeof(int32_t) * 2, st));
    k_seg_prep<<<cx_blocks(nbe, 256, ctx->num_sms * 16), 256, 0, st>>>(*A, s->rec, s->plane);
    CX_LAUNCH_CHECK(ctx, "k_seg_prep");
    if (cnbe > 0) {
      k_fs_desc<<<cx_blocks(cnbe, 128), 128, 0, st>>>(*A, s->fs);
      CX_LAUNCH_CHECK(ctx, "k_fs_desc");
    }
    const int fgrid = cx_blocks(p->gridn[3], 256, ctx->num_sms * 16);
    k_cell3_flags<<<fgrid, 256, 0, st>>>(*p, 2, cell_idx_d, nullptr, s->nbflag);
    k_cell3_flags<<<fgrid, 256, 0, st>>>(*p, 1, cell_idx_d, s->nbflag, s->nbtmp);
    k_cell3_flags<<<fgrid, 256, 0, st>>>(*p, 0, cell_idx_d, s->nbtmp, s->nbflag);
    ctx->launches += 2;
    CX_LAUNCH_CHECK(ctx, "k_cell3_flags");
    if (nbe - cnbe > 0) {
      k_cell_planes<<<cx_blocks(nbe - cnbe, 256, ctx->num_sms * 16), 256, 0, st>>>(*A, s->nbflag);
      CX_LAUNCH_CHECK(ctx, "k_cell_planes");
    }
  }
  *out = s;
  return CX_OK;
}

// directed tests of the planes [A->k_lo, A->k_hi), share (res_part of res_nparts), into the B planes
static int fill_resolve(cx_ctx *ctx, FillArgs *A, FillState *s)
{
  cudaStream_t st = ctx->stream;
  if (A->k_hi <= A->k_lo) return CX_OK;
  // work list: count, (grow the buffer,) fill -- the order of the entries is irrelevant (results are OR-ed bits)
  const int64_t ncell = A->p.gridn[3];
  const int lgrid = cx_blocks(ncell, 256, ctx->num_sms * 16);
  int64_t nrest = 0;   // wall records the record-driven kernels do not take
  for (int pass = 0; pass < 2; pass++) {
    CX_CUDA(ctx, cudaMemsetAsync(A->count64, 0, sizeof(unsigned long long), st));
    k_block_list<<<lgrid, 256, 0, st>>>(A->p, A->nbflag, A->res_nparts, A->res_part, pass ? s->blist : nullptr, A->count64);
    CX_LAUNCH_CHECK(ctx, "k_block_list");
    if (pass == 0) {
      CX_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail + 30, A->count64, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
      CX_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail + 31, A->nrest, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
      CX_CUDA(ctx, cudaStreamSynchronize(st));
      A->nblist = ctx->h_mail[30];
      nrest = (int64_t)(ctx->h_mail[31] & 0xffffffffll);
      if (A->nblist > s->cap_blist) {
        cudaFree(s->blist); s->blist = nullptr; s->cap_blist = 0;
        const int64_t want = A->nblist + A->nblist / 8 + 1024;
        CX_CUDA(ctx, cudaMalloc(&s->blist, sizeof(int32_t) * (size_t)want));
        s->cap_blist = want;
      }
      A->blist = s->blist;
      if (A->nblist == 0) break;
    }
  }
  // fine dr (dead distance test): the segments enumerate the few cells they can block, the in-plane rule is resolved per 3dr-cell
  // from plane masks; what they do not take (records that are not well-conditioned, or large against the 3dr-cells) stays with
  // the cell-driven kernel.  cx_fill_set_resolve_mode(ctx, 1) / CX_FILL_RESOLVE=blocks: the cell-driven kernel for everything.
  static const bool env_blocks = getenv("CX_FILL_RESOLVE") != nullptr && !strcmp(getenv("CX_FILL_RESOLVE"), "blocks");
  const int64_t nwall = A->nbe - A->cnbe;
  const bool rec_driven = A->dist_dead && !A->nofilter && !env_blocks && ctx->fill_resolve_mode != 1 && nwall > 0;
  if (rec_driven) {
    static int cross_bpm = 0, inplane_bpm = 0;
    if (!cross_bpm) {
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cross_bpm, k_resolve_cross, 128, 0) != cudaSuccess || cross_bpm < 1) cross_bpm = 4;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&inplane_bpm, k_resolve_inplane, RB_WARPS * 32, 0) != cudaSuccess || inplane_bpm < 1)
        inplane_bpm = 2;
    }
    k_resolve_cross<<<cx_blocks(nwall, 128, (int64_t)ctx->num_sms * cross_bpm * 4), 128, 0, st>>>(*A);
    CX_LAUNCH_CHECK(ctx, "k_resolve_cross");
    if (A->nblist > 0) {
      CX_CUDA(ctx, cudaMemsetAsync(A->count64, 0, sizeof(unsigned long long), st));
      k_resolve_inplane<<<ctx->num_sms * inplane_bpm, RB_WARPS * 32, 0, st>>>(*A);
      CX_LAUNCH_CHECK(ctx, "k_resolve_inplane");
    }
    if (A->nblist > 0 && nrest > 0) {
      A->skip3 = 1;
      CX_CUDA(ctx, cudaMemsetAsync(A->count64, 0, sizeof(unsigned long long), st));
      k_resolve_blocks<<<ctx->num_sms * 2, RB_WARPS * 32, 0, st>>>(*A);
      CX_LAUNCH_CHECK(ctx, "k_resolve_blocks");
      A->skip3 = 0;
    }
  } else if (A->nblist > 0) {
    CX_CUDA(ctx, cudaMemsetAsync(A->count64, 0, sizeof(unsigned long long), st));
    k_resolve_blocks<<<ctx->num_sms * 2, RB_WARPS * 32, 0, st>>>(*A);
    CX_LAUNCH_CHECK(ctx, "k_resolve_blocks");
  }
  if (A->cnbe > 0) {
    const int64_t total = (int64_t)A->dimg[1] * (A->k_hi - A->k_lo) * A->cnbe;
    k_resolve_fs<<<cx_blocks(total, 128, ctx->num_sms * 32), 128, 0, st>>>(*A);
    CX_LAUNCH_CHECK(ctx, "k_resolve_fs");
  }
  return CX_OK;
}

// propagation rounds.  max_rounds == 0: until no tile is active (the active count is read back every 4 rounds);
// max_rounds > 0: exactly that many rounds without any host synchronisation (empty rounds are cheap); the number of tiles
// still active afterwards is left in count[0] on the device.
static int fill_rounds(cx_ctx *ctx, FillArgs *A, FillState *s, int max_rounds, int32_t *rounds_out)
{
  cudaStream_t st = ctx->stream;
  int rounds = 0;
  const int grid_tiles = ctx->num_sms * 4;
  int32_t *hcount = (int32_t *)(ctx->h_mail + 28);
  for (;;) {
    const int batch = max_rounds > 0 ? max_rounds : 4;
    for (int r = 0; r < batch; r++) {
      if (rounds == 0) {   // later rounds find their pair zeroed by the k_fill_tiles before them
        A->count = s->count; A->count_next = s->count + 2;
        k_round_prep<<<1, 1, 0, st>>>(A->count);
        CX_LAUNCH_CHECK(ctx, "k_round_prep");
      }
      k_build_list<<<cx_blocks(A->ntiles, 256, ctx->num_sms * 8), 256, 0, st>>>(*A);
      CX_LAUNCH_CHECK(ctx, "k_build_list");
      A->force_activate = (rounds == 0);
      k_fill_tiles<<<grid_tiles, FT_THREADS, 0, st>>>(*A);
      CX_LAUNCH_CHECK(ctx, "k_fill_tiles");
      std::swap(A->act_cur, A->act_next);
      s->act_swapped ^= 1;
      rounds++;
      std::swap(A->count, A->count_next);   // the pair of the round just launched is now A->count_next
    }
    if (max_rounds > 0) break;
    CX_CUDA(ctx, cudaMemcpyAsync(hcount, A->count_next, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CX_CUDA(ctx, cudaStreamSynchronize(st));
    if (hcount[0] == 0) break;  // the last round of the batch found no active tile
    if (rounds > 4000000) return cx_fail(ctx, CX_INTERNAL_ERROR, "fill: round limit");
  }
  std::swap(A->count, A->count_next);   // callers read the number of tiles the last round found active from A->count[0]
  if (rounds_out) *rounds_out = rounds;
  return CX_OK;
}

static double dbg_now()
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

extern "C" int cx_fill_geometry_slab(cx_ctx *ctx, const cx_params *p, uint32_t *fpos_d, const float *norm_d, const int32_t *ep_d,
                                     const float *pos_d, int64_t nbe, int64_t cnbe, const int32_t *cell_idx_d, int64_t seed,
                                     int64_t k_begin, int64_t k_end, int first_call, int64_t *changed_host, int32_t *rounds_host)
{
  if (!ctx || !p || !fpos_d || !norm_d || !ep_d || !pos_d || !cell_idx_d || cnbe < 0 || cnbe > nbe) return CX_WRONG_INPUT;
  FillArgs A;
  FillState *s = nullptr;
  const bool dbg = getenv("CX_DEBUG_TIMING") != nullptr;
  double t0 = 0, t1 = 0, t2 = 0, t3 = 0;
  if (dbg) { cudaStreamSynchronize(ctx->stream); t0 = dbg_now(); }
  // first_call: 1 = start of a fill (state is rebuilt), 2 = start of a fill whose directed tests were resolved by
  // cx_fill_resolve_part (+ the caller's merge of the blocked planes): the state is kept, the seed is still to be set
  const bool pre_resolved = first_call == 2;
  int rc = fill_prepare(ctx, p, fpos_d, norm_d, ep_d, pos_d, nbe, cnbe, cell_idx_d, first_call == 1, &A, &s);
  if (rc != CX_OK) return rc;
  if (k_begin < 0 || k_end > A.dimg[2] || k_begin >= k_end) return cx_fail(ctx, CX_WRONG_INPUT, "fill: bad slab");
  if (pre_resolved && !(s->resolved_lo <= k_begin && s->resolved_hi >= k_end))
    return cx_fail(ctx, CX_WRONG_INPUT, "fill: first_call = 2 without cx_fill_resolve_part");
  A.k_lo = (int)k_begin; A.k_hi = (int)k_end;
  // only the owned planes (and one halo plane either side) are touched: a rank of an N-GPU run does not pay for the lattice
  const int64_t wplane = (int64_t)A.wr * A.dimg[1];
  A.w_lo = wplane * (k_begin > 0 ? k_begin - 1 : 0);
  A.w_hi = wplane * (k_end < A.dimg[2] ? k_end + 1 : k_end);
  cudaStream_t st = ctx->stream;
  CX_CUDA(ctx, cudaMemsetAsync(A.changed, 0, sizeof(int64_t), st));
  // a previous call that stopped at the round limit left its pending tile activations in the current half of s->act, and
  // cx_fill_resolve_part (first_call = 2) left the activations of the tiles that already hold set cells
  if (first_call == 1 || (first_call == 0 && !ctx->fill_unfinished)) {
    CX_CUDA(ctx, cudaMemsetAsync(s->act, 0, (size_t)A.ntiles * 2, st));
  }
  A.merge = first_call == 1 ? 0 : 1;
  k_unpack<<<cx_blocks(A.w_hi - A.w_lo, 256, ctx->num_sms * 16), 256, 0, st>>>(A, fpos_d);
  CX_LAUNCH_CHECK(ctx, "k_unpack");
  if (first_call) {
    k_seed<<<1, 1, 0, st>>>(A, seed);
    CX_LAUNCH_CHECK(ctx, "k_seed");
  }
  if (dbg) { cudaStreamSynchronize(ctx->stream); t1 = dbg_now(); }
  // Resolve every directed test of the owned planes up front, in one balanced launch: they are pure functions of the
  // geometry, so nothing is gained by waiting for the flood to arrive.
  if (!(s->resolved_lo <= A.k_lo && s->resolved_hi >= A.k_hi)) {
    rc = fill_resolve(ctx, &A, s);
    if (rc != CX_OK) return rc;
    s->resolved_lo = A.k_lo; s->resolved_hi = A.k_hi;
  }
  if (dbg) { cudaStreamSynchronize(ctx->stream); t2 = dbg_now(); }
  int32_t rounds = 0;
  rc = fill_rounds(ctx, &A, s, ctx->fill_max_rounds, &rounds);
  if (rc != CX_OK) return rc;
  if (rounds_host) *rounds_host = rounds;
  if (dbg) { cudaStreamSynchronize(ctx->stream); t3 = dbg_now(); }
  A.w_lo = wplane * k_begin; A.w_hi = wplane * k_end;   // only owned planes are written back
  k_pack<<<cx_blocks(A.w_hi - A.w_lo, 256, ctx->num_sms * 16), 256, 0, st>>>(A, fpos_d);
  CX_LAUNCH_CHECK(ctx, "k_pack");
  CX_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail + 24, A.changed, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  CX_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail + 28, A.count, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CX_CUDA(ctx, cudaStreamSynchronize(st));
  // with a round limit: the last round still had active tiles -> what they activated is pending in the current half of s->act
  ctx->fill_unfinished = (ctx->fill_max_rounds > 0 && *(int32_t *)(ctx->h_mail + 28) != 0) ? 1 : 0;
  if (dbg)
    fprintf(stderr, "[cx fill] prepare+unpack %.3f ms, resolve %.3f ms, %d rounds %.3f ms\n", 1e3 * (t1 - t0), 1e3 * (t2 - t1), rounds,
            1e3 * (t3 - t2));
  if (changed_host) *changed_host = ctx->h_mail[24];
  return CX_OK;
}

extern "C" int cx_fill_geometry(cx_ctx *ctx, const cx_params *p, uint32_t *fpos_d, const float *norm_d, const int32_t *ep_d,
                                const float *pos_d, int64_t nbe, int64_t cnbe, const int32_t *cell_idx_d, int64_t seed, int64_t *nfi_host,
                                int32_t *rounds_host)
{
  if (!ctx || !p) return CX_WRONG_INPUT;
  int64_t dimg[3];
  cx_fill_dims(p, dimg);
  if (seed >= dimg[0] * dimg[1] * dimg[2]) return cx_fail(ctx, CX_SEED_OUTSIDE, "fill: seed outside the lattice");
  const int keep = ctx->fill_max_rounds;
  ctx->fill_max_rounds = 0;
  const int rc = cx_fill_geometry_slab(ctx, p, fpos_d, norm_d, ep_d, pos_d, nbe, cnbe, cell_idx_d, seed, 0, dimg[2], 1, nfi_host, rounds_host);
  ctx->fill_max_rounds = keep;
  return rc;
}

// ---- multi-GPU building blocks (stage level; cx_dist_fill_geometry below strings them together) ---------------------
// The directed tests are pure functions of the geometry: each rank resolves its share, the six blocked bit planes are merged
// by the caller (every lattice cell is resolved by exactly one rank, so the shares are disjoint bit sets and a SUM
// all-reduce is an OR), and the cheap propagation then runs on the whole lattice on every rank (cx_fill_propagate) or by
// z-slab (cx_fill_geometry_slab(first_call = 2)).
static int resolve_share(cx_ctx *ctx, const cx_params *p, uint32_t *fpos_d, const float *norm_d, const int32_t *ep_d, const float *pos_d,
                         int64_t nbe, int64_t cnbe, const int32_t *cell_idx_d, int64_t k_begin, int64_t k_end, int nparts, int part,
                         uint32_t **blocked_d, int64_t *blocked_words)
{
  FillArgs A;
  FillState *s = nullptr;
  int rc = fill_prepare(ctx, p, fpos_d, norm_d, ep_d, pos_d, nbe, cnbe, cell_idx_d, 1, &A, &s);
  if (rc != CX_OK) return rc;
  if (k_begin < 0 || k_end > A.dimg[2] || k_begin > k_end) return cx_fail(ctx, CX_WRONG_INPUT, "fill: bad slab");
  cudaStream_t st = ctx->stream;
  A.merge = 0;
  CX_CUDA(ctx, cudaMemsetAsync(s->act, 0, (size_t)A.ntiles * 2, st));
  k_unpack<<<cx_blocks(A.nw, 256, ctx->num_sms * 16), 256, 0, st>>>(A, fpos_d);   // tiles that hold set cells are activated
  CX_LAUNCH_CHECK(ctx, "k_unpack");
  A.k_lo = (int)k_begin; A.k_hi = (int)k_end;
  A.res_nparts = nparts; A.res_part = part;
  rc = fill_resolve(ctx, &A, s);
  if (rc != CX_OK) return rc;
  s->resolved_lo = 0; s->resolved_hi = A.dimg[2];   // after the caller's merge every plane is resolved
  if (blocked_d) *blocked_d = s->B;
  if (blocked_words) *blocked_words = 6 * A.nw;
  return CX_OK;
}

extern "C" int cx_fill_resolve_slab(cx_ctx *ctx, const cx_params *p, uint32_t *fpos_d, const float *norm_d, const int32_t *ep_d,
                                    const float *pos_d, int64_t nbe, int64_t cnbe, const int32_t *cell_idx_d, int64_t k_begin,
                                    int64_t k_end, uint32_t **blocked_d, int64_t *blocked_words)
{
  if (!ctx || !p || !fpos_d || !norm_d || !ep_d || !pos_d || !cell_idx_d || cnbe < 0 || cnbe > nbe) return CX_WRONG_INPUT;
  return resolve_share(ctx, p, fpos_d, norm_d, ep_d, pos_d, nbe, cnbe, cell_idx_d, k_begin, k_end, 1, 0, blocked_d, blocked_words);
}

// interleaved variant: the 3dr-cells are dealt to the ranks in chunks of 32 (z-slabs balance badly: the floor slab holds a
// whole wall); same contract otherwise
extern "C" int cx_fill_resolve_part(cx_ctx *ctx, const cx_params *p, uint32_t *fpos_d, const float *norm_d, const int32_t *ep_d,
                                    const float *pos_d, int64_t nbe, int64_t cnbe, const int32_t *cell_idx_d, int nparts, int part,
                                    uint32_t **blocked_d, int64_t *blocked_words)
{
  if (!ctx || !p || !fpos_d || !norm_d || !ep_d || !pos_d || !cell_idx_d || cnbe < 0 || cnbe > nbe || nparts < 1 || part < 0 || part >= nparts)
    return CX_WRONG_INPUT;
  int64_t dimg[3];
  cx_fill_dims(p, dimg);
  return resolve_share(ctx, p, fpos_d, norm_d, ep_d, pos_d, nbe, cnbe, cell_idx_d, 0, dimg[2], nparts, part, blocked_d, blocked_words);
}

extern "C" int cx_fill_propagate(cx_ctx *ctx, const cx_params *p, uint32_t *fpos_d, const float *norm_d, const int32_t *ep_d,
                                 const float *pos_d, int64_t nbe, int64_t cnbe, const int32_t *cell_idx_d, int64_t seed,
                                 int64_t *nfi_host, int32_t *rounds_host)
{
  if (!ctx || !p || !fpos_d) return CX_WRONG_INPUT;
  FillArgs A;
  FillState *s = nullptr;
  int rc = fill_prepare(ctx, p, fpos_d, norm_d, ep_d, pos_d, nbe, cnbe, cell_idx_d, 0, &A, &s);
  if (rc != CX_OK) return rc;
  if (s->resolved_hi != A.dimg[2] || s->resolved_lo != 0) return cx_fail(ctx, CX_WRONG_INPUT, "fill: cx_fill_propagate without cx_fill_resolve_slab");
  cudaStream_t st = ctx->stream;
  CX_CUDA(ctx, cudaMemsetAsync(A.changed, 0, sizeof(int64_t), st));
  CX_CUDA(ctx, cudaMemsetAsync(s->act, 0, (size_t)A.ntiles * 2, st));
  A.merge = 0;
  k_unpack<<<cx_blocks(A.nw, 256, ctx->num_sms * 16), 256, 0, st>>>(A, fpos_d);
  CX_LAUNCH_CHECK(ctx, "k_unpack");
  k_seed<<<1, 1, 0, st>>>(A, seed);
  CX_LAUNCH_CHECK(ctx, "k_seed");
  rc = fill_rounds(ctx, &A, s, 0, rounds_host);
  if (rc != CX_OK) return rc;
  k_pack<<<cx_blocks(A.nw, 256, ctx->num_sms * 16), 256, 0, st>>>(A, fpos_d);
  CX_LAUNCH_CHECK(ctx, "k_pack");
  CX_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail + 24, A.changed, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  CX_CUDA(ctx, cudaStreamSynchronize(st));
  if (nfi_host) *nfi_host = ctx->h_mail[24];
  return CX_OK;
}

// multi-GPU pipelining: bound the number of propagation rounds of one cx_fill_geometry_slab call (0 = until converged),
// and ask whether the last call stopped at that bound with active tiles left
extern "C" int cx_fill_set_max_rounds(cx_ctx *ctx, int max_rounds)
{
  if (!ctx || max_rounds < 0) return CX_WRONG_INPUT;
  ctx->fill_max_rounds = max_rounds;
  return CX_OK;
}
extern "C" int cx_fill_unfinished(cx_ctx *ctx) { return ctx ? ctx->fill_unfinished : 0; }
extern "C" int cx_fill_set_resolve_mode(cx_ctx *ctx, int mode)
{
  if (!ctx || mode < 0 || mode > 1) return CX_WRONG_INPUT;
  ctx->fill_resolve_mode = mode;
  return CX_OK;
}
extern "C" int cx_dist_set_fill_mode(cx_ctx *ctx, int mode)
{
  if (!ctx || mode < 0 || mode > 2) return CX_WRONG_INPUT;
  ctx->dist_fill_mode = mode;
  return CX_OK;
}

// ================================================================================================ multi-GPU geometry fill
// fill_fluid_complex to its fixed point over all ranks of the context's communicator (the loop of crixus.cu:934-947, sharded):
//   small lattices (blocked planes <= CX_DIST_SMALL_BYTES or fewer than 2 planes per rank): every rank resolves its share of
//      the directed tests (3dr-cell chunks dealt round-robin with a rotation), the planes are all-reduced and every rank floods
//      the whole lattice -- no halo exchange, no per-sweep collective;
//   large lattices: every rank resolves the directed tests of the z-slab it floods (the blocked bits of a cell live with the
//      cell: no merge.  The floor and lid slabs hold whole walls, but with the record-driven kernels the whole stage costs
//      milliseconds, and dealing would make every rank enumerate every segment), floods its slab for CX_DIST_FILL_ROUNDS rounds,
//      boundary planes (row-aligned words) go to the z-neighbours, and one all-reduce of two counters decides termination;
//      finally the owned planes are all-gathered (grouped broadcasts) and packed into the reference's bit order on every rank.
// k_flags: [0] = this rank still has work (active tiles or fresh halo bits)
__global__ void k_dist_flags(const int32_t *count, const int *fresh, const unsigned long long *changed, long long *out)
{
  out[0] = (count[0] != 0 || *fresh != 0) ? 1 : 0;
  out[1] = (long long)*changed;
}

extern "C" int cx_dist_fill_geometry(cx_ctx *ctx, const cx_params *p, uint32_t *fpos_d, const float *norm_d, const int32_t *ep_d,
                                     const float *pos_d, int64_t nbe, int64_t cnbe, const int32_t *cell_idx_d, int64_t seed,
                                     int64_t *nfi_host, int32_t *rounds_host)
{
  if (!ctx || !p || !fpos_d || !norm_d || !ep_d || !pos_d || !cell_idx_d || cnbe < 0 || cnbe > nbe) return CX_WRONG_INPUT;
  const int world = cx_world(ctx), rank = cx_rank(ctx);
  if (world == 1) return cx_fill_geometry(ctx, p, fpos_d, norm_d, ep_d, pos_d, nbe, cnbe, cell_idx_d, seed, nfi_host, rounds_host);
  CxColl *coll = cx_coll(ctx);
  int64_t dimg[3];
  cx_fill_dims(p, dimg);
  if (seed >= dimg[0] * dimg[1] * dimg[2]) return cx_fail(ctx, CX_SEED_OUTSIDE, "fill: seed outside the lattice");
  cudaStream_t st = ctx->stream;
  uint32_t *B = nullptr;
  int64_t bwords = 6 * (int64_t)((dimg[0] + 31) / 32) * dimg[1] * dimg[2];
  bool small = (size_t)bwords * sizeof(uint32_t) <= CX_DIST_SMALL_BYTES || dimg[2] < 2 * world;
  if (ctx->dist_fill_mode == 2 && dimg[2] >= 2 * world) small = false;   // cx_dist_set_fill_mode (tests, tuning)
  if (ctx->dist_fill_mode == 1) small = true;
  int64_t k0 = 0, k1 = dimg[2];
  int rc;
  if (small) rc = resolve_share(ctx, p, fpos_d, norm_d, ep_d, pos_d, nbe, cnbe, cell_idx_d, 0, dimg[2], world, rank, &B, &bwords);
  else {
    // the blocked bits of a cell live with the cell: a rank that resolves the planes it floods needs nothing from the others
    cx_split_range(dimg[2], world, rank, &k0, &k1);
    rc = resolve_share(ctx, p, fpos_d, norm_d, ep_d, pos_d, nbe, cnbe, cell_idx_d, k0, k1, 1, 0, &B, &bwords);
  }
  if (rc != CX_OK) return rc;
  FillArgs A;
  FillState *s = nullptr;
  rc = fill_prepare(ctx, p, fpos_d, norm_d, ep_d, pos_d, nbe, cnbe, cell_idx_d, 0, &A, &s);
  if (rc != CX_OK) return rc;
  CX_CUDA(ctx, cudaMemsetAsync(A.changed, 0, sizeof(int64_t), st));
  int32_t rounds = 0;
  if (small) {
    rc = coll->allreduce_sum_u32(ctx, B, bwords);
    if (rc != CX_OK) return rc;
    k_seed<<<1, 1, 0, st>>>(A, seed);
    CX_LAUNCH_CHECK(ctx, "k_seed");
    rc = fill_rounds(ctx, &A, s, 0, &rounds);
    if (rc != CX_OK) return rc;
    k_pack<<<cx_blocks(A.nw, 256, ctx->num_sms * 16), 256, 0, st>>>(A, fpos_d);
    CX_LAUNCH_CHECK(ctx, "k_pack");
    CX_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail + 24, A.changed, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CX_CUDA(ctx, cudaStreamSynchronize(st));
    if (nfi_host) *nfi_host = ctx->h_mail[24];
    if (rounds_host) *rounds_host = rounds;
    return CX_OK;
  }
  // ---- z-slabs
  const int64_t wplane = (int64_t)A.wr * A.dimg[1];
  std::vector<CxSeg> segs;
  A.k_lo = (int)k0; A.k_hi = (int)k1;
  if (wplane > s->cap_halo) {
    cudaFree(s->halo); s->halo = nullptr; s->cap_halo = 0;
    CX_CUDA(ctx, cudaMalloc(&s->halo, sizeof(uint32_t) * (size_t)wplane * 2));
    s->cap_halo = wplane;
  }
  k_seed<<<1, 1, 0, st>>>(A, seed);
  CX_LAUNCH_CHECK(ctx, "k_seed");
  int *fresh = (int *)(s->count + 6);
  long long *flags = (long long *)(ctx->d_mail + 36);
  for (int it = 0;; it++) {
    int32_t r = 0;
    rc = fill_rounds(ctx, &A, s, CX_DIST_FILL_ROUNDS, &r);   // force_activate in its first round: tiles holding fresh halo bits wake the slab
    if (rc != CX_OK) return rc;
    rounds += r;
    // boundary planes to the z-neighbours: my top plane goes up, my bottom plane goes down
    rc = coll->halo_exchange_u32(ctx, A.F + wplane * (k1 - 1), s->halo, A.F + wplane * k0, s->halo + wplane, wplane);
    if (rc != CX_OK) return rc;
    CX_CUDA(ctx, cudaMemsetAsync(fresh, 0, sizeof(int), st));
    if (rank + 1 < world) {
      k_merge_plane<<<cx_blocks(wplane, 256, ctx->num_sms * 8), 256, 0, st>>>(A, s->halo, (int)k1, fresh);
      CX_LAUNCH_CHECK(ctx, "k_merge_plane");
    }
    if (rank > 0) {
      k_merge_plane<<<cx_blocks(wplane, 256, ctx->num_sms * 8), 256, 0, st>>>(A, s->halo + wplane, (int)k0 - 1, fresh);
      CX_LAUNCH_CHECK(ctx, "k_merge_plane");
    }
    k_dist_flags<<<1, 1, 0, st>>>(A.count, fresh, A.changed, flags);
    CX_LAUNCH_CHECK(ctx, "k_dist_flags");
    rc = coll->allreduce_sum_i64(ctx, (int64_t *)flags, 2);
    if (rc != CX_OK) return rc;
    CX_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail + 36, flags, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CX_CUDA(ctx, cudaStreamSynchronize(st));
    if (ctx->h_mail[36] == 0) break;
    if (it > 1000000) return cx_fail(ctx, CX_INTERNAL_ERROR, "fill: exchange limit");
  }
  if (nfi_host) *nfi_host = ctx->h_mail[37];
  if (rounds_host) *rounds_host = rounds;
  // ---- every rank gets every slab, then packs the whole lattice into the reference's bit order
  segs.clear();
  for (int r = 0; r < world; r++) {
    int64_t a, b;
    cx_split_range(dimg[2], world, r, &a, &b);
    segs.push_back({wplane * a, wplane * (b - a), r});
  }
  rc = coll->bcast_segments_u32(ctx, A.F, segs.data(), (int)segs.size());
  if (rc != CX_OK) return rc;
  A.k_lo = 0; A.k_hi = A.dimg[2];
  k_pack<<<cx_blocks(A.nw, 256, ctx->num_sms * 16), 256, 0, st>>>(A, fpos_d);
  CX_LAUNCH_CHECK(ctx, "k_pack");
  return CX_OK;
}

// crixus_b200 -- fill_fluid (box) and fill_fluid_complex (geometry) on the dr lattice.
// Reference: /root/reference/src/crixus_d.cu:680-744 (box), :753-963 (checkTriangleCollision, checkCollision),
// :965-1058 (fill_fluid_complex, one full-grid sweep per launch, looped up to 10 000 times by crixus.cu:934-947).
//
// The fill result is the least fixed point of "s unset, a 6-neighbour e set, !checkCollision(s,e) => set s"
// (SURVEY.md A.6), so any schedule reaches the same bits.  B200 mapping (DESIGN.md "fill"):
//   * the packed reference bitfield is unpacked to a row-aligned layout (x-rows padded to 32-bit words);
//   * a classification pass marks the cells whose test can be non-trivial ("hard": a non-empty 27-cell
//     neighbourhood in the 3dr grid, or close to a free-surface triangle).  For every other cell
//     checkCollision is provably false, so propagation is pure bit arithmetic on 32-cell words;
//   * the lattice is cut into tiles of 1 word x 16 x 16 rows.  The first time the flood reaches a tile, ALL hard cells
//     of the tile are resolved in one bulk, perfectly parallel pass (k_resolve_tiles): one warp per cell, lanes
//     stride the cell-sorted segments of the 27 neighbour cells, the 6 directions of a cell share the distance
//     computation; the results are 6 "blocked" bit planes.  The directed tests are pure functions of (s,e), so
//     evaluating them before they are needed changes nothing but the schedule;
//   * propagation is then bit arithmetic for every cell: tiles are iterated to local convergence in shared memory
//     (k_fill_tiles); a tile that changed activates its 6 face neighbours for the next round; rounds continue
//     until no tile is active.  Only reached tiles are touched: work ~ filled volume + wall area.
// Bound: HBM, 0.25 B/cell + 8 B/filled cell -- in practice latency/round bound.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <algorithm>
#include "common.cuh"

#define FT_Y 16
#define FT_Z 16
#define FT_THREADS (FT_Y * FT_Z)
#define FS_MAX_DESC 512

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
// conservative description of a free-surface (fshape) triangle, used only to decide which cells are "hard"
struct FsDesc {
  float lo[3], hi[3];  // bounding box grown by the margin outside of which checkTriangleCollision cannot be true
  float n[3], v0[3];   // plane, for the "ray in plane => collision" rule (crixus_d.cu:766-769)
  int quirk;           // some axis direction can be parallel to the plane within eps
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
  uint32_t *H;       // hard mask, same shape
  uint32_t *B;       // 6 blocked planes, same shape each (direction d at B + d*nw)
  int64_t nw;        // words per plane set
  const uint8_t *nbflag;  // per 3dr-cell: 27-neighbourhood contains a segment
  const FsDesc *fs;       // fshape descriptors
  int nfs, all_hard;
  const int *all_hard_d;       // set by k_fs_desc when a free-surface triangle cannot be culled
  uint8_t *act_cur, *act_next;  // tile activation flags
  int32_t *list;                // active tile list
  int32_t *count;               // [0] = list length, [1] = cursor, [2] = resolve-list length, [3] = resolve cursor
  unsigned long long *count64;  // work counter of the up-front resolve (one item per 32-cell word of the lattice)
  uint8_t *tile_res;            // tile already resolved (blocked planes valid)
  int32_t *rlist;               // tiles to resolve this round
  unsigned long long *changed;  // newly set cells
  int res_nparts, res_part;     // up-front resolve of items part, part+nparts, ... in chunks of 64 (0/1: everything)
  int resolved_all;             // every hard cell of the owned planes was resolved up front (no per-round resolve)
  int force_activate;           // first round of a call: processed tiles wake their neighbours unconditionally
  int merge;                    // k_unpack: OR into the existing row-aligned state (later slab calls)
  int64_t w_lo, w_hi;           // word range k_unpack / k_pack walk (slab calls: the owned planes, +-1 halo plane for k_unpack)
  long long item_lo, item_hi;   // item range of the up-front resolve (slab calls: the tile layers of the owned planes)
  int classify_inline;          // up-front resolve computes the hard mask itself (no k_classify pass over the lattice)
  const float4 *rec;            // per-segment invariants of checkCollision (k_seg_prep), 8 x float4 per segment
  float rrej;                   // plane distance beyond which a well-formed triangle cannot matter
  float drw2;                   // dr_wall*dr_wall
  double one_eps;               // 1.0 + eps (double, crixus_d.cu:774)
};

// ---- the reference's directed test, warp-cooperative -------------------------------------------------------------
// Everything in checkCollision that depends on the triangle alone is hoisted into a per-segment record (k_seg_prep):
// the same operations in the same order, so the results are bit-identical to evaluating them per (cell, segment) pair.
//   R0 = (n, wf)   R1 = (v0, uu)   R2 = (v1, uv)   R3 = (v2, vv)   R4 = (segpos, D)   R5..R7 = (nr[0..2], -)
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
// vv, D), plus the "well-formed" flag that licenses the plane-distance reject in warp_check_collision.
__global__ void __launch_bounds__(256) k_seg_prep(FillArgs A, float4 *__restrict__ rec)
{
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
    // well-formed: |n| = 1 within 1 %, n perpendicular to both edges within 2e-3 rad (evaluated in double).  Then for
    // every point within the influence radius, |n.(s - v0)| is the distance to the triangle's plane within a few
    // per cent, and no distance the reference can compute (plane, vertex or edge-line distance) is smaller.
    const double nn = (double)n.x * n.x + (double)n.y * n.y + (double)n.z * n.z;
    const double lu = sqrt((double)u.x * u.x + (double)u.y * u.y + (double)u.z * u.z);
    const double lv = sqrt((double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z);
    const double nu = (double)n.x * u.x + (double)n.y * u.y + (double)n.z * u.z;
    const double nv = (double)n.x * v.x + (double)n.y * v.y + (double)n.z * v.z;
    const bool wf = nn > 0.98 && nn < 1.02 && lu > 0 && lv > 0 && fabs(nu) <= 2e-3 * lu && fabs(nv) <= 2e-3 * lv;
    float4 *r = rec + i * FREC;
    r[0] = make_float4(n.x, n.y, n.z, wf ? 1.f : 0.f);
    r[1] = make_float4(v0.x, v0.y, v0.z, uu);
    r[2] = make_float4(v1.x, v1.y, v1.z, uv);
    r[3] = make_float4(v2.x, v2.y, v2.z, vv);
    r[4] = make_float4(segpos.x, segpos.y, segpos.z, D);
    r[5] = make_float4(nr[0].x, nr[0].y, nr[0].z, 0.f);
    r[6] = make_float4(nr[1].x, nr[1].y, nr[1].z, 0.f);
    r[7] = make_float4(nr[2].x, nr[2].y, nr[2].z, 0.f);
  }
}

// checkCollision (crixus_d.cu:801-963) for cell s=(si,sj,sk) and every direction in dmask at once (bit d: 0:-x 1:+x
// 2:-y 3:+y 4:-z 5:+z, the reference's neighbour order).  Called by a full warp; returns the directions that collide.
//
// Plane-distance reject (exact): for a well-formed triangle, a cell further than rrej = 1.2 max(dr_wall, dr) from the
// triangle's plane can neither pass the distance test (every distance the reference computes is >= the plane
// distance up to a few per cent) nor be hit by a lattice step of length dr, and the "ray in plane" rule needs
// |n.(s-v0)| < eps.  Such pairs are skipped after two loads; everything else runs the reference's arithmetic.
__device__ uint32_t warp_check_collision(const FillArgs &A, int si, int sj, int sk, uint32_t dmask, int lane)
{
  const cx_params &p = A.p;
  const float dr = p.dr, eps = p.eps;
  const int sidx[3] = {si, sj, sk};
  float sc[3], ec[6], dirc[6], infl[6], midc[6], mids[3];
#pragma unroll
  for (int a = 0; a < 3; a++) {
    sc[a] = ffma((float)sidx[a], dr, p.fcmin[a]);  // crixus_d.cu:805-807 (R5)
    mids[a] = fadd(sc[a], sc[a]);
  }
#pragma unroll
  for (int d = 0; d < 6; d++) {
    const int a = d >> 1;
    ec[d] = ffma((float)(sidx[a] + ((d & 1) ? 1 : -1)), dr, p.fcmin[a]);
    dirc[d] = fsub(ec[d], sc[a]);
    const float hd = fmul(fsub(sc[a], ec[d]), 0.5f);  // (s-e)/2.0f; the two other components are exactly +0
    const f3 hv = mk3(a == 0 ? hd : 0.f, a == 1 ? hd : 0.f, a == 2 ? hd : 0.f);
    float l = dot3_r3(hv, hv);
    l = __fsqrt_rn(l);
    float r = fadd(ffma(dr, 3.0f, l), p.dr_wall);  // crixus_d.cu:815 (R6)
    infl[d] = fmul(r, r);
    midc[d] = fadd(sc[a], ec[d]);
  }
  const f3 s = mk3(sc[0], sc[1], sc[2]);
  const int gx = cell_coord(sc[0], p.dmin[0], p.griddr[0], p.gridn[0]);
  const int gy = cell_coord(sc[1], p.dmin[1], p.griddr[1], p.gridn[1]);
  const int gz = cell_coord(sc[2], p.dmin[2], p.griddr[2], p.gridn[2]);
  uint32_t coll = 0;
  // the 27 cells as 9 (x,y) columns; the z-neighbours of a column are adjacent in memory, so one column is one run
  // of segments (plus, at a periodic z border, the wrapped cell as a run of its own)
  for (int c9 = 0; c9 < 9 && coll != dmask; c9++) {
    const int cx = gx + c9 / 3 - 1, cy = gy + c9 % 3 - 1;
    for (int part = 0; part < 3 && coll != dmask; part++) {
      int b, e;
      if (part == 0) {
        const int z0 = max(gz - 1, 0), z1 = min(gz + 1, p.gridn[2] - 1);
        const int64_t c0 = nb_cell(p, cx, cy, z0);
        if (c0 < 0) break;   // column outside the grid (not periodic): nothing in the other parts either
        b = __ldg(&A.cell_idx[c0]); e = __ldg(&A.cell_idx[c0 + (z1 - z0) + 1]);
      } else {
        const int zz = part == 1 ? gz - 1 : gz + 1;
        if (zz >= 0 && zz < p.gridn[2]) continue;   // in range: part of the run above
        const int64_t c0 = nb_cell(p, cx, cy, zz);  // wrapped (periodic z) or -1
        if (c0 < 0) continue;
        b = __ldg(&A.cell_idx[c0]); e = __ldg(&A.cell_idx[c0 + 1]);
      }
      for (int base = b; base < e && coll != dmask; base += 32) {
        const int i = base + lane;
        uint32_t mine = 0;
        if (i < e) {
          const float4 *r = A.rec + (int64_t)i * FREC;
          const float4 R0 = __ldg(r), R1 = __ldg(r + 1);
          const f3 n = mk3(R0), v0 = mk3(R1);
          const float t = (s.x - v0.x) * n.x + (s.y - v0.y) * n.y + (s.z - v0.z) * n.z;
          if (!(R0.w != 0.f && fabsf(t) > A.rrej)) {
            uint32_t pass = 0;
#pragma unroll
            for (int d = 0; d < 6; d++) {
              if (!((dmask & ~coll) >> d & 1)) continue;
              const int a = d >> 1;
              const f3 tt = mk3(ffma(a == 0 ? midc[d] : mids[0], 0.5f, -v0.x), ffma(a == 1 ? midc[d] : mids[1], 0.5f, -v0.y),
                                ffma(a == 2 ? midc[d] : mids[2], 0.5f, -v0.z));
              if (!(dot3_r3(tt, tt) > infl[d])) pass |= 1u << d;  // crixus_d.cu:855-857
            }
            if (pass) {
              const float4 R2 = __ldg(r + 2), R3 = __ldg(r + 3), R4 = __ldg(r + 4), R5 = __ldg(r + 5), R6 = __ldg(r + 6), R7 = __ldg(r + 7);
              const f3 v1 = mk3(R2), v2 = mk3(R3);
              const float dist = tri_dist2(A, s, n, v0, v1, v2, mk3(R4), mk3(R5), mk3(R6), mk3(R7));
              if (fadd(dist, eps) < A.drw2) mine = pass;  // crixus_d.cu:940
              else {
                const f3 u = sub3(v1, v0), v = sub3(v2, v0);
#pragma unroll
                for (int d = 0; d < 6; d++) {
                  if (!(pass >> d & 1)) continue;
                  const int a = d >> 1;
                  const f3 dir = mk3(a == 0 ? dirc[d] : 0.f, a == 1 ? dirc[d] : 0.f, a == 2 ? dirc[d] : 0.f);
                  if (tri_collision(A, s, dir, n, v0, u, v, R1.w, R2.w, R3.w, R4.w)) mine |= 1u << d;
                }
              }
            }
          }
        }
        coll |= __reduce_or_sync(0xffffffffu, mine);
      }
    }
  }
  // free-surface triangles: all of them, segment test only (crixus_d.cu:951-960)
  for (int64_t base = A.nbe - A.cnbe; base < A.nbe && coll != dmask; base += 32) {
    const int64_t i = base + lane;
    uint32_t mine = 0;
    if (i < A.nbe) {
      const float4 *r = A.rec + i * FREC;
      const float4 R0 = __ldg(r), R1 = __ldg(r + 1), R2 = __ldg(r + 2), R3 = __ldg(r + 3), R4 = __ldg(r + 4);
      const f3 n = mk3(R0), v0 = mk3(R1), u = sub3(mk3(R2), v0), v = sub3(mk3(R3), v0);
#pragma unroll
      for (int d = 0; d < 6; d++) {
        if (!((dmask & ~coll) >> d & 1)) continue;
        const int a = d >> 1;
        const f3 dir = mk3(a == 0 ? dirc[d] : 0.f, a == 1 ? dirc[d] : 0.f, a == 2 ? dirc[d] : 0.f);
        if (tri_collision(A, s, dir, n, v0, u, v, R1.w, R2.w, R3.w, R4.w)) mine |= 1u << d;
      }
    }
    coll |= __reduce_or_sync(0xffffffffu, mine);
  }
  return coll & dmask;
}

// ---- preparation kernels ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_cell3_flags(cx_params p, const int32_t *__restrict__ cell_idx, uint8_t *__restrict__ flag)
{
  const int64_t n = p.gridn[3];
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < n; c += (int64_t)gridDim.x * blockDim.x) {
    const int gz = (int)(c % p.gridn[2]), gy = (int)((c / p.gridn[2]) % p.gridn[1]), gx = (int)(c / ((int64_t)p.gridn[2] * p.gridn[1]));
    int any = 0;
    for (int c27 = 0; c27 < 27; c27++) {
      const int64_t q = nb_cell(p, gx + c27 / 9 - 1, gy + (c27 / 3) % 3 - 1, gz + c27 % 3 - 1);
      if (q >= 0 && cell_idx[q + 1] > cell_idx[q]) { any = 1; break; }
    }
    flag[c] = (uint8_t)any;
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

// hard mask: cell needs the real test (otherwise checkCollision(s, .) is false for every direction)
__global__ void __launch_bounds__(256) k_classify(FillArgs A)
{
  const cx_params &p = A.p;
  const int64_t rows = (int64_t)A.dimg[1] * A.dimg[2];
  for (int64_t wi = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; wi < A.nw; wi += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = wi / A.wr;
    const int wx = (int)(wi - row * A.wr);
    if (row >= rows) continue;
    const int j = (int)(row % A.dimg[1]), k = (int)(row / A.dimg[1]);
    const int nvalid = min(32, A.dimg[0] - wx * 32);
    uint32_t h = 0;
    if (A.all_hard || *A.all_hard_d) h = 0xffffffffu;
    else {
      const float sy = ffma((float)j, p.dr, p.fcmin[1]), sz = ffma((float)k, p.dr, p.fcmin[2]);
      const int gy = cell_coord(sy, p.dmin[1], p.griddr[1], p.gridn[1]), gz = cell_coord(sz, p.dmin[2], p.griddr[2], p.gridn[2]);
      for (int b = 0; b < nvalid; b++) {
        const float sx = ffma((float)(wx * 32 + b), p.dr, p.fcmin[0]);
        const int gx = cell_coord(sx, p.dmin[0], p.griddr[0], p.gridn[0]);
        int hard = A.nbflag[((int64_t)gx * p.gridn[1] + gy) * p.gridn[2] + gz];
        for (int t = 0; t < A.nfs && !hard; t++) {
          const FsDesc &f = A.fs[t];
          if (sx >= f.lo[0] && sx <= f.hi[0] && sy >= f.lo[1] && sy <= f.hi[1] && sz >= f.lo[2] && sz <= f.hi[2]) hard = 1;
          else if (f.quirk) {
            const float a = f.n[0] * (sx - f.v0[0]) + f.n[1] * (sy - f.v0[1]) + f.n[2] * (sz - f.v0[2]);
            if (fabsf(a) < 4.f * p.eps) hard = 1;
          }
        }
        h |= (uint32_t)hard << b;
      }
    }
    if (nvalid < 32) h &= (1u << nvalid) - 1u;
    A.H[wi] = h;
  }
}

// the same hard mask for one 32-cell word, computed by a warp (lane = cell): used by the up-front resolve so that no
// separate pass over the whole lattice is needed (at 1e10 cells k_classify alone is ~0.1 s, and it does not shard)
__device__ __forceinline__ uint32_t classify_word_warp(const FillArgs &A, int wx, int j, int k, int lane)
{
  const cx_params &p = A.p;
  const int nvalid = min(32, A.dimg[0] - wx * 32);
  const uint32_t vmask = nvalid < 32 ? (1u << nvalid) - 1u : 0xffffffffu;
  if (A.all_hard || *A.all_hard_d) return vmask;
  int hard = 0;
  if (lane < nvalid) {
    const float sy = ffma((float)j, p.dr, p.fcmin[1]), sz = ffma((float)k, p.dr, p.fcmin[2]);
    const int gy = cell_coord(sy, p.dmin[1], p.griddr[1], p.gridn[1]), gz = cell_coord(sz, p.dmin[2], p.griddr[2], p.gridn[2]);
    const float sx = ffma((float)(wx * 32 + lane), p.dr, p.fcmin[0]);
    const int gx = cell_coord(sx, p.dmin[0], p.griddr[0], p.gridn[0]);
    hard = A.nbflag[((int64_t)gx * p.gridn[1] + gy) * p.gridn[2] + gz];
    for (int t = 0; t < A.nfs && !hard; t++) {
      const FsDesc &f = A.fs[t];
      if (sx >= f.lo[0] && sx <= f.hi[0] && sy >= f.lo[1] && sy <= f.hi[1] && sz >= f.lo[2] && sz <= f.hi[2]) hard = 1;
      else if (f.quirk) {
        const float a = f.n[0] * (sx - f.v0[0]) + f.n[1] * (sy - f.v0[1]) + f.n[2] * (sz - f.v0[2]);
        if (fabsf(a) < 4.f * p.eps) hard = 1;
      }
    }
  }
  return __ballot_sync(0xffffffffu, hard) & vmask;
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

// ---- one round -------------------------------------------------------------------------------------------------
__global__ void k_round_prep(int32_t *count) { count[0] = 0; count[1] = 0; count[2] = 0; count[3] = 0; }

__global__ void __launch_bounds__(256) k_build_list(FillArgs A)
{
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < A.ntiles; t += (int64_t)gridDim.x * blockDim.x)
    if (A.act_cur[t]) {
      A.act_cur[t] = 0;
      A.list[atomicAdd(&A.count[0], 1)] = (int32_t)t;
      if (!A.tile_res[t]) {   // first visit: its hard cells are resolved before k_fill_tiles runs
        A.tile_res[t] = 1;
        A.rlist[atomicAdd(&A.count[2], 1)] = (int32_t)t;
      }
    }
}

// bulk resolution of the hard cells of newly reached tiles: one warp per 32-cell word, one warp-cooperative
// checkCollision per hard cell for all of its in-lattice directions at once
template <bool ALL>
__global__ void __launch_bounds__(256, 3) k_resolve_tiles(FillArgs A)
{
  const int lane = threadIdx.x & 31;
  for (;;) {
    long long item = 0;
    if (lane == 0) item = ALL ? (long long)atomicAdd((unsigned long long *)A.count64, 1ull) : (long long)atomicAdd(&A.count[3], 1);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (ALL && A.res_nparts > 1) {   // interleaved sharding: every group of nparts 64-item chunks is dealt one chunk per rank, the
      const long long g = item >> 6;   // order rotated by a hash of the group (plane-aligned walls would otherwise land on few ranks)
      const int rot = (int)((((unsigned)g * 2654435761u) >> 16) % (unsigned)A.res_nparts);
      item = (g * A.res_nparts + (A.res_part + rot) % A.res_nparts) * 64 + (item & 63);
    }
    if (ALL) item += A.item_lo;
    if (item >= (ALL ? A.item_hi : (long long)A.count[2] * FT_THREADS)) break;
    const int tile = ALL ? (int)(item / FT_THREADS) : A.rlist[item / FT_THREADS], wt = (int)(item % FT_THREADS);
    const int wx = tile % A.wr, tyi = (tile / A.wr) % A.ty, tzi = tile / (A.wr * A.ty);
    const int j = tyi * FT_Y + wt % FT_Y, k = tzi * FT_Z + wt / FT_Y;
    if (j >= A.dimg[1] || k >= A.dimg[2] || k < A.k_lo || k >= A.k_hi) continue;
    const int64_t wi = ((int64_t)k * A.dimg[1] + j) * A.wr + wx;
    uint32_t hw;
    if (ALL && A.classify_inline) {
      hw = classify_word_warp(A, wx, j, k, lane);
      if (lane == 0) A.H[wi] = hw;
    } else hw = A.H[wi];
    uint32_t todo = hw & ~A.F[wi];
    if (!todo) continue;
    uint32_t b0 = 0, b1 = 0, b2 = 0, b3 = 0, b4 = 0, b5 = 0;
    while (todo) {
      const int b = __ffs(todo) - 1;
      todo &= todo - 1;
      const int i = wx * 32 + b;
      const uint32_t dm = (i > 0 ? 1u : 0u) | (i + 1 < A.dimg[0] ? 2u : 0u) | (j > 0 ? 4u : 0u) | (j + 1 < A.dimg[1] ? 8u : 0u) |
                          (k > 0 ? 16u : 0u) | (k + 1 < A.dimg[2] ? 32u : 0u);
      const uint32_t coll = warp_check_collision(A, i, j, k, dm, lane);
      const uint32_t bit = 1u << b;
      if (coll & 1u) b0 |= bit;
      if (coll & 2u) b1 |= bit;
      if (coll & 4u) b2 |= bit;
      if (coll & 8u) b3 |= bit;
      if (coll & 16u) b4 |= bit;
      if (coll & 32u) b5 |= bit;
    }
    if (lane == 0) {
      A.B[0 * A.nw + wi] = b0; A.B[1 * A.nw + wi] = b1; A.B[2 * A.nw + wi] = b2;
      A.B[3 * A.nw + wi] = b3; A.B[4 * A.nw + wi] = b4; A.B[5 * A.nw + wi] = b5;
    }
  }
}

struct TileSM {
  uint32_t F[FT_Z + 2][FT_Y + 2];  // filled, with y/z halo
  int changed, newbits;
};

// propagation inside active tiles: pure bit arithmetic (all directed tests of the tile are already in the B planes)
__global__ void __launch_bounds__(FT_THREADS) k_fill_tiles(FillArgs A)
{
  __shared__ TileSM S;
  __shared__ int s_t;
  const int tid = threadIdx.x;
  const int ly = tid % FT_Y, lz = tid / FT_Y;
  for (;;) {
    if (tid == 0) {
      const int q = atomicAdd(&A.count[1], 1);
      s_t = q < A.count[0] ? A.list[q] : -1;
      S.newbits = 0;
    }
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
    uint32_t xl = 0, xr = 0, hmask = 0;
    // per direction: cells that may be entered from that neighbour (easy cells always, hard cells unless blocked)
    uint32_t o0 = vmask, o1 = vmask, o2 = vmask, o3 = vmask, o4 = vmask, o5 = vmask;
    if (inrow) {
      if (wx > 0) xl = A.F[wi - 1] >> 31;
      if (wx + 1 < A.wr) xr = A.F[wi + 1] << 31;
      hmask = A.H[wi];
      if (hmask) {
        o0 &= ~(hmask & A.B[0 * A.nw + wi]); o1 &= ~(hmask & A.B[1 * A.nw + wi]); o2 &= ~(hmask & A.B[2 * A.nw + wi]);
        o3 &= ~(hmask & A.B[3 * A.nw + wi]); o4 &= ~(hmask & A.B[4 * A.nw + wi]); o5 &= ~(hmask & A.B[5 * A.nw + wi]);
      }
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
      A.F[wi] = f_end;
      atomicAdd(&S.newbits, __popc(f_end & ~f_start));
    }
    __syncthreads();
    if (tid == 0 && S.newbits) atomicAdd(A.changed, (unsigned long long)S.newbits);
    if (tid == 0 && (S.newbits || A.force_activate)) {
      if (wx > 0) A.act_next[tile - 1] = 1;
      if (wx + 1 < A.wr) A.act_next[tile + 1] = 1;
      if (tyi > 0) A.act_next[tile - A.wr] = 1;
      if (tyi + 1 < A.ty) A.act_next[tile + A.wr] = 1;
      if (tzi > 0) A.act_next[tile - A.wr * A.ty] = 1;
      if (tzi + 1 < A.tz) A.act_next[tile + A.wr * A.ty] = 1;
    }
    __syncthreads();
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------
struct FillState {
  const uint32_t *fpos_key = nullptr;
  int64_t dimg[3] = {0, 0, 0};
  int64_t nbe = 0, cnbe = 0;
  uint32_t *F = nullptr, *B = nullptr;   // B: 6 blocked planes + the hard mask
  float4 *rec = nullptr;
  int64_t cap_nw = 0, cap_ntiles = 0, cap_ncell = 0, cap_rec = 0;
  uint8_t *nbflag = nullptr, *act = nullptr, *tile_res = nullptr;
  int32_t *list = nullptr, *count = nullptr, *rlist = nullptr;
  FsDesc *fs = nullptr;
  int nfs = 0, all_hard = 0;
  int64_t nw = 0, ntiles = 0;
  int resolved_lo = 0, resolved_hi = 0;   // planes whose hard cells were resolved up front
};

void cx_fill_state_free(cx_ctx *ctx)
{
  FillState *s = (FillState *)ctx->fill_state;
  if (!s) return;
  cudaFree(s->F); cudaFree(s->B); cudaFree(s->nbflag); cudaFree(s->act); cudaFree(s->list);
  cudaFree(s->count); cudaFree(s->fs); cudaFree(s->tile_res); cudaFree(s->rlist); cudaFree(s->rec);
  delete s;
  ctx->fill_state = nullptr;
}

// conservative fshape descriptors, one thread per free-surface triangle (DESIGN.md "fill / easy cells").
// A triangle is "cullable" when its stored normal is consistent with its plane and it is well conditioned; then
// checkTriangleCollision(s, dir, .) can only be true for s inside the triangle's bounding box grown by
// 1.5 dr + 1 % of its edge lengths + 10 eps, or (ray-in-plane rule, crixus_d.cu:766-769) within 4 eps of its plane
// when some axis direction is parallel to the plane within eps.  One non-cullable triangle makes every cell hard.
__global__ void k_fs_desc(FillArgs A, FsDesc *out, int *all_hard)
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
  const bool cullable = ln > 0 && lu > 0 && lv > 0 && fabs(nu) <= 1e-3 * ln * lu && fabs(nv) <= 1e-3 * ln * lv &&
                        (lu * lu * lv * lv - uv * uv) >= 1e-4 * lu * lu * lv * lv;
  if (!cullable) atomicExch(all_hard, 1);
  const double margin = 1.5 * A.p.dr + 0.01 * (lu + lv) + 10.0 * A.p.eps;
  FsDesc d;
  d.quirk = 0;
  for (int k = 0; k < 3; k++) {
    const float mn = fminf(a[k], fminf(b[k], c[k])), mx = fmaxf(a[k], fmaxf(b[k], c[k]));
    d.lo[k] = (float)(mn - margin); d.hi[k] = (float)(mx + margin);
    d.n[k] = n[k]; d.v0[k] = a[k];
    if (fabs((double)n[k]) * A.p.dr * 0.5 < A.p.eps) d.quirk = 1;
  }
  out[t] = d;
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
      // the hard mask lives right behind the six blocked planes: one buffer of 7 planes, so that a multi-GPU caller merges
      // mask and planes with a single all-reduce (cx_fill_resolve_part)
      cudaFree(s->F); cudaFree(s->B);
      s->F = s->B = nullptr; s->cap_nw = 0;
      CX_CUDA(ctx, cudaMalloc(&s->F, sizeof(uint32_t) * (size_t)nw));
      CX_CUDA(ctx, cudaMalloc(&s->B, sizeof(uint32_t) * (size_t)nw * 7));
      s->cap_nw = nw;
    }
    if (ntiles > s->cap_ntiles) {
      cudaFree(s->act); cudaFree(s->list); cudaFree(s->tile_res); cudaFree(s->rlist);
      s->act = s->tile_res = nullptr; s->list = s->rlist = nullptr; s->cap_ntiles = 0;
      CX_CUDA(ctx, cudaMalloc(&s->act, (size_t)ntiles * 2));
      CX_CUDA(ctx, cudaMalloc(&s->list, sizeof(int32_t) * (size_t)ntiles));
      CX_CUDA(ctx, cudaMalloc(&s->tile_res, (size_t)ntiles));
      CX_CUDA(ctx, cudaMalloc(&s->rlist, sizeof(int32_t) * (size_t)ntiles));
      s->cap_ntiles = ntiles;
    }
    if (p->gridn[3] > s->cap_ncell) {
      cudaFree(s->nbflag); s->nbflag = nullptr; s->cap_ncell = 0;
      CX_CUDA(ctx, cudaMalloc(&s->nbflag, (size_t)p->gridn[3]));
      s->cap_ncell = p->gridn[3];
    }
    if (nbe > s->cap_rec) {
      cudaFree(s->rec); s->rec = nullptr; s->cap_rec = 0;
      CX_CUDA(ctx, cudaMalloc(&s->rec, sizeof(float4) * FREC * (size_t)nbe));
      s->cap_rec = nbe;
    }
    if (!s->count) CX_CUDA(ctx, cudaMalloc(&s->count, sizeof(int32_t) * 8));
    if (!s->fs) CX_CUDA(ctx, cudaMalloc(&s->fs, sizeof(FsDesc) * FS_MAX_DESC + 16));
    s->fpos_key = fpos_d; s->nbe = nbe; s->cnbe = cnbe; s->nw = nw; s->ntiles = ntiles;
    for (int k = 0; k < 3; k++) s->dimg[k] = dimg[k];
    CX_CUDA(ctx, cudaMemsetAsync(s->tile_res, 0, (size_t)ntiles, st));
    CX_CUDA(ctx, cudaMemsetAsync(s->B, 0, sizeof(uint32_t) * (size_t)nw * 7, st));
    CX_CUDA(ctx, cudaMemsetAsync(s->act, 0, (size_t)ntiles * 2, st));
    // fshape descriptors are built on the device (k_fs_desc): no host round trip
    s->nfs = (cnbe > 0 && cnbe <= FS_MAX_DESC) ? (int)cnbe : 0;
    s->all_hard = (cnbe > FS_MAX_DESC) ? 1 : 0;
  }
  memset(A, 0, sizeof(*A));
  A->p = *p;
  A->pos = (const float4 *)pos_d; A->norm = (const float4 *)norm_d; A->ep = (const int4 *)ep_d; A->cell_idx = cell_idx_d;
  A->nbe = nbe; A->cnbe = cnbe;
  for (int k = 0; k < 3; k++) A->dimg[k] = (int)dimg[k];
  A->wr = wr; A->ty = ty; A->tz = tz; A->ntiles = ntiles; A->nw = nw;
  A->k_lo = 0; A->k_hi = (int)dimg[2];
  A->w_lo = 0; A->w_hi = nw;
  A->item_lo = 0; A->item_hi = (long long)ntiles * FT_THREADS;
  A->classify_inline = getenv("CX_FILL_LAZY") ? 0 : 1;
  A->F = s->F; A->H = s->B + 6 * nw; A->B = s->B; A->nbflag = s->nbflag; A->fs = s->fs; A->nfs = s->nfs; A->all_hard = s->all_hard;
  A->act_cur = s->act; A->act_next = s->act + ntiles; A->list = s->list; A->count = s->count;
  A->tile_res = s->tile_res; A->rlist = s->rlist;
  A->count64 = (unsigned long long *)(s->count + 4);
  A->changed = (unsigned long long *)(ctx->d_mail + 24);
  A->rec = s->rec;
  A->rrej = 1.2f * fmaxf(p->dr_wall, p->dr);
  A->drw2 = p->dr_wall * p->dr_wall;
  A->one_eps = 1.0 + (double)p->eps;
  A->all_hard_d = (const int *)((const char *)s->fs + sizeof(FsDesc) * FS_MAX_DESC);
  if (fresh) {
    k_seg_prep<<<cx_blocks(nbe, 256, ctx->num_sms * 16), 256, 0, st>>>(*A, s->rec);
    CX_LAUNCH_CHECK(ctx, "k_seg_prep");
    CX_CUDA(ctx, cudaMemsetAsync((void *)A->all_hard_d, 0, sizeof(int), st));
    if (s->nfs > 0) {
      k_fs_desc<<<cx_blocks(s->nfs, 128), 128, 0, st>>>(*A, s->fs, (int *)A->all_hard_d);
      CX_LAUNCH_CHECK(ctx, "k_fs_desc");
    }
    k_cell3_flags<<<cx_blocks(p->gridn[3], 256, ctx->num_sms * 16), 256, 0, st>>>(*p, cell_idx_d, s->nbflag);
    CX_LAUNCH_CHECK(ctx, "k_cell3_flags");
    if (!A->classify_inline) {   // lazy per-round resolution (CX_FILL_LAZY=1, kept for comparison) reads the mask tile by tile
      k_classify<<<cx_blocks(nw, 256, ctx->num_sms * 16), 256, 0, st>>>(*A);
      CX_LAUNCH_CHECK(ctx, "k_classify");
    }
  }
  *out = s;
  return CX_OK;
}

// runs rounds until no tile is active; returns the number of rounds
static int fill_rounds(cx_ctx *ctx, FillArgs *A, FillState *s, int32_t *rounds_out)
{
  cudaStream_t st = ctx->stream;
  int rounds = 0;
  const int grid_tiles = ctx->num_sms * 4;
  int32_t *hcount = (int32_t *)(ctx->h_mail + 28);
  for (;;) {
    // a batch of rounds without host synchronisation; empty rounds are cheap
    const int batch = 4;
    for (int r = 0; r < batch; r++) {
      k_round_prep<<<1, 1, 0, st>>>(A->count);
      CX_LAUNCH_CHECK(ctx, "k_round_prep");
      k_build_list<<<cx_blocks(A->ntiles, 256, ctx->num_sms * 8), 256, 0, st>>>(*A);
      CX_LAUNCH_CHECK(ctx, "k_build_list");
      if (!A->resolved_all) {
        k_resolve_tiles<false><<<ctx->num_sms * 4, 256, 0, st>>>(*A);
        CX_LAUNCH_CHECK(ctx, "k_resolve_tiles");
      }
      A->force_activate = (rounds == 0);
      k_fill_tiles<<<grid_tiles, FT_THREADS, 0, st>>>(*A);
      CX_LAUNCH_CHECK(ctx, "k_fill_tiles");
      std::swap(A->act_cur, A->act_next);
      rounds++;
    }
    CX_CUDA(ctx, cudaMemcpyAsync(hcount, A->count, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CX_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->fill_unfinished = 0;
    if (hcount[0] == 0) break;  // the last round of the batch found no active tile
    if (ctx->fill_max_rounds > 0 && rounds >= ctx->fill_max_rounds) { ctx->fill_unfinished = 1; break; }   // caller exchanges halos first
    if (rounds > 4000000) return cx_fail(ctx, CX_INTERNAL_ERROR, "fill: round limit");
  }
  (void)s;
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
  double t0 = 0, t1 = 0, t2 = 0;
  if (dbg) { cudaStreamSynchronize(ctx->stream); t0 = dbg_now(); }
  // first_call: 1 = start of a fill (state is rebuilt), 2 = start of a fill whose directed tests were resolved by
  // cx_fill_resolve_part (+ the caller's merge of the blocked planes): the state is kept, the seed is still to be set
  const bool pre_resolved = first_call == 2;
  int rc = fill_prepare(ctx, p, fpos_d, norm_d, ep_d, pos_d, nbe, cnbe, cell_idx_d, first_call == 1, &A, &s);
  if (rc != CX_OK) return rc;
  if (dbg) { cudaStreamSynchronize(ctx->stream); t1 = dbg_now(); }
  if (k_begin < 0 || k_end > A.dimg[2] || k_begin >= k_end) return cx_fail(ctx, CX_WRONG_INPUT, "fill: bad slab");
  if (pre_resolved && !(s->resolved_lo <= k_begin && s->resolved_hi >= k_end))
    return cx_fail(ctx, CX_WRONG_INPUT, "fill: first_call = 2 without cx_fill_resolve_part");
  A.k_lo = (int)k_begin; A.k_hi = (int)k_end;
  // only the owned planes (and one halo plane either side) are touched: a rank of an N-GPU run does not pay for the lattice
  const int64_t wplane = (int64_t)A.wr * A.dimg[1];
  A.w_lo = wplane * (k_begin > 0 ? k_begin - 1 : 0);
  A.w_hi = wplane * (k_end < A.dimg[2] ? k_end + 1 : k_end);
  A.item_lo = (long long)(k_begin / FT_Z) * A.ty * A.wr * FT_THREADS;
  A.item_hi = (long long)((k_end + FT_Z - 1) / FT_Z) * A.ty * A.wr * FT_THREADS;
  cudaStream_t st = ctx->stream;
  CX_CUDA(ctx, cudaMemsetAsync(A.changed, 0, sizeof(int64_t), st));
  // a previous call that stopped at the round limit left its pending tile activations in the first half of s->act
  if (first_call || !ctx->fill_unfinished) CX_CUDA(ctx, cudaMemsetAsync(s->act, 0, (size_t)A.ntiles * 2, st));
  A.merge = first_call == 1 ? 0 : 1;
  k_unpack<<<cx_blocks(A.w_hi - A.w_lo, 256, ctx->num_sms * 16), 256, 0, st>>>(A, fpos_d);
  CX_LAUNCH_CHECK(ctx, "k_unpack");
  if (first_call) {
    k_seed<<<1, 1, 0, st>>>(A, seed);
    CX_LAUNCH_CHECK(ctx, "k_seed");
  }
  // Resolve every hard cell of the owned planes up front, in one balanced launch: the directed tests are pure functions
  // of the geometry, so nothing is gained by waiting for the flood to arrive (24 small launches with tails before).
  // CX_FILL_LAZY=1 restores the per-round resolution of newly reached tiles.
  static const bool lazy = getenv("CX_FILL_LAZY") != nullptr;
  if (!lazy) {
    if (first_call == 1) { s->resolved_lo = s->resolved_hi = 0; }
    if (!(s->resolved_lo <= A.k_lo && s->resolved_hi >= A.k_hi)) {
      CX_CUDA(ctx, cudaMemsetAsync(A.count64, 0, sizeof(unsigned long long), st));
      k_resolve_tiles<true><<<ctx->num_sms * 6, 256, 0, st>>>(A);
      CX_LAUNCH_CHECK(ctx, "k_resolve_all");
      s->resolved_lo = A.k_lo; s->resolved_hi = A.k_hi;
    }
    A.resolved_all = 1;
  }
  rc = fill_rounds(ctx, &A, s, rounds_host);
  if (rc != CX_OK) return rc;
  if (dbg) {
    cudaStreamSynchronize(ctx->stream);
    t2 = dbg_now();
    fprintf(stderr, "[cx fill] prepare %.3f ms, rounds %.3f ms (%d rounds)\n", 1e3 * (t1 - t0), 1e3 * (t2 - t1), rounds_host ? *rounds_host : -1);
  }
  A.w_lo = wplane * k_begin; A.w_hi = wplane * k_end;   // only owned planes are written back
  k_pack<<<cx_blocks(A.w_hi - A.w_lo, 256, ctx->num_sms * 16), 256, 0, st>>>(A, fpos_d);
  CX_LAUNCH_CHECK(ctx, "k_pack");
  CX_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail + 24, A.changed, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  CX_CUDA(ctx, cudaStreamSynchronize(st));
  if (changed_host) *changed_host = ctx->h_mail[24];
  return CX_OK;
}

extern "C" int cx_fill_geometry(cx_ctx *ctx, const cx_params *p, uint32_t *fpos_d, const float *norm_d, const int32_t *ep_d,
                                const float *pos_d, int64_t nbe, int64_t cnbe, const int32_t *cell_idx_d, int64_t seed, int64_t *nfi_host,
                                int32_t *rounds_host)
{
  if (!p) return CX_WRONG_INPUT;
  int64_t dimg[3];
  cx_fill_dims(p, dimg);
  if (seed >= dimg[0] * dimg[1] * dimg[2]) return cx_fail(ctx, CX_SEED_OUTSIDE, "fill: seed outside the lattice");
  return cx_fill_geometry_slab(ctx, p, fpos_d, norm_d, ep_d, pos_d, nbe, cnbe, cell_idx_d, seed, 0, dimg[2], 1, nfi_host, rounds_host);
}

// ---- multi-GPU helpers: resolve a slab, merge the blocked planes across ranks, propagate everywhere ---------------
// The directed tests are pure functions of the geometry: each rank resolves the hard cells of its own planes
// (the expensive part, shards perfectly), the six blocked bit planes are merged by the caller (disjoint rows: a SUM
// all-reduce is an OR), and the cheap propagation then runs on the whole lattice on every rank -- no halo exchange and
// no per-sweep collective.  Meant for lattices whose bit planes are small; large lattices use cx_fill_geometry_slab.
extern "C" int cx_fill_resolve_slab(cx_ctx *ctx, const cx_params *p, uint32_t *fpos_d, const float *norm_d, const int32_t *ep_d,
                                    const float *pos_d, int64_t nbe, int64_t cnbe, const int32_t *cell_idx_d, int64_t k_begin,
                                    int64_t k_end, uint32_t **blocked_d, int64_t *blocked_words)
{
  if (!ctx || !p || !fpos_d || !norm_d || !ep_d || !pos_d || !cell_idx_d || cnbe < 0 || cnbe > nbe) return CX_WRONG_INPUT;
  FillArgs A;
  FillState *s = nullptr;
  int rc = fill_prepare(ctx, p, fpos_d, norm_d, ep_d, pos_d, nbe, cnbe, cell_idx_d, 1, &A, &s);
  if (rc != CX_OK) return rc;
  if (k_begin < 0 || k_end > A.dimg[2] || k_begin > k_end) return cx_fail(ctx, CX_WRONG_INPUT, "fill: bad slab");
  cudaStream_t st = ctx->stream;
  CX_CUDA(ctx, cudaMemsetAsync(s->act, 0, (size_t)A.ntiles * 2, st));
  A.merge = 0;
  k_unpack<<<cx_blocks(A.nw, 256, ctx->num_sms * 16), 256, 0, st>>>(A, fpos_d);   // cells that are already set need no test
  CX_LAUNCH_CHECK(ctx, "k_unpack");
  A.k_lo = (int)k_begin; A.k_hi = (int)k_end;
  if (k_end > k_begin) {
    CX_CUDA(ctx, cudaMemsetAsync(A.count64, 0, sizeof(unsigned long long), st));
    k_resolve_tiles<true><<<ctx->num_sms * 6, 256, 0, st>>>(A);
    CX_LAUNCH_CHECK(ctx, "k_resolve_all");
  }
  s->resolved_lo = 0; s->resolved_hi = A.dimg[2];   // after the caller's merge every plane is resolved
  if (blocked_d) *blocked_d = s->B;
  if (blocked_words) *blocked_words = 7 * A.nw;
  return CX_OK;
}

// interleaved variant: the 32-cell words of the lattice are dealt to the ranks in chunks of 64 (z-slabs balance badly:
// the floor slab holds most of the hard cells); same contract otherwise
extern "C" int cx_fill_resolve_part(cx_ctx *ctx, const cx_params *p, uint32_t *fpos_d, const float *norm_d, const int32_t *ep_d,
                                    const float *pos_d, int64_t nbe, int64_t cnbe, const int32_t *cell_idx_d, int nparts, int part,
                                    uint32_t **blocked_d, int64_t *blocked_words)
{
  if (!ctx || !p || !fpos_d || !norm_d || !ep_d || !pos_d || !cell_idx_d || cnbe < 0 || cnbe > nbe || nparts < 1 || part < 0 || part >= nparts)
    return CX_WRONG_INPUT;
  FillArgs A;
  FillState *s = nullptr;
  int rc = fill_prepare(ctx, p, fpos_d, norm_d, ep_d, pos_d, nbe, cnbe, cell_idx_d, 1, &A, &s);
  if (rc != CX_OK) return rc;
  cudaStream_t st = ctx->stream;
  CX_CUDA(ctx, cudaMemsetAsync(s->act, 0, (size_t)A.ntiles * 2, st));
  A.merge = 0;
  k_unpack<<<cx_blocks(A.nw, 256, ctx->num_sms * 16), 256, 0, st>>>(A, fpos_d);
  CX_LAUNCH_CHECK(ctx, "k_unpack");
  A.res_nparts = nparts; A.res_part = part;
  CX_CUDA(ctx, cudaMemsetAsync(A.count64, 0, sizeof(unsigned long long), st));
  k_resolve_tiles<true><<<ctx->num_sms * 6, 256, 0, st>>>(A);
  CX_LAUNCH_CHECK(ctx, "k_resolve_all");
  s->resolved_lo = 0; s->resolved_hi = A.dimg[2];
  if (blocked_d) *blocked_d = s->B;
  if (blocked_words) *blocked_words = 7 * A.nw;
  return CX_OK;
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
  A.resolved_all = 1;
  rc = fill_rounds(ctx, &A, s, rounds_host);
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

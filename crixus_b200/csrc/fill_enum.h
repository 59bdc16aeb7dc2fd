/* crixus_b200 -- geometry fill, record-driven resolve: which lattice cells can a wall segment block?
 *
 * The culling rules of k_resolve_cross / k_resolve_inplane (fill.cu), written once for the device and for the host: the
 * CPU test tests/test_resolve_enum.py compiles this header with gcc next to the oracle and checks, on seeded scenes,
 * that every (lattice cell, direction, segment) triple for which the reference's checkCollision arithmetic
 * (crixus_d.cu:753-799, 846-857) reports a collision is enumerated here.  Nothing in this file decides a collision:
 * it only bounds where the reference's arithmetic has to be evaluated.
 *
 * Only for "well-conditioned" records (flag 2 of k_seg_prep) and a dead distance test (eps >= dr_wall^2): then a
 * collision of cell s along direction d (axis ax, step dc) with triangle (v0, u, v, normal n) is one of
 *   crossing : |b| >= eps, r = a / b in [-eps, 1 + eps], and the barycentric coordinates of the hit point s + r dir
 *              within eps of the triangle.   (a = -n.(s - v0), b = n_ax dc)
 *              => every coordinate c != ax of s lies in the triangle's bounding box grown by its tolerances, the
 *                 coordinate ax within one lattice step (1 + eps) of it, and |a| <= |n_ax| (1 + eps) |dc|;
 *   in plane : |b| < eps and |a| < eps, whatever the triangle's extent (crixus_d.cu:766-769)
 *              => |n_ax| lstep < 2 eps ("capable" axis) and |n.(s - v0)| < 2 eps.
 * Both only inside the influence radius and for cells whose 3dr-cell has the segment's cell among its 27 neighbours. */
#ifndef CX_FILL_ENUM_H
#define CX_FILL_ENUM_H
#include <math.h>
#include <stdint.h>
#include "../../include/crixus_b200.h"

#ifdef __CUDACC__
#define FE_FN __host__ __device__ __forceinline__
#define FE_UNROLL _Pragma("unroll")
#else
#define FE_FN static inline
#define FE_UNROLL
#endif
#define FE_PICK(a, i) ((i) == 0 ? (a)[0] : ((i) == 1 ? (a)[1] : (a)[2])) /* no dynamically indexed register arrays on the device */

/* single correctly rounded operations (the lattice positions and cell coordinates must be the reference's bits) */
FE_FN float fe_sub(float a, float b)
{
#ifdef __CUDA_ARCH__
  return __fsub_rn(a, b);
#else
  volatile float r = a - b;
  return r;
#endif
}
FE_FN float fe_div(float a, float b)
{
#ifdef __CUDA_ARCH__
  return __fdiv_rn(a, b);
#else
  volatile float r = a / b;
  return r;
#endif
}
/* lattice position, crixus_d.cu:805-807 (one fma, DESIGN.md "Numerics" R5) */
FE_FN float fe_lat_pos(const cx_params *p, int a, int idx) { return fmaf((float)idx, p->dr, p->fcmin[a]); }
/* 3dr-cell coordinate of a lattice index, crixus_d.cu:821 */
FE_FN int fe_lat_cell(const cx_params *p, int a, int idx)
{
  int g = (int)fe_div(fe_sub(fe_lat_pos(p, a, idx), p->dmin[a]), p->griddr[a]);
  g = g < p->gridn[a] - 1 ? g : p->gridn[a] - 1;
  return g > 0 ? g : 0;
}

/* reach of one lattice step along an axis with the tolerance of the ray parameter, 2 % on top: >= 1.02 (1 + eps) |dc| for
 * every step dc = lat_pos(i +- 1) - lat_pos(i) of the lattice (positions are rounded to the grid of the largest coordinate) */
FE_FN float fe_lstep(const cx_params *p, const int dim[3])
{
  float big = 0.f;
  FE_UNROLL
  for (int a = 0; a < 3; a++) {
    const float e = fabsf(p->fcmin[a]) + (float)(dim[a] + 1) * p->dr;
    big = e > big ? e : big;
  }
  return 1.02f * (1.0f + p->eps) * (p->dr + 2.5e-7f * big);
}

/* what the pre-filters may assume about a triangle (k_seg_prep; evaluated in double):
 *   1 well-formed: |n| = 1 within 1 %, n perpendicular to both edges within 2e-3 rad.  Then for every point within the
 *     influence radius |n.(s - v0)| is the distance to the triangle's plane within a few per cent, and no distance the
 *     reference can compute (plane, vertex or edge-line distance) is smaller;
 *   2 and well-conditioned: sin^2 of the angle between the edges >= 1e-2 and the float D (crixus_d.cu:783) agrees with it:
 *     the barycentric coordinates (t1, t2) the reference computes describe the hit point within ~1e-4, and
 *     |t1 u + t2 v| <= (1 + 3 eps) max(|u|, |v|) for every accepted (t1, t2).
 * n = stored normal, u = v1 - v0, v = v2 - v0 (as the reference subtracts them), D = uv*uv - uu*vv in float. */
FE_FN int fe_classify(const float n[3], const float u[3], const float v[3], float D)
{
  const double nn = (double)n[0] * n[0] + (double)n[1] * n[1] + (double)n[2] * n[2];
  const double lu2 = (double)u[0] * u[0] + (double)u[1] * u[1] + (double)u[2] * u[2];
  const double lv2 = (double)v[0] * v[0] + (double)v[1] * v[1] + (double)v[2] * v[2];
  const double lu = sqrt(lu2), lv = sqrt(lv2);
  const double nu = (double)n[0] * u[0] + (double)n[1] * u[1] + (double)n[2] * u[2];
  const double nv = (double)n[0] * v[0] + (double)n[1] * v[1] + (double)n[2] * v[2];
  const double duv = (double)u[0] * v[0] + (double)u[1] * v[1] + (double)u[2] * v[2];
  const int wf = nn > 0.98 && nn < 1.02 && lu > 0 && lv > 0 && fabs(nu) <= 2e-3 * lu && fabs(nv) <= 2e-3 * lv;
  const double Dd = lu2 * lv2 - duv * duv;
  const int wc = wf && Dd >= 1e-2 * lu2 * lv2 && isfinite(D) && fabs((double)(-D) - Dd) <= 1e-3 * Dd;
  return wc ? 2 : (wf ? 1 : 0);
}

typedef struct FeRec {
  float n[3], v0[3];
  float blo[3], bhi[3]; /* where the hit point of an accepted crossing can lie (bounding box of the triangle + tolerances) */
  float rad;            /* ... and its distance from v0 is at most rad */
  int g[3];             /* the segment's 3dr-cell */
  int safe;             /* every lattice point the boxes below can produce is a 27-neighbour of g: no per-cell check needed */
} FeRec;

/* 3dr-cell index (< 2^31: cell_idx is int32) -> coordinates */
FE_FN void fe_cell_coords(const cx_params *p, int cell, int g[3])
{
  const unsigned c = (unsigned)cell, gz = (unsigned)p->gridn[2], gy = (unsigned)p->gridn[1];
  const unsigned cz = c / gz;
  g[2] = (int)(c - cz * gz);
  const unsigned cy = cz / gy;
  g[1] = (int)(cz - cy * gy);
  g[0] = (int)cy;
}

/* n, v0, v1, v2 of the record; uu, vv = |v1 - v0|^2, |v2 - v0|^2 as k_seg_prep stores them; g = its 3dr-cell */
FE_FN void fe_rec_setup(const cx_params *p, float lstep, const float n[3], const float v0[3], const float v1[3], const float v2[3],
                        float uu, float vv, const int g[3], FeRec *T)
{
  const float eps = p->eps;
  const float maxlen = sqrtf(uu > vv ? uu : vv);
  T->rad = 1.01f * (1.0f + 3.0f * eps) * maxlen;
  T->g[0] = g[0]; T->g[1] = g[1]; T->g[2] = g[2];
  T->safe = 1;
  FE_UNROLL
  for (int c = 0; c < 3; c++) {
    T->n[c] = n[c];
    T->v0[c] = v0[c];
    const float uc = v1[c] - v0[c], vc = v2[c] - v0[c];
    const float mn = fminf(0.f, fminf(uc, vc)), mx = fmaxf(0.f, fmaxf(uc, vc));
    /* hit point - v0 = t1 u + t2 v with t1, t2 >= -eps, t1 + t2 <= 1 + eps: per coordinate within [mn, mx] grown by
     * 3 eps max(|u_c|, |v_c|); 1 % of the longer edge for what the float (t1, t2) of a well-conditioned triangle and the
     * stored normal can add (the same margin as rad), and the rounding of the coordinates themselves */
    const float m = (0.01f + 3.1f * eps) * (1.0f + 3.0f * eps) * maxlen + 1e-6f * (fabsf(v0[c]) + maxlen);
    T->blo[c] = v0[c] + mn - m;
    T->bhi[c] = v0[c] + mx + m;
    /* the 27 neighbours of g along this axis hold every point of [dmin + (g-1) griddr, dmin + (g+2) griddr) -- without a
     * lower (upper) end when the row of cells ends there, because cell coordinates are clamped to the grid */
    const float reach = lstep + 1e-5f * (p->griddr[c] + fabsf(T->blo[c]) + fabsf(T->bhi[c]));
    if (T->g[c] - 1 > 0 && !(T->blo[c] - reach > p->dmin[c] + (float)(T->g[c] - 1) * p->griddr[c])) T->safe = 0;
    if (T->g[c] + 1 < p->gridn[c] - 1 && !(T->bhi[c] + reach < p->dmin[c] + (float)(T->g[c] + 2) * p->griddr[c])) T->safe = 0;
  }
}

/* can a step along axis ax cross the triangle's plane at all (|b| >= eps for some cell)? */
FE_FN int fe_axis_crossing(const cx_params *p, float lstep, const FeRec *T, int ax) { return !(fabsf(FE_PICK(T->n, ax)) * lstep < p->eps); }
/* directions (bit d: 0:-x 1:+x 2:-y 3:+y 4:-z 5:+z) along which the "ray in plane" rule can apply */
FE_FN uint32_t fe_capable(const cx_params *p, float lstep, const float n[3])
{
  const float e2 = 2.0f * p->eps;
  return (fabsf(n[0]) * lstep < e2 ? 3u : 0u) | (fabsf(n[1]) * lstep < e2 ? 12u : 0u) | (fabsf(n[2]) * lstep < e2 ? 48u : 0u);
}

/* inclusive lattice index range along coordinate c of the cells a crossing can start from (grow = lstep along the axis of
 * the step, 0 across it): exactly the indices i in [cut_lo, cut_hi) with blo - grow <= lat_pos(i) <= bhi + grow (lat_pos is
 * monotone in i); empty when lo > hi */
FE_FN void fe_axis_range(const cx_params *p, const FeRec *T, int c, float grow, int cut_lo, int cut_hi, int *lo, int *hi)
{
  const float inv = 1.0f / p->dr, lob = T->blo[c] - grow, hib = T->bhi[c] + grow;
  float a = floorf((lob - p->fcmin[c]) * inv) - 1.0f;
  float b = ceilf((hib - p->fcmin[c]) * inv) + 1.0f;
  if (!(a > (float)cut_lo)) a = (float)cut_lo;             /* also takes NaN to the whole range */
  if (!(b < (float)(cut_hi - 1))) b = (float)(cut_hi - 1);
  int ia = (int)a, ib = (int)b;
  ia = ia > cut_lo ? ia : cut_lo;             /* (float)(cut_hi - 1) may have rounded up on lattices of more than 2^24 planes */
  ib = ib < cut_hi - 1 ? ib : cut_hi - 1;
  while (ia <= ib && fe_lat_pos(p, c, ia) < lob) ia++;
  while (ib >= ia && fe_lat_pos(p, c, ib) > hib) ib--;
  *lo = ia;
  *hi = ib;
}

/* is g a neighbour (with periodic wrap) of gc along axis ax?  crixus_d.cu:822-846 */
FE_FN int fe_nb_axis(const cx_params *p, int ax, int gs, int gc)
{
  const int d = gc - gs;
  if (d >= -1 && d <= 1) return 1;
  return p->per[ax] && ((gs == 0 && gc == p->gridn[ax] - 1) || (gs == p->gridn[ax] - 1 && gc == 0));
}

/* one lattice cell of the box of axis ax: can a step along ax from (si, sj, sk) cross the triangle?  (conservative; records
 * with T->safe only -- the others stay with the cell-driven kernel, which looks segments up the way the reference does) */
FE_FN int fe_cross_candidate(const cx_params *p, float lstep, const FeRec *T, int ax, int si, int sj, int sk)
{
  const int idx[3] = {si, sj, sk};
  float w[3];
  FE_UNROLL
  for (int c = 0; c < 3; c++) {
    const float s = fe_lat_pos(p, c, idx[c]);
    const float grow = c == ax ? lstep : 0.f;
    if (!(s >= T->blo[c] - grow && s <= T->bhi[c] + grow)) return 0;
    w[c] = s - T->v0[c];
  }
  const float aa = fabsf(w[0] * T->n[0] + w[1] * T->n[1] + w[2] * T->n[2]);
  /* |a| <= |b| (1 + eps) <= |n_ax| lstep / 1.02; the slack covers the rounding of this dot product many times over */
  if (!(aa <= fabsf(FE_PICK(T->n, ax)) * lstep + 2e-6f * (fabsf(w[0]) + fabsf(w[1]) + fabsf(w[2])))) return 0;
  return 1;   /* T->safe: the cell is a 27-neighbour of the segment's cell */
}

/* |s - v0| of every (cell, segment) pair that passes the influence-radius test (crixus_d.cu:815, 855-857: the midpoint of the
 * step within 3 dr + |step| / 2 + dr_wall of v0), 2 % on top */
FE_FN float fe_infl_reach(const cx_params *p, float lstep) { return 4.0f * lstep + 1.01f * p->dr_wall; }

/* in-plane rule.  Planes are kept relative to the lattice origin, (n, d = n.(v0 - fcmin)), and tested at q = lat_pos - fcmin
 * (the rounded lattice position the reference uses, minus the origin): every term is then of the size of the domain, so that the
 * rounding of n.q - d is ~1e-7 x extent = 0.01 eps whatever the offset of the geometry from the coordinate origin. */
FE_FN float fe_plane_offset(const cx_params *p, const float n[3], const float v0[3])
{
  return n[0] * (v0[0] - p->fcmin[0]) + n[1] * (v0[1] - p->fcmin[1]) + n[2] * (v0[2] - p->fcmin[2]);
}
/* can the cell at q lie in the plane (n, d), within the influence radius of v0?  conservative */
FE_FN int fe_inplane_candidate(const cx_params *p, const float n[3], float d, const float q[3], const float v0[3], float infl_reach)
{
  if (!(fabsf(q[0] * n[0] + q[1] * n[1] + q[2] * n[2] - d) < 2.0f * p->eps)) return 0;
  const float wx = q[0] - (v0[0] - p->fcmin[0]), wy = q[1] - (v0[1] - p->fcmin[1]), wz = q[2] - (v0[2] - p->fcmin[2]);
  return wx * wx + wy * wy + wz * wz <= infl_reach * infl_reach;
}
#endif

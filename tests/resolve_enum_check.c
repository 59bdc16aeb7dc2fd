/* TEST INFRASTRUCTURE.  Checks the culling rules of the record-driven fill resolve (crixus_b200/csrc/fill_enum.h, the very
 * header fill.cu compiles for the device) against the oracle's restatement of checkCollision, record by record:
 * every (lattice cell, direction, wall segment) triple that the reference's arithmetic blocks must be among the
 * candidates the header enumerates -- crossing candidates (k_resolve_cross) or in-plane candidates (k_resolve_inplane).
 * Built and called by tests/test_resolve_enum.py (gcc, -ffp-contract=off like the oracle). */
#include "../oracle/crixus_oracle.c"
#include "../crixus_b200/csrc/fill_enum.h"
#include <stdio.h>

/* out: [0] triples the reference blocks (eligible records)  [1] of those NOT enumerated (must be 0)
 *      [2] crossing candidates enumerated  [3] in-plane candidates enumerated  [4] eligible records (flag 3)
 *      [5] records  [6] well-conditioned but not "safe" records  [7] cells x directions looked at
 * returns 0, or -1 when the distance test is alive (the record-driven path is not used then) */
int64_t rec_enum_check(const orc_params *op, const float *pos, const float *norm, const int32_t *ep, int64_t nwall,
                       const int32_t *cell_idx, int64_t *out, int verbose)
{
  const cx_params *p = (const cx_params *)op; /* same layout (oracle/orc.py: Params) */
  if (!(op->eps >= op->dr_wall * op->dr_wall)) return -1;
  int64_t dimg[3];
  orc_fill_dims(op, dimg);
  const int dim[3] = {(int)dimg[0], (int)dimg[1], (int)dimg[2]};
  const float lstep = fe_lstep(p, dim);
  const int64_t G = dimg[0] * dimg[1] * dimg[2];
  uint8_t *mark = (uint8_t *)calloc((size_t)G, 1);
  for (int k = 0; k < 8; k++) out[k] = 0;
  const int reach = 6 + (int)ceilf(op->dr_wall / op->dr);
  for (int64_t c = 0; c < op->gridn[3]; c++) {
    int g3[3];
    fe_cell_coords(p, (int)c, g3);
    if (((int64_t)g3[0] * op->gridn[1] + g3[1]) * op->gridn[2] + g3[2] != c) return -2;
    for (int64_t i = cell_idx[c]; i < cell_idx[c + 1] && i < nwall; i++) {
      out[5]++;
      const float *n = norm + 4 * i;
      const float *v0 = pos + 4 * (int64_t)ep[4 * i], *v1 = pos + 4 * (int64_t)ep[4 * i + 1], *v2 = pos + 4 * (int64_t)ep[4 * i + 2];
      float u[3], v[3];
      for (int j = 0; j < 3; j++) { u[j] = v1[j] - v0[j]; v[j] = v2[j] - v0[j]; }
      const float uu = dot_r3(u, u), uv = dot_r3(u, v), vv = dot_r3(v, v);
      const float D = det_r2(uv, uv, uu, vv);
      if (fe_classify(n, u, v, D) != 2) continue;
      FeRec T;
      fe_rec_setup(p, lstep, n, v0, v1, v2, uu, vv, g3, &T);
      if (!T.safe) { out[6]++; continue; }
      out[4]++;
      /* the box of cells looked at: everything the influence radius can reach */
      int blo[3], bhi[3];
      for (int a = 0; a < 3; a++) {
        const int ctr = (int)floor(((double)v0[a] - op->fcmin[a]) / op->dr);
        blo[a] = imax_(ctr - reach, 0);
        bhi[a] = imin_(ctr + reach + 1, dim[a] - 1);
      }
      /* ---- candidates, as the kernels enumerate them */
      for (int ax = 0; ax < 3; ax++) {
        if (!fe_axis_crossing(p, lstep, &T, ax)) continue;
        int lo[3], hi[3];
        for (int a = 0; a < 3; a++) fe_axis_range(p, &T, a, a == ax ? lstep : 0.f, 0, dim[a], &lo[a], &hi[a]);
        for (int sk = lo[2]; sk <= hi[2]; sk++)
          for (int sj = lo[1]; sj <= hi[1]; sj++)
            for (int si = lo[0]; si <= hi[0]; si++)
              if (fe_cross_candidate(p, lstep, &T, ax, si, sj, sk)) {
                out[2]++;
                if (si < blo[0] || si > bhi[0] || sj < blo[1] || sj > bhi[1] || sk < blo[2] || sk > bhi[2]) continue;
                mark[((int64_t)sk * dim[1] + sj) * dim[0] + si] |= (uint8_t)(3u << (2 * ax));
              }
      }
      const uint32_t cap = fe_capable(p, lstep, T.n);
      if (cap) {
        const float d = fe_plane_offset(p, T.n, v0);
        for (int sk = blo[2]; sk <= bhi[2]; sk++)
          for (int sj = blo[1]; sj <= bhi[1]; sj++)
            for (int si = blo[0]; si <= bhi[0]; si++) {
              const float q[3] = {fe_lat_pos(p, 0, si) - p->fcmin[0], fe_lat_pos(p, 1, sj) - p->fcmin[1], fe_lat_pos(p, 2, sk) - p->fcmin[2]};
              if (fe_inplane_candidate(p, T.n, d, q, v0, fe_infl_reach(p, lstep))) {
                out[3]++;
                mark[((int64_t)sk * dim[1] + sj) * dim[0] + si] |= (uint8_t)cap;
              }
            }
      }
      /* ---- the reference's arithmetic for this record alone (crixus_d.cu:846-857 influence radius, :942 segment test) */
      for (int sk = blo[2]; sk <= bhi[2]; sk++)
        for (int sj = blo[1]; sj <= bhi[1]; sj++)
          for (int si = blo[0]; si <= bhi[0]; si++) {
            const int64_t q = ((int64_t)sk * dim[1] + sj) * dim[0] + si;
            const int sidx[3] = {si, sj, sk};
            float s[3];
            for (int a = 0; a < 3; a++) s[a] = fmaf((float)sidx[a], op->dr, op->fcmin[a]);
            int gridi[3], seen = 0;
            cell_of(op, s, gridi);
            for (int gi = -1; gi <= 1 && !seen; gi++)
              for (int gj = -1; gj <= 1 && !seen; gj++)
                for (int gk = -1; gk <= 1 && !seen; gk++)
                  if (nb_cell(op, gridi, gi, gj, gk) == c) seen = 1;
            if (seen)
              for (int d = 0; d < 6; d++) {
                const int a = d >> 1;
                int eidx[3] = {si, sj, sk};
                eidx[a] += (d & 1) ? 1 : -1;
                if (eidx[a] < 0 || eidx[a] >= dim[a]) continue;
                out[7]++;
                float e[3], dir[3], hd[3], mid[3], t[3];
                for (int b = 0; b < 3; b++) {
                  e[b] = fmaf((float)eidx[b], op->dr, op->fcmin[b]);
                  dir[b] = e[b] - s[b]; hd[b] = (s[b] - e[b]) * 0.5f; mid[b] = s[b] + e[b];
                }
                float infl = fmaf(op->dr, 3.0f, sqrtf(dot_r3(hd, hd))) + op->dr_wall;
                infl *= infl;
                for (int b = 0; b < 3; b++) t[b] = fmaf(mid[b], 0.5f, -v0[b]);
                if (dot_r3(t, t) > infl) continue;
                if (!tri_collision(op, s, dir, n, v0, v1, v2)) continue;
                out[0]++;
                if (!(mark[q] >> d & 1)) {
                  out[1]++;
                  if (verbose && out[1] <= 10)
                    fprintf(stderr, "missed: record %lld cell (%d,%d,%d) dir %d  n=(%g,%g,%g)\n", (long long)i, si, sj, sk, d, n[0], n[1], n[2]);
                }
              }
            mark[q] = 0;
          }
    }
  }
  free(mark);
  return 0;
}

/* The "remembered plane" of k_resolve_inplane: a segment j is skipped when a plane with the same normal bits and an offset within
 * `tol` x eps of its own (tol = 0.5 per lane; 0.75 through the per-cell summaries) has all the cells of ITS 2-eps mask blocked.
 * That is sound if every cell the reference puts in the plane of j (|a| < eps with the reference's arithmetic, crixus_d.cu:757,766)
 * is in the 2-eps mask of that other plane i, computed the way the kernel computes it (fe_plane_offset, lattice positions relative
 * to the lattice origin).  Checked here for every pair (i, j) of eligible capable segments of the same or neighbouring 3dr-cells.
 * out: [0] pairs within the tolerance  [1] in-plane cells of j looked at  [2] of those outside i's mask (must be 0) */
int64_t plane_cache_check(const orc_params *op, const float *pos, const float *norm, const int32_t *ep, int64_t nwall,
                          const int32_t *cell_idx, double tol, int64_t *out)
{
  const cx_params *p = (const cx_params *)op;
  if (!(op->eps >= op->dr_wall * op->dr_wall)) return -1;
  int64_t dimg[3];
  orc_fill_dims(op, dimg);
  const int dim[3] = {(int)dimg[0], (int)dimg[1], (int)dimg[2]};
  const float lstep = fe_lstep(p, dim);
  out[0] = out[1] = out[2] = 0;
  /* eligible capable segments with their planes */
  int64_t *idx = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nwall + 1));
  float *dd = (float *)malloc(sizeof(float) * (size_t)(nwall + 1));
  int *cellof = (int *)malloc(sizeof(int) * (size_t)(nwall + 1));
  int64_t m = 0;
  for (int64_t c = 0; c < op->gridn[3]; c++) {
    int g3[3];
    fe_cell_coords(p, (int)c, g3);
    for (int64_t i = cell_idx[c]; i < cell_idx[c + 1] && i < nwall; i++) {
      const float *n = norm + 4 * i;
      const float *v0 = pos + 4 * (int64_t)ep[4 * i], *v1 = pos + 4 * (int64_t)ep[4 * i + 1], *v2 = pos + 4 * (int64_t)ep[4 * i + 2];
      float u[3], v[3];
      for (int j = 0; j < 3; j++) { u[j] = v1[j] - v0[j]; v[j] = v2[j] - v0[j]; }
      const float uu = dot_r3(u, u), uv = dot_r3(u, v), vv = dot_r3(v, v);
      if (fe_classify(n, u, v, det_r2(uv, uv, uu, vv)) != 2) continue;
      FeRec T;
      fe_rec_setup(p, lstep, n, v0, v1, v2, uu, vv, g3, &T);
      if (!T.safe || !fe_capable(p, lstep, n)) continue;
      idx[m] = i; dd[m] = fe_plane_offset(p, n, v0); cellof[m] = (int)c; m++;
    }
  }
  const int reach = 6 + (int)ceilf(op->dr_wall / op->dr);
  for (int64_t a = 0; a < m; a++)
    for (int64_t b = 0; b < m; b++) {
      const float *ni = norm + 4 * idx[a], *nj = norm + 4 * idx[b];
      if (!(ni[0] == nj[0] && ni[1] == nj[1] && ni[2] == nj[2])) continue;
      if (!(fabs((double)dd[a] - (double)dd[b]) < tol * op->eps)) continue;
      int ga[3], gb[3], near = 1;
      fe_cell_coords(p, cellof[a], ga); fe_cell_coords(p, cellof[b], gb);
      for (int k = 0; k < 3; k++) if (abs(ga[k] - gb[k]) > 2) near = 0;
      if (!near) continue;
      out[0]++;
      const float *v0 = pos + 4 * (int64_t)ep[4 * idx[b]];
      for (int sk = 0; sk < dim[2]; sk++) {
        const int ck = (int)floor(((double)v0[2] - op->fcmin[2]) / op->dr);
        if (abs(sk - ck) > reach) continue;
        for (int sj = 0; sj < dim[1]; sj++) {
          const int cj = (int)floor(((double)v0[1] - op->fcmin[1]) / op->dr);
          if (abs(sj - cj) > reach) continue;
          for (int si = 0; si < dim[0]; si++) {
            const int ci = (int)floor(((double)v0[0] - op->fcmin[0]) / op->dr);
            if (abs(si - ci) > reach) continue;
            const float s[3] = {fe_lat_pos(p, 0, si), fe_lat_pos(p, 1, sj), fe_lat_pos(p, 2, sk)};
            float w0[3];
            for (int k = 0; k < 3; k++) w0[k] = s[k] - v0[k];
            const float aj = -dot_r3(nj, w0);                  /* the reference's a for segment j */
            if (!(fabsf(aj) < op->eps)) continue;
            out[1]++;
            const float q[3] = {s[0] - p->fcmin[0], s[1] - p->fcmin[1], s[2] - p->fcmin[2]};
            if (!(fabsf(q[0] * ni[0] + q[1] * ni[1] + q[2] * ni[2] - dd[a]) < 2.0f * op->eps)) out[2]++;   /* i's mask, as the kernel */
          }
        }
      }
    }
  free(idx); free(dd); free(cellof);
  return 0;
}

/* TEST INFRASTRUCTURE -- not part of the product.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.
 *
 * CPU restatement (plain C + OpenMP, 64-bit indices, stable binning order) of the hot path of
 * Azrael3000/Crixus.  Every function cites the reference file:line it follows (paths relative to
 * /root/reference).  Floating-point expression trees are kept as in the reference *including the
 * fused-multiply-add pattern nvcc 12.9 (-fmad=true) emits for them on sm_100* -- determined by reading the
 * SASS of the reference build (oracle/build_ref.sh, DESIGN.md "Numerics") -- so this file must be
 * compiled with -ffp-contract=off: every fusion is an explicit fmaf().
 *
 * Observed contraction rules (reference SASS, sm_100):
 *   R1  sp = 0; sp += a[o]*b[o] (o=0..2)       -> fma(a2,b2, fma(a1,b1, fma(a0,b0, 0)))
 *   R2  a*b - c*d                              -> fma(a, b, -(c*d))          (c*d rounded first)
 *   R3  a0*b0 + a1*b1 + a2*b2 (dot3 / dot())   -> fma(a2,b2, fma(a0,b0, a1*b1))
 *   R4  x - n*d   (uf4 operator* then -)       -> fma(-n, d, x)
 *   R5  (float)i*dr + fcmin                    -> fma((float)i, dr, fcmin)
 *   R6  len + 3.0f*dr                          -> fma(dr, 3, len)
 *   pow(float,2) is double pow: the square of a float is exact in double, so it is restated as x*x in double.
 *
 * Pinning: checked in tests/ against (a) oracle/_ref/libcrixus_refshim.so (the reference's own kernels on the
 * host, g++ contraction) and (b) golden outputs of the reference CUDA binary run on a B200 (tests/golden/).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define GRES 20   /* src/crixus.h:26 */
#define TRIMAX 50 /* src/crixus.h:27 */
#define MAX_ITERATIONS 10000 /* src/crixus.h:30 */

/* mirrors the __constant__ symbols of src/crixus_d.cu:13-32 */
typedef struct {
  float dr, dr_wall, eps;
  float dmin[4], dmax[4];
  float griddr[4];
  int gridn[4];
  float fcmin[4], fcmax[4];
  int per[3];
} orc_params;

static inline float sgnf_(float x) { return (float)((x > 0.) - (x < 0.)); } /* src/crixus.h:22 */
static inline float dot_r1(const float *a, const float *b)
{
  float sp = fmaf(a[0], b[0], 0.0f);
  sp = fmaf(a[1], b[1], sp);
  return fmaf(a[2], b[2], sp);
}
static inline float dot_r3(const float *a, const float *b) { return fmaf(a[2], b[2], fmaf(a[0], b[0], a[1] * b[1])); }
static inline float det_r2(float a, float b, float c, float d) { return fmaf(a, b, -(c * d)); }
static inline void cross_r2(const float *a, const float *b, float *c) /* src/vector_math.h:48-55 */
{
  c[0] = det_r2(a[1], b[2], a[2], b[1]);
  c[1] = det_r2(a[2], b[0], a[0], b[2]);
  c[2] = det_r2(a[0], b[1], a[1], b[0]);
}
static inline int imin_(int a, int b) { return a < b ? a : b; }
static inline int imax_(int a, int b) { return a > b ? a : b; }

/* ------------------------------------------------------------------------------------------------ params
 * src/crixus.cu:262-283: eps#1, 3dr cell grid.  bbox = min/max over all raw STL corners. */
void orc_params_domain(orc_params *p, float dr, const float *bbmin, const float *bbmax)
{
  memset(p, 0, sizeof(*p));
  p->dr = dr;
  p->dr_wall = dr;
  for (int j = 0; j < 3; j++) {
    p->dmin[j] = bbmin[j];
    p->dmax[j] = bbmax[j];
  }
  p->eps = (float)(dr / (float)GRES * 1e-4); /* :265 */
  for (int j = 0; j < 3; j++) p->gridn[j] = imax_((int)((bbmax[j] - bbmin[j]) / dr / 3.0), 1); /* :270-272 */
  p->gridn[3] = p->gridn[0] * p->gridn[1] * p->gridn[2];
  for (int j = 0; j < 3; j++) p->griddr[j] = (p->dmax[j] - p->dmin[j] + p->eps) / (float)p->gridn[j]; /* :276 */
  for (int j = 0; j < 3; j++) { p->fcmin[j] = p->dmin[j]; p->fcmax[j] = p->dmax[j]; }
}

/* src/crixus.cu:464-467: eps#2 from the domain extent (used by all fill code) */
void orc_params_fill_eps(orc_params *p)
{
  float eps = 1e-10f;
  for (int i = 0; i < 3; i++) {
    float e = (p->dmax[i] - p->dmin[i]) * 1e-5f;
    eps = e > eps ? e : eps;
  }
  p->eps = eps;
}

/* src/crixus.cu:698-722: fluid container.  use_container: explicit box; else bbox minus dr in periodic dims. */
void orc_params_container(orc_params *p, int use_container, const float *cmin, const float *cmax)
{
  for (int i = 0; i < 3; i++) {
    if (use_container) {
      p->fcmin[i] = cmin[i];
      p->fcmax[i] = cmax[i];
    } else {
      p->fcmin[i] = p->dmin[i];
      p->fcmax[i] = p->dmax[i];
      if (p->per[i]) p->fcmax[i] -= p->dr;
    }
  }
}

/* src/crixus_d.cu:985-987 (device) == src/crixus.cu:802-803,987-989 (host): dr-lattice dimensions */
void orc_fill_dims(const orc_params *p, int64_t dimg[3])
{
  for (int j = 0; j < 3; j++) dimg[j] = (int)floorf((p->fcmax[j] - p->fcmin[j] + p->eps) / p->dr + 1);
}

/* src/crixus.cu:799-804: seed cell index; returns -1 if outside the lattice (the reference does not check) */
int64_t orc_seed_index(const orc_params *p, const float *spos)
{
  int64_t dimg[3], s[3];
  orc_fill_dims(p, dimg);
  for (int j = 0; j < 3; j++) {
    s[j] = (int)roundf((spos[j] - p->fcmin[j] + p->eps) / p->dr);
    if (s[j] < 0 || s[j] >= dimg[j]) return -1;
  }
  return s[0] + s[1] * dimg[0] + s[2] * dimg[0] * dimg[1];
}

/* ------------------------------------------------------------------------------------------------ weld
 * src/crixus.cpp:343-455 remove_duplicate_vertices, outcome restated:
 *  - corners are grouped by cell ((int)(x/dr),(int)(y/dr),(int)(z/dr)) (float divide, truncation) :354
 *  - cells are visited in lexicographic (x,y,z) order (std::map<mint3>), corners of a cell in insertion order
 *  - a corner not yet "found" becomes a new vertex and claims every LATER corner of the same cell with
 *    dist^2 < 1e-10*dr*dr (float dist, double threshold) :317-329 -- including corners already claimed
 *  - the neighbour-cell searches operate on a by-value copy of the map (:331): they only write new_ids_map,
 *    which is overwritten when that later cell is visited, so they have no lasting effect. */
typedef struct { int cx, cy, cz; int64_t idx; } weld_key;
static int weld_cmp(const void *a, const void *b)
{
  const weld_key *p = (const weld_key *)a, *q = (const weld_key *)b;
  if (p->cx != q->cx) return p->cx < q->cx ? -1 : 1;
  if (p->cy != q->cy) return p->cy < q->cy ? -1 : 1;
  if (p->cz != q->cz) return p->cz < q->cz ? -1 : 1;
  return p->idx < q->idx ? -1 : (p->idx > q->idx);
}
int64_t orc_weld(const float *corners, int64_t ncorner, float dr, float *out_verts, int32_t *out_ids)
{
  weld_key *keys = (weld_key *)malloc(sizeof(weld_key) * (size_t)ncorner);
  char *found = (char *)calloc((size_t)ncorner, 1);
  for (int64_t c = 0; c < ncorner; c++) {
    keys[c].cx = (int)(corners[3 * c + 0] / dr);
    keys[c].cy = (int)(corners[3 * c + 1] / dr);
    keys[c].cz = (int)(corners[3 * c + 2] / dr);
    keys[c].idx = c;
  }
  qsort(keys, (size_t)ncorner, sizeof(weld_key), weld_cmp);
  int64_t nvert = 0;
  const double thr = 1e-10 * dr * dr;
  for (int64_t a = 0; a < ncorner;) {
    int64_t b = a;
    while (b < ncorner && keys[b].cx == keys[a].cx && keys[b].cy == keys[a].cy && keys[b].cz == keys[a].cz) b++;
    for (int64_t i = a; i < b; i++) {
      if (found[i]) continue;
      found[i] = 1;
      const float *pi = corners + 3 * keys[i].idx;
      out_verts[3 * nvert + 0] = pi[0];
      out_verts[3 * nvert + 1] = pi[1];
      out_verts[3 * nvert + 2] = pi[2];
      out_ids[keys[i].idx] = (int32_t)nvert;
      for (int64_t k = i + 1; k < b; k++) {
        const float *pk = corners + 3 * keys[k].idx;
        const float xij = pk[0] - pi[0], yij = pk[1] - pi[1], zij = pk[2] - pi[2];
        const float dist = xij * xij + yij * yij + zij * zij; /* :323, g++ -O2 host code: no contraction */
        if (dist < thr) {
          found[k] = 1;
          out_ids[keys[k].idx] = (int32_t)nvert;
        }
      }
      nvert++;
    }
    a = b;
  }
  free(keys);
  free(found);
  return nvert;
}

/* ------------------------------------------------------------------------------------------------ swap_normals
 * src/crixus_d.cu:203-211 */
void orc_swap_normals(float *norm, int64_t nbe)
{
  for (int64_t i = 0; i < nbe; i++)
    for (int j = 0; j < 3; j++) norm[4 * i + j] = (float)(norm[4 * i + j] * -1.);
}

/* ------------------------------------------------------------------------------------------------ set_bound_elem
 * src/crixus_d.cu:105-201.  pos: [nvert+nbe] float4; writes pos[nvert+i], surf[i], may swap ep[i].a[1]<->a[2].
 * zstat = {xminp, xminn, nminp, nminn}; ties on z resolved towards the lowest segment index (the reference's
 * tie order is schedule dependent). */
void orc_set_bound_elem(float *pos, const float *norm, float *surf, int32_t *ep, int64_t nbe, int64_t nvert,
                        float zstat[4])
{
  float xminp = 1e10f, xminn = 1e10f, nminp = 0.f, nminn = 0.f; /* src/crixus.cu:325-326 */
  for (int64_t i = 0; i < nbe; i++) {
    const float *v0 = pos + 4 * (int64_t)ep[4 * i + 0], *v1 = pos + 4 * (int64_t)ep[4 * i + 1],
                *v2 = pos + 4 * (int64_t)ep[4 * i + 2];
    float a2 = 0.f, b2 = 0.f, c2 = 0.f, ddum[3] = {0.f, 0.f, 0.f}, ndir[3];
    for (int j = 0; j < 3; j++) {
      ddum[j] = (float)((double)ddum[j] + (double)v0[j] / 3.); /* :132-134 */
      ddum[j] = (float)((double)ddum[j] + (double)v1[j] / 3.);
      ddum[j] = (float)((double)ddum[j] + (double)v2[j] / 3.);
      double d;
      d = (double)(v0[j] - v1[j]); a2 = (float)((double)a2 + d * d); /* :135 pow(float,2) in double */
      d = (double)(v1[j] - v2[j]); b2 = (float)((double)b2 + d * d);
      d = (double)(v2[j] - v0[j]); c2 = (float)((double)c2 + d * d);
      const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      ndir[j] = det_r2(v1[j1] - v0[j1], v2[j2] - v0[j2], v1[j2] - v0[j2], v2[j1] - v0[j1]); /* :139-142 */
    }
    const float *n = norm + 4 * i;
    if (dot_r3(n, ndir) < 0.0f) { /* :148 */
      int32_t tmp = ep[4 * i + 1];
      ep[4 * i + 1] = ep[4 * i + 2];
      ep[4 * i + 2] = tmp;
    }
    if ((double)n[2] > 1e-5 && xminp > ddum[2]) { xminp = ddum[2]; nminp = n[2]; }  /* :153 */
    if ((double)n[2] < -1e-5 && xminn > ddum[2]) { xminn = ddum[2]; nminn = n[2]; } /* :157 */
    const double s = (double)(a2 + b2 - c2);
    surf[i] = (float)(0.25 * sqrt(4. * (double)a2 * (double)b2 - s * s)); /* :161 */
    for (int j = 0; j < 3; j++) pos[4 * (i + nvert) + j] = ddum[j];
  }
  zstat[0] = xminp; zstat[1] = xminn; zstat[2] = nminp; zstat[3] = nminn;
}

/* ------------------------------------------------------------------------------------------------ binning
 * src/crixus_d.cu:37-103 (init_cell_idx, count_cell_bes, add_up_indices, sort_bes) with a STABLE scatter. */
static inline int64_t cell_of(const orc_params *p, const float *x, int g[3])
{
  for (int j = 0; j < 3; j++)
    g[j] = imax_(imin_((int)((x[j] - p->dmin[j]) / p->griddr[j]), p->gridn[j] - 1), 0); /* :55 */
  return (int64_t)g[0] * p->gridn[1] * p->gridn[2] + (int64_t)g[1] * p->gridn[2] + g[2]; /* :56 */
}
void orc_bin_segments(const orc_params *p, const float *pre_pos, float *pos, const float *pre_norm, float *norm,
                      const int32_t *pre_ep, int32_t *ep, const float *pre_surf, float *surf, int32_t *cell_idx,
                      int64_t nbe, int64_t nvert)
{
  const int64_t ncell = p->gridn[3];
  int g[3];
  memset(cell_idx, 0, sizeof(int32_t) * (size_t)(ncell + 1));
  for (int64_t i = 0; i < nbe; i++) cell_idx[cell_of(p, pre_pos + 4 * (nvert + i), g)]++;
  int32_t run = 0;
  for (int64_t c = 0; c <= ncell; c++) { int32_t n = cell_idx[c]; cell_idx[c] = run; run += n; } /* :62-78 */
  int32_t *cur = (int32_t *)malloc(sizeof(int32_t) * (size_t)(ncell + 1));
  memcpy(cur, cell_idx, sizeof(int32_t) * (size_t)(ncell + 1));
  memcpy(pos, pre_pos, sizeof(float) * 4 * (size_t)nvert); /* :85-86 */
  for (int64_t i = 0; i < nbe; i++) {
    int64_t j = cur[cell_of(p, pre_pos + 4 * (nvert + i), g)]++;
    memcpy(pos + 4 * (nvert + j), pre_pos + 4 * (nvert + i), 16);
    memcpy(norm + 4 * j, pre_norm + 4 * i, 16);
    memcpy(ep + 4 * j, pre_ep + 4 * i, 16);
    surf[j] = pre_surf[i];
  }
  free(cur);
}

/* ------------------------------------------------------------------------------------------------ periodicity
 * src/crixus_d.cu:213-237 find_links.  Returns the number of max-face vertices left without a partner
 * (the reference silently abandons the thread there, :229-232).  powf(x,2.f) restated as x*x. */
int64_t orc_find_links(const orc_params *p, float *pos, int64_t nvert, int32_t *newlink, int idim)
{
  int64_t missing = 0;
  const double thr = 1e-4 * p->dr;
  const int d1 = (idim + 1) % 3, d2 = (idim + 2) % 3;
  for (int64_t i = 0; i < nvert; i++) {
    if (!((double)fabsf(pos[4 * i + idim] - p->dmax[idim]) < thr)) continue;
    int64_t hit = -1;
    for (int64_t j = 0; j < nvert; j++) {
      if (j == i) continue;
      const float a = pos[4 * i + d1] - pos[4 * j + d1], b = pos[4 * i + d2] - pos[4 * j + d2],
                  c = pos[4 * j + idim] - p->dmin[idim];
      if ((double)sqrtf(a * a + b * b + c * c) < thr) { hit = j; break; }
    }
    if (hit < 0) { missing++; continue; }
    newlink[i] = (int32_t)hit;
    for (int k = 0; k < 3; k++) pos[4 * i + k] = -1e10f; /* :225-226 */
  }
  return missing;
}
/* src/crixus_d.cu:240-256 */
void orc_periodicity_links(int32_t *ep, int64_t nbe, const int32_t *newlink)
{
  for (int64_t i = 0; i < nbe; i++)
    for (int j = 0; j < 3; j++)
      if (newlink[ep[4 * i + j]] != -1) ep[4 * i + j] = newlink[ep[4 * i + j]];
}

/* ------------------------------------------------------------------------------------------------ special boundaries
 * identifySpecialBoundarySegments + segInTri, src/crixus_d.cu:1060-1112, with the multiply-add fusion of the reference's
 * sm_100 SASS: cross products a*b-c*d -> fma(a,b,-(c*d)); every "dot += x*y" chain -> fma chain from 0;
 * 1.0/det -> correctly rounded float reciprocal; eps = the fill eps (crixus.cu:464-467 runs before the loop :469-642).
 * sbid: int[nvert+nbe] (vertices, then segments); a segment whose barycentre lies in a special-boundary triangle gets sbi,
 * and so do its three vertices. */
static int orc_seg_in_tri(const float vb[3][3], const float spos[3], const float norm[4], float eps)
{
  float pa[3], ba[3], ca[3];
  for (int i = 0; i < 3; i++) { ba[i] = vb[1][i] - vb[0][i]; ca[i] = vb[2][i] - vb[0][i]; pa[i] = spos[i] - vb[0][i]; }
  float d = fmaf(norm[2], pa[2], fmaf(norm[1], pa[1], fmaf(norm[0], pa[0], 0.0f)));
  if (fabsf(d) > eps * norm[3]) return 0;                                     /* :1065 */
  const float dot00 = fmaf(ba[2], ba[2], fmaf(ba[1], ba[1], fmaf(ba[0], ba[0], 0.0f)));
  const float dot01 = fmaf(ca[2], ba[2], fmaf(ba[1], ca[1], fmaf(ca[0], ba[0], 0.0f)));
  const float dot02 = fmaf(ba[2], pa[2], fmaf(ba[1], pa[1], fmaf(ba[0], pa[0], 0.0f)));
  const float dot11 = fmaf(ca[2], ca[2], fmaf(ca[1], ca[1], fmaf(ca[0], ca[0], 0.0f)));
  const float dot12 = fmaf(ca[2], pa[2], fmaf(ca[1], pa[1], fmaf(ca[0], pa[0], 0.0f)));
  const float det = fmaf(dot00, dot11, -(dot01 * dot01));
  const float invdet = 1.0f / det;                                            /* :1078 */
  const float u = fmaf(dot02, dot11, -(dot01 * dot12)) * invdet;
  const float v = fmaf(dot00, dot12, -(dot01 * dot02)) * invdet;
  if (u < -eps || v < -eps || u + v > 1 + eps) return 0;                      /* :1081 */
  return 1;
}

void orc_identify_special_boundary(const orc_params *p, const float *pos, const int32_t *ep, int64_t nvert, int64_t nbe,
                                   const float *sbpos, const int32_t *sbep, int64_t sbnbe, int32_t *sbid, int32_t sbi)
{
  const float eps = p->eps;
#pragma omp parallel for schedule(static)
  for (int64_t id = 0; id < nbe; id++) {
    const float *spos = pos + 4 * (nvert + id);
    for (int64_t i = 0; i < sbnbe; i++) {
      float vb[3][3], norm[4];
      for (int j = 0; j < 3; j++)
        for (int k = 0; k < 3; k++) vb[j][k] = sbpos[4 * (int64_t)sbep[4 * i + j] + k];
      norm[3] = 0.0f;
      for (int j = 0; j < 3; j++) {                                           /* :1096-1100 */
        const int a = (j + 1) % 3, b = (j + 2) % 3;
        norm[j] = fmaf(vb[1][a] - vb[0][a], vb[2][b] - vb[0][b], -((vb[1][b] - vb[0][b]) * (vb[2][a] - vb[0][a])));
        norm[3] += norm[j];
      }
      norm[3] = fabsf(norm[3]);
      if (orc_seg_in_tri(vb, spos, norm, eps)) {
        sbid[nvert + id] = sbi;
        for (int j = 0; j < 3; j++) {
#pragma omp atomic write
          sbid[ep[4 * id + j]] = sbi;
        }
        break;
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------------ mesh precondition
 * resources/test-triangle-size.c:24,38-58: every vertex-barycentre distance must be <= dr - 1e-4*dr.
 * corners: ntri*9 floats (STL order).  out[0] = min distance, out[1] = max distance; returns the number of failures. */
int64_t orc_triangle_size(const float *corners, int64_t ntri, float dr, float *out)
{
  const float eps = dr * 1e-4;                    /* :24  (double product narrowed to float) */
  float mind = 1e10, maxd = 0e10;                 /* :33-34 */
  int64_t failed = 0;
  for (int64_t i = 0; i < ntri; i++) {
    const float *v = corners + 9 * i;
    float s[3] = {0.0, 0.0, 0.0};
    for (int j = 0; j < 3; j++)
      for (int k = 0; k < 3; k++) s[k] += v[3 * j + k] / 3.0;      /* :43-46 double division, float accumulator */
    for (int j = 0; j < 3; j++) {
      const float dist = sqrt(pow(s[0] - v[3 * j + 0], 2) + pow(s[1] - v[3 * j + 1], 2) + pow(s[2] - v[3 * j + 2], 2.0));  /* :48 */
      if (dist > maxd) maxd = dist;
      if (dist < mind) mind = dist;
      if (dist > dr - eps) failed++;              /* :53 */
    }
  }
  out[0] = mind; out[1] = maxd;
  return failed;
}

/* ------------------------------------------------------------------------------------------------ trisize
 * src/crixus_d.cu:258-267 */
void orc_calc_trisize(const int32_t *ep, int32_t *trisize, int64_t nbe)
{
  for (int64_t i = 0; i < nbe; i++)
    for (int j = 0; j < 3; j++) trisize[ep[4 * i + j]]++;
}

/* ------------------------------------------------------------------------------------------------ neighbour cells
 * the 27-cell walk shared by calc_vert_volume (:323-347) and checkCollision (:824-848); returns cell id or -1 */
static inline int64_t nb_cell(const orc_params *p, const int gi[3], int a, int b, int c)
{
  int g[3] = {gi[0] + a, gi[1] + b, gi[2] + c};
  for (int j = 0; j < 3; j++) {
    if (g[j] == -1 || g[j] == p->gridn[j]) {
      if (p->per[j]) g[j] = (g[j] == -1) ? p->gridn[j] - 1 : 0;
      else return -1;
    }
  }
  return (int64_t)g[0] * p->gridn[1] * p->gridn[2] + (int64_t)g[1] * p->gridn[2] + g[2];
}

/* periodic unwrap of a vertex-to-vertex offset, src/crixus_d.cu:513 */
static inline float unwrap(const orc_params *p, float c, int n)
{
  if (p->per[n] && fabsf(c) > 2 * p->dr) c = fmaf(sgnf_(c), (-p->dmax[n] + p->dmin[n]), c);
  return c;
}

/* ------------------------------------------------------------------------------------------------ calc_vert_volume
 * src/crixus_d.cu:271-678, one vertex.  Returns 0, or 1 if trisize > trimax (reference: thread returns, :306). */
static int vert_volume_one(const orc_params *p, const float *pos, const float *norm, const int32_t *ep, float *vol,
                           const int32_t *trisize, const int32_t *cell_idx, int64_t i, int zeroOpen)
{
  const unsigned gsize = GRES * 2 + 1;
  const float gdr = (float)(p->dr / 2.0 / (float)GRES); /* :281 */
  const float eps = p->eps, dr = p->dr;
  const double cube_thr = dr / 2. + eps; /* :535, double */
  float cvec[TRIMAX][12][3];
  int tri[TRIMAX][3];
  float avnorm[3] = {0.f, 0.f, 0.f};
  char first[TRIMAX];
  float edgen[TRIMAX][4];
  int closed = 1;
  float volacc = 0.f;
  const float *pi = pos + 4 * i;

  const unsigned tris = (unsigned)trisize[i];
  if (tris > TRIMAX) return 1;
  memset(tri, 0, sizeof(tri));
  for (unsigned j = 0; j < tris; j++) {
    first[j] = 1;
    for (int k = 0; k < 4; k++) edgen[j][k] = 0.f;
  }
  /* fan gather :316-369 */
  unsigned itris = 0;
  int gridi[3];
  cell_of(p, pi, gridi);
  for (int gi = -1; gi <= 1; gi++)
    for (int gj = -1; gj <= 1; gj++)
      for (int gk = -1; gk <= 1; gk++) {
        const int64_t c = nb_cell(p, gridi, gi, gj, gk);
        if (c < 0) continue;
        for (int64_t j = cell_idx[c]; j < cell_idx[c + 1]; j++)
          for (int k = 0; k < 3; k++)
            if (ep[4 * j + k] == i) {
              if (itris >= TRIMAX) return 1;
              tri[itris][0] = ep[4 * j + (k + 1) % 3];
              tri[itris][1] = ep[4 * j + (k + 2) % 3];
              tri[itris][2] = (int)j;
              itris++;
            }
      }
  /* chain ordering :371-399 */
  for (unsigned j = 0; j < tris; j++) {
    for (unsigned k = j + 1; k < tris; k++) {
      if (tri[j][1] == tri[k][0]) {
        if (k != j + 1)
          for (int l = 0; l < 3; l++) { int t = tri[j + 1][l]; tri[j + 1][l] = tri[k][l]; tri[k][l] = t; }
        break;
      }
      if (tri[j][1] == tri[k][1]) {
        int id[3] = {tri[k][1], tri[k][0], tri[k][2]};
        for (int l = 0; l < 3; l++) { tri[k][l] = tri[j + 1][l]; tri[j + 1][l] = id[l]; }
        break;
      }
      if (k == tris - 1) closed = 0;
    }
  }
  if (tris > 0 && tri[0][0] != tri[tris - 1][1]) closed = 0;
  if (zeroOpen && !closed) { vol[i] = 0.0f; return 0; } /* :403-407 */

  /* edge normals :409-473 */
  itris = 0;
  for (int gi = -1; gi <= 1; gi++)
    for (int gj = -1; gj <= 1; gj++)
      for (int gk = -1; gk <= 1; gk++) {
        const int64_t c = nb_cell(p, gridi, gi, gj, gk);
        if (c < 0) continue;
        for (int64_t j = cell_idx[c]; j < cell_idx[c + 1]; j++) {
          for (unsigned k = 0; k < tris; k++) {
            if ((int)(edgen[k][3] + eps) == 2) continue;
            int vfound = 0;
            for (int l = 0; l < 3; l++)
              if (ep[4 * j + l] == tri[k][0] || ep[4 * j + l] == tri[k][1]) vfound++;
            if (vfound == 2) {
              for (int l = 0; l < 3; l++) edgen[k][l] += norm[4 * j + l];
              edgen[k][3] = (float)(edgen[k][3] + 1.);
            }
            if ((int)(edgen[k][3] + eps) == 2) {
              float edge[3], en[3];
              for (int n = 0; n < 3; n++) {
                edge[n] = pos[4 * (int64_t)tri[k][0] + n] - pos[4 * (int64_t)tri[k][1] + n];
                edge[n] = unwrap(p, edge[n], n); /* :456 */
              }
              cross_r2(edgen[k], edge, en);
              edgen[k][0] = en[0]; edgen[k][1] = en[1]; edgen[k][2] = en[2]; edgen[k][3] = 0.f;
              itris++;
            }
          }
          if (itris == tris) break;
        }
      }
  for (unsigned k = 0; k < tris; k++)
    if ((int)(edgen[k][3] + eps) == 1) { /* :465-473, no periodic unwrap */
      float edge[3], en[3];
      for (int n = 0; n < 3; n++) edge[n] = pos[4 * (int64_t)tri[k][0] + n] - pos[4 * (int64_t)tri[k][1] + n];
      cross_r2(edgen[k], edge, en);
      edgen[k][0] = en[0]; edgen[k][1] = en[1]; edgen[k][2] = en[2]; edgen[k][3] = 0.f;
    }

  int cubesInitialized = 0;
  for (unsigned k = 0; k < gsize; k++)
    for (unsigned l = 0; l < gsize; l++)
      for (unsigned m = 0; m < gsize; m++) {
        float gp[3];
        gp[0] = (((float)k - (float)(gsize - 1) / 2)) * gdr; /* :493-495 */
        gp[1] = (((float)l - (float)(gsize - 1) / 2)) * gdr;
        gp[2] = (((float)m - (float)(gsize - 1) / 2)) * gdr;
        float vgrid = 0.f;
        /* cubes :506-544 */
        for (unsigned j = 0; j < tris; j++) {
          if (!cubesInitialized) {
            const float *nj = norm + 4 * (int64_t)tri[j][2];
            for (int n = 0; n < 3; n++) cvec[j][2][n] = nj[n];
            float vnorm = 0.f;
            for (int n = 0; n < 3; n++) {
              cvec[j][0][n] = unwrap(p, pos[4 * (int64_t)tri[j][0] + n] - pi[n], n);
              vnorm = (float)((double)vnorm + (double)cvec[j][0][n] * (double)cvec[j][0][n]); /* :514 */
            }
            vnorm = sqrtf(vnorm);
            for (int n = 0; n < 3; n++) cvec[j][0][n] /= vnorm;
            for (int n = 0; n < 3; n++)
              cvec[j][1][n] = det_r2(cvec[j][0][(n + 1) % 3], cvec[j][2][(n + 2) % 3], cvec[j][0][(n + 2) % 3],
                                     cvec[j][2][(n + 1) % 3]); /* :518 */
            vnorm = 0.f;
            for (int n = 0; n < 3; n++) {
              cvec[j][3][n] = unwrap(p, pos[4 * (int64_t)tri[j][1] + n] - pi[n], n);
              vnorm = (float)((double)vnorm + (double)cvec[j][3][n] * (double)cvec[j][3][n]);
              avnorm[n] -= nj[n]; /* :524 */
            }
            vnorm = sqrtf(vnorm);
            for (int n = 0; n < 3; n++) cvec[j][3][n] /= vnorm;
            for (int n = 0; n < 3; n++)
              cvec[j][4][n] = det_r2(cvec[j][3][(n + 1) % 3], cvec[j][2][(n + 2) % 3], cvec[j][3][(n + 2) % 3],
                                     cvec[j][2][(n + 1) % 3]); /* :528 */
          }
          int incube[5];
          for (int n = 0; n < 5; n++) incube[n] = ((double)fabsf(dot_r1(gp, cvec[j][n])) <= cube_thr); /* :532-536 */
          if ((incube[0] && incube[1] && incube[2]) || (incube[2] && incube[3] && incube[4])) {
            vgrid = 1.f;
            if (cubesInitialized) break;
          }
        }
        cubesInitialized = 1;

        /* planes :549-666 */
        float tvec[3][3];
        for (unsigned j = 0; j < tris; j++) {
          if (vgrid < eps) break;
          if (first[j]) {
            first[j] = 0;
            for (int n = 0; n < 3; n++) {
              cvec[j][5][n] = unwrap(p, pos[4 * (int64_t)tri[j][0] + n] - pi[n], n);
              cvec[j][6][n] = (float)((double)pi[n] + (double)cvec[j][5][n] / 2.); /* :558 */
              tvec[0][n] = cvec[j][5][n];
              tvec[1][n] = unwrap(p, pos[4 * (int64_t)tri[j][1] + n] - pi[n], n);
              if (!closed) {
                cvec[j][7][n] = tvec[1][n];
                cvec[j][8][n] = (float)((double)pi[n] + (double)cvec[j][7][n] / 2.);
              }
              tvec[2][n] = avnorm[n];
            }
            for (int n = 0; n < 3; n++)
              for (int kk = 0; kk < 3; kk++)
                cvec[j][kk + 9][n] = det_r2(tvec[kk][(n + 1) % 3], tvec[(kk + 1) % 3][(n + 2) % 3],
                                            tvec[kk][(n + 2) % 3], tvec[(kk + 1) % 3][(n + 1) % 3]); /* :570 */
            float sp = dot_r1(norm + 4 * (int64_t)tri[j][2], cvec[j][9]);
            if (sp > 0.f)
              for (int kk = 0; kk < 3; kk++)
                for (int n = 0; n < 3; n++) cvec[j][kk + 9][n] = -cvec[j][kk + 9][n];
            sp = dot_r1(edgen[j], cvec[j][5]);
            if (sp < 0.f)
              for (int n = 0; n < 3; n++) edgen[j][n] = -edgen[j][n];
          }
          float sp;
          /* voronoi plane :600-612 */
          for (int n = 0; n < 3; n++) tvec[0][n] = gp[n] + pi[n] - cvec[j][6][n];
          sp = dot_r1(tvec[0], cvec[j][5]);
          if (sp > eps) { vgrid = 0.f; break; }
          else if (fabsf(sp) < eps) vgrid /= 2.f;
          /* voronoi plane 2 :614-625 */
          if (!closed) {
            for (int n = 0; n < 3; n++) tvec[0][n] = gp[n] + pi[n] - cvec[j][8][n];
            sp = dot_r1(tvec[0], cvec[j][7]);
            if (sp > eps) { vgrid = 0.f; break; }
            else if (fabsf(sp) < eps) vgrid /= 2.f;
          }
          /* walls :627-643 */
          int half = 0;
          for (int o = 0; o < 3; o++) {
            sp = dot_r1(gp, cvec[j][9 + o]);
            if (sp < -eps) break;
            if (fabsf(sp) < eps && o == 0) half = 1;
            if (o == 2 && !half) { vgrid = 0.f; break; }
            else if (o == 2 && half) vgrid /= 2.f;
          }
          /* edges :645-661 */
          for (int n = 0; n < 3; n++) tvec[0][n] = gp[n] - cvec[j][5][n];
          sp = dot_r1(tvec[0], edgen[j]);
          if (sp > eps) { vgrid = 0.f; break; }
          else if (fabsf(sp) < eps) vgrid /= 2.f;
          if (vgrid < eps) break;
          if (j == tris - 1) volacc += vgrid; /* :665 */
        }
      }
  const double gd = p->dr / 2.0 / (float)GRES;
  vol[i] = (float)((double)volacc * (gd * gd * gd)); /* :674 pow(double,3) */
  return 0;
}

/* all vertices (or the index range [v0,v1) for sampled timing).  Returns #vertices with trisize>trimax. */
int64_t orc_calc_vert_volume(const orc_params *p, const float *pos, const float *norm, const int32_t *ep, float *vol,
                             const int32_t *trisize, const int32_t *cell_idx, int64_t v0, int64_t v1, int zeroOpen)
{
  int64_t bad = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : bad)
  for (int64_t i = v0; i < v1; i++) {
    if (pos[4 * i] < -1e9f) continue; /* deleted vertex :297 */
    vol[i] = 0.f;
    bad += vert_volume_one(p, pos, norm, ep, vol, trisize, cell_idx, i, zeroOpen);
  }
  return bad;
}

/* ------------------------------------------------------------------------------------------------ fill (box)
 * src/crixus_d.cu:680-744.  Returns nfi = number of in-box lattice cells (counted even if already set, :713-715). */
int64_t orc_fill_box(const orc_params *p, uint32_t *fpos, float xmin, float xmax, float ymin, float ymax, float zmin,
                     float zmax)
{
  const float eps = p->eps, dr = p->dr;
  const float lo[3] = {xmin, ymin, zmin}, hi[3] = {xmax, ymax, zmax};
  int64_t dimg[3];
  orc_fill_dims(p, dimg);
  uint32_t dim[3], mn[3];
  for (int j = 0; j < 3; j++) {
    dim[j] = (uint32_t)(floorf((hi[j] + eps - lo[j]) / dr) + 1); /* :686-688 */
    mn[j] = (uint32_t)round(((float)(uint32_t)dimg[j] - 1.) * (lo[j] - p->fcmin[j]) / (p->fcmax[j] - p->fcmin[j]));
  }
  int64_t nfi = 0;
  for (int64_t k = 0; k < dimg[2]; k++)
    for (int64_t j = 0; j < dimg[1]; j++)
      for (int64_t i = 0; i < dimg[0]; i++) {
        const uint32_t ui = (uint32_t)i, uj = (uint32_t)j, uk = (uint32_t)k;
        if ((ui >= mn[0] && uj >= mn[1] && uk >= mn[2]) && (ui - mn[0] < dim[0] && uj - mn[1] < dim[1] && uk - mn[2] < dim[2])) {
          const int64_t b = i + j * dimg[0] + k * dimg[0] * dimg[1];
          fpos[b >> 5] |= 1u << (b & 31);
          nfi++;
        }
      }
  return nfi;
}

/* ------------------------------------------------------------------------------------------------ fill (geometry)
 * checkTriangleCollision, src/crixus_d.cu:753-799 (softSurfer segment/triangle test with eps tolerances) */
static int tri_collision(const orc_params *p, const float *s, const float *dir, const float *n, const float *v0,
                         const float *v1, const float *v2)
{
  const float eps = p->eps;
  float u[3], v[3], w0[3], w[3];
  for (int j = 0; j < 3; j++) { u[j] = v1[j] - v0[j]; v[j] = v2[j] - v0[j]; w0[j] = s[j] - v0[j]; }
  const float a = -dot_r3(n, w0), b = dot_r3(n, dir);
  if (fabsf(b) < eps) return fabsf(a) < eps; /* :766-769 in-plane: infinite plane collides */
  const float r = a / b;
  if (r < -eps || (double)r > 1.0 + (double)eps) return 0;
  for (int j = 0; j < 3; j++) w[j] = fmaf(r, dir[j], s[j]) - v0[j]; /* :778 */
  const float uu = dot_r3(u, u), uv = dot_r3(u, v), vv = dot_r3(v, v), wu = dot_r3(w, u), wv = dot_r3(w, v);
  const float D = det_r2(uv, uv, uu, vv);
  const float t1 = det_r2(uv, wv, vv, wu) / D;
  if (t1 < -eps || (double)t1 > 1.0 + (double)eps) return 0;
  const float t2 = det_r2(uv, wu, uu, wv) / D;
  if (t2 < -eps || (double)(t1 + t2) > 1.0 + (double)eps) return 0;
  return 1;
}

/* checkCollision, src/crixus_d.cu:801-963.  nbe counts wall segments + fshape segments; the last cnbe are fshape. */
static int check_collision(const orc_params *p, int64_t si, int64_t sj, int64_t sk, int64_t ei, int64_t ej, int64_t ek,
                           const float *norm, const int32_t *ep, const float *pos, int64_t nbe, int64_t cnbe,
                           const int32_t *cell_idx)
{
  const float eps = p->eps, dr = p->dr, dr_wall = p->dr_wall;
  float s[3], e[3], dir[3], hd[3], mid[3];
  s[0] = fmaf((float)si, dr, p->fcmin[0]); s[1] = fmaf((float)sj, dr, p->fcmin[1]); s[2] = fmaf((float)sk, dr, p->fcmin[2]);
  e[0] = fmaf((float)ei, dr, p->fcmin[0]); e[1] = fmaf((float)ej, dr, p->fcmin[1]); e[2] = fmaf((float)ek, dr, p->fcmin[2]);
  for (int i = 0; i < 3; i++) { dir[i] = e[i] - s[i]; hd[i] = (s[i] - e[i]) * 0.5f; mid[i] = s[i] + e[i]; }
  float infl = fmaf(dr, 3.0f, sqrtf(dot_r3(hd, hd))) + dr_wall; /* :815 */
  infl *= infl;
  int gridi[3];
  cell_of(p, s, gridi); /* :821 */
  for (int gi = -1; gi <= 1; gi++)
    for (int gj = -1; gj <= 1; gj++)
      for (int gk = -1; gk <= 1; gk++) {
        const int64_t c = nb_cell(p, gridi, gi, gj, gk);
        if (c < 0) continue;
        for (int64_t i = cell_idx[c]; i < cell_idx[c + 1]; i++) {
          const float *n = norm + 4 * i;
          const float *v[3] = {pos + 4 * (int64_t)ep[4 * i], pos + 4 * (int64_t)ep[4 * i + 1], pos + 4 * (int64_t)ep[4 * i + 2]};
          float t[3];
          for (int j = 0; j < 3; j++) t[j] = fmaf(mid[j], 0.5f, -v[0][j]);
          float dist = dot_r3(t, t); /* :855 */
          if (dist > infl) continue;
          float segpos[3], nr[3][3], nd[3];
          for (int j = 0; j < 3; j++) segpos[j] = ((v[0][j] + v[1][j]) + v[2][j]) / 3.0f; /* :863 */
          for (int k = 0; k < 3; k++) {
            const float *a = v[k], *b = v[(k + 1) % 3];
            float ed[3], sv[3];
            for (int j = 0; j < 3; j++) ed[j] = b[j] - a[j];
            const float len = sqrtf(dot_r3(ed, ed));
            for (int j = 0; j < 3; j++) { ed[j] = ed[j] / len; sv[j] = segpos[j] - a[j]; }
            nd[k] = dot_r3(ed, sv);                                     /* :879-881 */
            for (int j = 0; j < 3; j++) nr[k][j] = fmaf(-ed[j], nd[k], sv[j]); /* :884-886 */
          }
          float pp[3], sm[3];
          for (int j = 0; j < 3; j++) sm[j] = s[j] - segpos[j];
          const float pd = dot_r3(sm, n); /* :893 */
          for (int j = 0; j < 3; j++) pp[j] = fmaf(-n[j], pd, sm[j]); /* :895 */
          float orient[3];
          for (int k = 0; k < 3; k++) {
            float q[3];
            for (int j = 0; j < 3; j++) q[j] = (pp[j] + segpos[j]) - v[k][j];
            orient[k] = dot_r3(q, nr[k]); /* :899-901 */
          }
          dist = 0.0f;
          if (orient[0] > -eps && orient[1] > -eps && orient[2] > -eps) {
            for (int j = 0; j < 3; j++) t[j] = pp[j] - sm[j];
            dist = dot_r3(t, t);
          } else if (orient[0] < eps && orient[1] < eps) {
            for (int j = 0; j < 3; j++) t[j] = s[j] - v[2][j];
            dist = dot_r3(t, t);
          } else if (orient[1] < eps && orient[2] < eps) {
            for (int j = 0; j < 3; j++) t[j] = s[j] - v[0][j];
            dist = dot_r3(t, t);
          } else if (orient[2] < eps && orient[0] < eps) {
            for (int j = 0; j < 3; j++) t[j] = s[j] - v[1][j];
            dist = dot_r3(t, t);
          } else {
            const float *x1, *x2;
            if (orient[0] < eps) { x1 = v[0]; x2 = v[1]; }
            else if (orient[1] < eps) { x1 = v[1]; x2 = v[2]; }
            else { x1 = v[2]; x2 = v[0]; }
            float x21[3], sx1[3];
            for (int j = 0; j < 3; j++) { x21[j] = x2[j] - x1[j]; sx1[j] = s[j] - x1[j]; }
            const float tdot = dot_r3(x21, sx1), l2 = dot_r3(x21, x21);
            for (int j = 0; j < 3; j++) t[j] = fmaf(-(x21[j] / l2), tdot, sx1[j]); /* :938 */
            dist = dot_r3(t, t);
          }
          if (dist + eps < dr_wall * dr_wall) return 1; /* :940 */
          if (tri_collision(p, s, dir, n, v[0], v[1], v[2])) return 1;
        }
      }
  for (int64_t i = nbe - cnbe; i < nbe; i++) { /* fshape: all of them, :951-960 */
    const float *n = norm + 4 * i;
    if (tri_collision(p, s, dir, n, pos + 4 * (int64_t)ep[4 * i], pos + 4 * (int64_t)ep[4 * i + 1], pos + 4 * (int64_t)ep[4 * i + 2]))
      return 1;
  }
  return 0;
}

static inline int getbit(const uint32_t *f, int64_t b) { return (f[b >> 5] >> (b & 31)) & 1u; }

/* fill_fluid_complex driven to its fixed point (src/crixus_d.cu:965-1058, loop src/crixus.cu:934-947).
 * The _slab variant sweeps only the lattice planes [k0,k1) (cells outside are read-only): the unit of the z-slab
 * sharded multi-GPU fill.  Sweeps in cell-index order on one thread (a legal schedule of the reference: its result is the unique least
 * fixed point).  seed < 0: no seed.  Returns the number of newly set cells; *sweeps_out = sweeps. */
int64_t orc_fill_geometry_slab(const orc_params *p, uint32_t *fpos, const float *norm, const int32_t *ep, const float *pos,
                               int64_t nbe, int64_t cnbe, const int32_t *cell_idx, int64_t seed, int64_t k0, int64_t k1,
                               int *sweeps_out)
{
  int64_t dimg[3], total = 0;
  orc_fill_dims(p, dimg);
  const int64_t plane = dimg[0] * dimg[1];
  int it = 0;
  int64_t nfi;
  do {
    it++;
    nfi = 0;
    if (it == 1 && seed >= 0 && seed / plane >= k0 && seed / plane < k1 && !getbit(fpos, seed)) { /* :972-980 */
      fpos[seed >> 5] |= 1u << (seed & 31);
      nfi++;
    }
    for (int64_t b = k0 * plane; b < k1 * plane; b++) {
      if (getbit(fpos, b)) continue;
      const int64_t k = b / plane, tmp = b % plane, j = tmp / dimg[0], i = tmp % dimg[0];
      for (int ij = 0; ij < 6; ij++) { /* :1005-1028, order -x +x -y +y -z +z */
        int64_t io = i, jo = j, ko = k;
        if (ij <= 1) io += (ij == 0) ? -1 : 1;
        else if (ij <= 3) jo += (ij == 2) ? -1 : 1;
        else ko += (ij == 4) ? -1 : 1;
        if (io < 0 || io >= dimg[0] || jo < 0 || jo >= dimg[1] || ko < 0 || ko >= dimg[2]) continue;
        const int64_t o = io + jo * dimg[0] + ko * plane;
        if (getbit(fpos, o) && !check_collision(p, i, j, k, io, jo, ko, norm, ep, pos, nbe, cnbe, cell_idx)) {
          fpos[b >> 5] |= 1u << (b & 31);
          nfi++;
          break;
        }
      }
    }
    total += nfi;
  } while (nfi > 0 && it < MAX_ITERATIONS);
  if (sweeps_out) *sweeps_out = it;
  return total;
}

/* whole lattice: the slab [0, kdim) */
int64_t orc_fill_geometry(const orc_params *p, uint32_t *fpos, const float *norm, const int32_t *ep, const float *pos,
                          int64_t nbe, int64_t cnbe, const int32_t *cell_idx, int64_t seed, int *sweeps_out)
{
  int64_t dimg[3];
  orc_fill_dims(p, dimg);
  return orc_fill_geometry_slab(p, fpos, norm, ep, pos, nbe, cnbe, cell_idx, seed, 0, dimg[2], sweeps_out);
}

/* one directed test, exported so tests can probe single (s,e) pairs against the shim build */
int orc_check_collision(const orc_params *p, const int64_t s[3], const int64_t e[3], const float *norm, const int32_t *ep,
                        const float *pos, int64_t nbe, int64_t cnbe, const int32_t *cell_idx)
{
  return check_collision(p, s[0], s[1], s[2], e[0], e[1], e[2], norm, ep, pos, nbe, cnbe, cell_idx);
}

int orc_num_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// crixus_b200 -- calc_vert_volume: vertex-particle volume by the reference's 41^3-voxel quadrature.
// Reference: /root/reference/src/crixus_d.cu:271-678 (one thread per vertex, 8.6 KB of local memory per thread).
//
// B200 mapping (DESIGN.md "calc_vert_volume"):
//   * one WARP per vertex, vertices handed out by an atomic work counter (persistent CTAs, 4 warps each);
//   * phase A (warp-cooperative, once per vertex): fan gather over the 27 neighbouring 3dr-cells with lanes striding
//     the cell-sorted segments (ballot/scan compaction keeps the reference's discovery order), chain ordering,
//     edge-normal pairing, and the 12 plane vectors per fan triangle -> 160 B per triangle in shared memory;
//   * phase B: the voxel cube is cut into 41x41 LINES of 41 voxels and the lines are dealt to the lanes.  Every plane
//     expression of the reference, sp = fma(gz,cz, fma(gy,cy, fma(gx,cx,0))) (R1), is a chain of correctly rounded, hence
//     monotone, operations of each of its three grid coordinates: along a line a threshold predicate flips once, and the flip
//     is found by a six-step search that evaluates the SAME arithmetic (exact).  Per-voxel state lives in 41-bit masks
//     (hit / alive / bit-sliced halving counters).
//   * the direction of the lines is chosen PER VERTEX (template parameter AX of the kernel, one launch per direction over a
//     work list built by a classification pass): along the dominant axis of the vertex normal.  The voronoi, edge and wedge-side
//     planes of a fan all contain (nearly) the normal direction, so lines parallel to it are removed whole by the row pass or lie
//     inside ONE wall wedge, instead of crossing every sector of the fan: on an axis-aligned wall 2-3x fewer plane searches.
//     Any direction gives the same bits; the choice only changes the amount of work.
//   * the result is summed in integers (weights are 2^-k) so it is independent of the order of summation.
// Bound: FP32 ALU (non-tensor).  Algorithmic flops F = 68921 * T * 56 per vertex.
#include <stdlib.h>
#include "common.cuh"

#define VV_WARPS 4
#define VV_GS 41      // gres*2+1 voxels per axis (crixus_d.cu:280)
#define VV_LINES (VV_GS * VV_GS)
#define VV_PLANE_F4 10
#define VV_LSTRIDE 1688  // entries of the per-vertex line list (>= 41 * 41, 16-byte multiple)

#define VVE_TRIMAX 1
#define VVE_FAN 2
#define VVE_NONMANIFOLD 4

// device-side control block of one call: work counters (0: k_vv_fans, 1..3: k_vv_rows<x,y,z>, 4..6: k_vv_lines<x,y,z>, 7: the
// monolithic kernel), list lengths (0..2 = line direction x, y, z; 3 = fans for the monolithic kernel), bump pointer of the plane
// store, error bits.  Everything except count[3] and err is reset for every chunk of vertices.
struct VVCtl {
  unsigned long long work[8];
  int count[4];
  int bump;
  int err;
  int pad[2];
};

struct VVArgs {
  cx_params p;
  const float4 *pos;
  const float4 *norm;
  const int4 *ep;
  float *vol;
  const int32_t *trisize;
  const int32_t *cell_idx;
  int64_t v_begin, v_end;
  int64_t chunk, nown;       // interleaved sharding: this call owns chunks part, part+nparts, ... of `chunk` vertices;
  int nparts, part;          // nown = number of owned vertices in [v_begin, v_end)
  int zero_open;
  float thr_cube;  // largest float <= dr/2.+eps (double), so |sp| <= thr_cube  <=>  (double)|sp| <= dr/2.+eps
  float eps;
  float two_dr;    // 2*dr (crixus_d.cu:513)
  double gd3;      // pow(dr/2.0/(float)gres, 3) (crixus_d.cu:674)
  int cmax;        // largest k with 2^-k >= eps (vgrid < eps -> voxel dropped, crixus_d.cu:551,662)
  VVCtl *ctl;
  int *list;                 // 4 work lists of nown entries each (slot s at list + s*nown): 0..2 hold chunk slots, 3 vertices
  int64_t c_begin, c_end;    // the chunk of owned vertices (indices into the owned set) the split kernels work on
  int4 *heads;               // per chunk slot: plane offset, fan size, closed, vertex
  float4 *planes;            // plane store of the chunk: 10 float4 per fan triangle
  int plane_cap;             // its capacity in fan triangles
  unsigned short *lines;     // per chunk slot: VV_LSTRIDE entries, the two lists of surviving lines (vv_rows)
  unsigned short *lines_big; // the same per warp of the monolithic kernel's grid
  int force_axis;            // CX_VV_AXIS: one line direction for every vertex (-1: by the fan normal)
  const int *adj_start;      // vertex -> (segment, corner) adjacency (k_vv_adj_*): entries adj[adj_start[v] .. adj_start[v+1])
  const int *adj;            // entry = segment * 4 + corner
  float gz[VV_GS];           // ((float)m - 20.f) * gdr   (crixus_d.cu:493-495)
};

// per-warp shared-memory working set; TCAP = max fan size it can hold.  Almost every vertex has <= 16 triangles, so the
// main launch uses TCAP=16 (3 KB per warp, occupancy limited by registers only); the few larger fans are deferred to
// a second launch with TCAP = trimax = 50.
template <int TCAP>
struct VVWarpT {
  float4 plane[TCAP][VV_PLANE_F4];
  int tri[TCAP][3];
  int ecount[TCAP];
  int eseg[TCAP][2];
  int closed;
};
#define VV_TCAP_SMALL 16

__device__ __forceinline__ float sgnf(float x) { return (float)((x > 0.f) - (x < 0.f)); }
// periodic unwrap of a vertex-to-vertex offset component (crixus_d.cu:513): c += sgn(c)*(-dmax+dmin)
__device__ __forceinline__ float unwrap1(const cx_params &p, float c, int n, float two_dr)
{
  if (p.per[n] && fabsf(c) > two_dr) c = ffma(sgnf(c), fadd(-p.dmax[n], p.dmin[n]), c);
  return c;
}
__device__ __forceinline__ f3 unwrap3(const cx_params &p, f3 c, float two_dr)
{
  return mk3(unwrap1(p, c.x, 0, two_dr), unwrap1(p, c.y, 1, two_dr), unwrap1(p, c.z, 2, two_dr));
}
// normalise like crixus_d.cu:510-517: vnorm accumulated in double from float squares, narrowed every step
__device__ __forceinline__ f3 normalise_ref(f3 c)
{
  float vn = (float)(0.0 + (double)c.x * (double)c.x);
  vn = (float)((double)vn + (double)c.y * (double)c.y);
  vn = (float)((double)vn + (double)c.z * (double)c.z);
  vn = __fsqrt_rn(vn);
  return mk3(__fdiv_rn(c.x, vn), __fdiv_rn(c.y, vn), __fdiv_rn(c.z, vn));
}
__device__ __forceinline__ float halfway(float pi, float c) { return (float)((double)pi + (double)c / 2.); }  // :558

__device__ __forceinline__ int warp_incl_scan_i(int v, int lane)
{
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// ---------------------------------------------------------------------------------------------------------------
// Phase B building blocks.  A z-line has 41 voxels; all per-voxel state is a 41-bit mask in a uint64_t.
#define VV_ALL 0x1FFFFFFFFFFull
#define VV_FULLWARP 0xffffffffu

// bit-sliced +1 on the 5-bit per-voxel halving counters (saturating at 31) for every voxel whose bit is set in `m`
__device__ __forceinline__ void half_inc(uint64_t (&h)[5], uint64_t m)
{
  uint64_t c = m;
#pragma unroll
  for (int b = 0; b < 5; b++) {
    const uint64_t t = h[b] & c;
    h[b] ^= c;
    c = t;
  }
#pragma unroll
  for (int b = 0; b < 5; b++) h[b] |= c;  // overflow -> stay at 31
}

// same, called by the converged warp: almost every voxel is halved at most once, so the carry out of bit 0 is empty
__device__ __forceinline__ void half_inc_warp(uint64_t (&h)[5], uint64_t m)
{
  uint64_t c = h[0] & m;
  h[0] ^= m;
  if (__any_sync(VV_FULLWARP, c != 0ull)) {
#pragma unroll
    for (int b = 1; b < 5; b++) {
      const uint64_t t = h[b] & c;
      h[b] ^= c;
      c = t;
    }
#pragma unroll
    for (int b = 0; b < 5; b++) h[b] |= c;
  }
}

__device__ __forceinline__ uint64_t lowmask(int b) { return (1ull << b) - 1ull; }   // b in [0, 63]
__device__ __forceinline__ uint64_t bit64(int b) { return 1ull << b; }

#ifdef VV_STATS
__device__ unsigned long long vv_stats[16];
#define VV_STAT(i, n) do { if ((threadIdx.x & 31) == 0) atomicAdd(&vv_stats[i], (unsigned long long)(n)); } while (0)
#else
#define VV_STAT(i, n) do { } while (0)
#endif

// Every plane expression of the reference is, along a z-line, a composition of correctly rounded (hence monotone)
// operations of gz, and gz[m] is non-decreasing in m.  A threshold predicate of such an expression is therefore a
// monotone step function of m: its truth values are pa..pa pb..pb with ONE flip.  mono_search finds the flip with six
// evaluations of the very same arithmetic (exact, no approximation): the largest s in [0,39] with pred(gz[s]) == pa,
// given pred(gz[0]) == pa and pred(gz[40]) != pa.  Lanes whose ends agree run along (the warp stays converged) and
// ignore the result.
#ifndef VV_EXP
#define VV_EXP 0
#endif
// resident blocks per SM the kernels are compiled for (register budget = 65536 / (128 * blocks))
#ifndef VV_LB_LINES
#define VV_LB_LINES 7
#endif
#ifndef VV_LB_FANS
#define VV_LB_FANS 10
#endif
#define VV_UNLIKELY(x) ((VV_EXP & 1) ? __builtin_expect(!!(x), 0) : (x))
template <class Pred>
__device__ __forceinline__ int mono_search(const float *gzt, bool pa, Pred pred)
{
#if (VV_EXP & 8)
  int s = 0;
#pragma unroll 1
  for (int st = 32; st >= 1; st >>= 1) {
    const int i = min(s + st, VV_GS - 1);
    if (pred(gzt[i]) == pa) s = i;
  }
  return min(s, VV_GS - 2);
#else
  int s = (pred(gzt[32]) == pa) ? 32 : 0;
  {
    const int i = min(s + 16, VV_GS - 1);
    if (pred(gzt[i]) == pa) s = i;
  }
#if (VV_EXP & 4)
#pragma unroll
#else
#pragma unroll 1   // rolled on purpose: ~40 inlined instances; the hot loops sit at the edge of the instruction cache (2 % faster)
#endif
  for (int st = 8; st >= 1; st >>= 1)
    if (pred(gzt[s + st]) == pa) s += st;   // never beyond 39 for a lane that is really cut; the table has 64 entries
  return s;
#endif
}

// mask of the voxels where pv(f(g[m])) holds.  fa, fb = f at the two ends.  lo = last index that agrees with end A
// (40 if the line is not cut by the predicate), pa = the predicate at end A.
template <class F, class PV>
__device__ __forceinline__ uint64_t mono_mask(const float *gzt, float fa, float fb, F f, PV pv, int &lo, bool &pa)
{
  pa = pv(fa);
  const bool pb = pv(fb);
  lo = VV_GS - 1;
  const unsigned cutm = __ballot_sync(VV_FULLWARP, pa != pb);
  if (cutm) {
    VV_STAT(2, 1); VV_STAT(3, __popc(cutm));
    const bool pa_ = pa;
    const int s = mono_search(gzt, pa_, [&](float g) { return pv(f(g)); });
    if (pa != pb) lo = s;
  }
  // lmt[1][b] = the b lowest voxels, lmt[0][b] = the others (gzt + 64 floats = two tables of 64 x uint64)
  const uint64_t *lmt = reinterpret_cast<const uint64_t *>(gzt + 64);
  return lmt[(pa ? 64 : 0) + lo + 1];
}

// ---------------------------------------------------------------------------------------------------------------
// One plane expression of the reference as a function of ONE grid coordinate (axis VAR), the two others fixed.
// The reference evaluates sp = fma(tz,cz, fma(ty,cy, fma(tx,cx, 0))) (R1) where the term inputs are
//   KIND 0 (voronoi planes, crixus_d.cu:604-611):  t = (g + pi) - off      g = grid offset, pi = vertex, off = point of the plane
//   KIND 1 (edge plane, :629-633):                 t = g - off
//   KIND 2 (wall planes :640-648, cubes :531-540):  t = g
// Whatever sits outside the varying term is hoisted (same operations, same order -> same bits):
//   VAR 2: pre = fma(ty,cy, fma(tx,cx,0));             f(g) = fma(t(g), cz, pre)
//   VAR 1: pre = fma(tx,cx,0), a2 = tz;                f(g) = fma(a2, cz, fma(t(g), cy, pre))
//   VAR 0: a1 = ty, a2 = tz;                           f(g) = fma(a2, cz, fma(a1, cy, fma(t(g), cx, 0)))
// Every step is a correctly rounded operation that is monotone in its varying argument, so f is a monotone function of g
// (rising or falling with the sign of the coefficient of axis VAR) for every VAR -- the premise of mono_search.
template <int D> __device__ __forceinline__ float cget(const f3 &v) { return D == 0 ? v.x : (D == 1 ? v.y : v.z); }

template <int VAR, int KIND>
struct PlaneF {
  float cv, add, off, pre, a1, k1, a2, k2;
  __device__ __forceinline__ float operator()(float g) const
  {
    const float t = KIND == 0 ? fsub(fadd(g, add), off) : (KIND == 1 ? fsub(g, off) : g);
    if (VAR == 2) return ffma(t, cv, pre);
    if (VAR == 1) return ffma(a2, k2, ffma(t, cv, pre));
    return ffma(a2, k2, ffma(a1, k1, ffma(t, cv, 0.f)));
  }
};
// c = plane vector, off = point of the plane (KIND 0, 1), pi = vertex (KIND 0), gp = grid offsets (component VAR is ignored)
template <int VAR, int KIND>
__device__ __forceinline__ PlaneF<VAR, KIND> make_plane(const f3 &c, const f3 &off, const f3 &pi, const f3 &gp)
{
  PlaneF<VAR, KIND> F;
  const float tx = KIND == 0 ? fsub(fadd(gp.x, pi.x), off.x) : (KIND == 1 ? fsub(gp.x, off.x) : gp.x);
  const float ty = KIND == 0 ? fsub(fadd(gp.y, pi.y), off.y) : (KIND == 1 ? fsub(gp.y, off.y) : gp.y);
  const float tz = KIND == 0 ? fsub(fadd(gp.z, pi.z), off.z) : (KIND == 1 ? fsub(gp.z, off.z) : gp.z);
  F.cv = cget<VAR>(c); F.add = cget<VAR>(pi); F.off = cget<VAR>(off);
  F.pre = 0.f; F.a1 = 0.f; F.k1 = 0.f; F.a2 = 0.f; F.k2 = 0.f;
  if (VAR == 2) F.pre = ffma(ty, c.y, ffma(tx, c.x, 0.f));
  if (VAR == 1) { F.pre = ffma(tx, c.x, 0.f); F.a2 = tz; F.k2 = c.z; }
  if (VAR == 0) { F.a1 = ty; F.k1 = c.y; F.a2 = tz; F.k2 = c.z; }
  return F;
}
// voxel-cube coordinates of a line: a along the line axis AX, (u, v) on the two other axes in ascending order
// (AX 2: u = x, v = y;  AX 1: u = x, v = z;  AX 0: u = y, v = z).  The lines of one u form a "row".
template <int AX> __device__ __forceinline__ f3 place3(float a, float u, float v)
{
  return AX == 2 ? mk3(u, v, a) : (AX == 1 ? mk3(u, a, v) : mk3(a, u, v));
}

// the plane vectors of one fan triangle as the kernel keeps them in shared memory (phase A)
struct TriPlanes {
  f3 c0, c1, n, c3, c4;     // axes of the two sampling cubes (crixus_d.cu:507-529)
  f3 e1, c6, e2, c8;        // voronoi planes: vector and midpoint (:552-561)
  f3 c9, c10, c11, en;      // wall wedge (:562-579) and edge normal (:581-596)
};
#define VV_LOAD_VE(P, t)                                                                                      \
  {                                                                                                           \
    const float4 P4 = (P)[4], P5 = (P)[5], P6 = (P)[6], P9 = (P)[9];                                          \
    t.e1 = mk3(P4.x, P4.y, P4.z); t.c6 = mk3(P4.w, P5.x, P5.y); t.e2 = mk3(P5.z, P5.w, P6.x);                 \
    t.c8 = mk3(P6.y, P6.z, P6.w); t.en = mk3(P9.y, P9.z, P9.w);                                               \
  }
#define VV_LOAD_W(P, t)                                                                                       \
  {                                                                                                           \
    const float4 P7 = (P)[7], P8 = (P)[8], P9 = (P)[9];                                                       \
    t.c9 = mk3(P7.x, P7.y, P7.z); t.c10 = mk3(P7.w, P8.x, P8.y); t.c11 = mk3(P8.z, P8.w, P9.x);               \
  }
#define VV_LOAD_CUBES(P, t)                                                                                   \
  {                                                                                                           \
    const float4 P0 = (P)[0], P1 = (P)[1], P2 = (P)[2], P3 = (P)[3];                                          \
    t.c0 = mk3(P0.x, P0.y, P0.z); t.c1 = mk3(P0.w, P1.x, P1.y); t.n = mk3(P1.z, P1.w, P2.x);                  \
    t.c3 = mk3(P2.y, P2.z, P2.w); t.c4 = mk3(P3.x, P3.y, P3.z);                                               \
  }

// Exact per-voxel evaluation of triangle j on one line (the reference's loop body, crixus_d.cu:598-662): kill mask and
// halving events.  Used for the lanes where the threshold shortcut cannot prove its result (two voxels within eps of a
// plane on a line that is cut by it, or an expression exactly equal to a threshold).  One function per plane family,
// because pass 2 visits the families in two separate triangle loops.
// Results go to a per-warp shared-memory scratch (res[lane] = kill mask, res[32 k + lane] = voxels with at least k halvings,
// k = 1..3): no address of a caller's register variable escapes into this out-of-line function, so the
// per-voxel state of the hot loop (alive, h[5]) stays in registers (with reference parameters it lived in local memory:
// 2.5 % of the executed instructions were LDL/STL).
template <int AX> struct TriCtxVE {
  PlaneF<AX, 0> V, V2;
  PlaneF<AX, 1> E;
  float eps;
  int open;
};
template <int AX> struct TriCtxW {
  PlaneF<AX, 2> W0, W1, W2;
  float eps;
};
template <int AX>
__device__ __noinline__ void exact_tri_ve(const VVArgs &a, const TriCtxVE<AX> &c, bool mine, unsigned long long *res, int lane)
{
  if (!mine) return;   // called by the whole warp (keeps it converged); only the flagged lanes do the work
  const float eps = c.eps;
  uint64_t k = 0ull, c1 = 0ull, c2 = 0ull, c3 = 0ull;
  for (int m = 0; m < VV_GS; m++) {
    const float g = a.gz[m];
    const float sV = c.V(g);
    bool kill = sV > eps;
    int nh = fabsf(sV) < eps;
    if (c.open) {
      const float sV2 = c.V2(g);
      kill |= sV2 > eps;
      nh += fabsf(sV2) < eps;
    }
    const float sE = c.E(g);
    kill |= sE > eps;
    nh += fabsf(sE) < eps;
    if (kill) k |= 1ull << m;
    if (nh >= 1) c1 |= 1ull << m;
    if (nh >= 2) c2 |= 1ull << m;
    if (nh >= 3) c3 |= 1ull << m;
  }
  res[lane] = k; res[32 + lane] = c1; res[64 + lane] = c2; res[96 + lane] = c3;
}
template <int AX>
__device__ __noinline__ void exact_tri_w(const VVArgs &a, const TriCtxW<AX> &c, bool mine, unsigned long long *res, int lane)
{
  if (!mine) return;
  const float eps = c.eps, neps = -c.eps;
  uint64_t k = 0ull, n0 = 0ull;
  for (int m = 0; m < VV_GS; m++) {
    const float g = a.gz[m];
    const float w0 = c.W0(g), w1 = c.W1(g), w2 = c.W2(g);
    const bool inside = !(w0 < neps) && !(w1 < neps) && !(w2 < neps);
    const bool hw = fabsf(w0) < eps;
    if (inside && !hw) k |= 1ull << m;
    if (inside && hw) n0 |= 1ull << m;
  }
  res[lane] = k; res[32 + lane] = n0;
}

// kill mask {f > eps} and halving mask {|f| < eps} of a voronoi / edge plane on one line.  trig is raised when the
// halving mask cannot be decided from the ends and the element next to the flip.
// One search serves both questions: a lane whose line is cut by the plane looks for the flip of {f > eps}; a lane whose
// line is not cut but whose ends disagree on {f > -eps} (e.g. the last voxel lies exactly in a plane through the face of
// the sampling cube -- common when h is commensurate with dr/40) looks for the flip of that predicate instead.
// WIDE (kernel-level template flag, chosen by the host from dr): the plane expressions are not normalised -- c is an
// edge vector of length ~h -- while eps = dr/20*1e-4 and the voxel step is dr/40, so a voxel step changes f by ~h*dr/40
// against a band of 2*eps: for dr below ~2e-3 (length units of the mesh) a cut line has two or more voxels within eps
// of a plane as a rule, and those lanes also need the flip of {f > -eps}.  The non-WIDE kernel leaves that (then
// rare) case to exact_tri; the extra gated search costs ~8 % when it is never needed, hence two kernels.
template <bool WIDE, class F>
__device__ __forceinline__ void plane_kill(const float *gzt, float fa, float fb, F f, float eps, uint64_t &K, uint64_t &H, bool &trig)
{
  const float neps = -eps;
  const bool pa = fa > eps, pb = fb > eps, Aa = fa > neps, Ab = fb > neps;
  const bool cutK = pa != pb;
  const bool cutA = !cutK && !pa && (Aa != Ab);
  const float th = cutK ? eps : neps;
  const bool qa = cutK ? pa : Aa;
  int lo = VV_GS - 1;
  const unsigned cutm = __ballot_sync(VV_FULLWARP, cutK || cutA);
  if (cutm) {
    VV_STAT(2, 1); VV_STAT(3, __popc(cutm));
    const int s = mono_search(gzt, qa, [&](float g) { return f(g) > th; });
    if (cutK || cutA) lo = s;
  }
  const uint64_t *lmt = reinterpret_cast<const uint64_t *>(gzt + 64);
  const uint64_t mask = lmt[(qa ? 64 : 0) + lo + 1];     // {f > th}
  bool need2 = false;
  if (cutK) {
    K |= mask;
    // the surviving voxels next to the flip carry the largest values: the first one(s) may lie within eps of the plane
    const int d = pa ? 1 : -1, t1 = pa ? lo + 1 : lo;
    const int t2 = min(max(t1 + d, 0), VV_GS - 1);
    const float v1 = f(gzt[t1]), v2 = f(gzt[t2]);
    if (v1 == eps) trig = true;                       // exactly on the threshold: exact_tri decides
    else if (v1 > neps) {
      if (t2 != t1 && v2 > neps) { if (WIDE) need2 = true; else trig = true; }   // two or more voxels in the band
      else H = bit64(t1);
    }
  } else if (pa) {
    K = VV_ALL;                   // both ends beyond the plane: the whole line is removed
  } else {
    // no voxel is removed (all f <= eps).  If both ends are < eps so is every voxel (monotone) and the halved voxels
    // are {f > -eps}: the searched mask, or all / none when the ends agree.  An end exactly equal to eps: exact_tri.
    if (fa < eps && fb < eps) {
      if (cutA) H = mask;
      else if (Aa) H = VV_ALL;    // the whole line lies within eps of the plane (line parallel to it)
    } else trig = true;
  }
  if (WIDE) {
    if (__any_sync(VV_FULLWARP, need2)) {   // wide band on a cut line: the band is {f > -eps} minus the removed voxels
      int l2;
      bool p2;
      const uint64_t mA = mono_mask(gzt, fa, fb, f, [&](float v) { return v > neps; }, l2, p2);
      if (need2) H = mA & ~mask;
    }
  }
}

// First pass, a whole row of 41 lines at a time: which lines are removed entirely by a single triangle?
// A line dies if both of its ends lie beyond a voronoi/edge plane or strictly inside a wall sector (exact, by
// monotonicity along the line).  Along the line each plane expression takes its smaller value at a FIXED end (A if its
// coefficient of the line axis is >= 0, else B), so "both ends pass" is one predicate of the in-row index l -- again a monotone
// step function, found by the same six-evaluation search over g[l].  One lane per (row, triangle): ~6 searches instead of
// 41 line tests.
template <int AX, bool OPEN, class VVWarp>
__device__ __forceinline__ void rows_dead(const VVArgs &a, const VVWarp &w, const float *gzt, int tris, f3 pi, int lane,
                                          unsigned long long *rd)
{
  constexpr int VX = (AX == 2) ? 1 : 2;     // the in-row axis (place3)
  const float eps = a.eps, neps = -a.eps;
  const float gA = a.gz[0], gB = a.gz[VV_GS - 1];
  const int nitems = VV_GS * tris;
  for (int base = 0; base < nitems; base += 32) {   // warp-uniform trip count
    const int item = base + lane;
    const bool act = item < nitems;
    const int j = act ? item / VV_GS : 0, k = act ? item - j * VV_GS : 0;   // consecutive lanes: same triangle (broadcast loads)
    const float gu = gzt[k];
    TriPlanes t;
    VV_LOAD_VE(w.plane[j], t);
    VV_LOAD_W(w.plane[j], t);
    // grid offsets of the row with the line axis at the end where the expression is smallest
    auto low_end = [&](const f3 &c) { return place3<AX>(cget<AX>(c) >= 0.f ? gA : gB, gu, 0.f); };
    int lo;
    bool pa;
    uint64_t dead;
    // the lines of the row a voronoi / edge plane TOUCHES (removes or halves a voxel of): its expression exceeds -eps
    // somewhere on the line, i.e. at the end where it is largest -- the same kind of predicate of l
    auto high_end = [&](const f3 &c) { return place3<AX>(cget<AX>(c) >= 0.f ? gB : gA, gu, 0.f); };
    uint64_t touch;
    {  // voronoi plane
      const PlaneF<VX, 0> f = make_plane<VX, 0>(t.e1, t.c6, pi, low_end(t.e1));
      dead = mono_mask(gzt, f(gA), f(gB), f, [&](float v) { return v > eps; }, lo, pa);
      const PlaneF<VX, 0> fh = make_plane<VX, 0>(t.e1, t.c6, pi, high_end(t.e1));
      touch = mono_mask(gzt, fh(gA), fh(gB), fh, [&](float v) { return v > neps; }, lo, pa);
    }
    if (OPEN) {
      const PlaneF<VX, 0> f = make_plane<VX, 0>(t.e2, t.c8, pi, low_end(t.e2));
      dead |= mono_mask(gzt, f(gA), f(gB), f, [&](float v) { return v > eps; }, lo, pa);
      const PlaneF<VX, 0> fh = make_plane<VX, 0>(t.e2, t.c8, pi, high_end(t.e2));
      touch |= mono_mask(gzt, fh(gA), fh(gB), fh, [&](float v) { return v > neps; }, lo, pa);
    }
    {  // edge plane
      const PlaneF<VX, 1> f = make_plane<VX, 1>(t.en, t.e1, pi, low_end(t.en));
      dead |= mono_mask(gzt, f(gA), f(gB), f, [&](float v) { return v > eps; }, lo, pa);
      const PlaneF<VX, 1> fh = make_plane<VX, 1>(t.en, t.e1, pi, high_end(t.en));
      touch |= mono_mask(gzt, fh(gA), fh(gB), fh, [&](float v) { return v > neps; }, lo, pa);
    }
    if (act && touch) atomicOr(&rd[48 + k], (unsigned long long)touch);
    {  // wall sector: w0 >= eps, w1 >= -eps, w2 >= -eps at both ends
      const PlaneF<VX, 2> f0 = make_plane<VX, 2>(t.c9, t.c9, pi, low_end(t.c9));
      const PlaneF<VX, 2> f1 = make_plane<VX, 2>(t.c10, t.c10, pi, low_end(t.c10));
      const PlaneF<VX, 2> f2 = make_plane<VX, 2>(t.c11, t.c11, pi, low_end(t.c11));
      uint64_t m = mono_mask(gzt, f0(gA), f0(gB), f0, [&](float v) { return v >= eps; }, lo, pa);
      m &= mono_mask(gzt, f1(gA), f1(gB), f1, [&](float v) { return v >= neps; }, lo, pa);
      m &= mono_mask(gzt, f2(gA), f2(gB), f2, [&](float v) { return v >= neps; }, lo, pa);
      dead |= m;
    }
    if (act && dead) atomicOr(&rd[k], (unsigned long long)dead);
  }
}

// contribution of the voxels in `va`, in units of 2^-31
__device__ __forceinline__ long long group_weight(uint64_t va, const uint64_t (&h)[5], int cmax)
{
  const uint64_t anyh = (h[0] | h[1] | h[2] | h[3] | h[4]) & va;
  long long s = (long long)__popcll(va & ~anyh) << 31;
  if (anyh) {
    const uint64_t big = (h[3] | h[4]) & va;                  // halved eight times or more: rare, one by one
    const uint64_t lo4 = va & ~big & ~h[2], hi4 = va & ~big & h[2];
    if (cmax >= 1) s += (long long)__popcll(lo4 & h[0] & ~h[1]) << 30;
    if (cmax >= 2) s += (long long)__popcll(lo4 & ~h[0] & h[1]) << 29;
    if (cmax >= 3) s += (long long)__popcll(lo4 & h[0] & h[1]) << 28;
    if (hi4) {                                                // four to seven halvings (voxels on several planes at once)
      if (cmax >= 4) s += (long long)__popcll(hi4 & ~h[0] & ~h[1]) << 27;
      if (cmax >= 5) s += (long long)__popcll(hi4 & h[0] & ~h[1]) << 26;
      if (cmax >= 6) s += (long long)__popcll(hi4 & ~h[0] & h[1]) << 25;
      if (cmax >= 7) s += (long long)__popcll(hi4 & h[0] & h[1]) << 24;
    }
    uint64_t r = big;
    while (r) {
      const int b = __ffsll((long long)r) - 1;
      r &= r - 1;
      const int c = (int)((h[0] >> b) & 1) | ((int)((h[1] >> b) & 1) << 1) | ((int)((h[2] >> b) & 1) << 2) |
                    ((int)((h[3] >> b) & 1) << 3) | ((int)((h[4] >> b) & 1) << 4);
      if (c <= cmax) s += 1ll << (31 - c);
    }
  }
  return s;
}

// Second pass: the weight of one surviving line (in units of 2^-31), every lane of the warp on a line of its own.
// Lanes without a line (valid == false) run along with an empty line.  gu, gv: grid offsets of the line on the two axes other
// than AX (place3).
template <int AX, bool OPEN, bool WIDE, class VVWarp>
__device__ __forceinline__ long long line_weight(const VVArgs &a, const VVWarp &w, const float *gzt, int tris, f3 pi, float gu,
                                                 float gv, bool valid, bool ve, unsigned long long *xres, int lane)
{
  const float thr = a.thr_cube, eps = a.eps, neps = -a.eps;
  const float gA = a.gz[0], gB = a.gz[VV_GS - 1];
  const f3 gp = place3<AX>(0.f, gu, gv);
  uint64_t alive = valid ? VV_ALL : 0ull;
  uint64_t h[5] = {0ull, 0ull, 0ull, 0ull, 0ull};
  // ---------------- planes (crixus_d.cu:598-662) in two triangle loops -- voronoi (+ second voronoi) and edge planes first,
  // wall tetrahedra second.  The per-voxel result is a product over triangles and families, so the order is free; two
  // small loop bodies instead of one large one keep the hot code inside the instruction cache (ncu: instruction-fetch
  // stalls were a quarter of the issue interval), and lines that the voronoi planes remove never reach the wall tests.
  for (int j = 0; j < (ve ? tris : 0); j++) {    // ve: warp-uniform (vv_lines: the ring of the lines those planes touch)
    if (!__any_sync(VV_FULLWARP, alive != 0ull)) break;
    TriPlanes t;
    VV_LOAD_VE(w.plane[j], t);
    const PlaneF<AX, 0> fV = make_plane<AX, 0>(t.e1, t.c6, pi, gp);
    const PlaneF<AX, 1> fE = make_plane<AX, 1>(t.en, t.e1, pi, gp);
    PlaneF<AX, 0> fV2 = fV;
    if (OPEN) fV2 = make_plane<AX, 0>(t.e2, t.c8, pi, gp);
    uint64_t K = 0ull, HV = 0ull, HV2 = 0ull, HE = 0ull;
    bool trig = false;
    const bool live = alive != 0ull;
    // A family is entered only if it can touch a live line of the warp: both ends <= -eps means every voxel of the
    // line is (by monotonicity) on the near side of the plane and further than eps from it.
    const float sVa = fV(gA), sVb = fV(gB);
    if (VV_UNLIKELY(__any_sync(VV_FULLWARP, live && !((sVa <= neps) && (sVb <= neps))))) plane_kill<WIDE>(gzt, sVa, sVb, fV, eps, K, HV, trig);
    if (OPEN) {
      const float s2a = fV2(gA), s2b = fV2(gB);
      if (VV_UNLIKELY(__any_sync(VV_FULLWARP, live && !((s2a <= neps) && (s2b <= neps))))) plane_kill<WIDE>(gzt, s2a, s2b, fV2, eps, K, HV2, trig);
    }
    const float sEa = fE(gA), sEb = fE(gB);
    if (VV_UNLIKELY(__any_sync(VV_FULLWARP, live && !((sEa <= neps) && (sEb <= neps))))) plane_kill<WIDE>(gzt, sEa, sEb, fE, eps, K, HE, trig);
    trig = trig && live;
    VV_STAT(1, 1);
    if (VV_UNLIKELY(__any_sync(VV_FULLWARP, trig))) {
      VV_STAT(4, 1); VV_STAT(5, __popc(__ballot_sync(VV_FULLWARP, trig)));
      TriCtxVE<AX> c;
      c.V = fV; c.V2 = fV2; c.E = fE; c.eps = eps; c.open = OPEN ? 1 : 0;
      exact_tri_ve<AX>(a, c, trig, xres, lane);
      __syncwarp();
      if (trig) {   // voxels with >= 1, >= 2, >= 3 halvings: one increment each (the third only with a second voronoi plane)
        K = xres[lane];
        HV = xres[32 + lane]; HE = xres[64 + lane]; HV2 = xres[96 + lane];
      }
      __syncwarp();
    }
    if (__any_sync(VV_FULLWARP, (HV | HE) != 0ull)) { VV_STAT(8, 1); half_inc_warp(h, HV); half_inc_warp(h, HE); }
    if (OPEN) if (__any_sync(VV_FULLWARP, HV2 != 0ull)) half_inc_warp(h, HV2);
    alive &= ~K;
  }
  for (int j = 0; j < tris; j++) {
    if (!__any_sync(VV_FULLWARP, alive != 0ull)) break;
    TriPlanes t;
    VV_LOAD_W(w.plane[j], t);
    const PlaneF<AX, 2> fW0 = make_plane<AX, 2>(t.c9, t.c9, pi, gp);
    const PlaneF<AX, 2> fW1 = make_plane<AX, 2>(t.c10, t.c10, pi, gp);
    const PlaneF<AX, 2> fW2 = make_plane<AX, 2>(t.c11, t.c11, pi, gp);
    const bool live = alive != 0ull;
    // ---- wall tetrahedron: inside = all three w >= -eps; inside and |w0| < eps -> halved, inside otherwise -> removed
    const float w0a = fW0(gA), w0b = fW0(gB);
    const float w1a = fW1(gA), w1b = fW1(gB);
    const float w2a = fW2(gA), w2b = fW2(gB);
    const bool noneW = ((w0a < neps) && (w0b < neps)) || ((w1a < neps) && (w1b < neps)) || ((w2a < neps) && (w2b < neps));
    if (!__any_sync(VV_FULLWARP, live && !noneW)) continue;
    uint64_t K, HW;
    bool trig = false;
    auto ge = [&](float v) { return !(v < neps); };
    int lo0, lo1, lo2;
    bool pa0, pa1, pa2;
    const uint64_t IN0 = mono_mask(gzt, w0a, w0b, fW0, ge, lo0, pa0);
    const uint64_t IN1 = mono_mask(gzt, w1a, w1b, fW1, ge, lo1, pa1);
    const uint64_t IN2 = mono_mask(gzt, w2a, w2b, fW2, ge, lo2, pa2);
    const uint64_t inside = IN0 & IN1 & IN2;
    // LT = {w0 < eps}.  On a line cut by IN0 it is the complement of IN0 plus the n voxels of IN0 next to the flip that
    // lie in the wall plane (n = 0 or 1 in all but degenerate cases).
    const bool cut0 = pa0 != !(w0b < neps);
    uint64_t LT;
    bool needLT;
    {
      const int d = pa0 ? -1 : 1, t1 = pa0 ? lo0 : lo0 + 1;
      const int t2 = min(max(t1 + d, 0), VV_GS - 1);
      const float v1 = fW0(gzt[t1]), v2 = fW0(gzt[t2]);
      const bool b1 = v1 < eps, b2 = v2 < eps;
      const bool Ba = w0a < eps, Bb = w0b < eps;
      if (cut0) {
        trig |= (v1 == neps);                     // |w0| == eps exactly: IN0 true but not "in the wall plane"
        needLT = b1 && b2 && (t2 != t1);          // two or more voxels in the plane band: search
        LT = (VV_ALL & ~IN0) | (b1 ? bit64(t1) : 0ull);
      } else {
        trig |= pa0 && (fminf(w0a, w0b) == neps);
        needLT = Ba != Bb;
        LT = Ba ? VV_ALL : 0ull;
      }
    }
    if (__any_sync(VV_FULLWARP, needLT)) {
      int lo_;
      bool pa_;
      const uint64_t m = mono_mask(gzt, w0a, w0b, fW0, [&](float v) { return v < eps; }, lo_, pa_);
      if (needLT) LT = m;
    }
    K = inside & ~LT;
    HW = inside & LT;
    trig = trig && live;
    if (VV_UNLIKELY(__any_sync(VV_FULLWARP, trig))) {
      TriCtxW<AX> c;
      c.W0 = fW0; c.W1 = fW1; c.W2 = fW2; c.eps = eps;
      exact_tri_w<AX>(a, c, trig, xres, lane);
      __syncwarp();
      if (trig) { K = xres[lane]; HW = xres[32 + lane]; }
      __syncwarp();
    }
    if (__any_sync(VV_FULLWARP, HW != 0ull)) { VV_STAT(7, 1); half_inc_warp(h, HW); }
    alive &= ~K;
  }
  // ---------------- cubes (crixus_d.cu:531-543): a voxel is sampled if it lies in one of the two dr-cubes of some j
  uint64_t hit = 0ull;
  for (int j = 0; j < tris; j++) {
    if (!__any_sync(VV_FULLWARP, (alive & ~hit) != 0ull)) break;
    TriPlanes t;
    VV_LOAD_CUBES(w.plane[j], t);
    const PlaneF<AX, 2> f0 = make_plane<AX, 2>(t.c0, t.c0, pi, gp), f1 = make_plane<AX, 2>(t.c1, t.c1, pi, gp),
                        f2 = make_plane<AX, 2>(t.n, t.n, pi, gp), f3_ = make_plane<AX, 2>(t.c3, t.c3, pi, gp),
                        f4 = make_plane<AX, 2>(t.c4, t.c4, pi, gp);
    const float a0 = f0(gA), b0 = f0(gB);
    const float a1 = f1(gA), b1 = f1(gB);
    const float a2 = f2(gA), b2 = f2(gB);
    const float a3 = f3_(gA), b3 = f3_(gB);
    const float a4 = f4(gA), b4 = f4(gB);
    // both ends inside a cube => the whole line is inside it
    const bool i0 = fmaxf(fabsf(a0), fabsf(b0)) <= thr, i1 = fmaxf(fabsf(a1), fabsf(b1)) <= thr, i2 = fmaxf(fabsf(a2), fabsf(b2)) <= thr,
               i3 = fmaxf(fabsf(a3), fabsf(b3)) <= thr, i4 = fmaxf(fabsf(a4), fabsf(b4)) <= thr;
    const bool full = (i0 && i1 && i2) || (i2 && i3 && i4);
    if (full) hit = VV_ALL;
    VV_STAT(10, 1);
    if (__any_sync(VV_FULLWARP, !full && (alive & ~hit) != 0ull)) {
      VV_STAT(11, 1);
      const float nthr = -thr;
      auto le = [&](float v) { return v <= thr; };
      auto ge = [&](float v) { return v >= nthr; };
      int lo_;
      bool pa_;
#define VV_SLAB(fa, fb, F) (mono_mask(gzt, fa, fb, F, le, lo_, pa_) & mono_mask(gzt, fa, fb, F, ge, lo_, pa_))
      const uint64_t s0 = VV_SLAB(a0, b0, f0);
      const uint64_t s1 = VV_SLAB(a1, b1, f1);
      const uint64_t s2 = VV_SLAB(a2, b2, f2);
      const uint64_t s3 = VV_SLAB(a3, b3, f3_);
      const uint64_t s4 = VV_SLAB(a4, b4, f4);
#undef VV_SLAB
      hit |= (s0 & s1 & s2) | (s2 & s3 & s4);
    }
  }
  return valid ? group_weight(hit & alive, h, a.cmax) : 0ll;
}

// ---------------------------------------------------------------------------------------------------------------
// Phase A of one vertex (warp-cooperative): fan gather, chain ordering, edge-normal pairing, plane vectors -> w.
// Returns false when the vertex is finished here (error flagged or volume zero written).  av = the reference's avnorm.
// CSR: the fan and the edge neighbours come from the vertex -> (segment, corner) adjacency built once per call (k_vv_adj_*)
// instead of two scans over every segment of the 27 cells (~1700 segments per vertex on configs[1]); the 27 cell ranges only
// serve to put the fan into the reference's discovery order and to reproduce what the reference's scan cannot see.
template <int TCAP, bool CSR, class VVWarp>
__device__ __forceinline__ bool vv_fan(const VVArgs &a, VVWarp &w, const int64_t i, const f3 pi, const int tris, const int lane,
                                       int &closed_out, f3 &av_out)
{
  const cx_params &p = a.p;
  const int gx = cell_coord(pi.x, p.dmin[0], p.griddr[0], p.gridn[0]);
  const int gy = cell_coord(pi.y, p.dmin[1], p.griddr[1], p.gridn[1]);
  const int gz = cell_coord(pi.z, p.dmin[2], p.griddr[2], p.gridn[2]);

  // the 27 neighbouring 3dr-cells (crixus_d.cu:327-347): lane c holds the segment range of cell c (empty if skipped)
  int cb = 0, ce = 0;
  if (lane < 27) {
    const int64_t c = nb_cell(p, gx + lane / 9 - 1, gy + (lane / 3) % 3 - 1, gz + lane % 3 - 1);
    if (c >= 0) { cb = __ldg(&a.cell_idx[c]); ce = __ldg(&a.cell_idx[c + 1]); }
  }
  // ---------------- fan gather (crixus_d.cu:316-369): discovery order = cell order, segment order, corner order
  int itris = 0;
  if (CSR) {
    // entry = segment * 4 + corner.  Its place in the discovery order: (position of its cell among the 27, segment, corner);
    // the cell is the one whose range holds the segment (a ballot over the lanes' ranges).  A segment outside the neighbourhood
    // is not found by the reference's scan, one whose cell the neighbourhood holds twice (periodic grid of < 3 cells) is found
    // twice: both end in the itris != tris error below.
    const int base = __ldg(&a.adj_start[i]), ni = __ldg(&a.adj_start[i + 1]) - base;
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(&w.plane[0][0]);   // scratch: the planes come later
    const int nstore = min(ni, TCAP);
    for (int eb = 0; eb < ni; eb += 32) {
      const int e = eb + lane;
      int ent = 0, rank = 0;
      if (e < ni) ent = __ldg(&a.adj[base + e]);
      const int nb = min(32, ni - eb);
      for (int t = 0; t < nb; t++) {
        const int st = __shfl_sync(0xffffffffu, ent, t) >> 2;
        const unsigned m = __ballot_sync(0xffffffffu, lane < 27 && st >= cb && st < ce);
        itris += __popc(m);
        if (lane == t) rank = m ? __ffs(m) - 1 : 27;
      }
      if (e < nstore) keys[e] = ((unsigned long long)(unsigned)rank << 32) | (unsigned)ent;
    }
    if (ni != tris) itris = -1;
    __syncwarp();
    if (itris == tris) {
      for (int eb = 0; eb < nstore; eb += 32) {
        const int e = eb + lane;
        if (e < nstore) {
          const unsigned long long key = keys[e];
          int off = 0;
          for (int t = 0; t < nstore; t++) off += keys[t] < key;
          const int ent = (int)(unsigned)key, j = ent >> 2, c = ent & 3;
          const int4 E = __ldg(&a.ep[j]);
          w.tri[off][0] = c == 0 ? E.y : (c == 1 ? E.z : E.x);
          w.tri[off][1] = c == 0 ? E.z : (c == 1 ? E.x : E.y);
          w.tri[off][2] = j;
        }
      }
    }
  } else {
    for (int c27 = 0; c27 < 27; c27++) {
      const int b = __shfl_sync(0xffffffffu, cb, c27), e = __shfl_sync(0xffffffffu, ce, c27);
      for (int base = b; base < e; base += 32) {
        const int j = base + lane;
        int4 E = make_int4(-1, -1, -1, -1);
        if (j < e) E = __ldg(&a.ep[j]);
        const int m0 = (E.x == (int)i), m1 = (E.y == (int)i), m2 = (E.z == (int)i);
        const int cnt = m0 + m1 + m2;
        if (!__any_sync(0xffffffffu, cnt != 0)) continue;
        const int incl = warp_incl_scan_i(cnt, lane);
        int off = itris + incl - cnt;
        if (m0 && off < TCAP) { w.tri[off][0] = E.y; w.tri[off][1] = E.z; w.tri[off][2] = j; off++; }
        if (m1 && off < TCAP) { w.tri[off][0] = E.z; w.tri[off][1] = E.x; w.tri[off][2] = j; off++; }
        if (m2 && off < TCAP) { w.tri[off][0] = E.x; w.tri[off][1] = E.y; w.tri[off][2] = j; off++; }
        itris += __shfl_sync(0xffffffffu, incl, 31);
      }
    }
  }
  if (itris != tris) {  // triangles outside the 27-cell neighbourhood: the reference reads garbage here
    if (lane == 0) { atomicOr(&a.ctl->err, VVE_FAN); a.vol[i] = 0.f; }
    return false;
  }
  __syncwarp();

  // ---------------- chain ordering + closedness (crixus_d.cu:371-399), serial on lane 0 (T <= 50)
  if (lane == 0) {
    int closed = 1;
    for (int j = 0; j < tris; j++) {
      for (int k = j + 1; k < tris; k++) {
        if (w.tri[j][1] == w.tri[k][0]) {
          if (k != j + 1)
            for (int l = 0; l < 3; l++) { int t = w.tri[j + 1][l]; w.tri[j + 1][l] = w.tri[k][l]; w.tri[k][l] = t; }
          break;
        }
        if (w.tri[j][1] == w.tri[k][1]) {
          const int d0 = w.tri[k][1], d1 = w.tri[k][0], d2 = w.tri[k][2];
          for (int l = 0; l < 3; l++) w.tri[k][l] = w.tri[j + 1][l];
          w.tri[j + 1][0] = d0; w.tri[j + 1][1] = d1; w.tri[j + 1][2] = d2;
          break;
        }
        if (k == tris - 1) closed = 0;
      }
    }
    if (w.tri[0][0] != w.tri[tris - 1][1]) closed = 0;
    w.closed = closed;
  }
  for (int k = lane; k < tris; k += 32) w.ecount[k] = 0;
  __syncwarp();
  const int closed = w.closed;
  if (a.zero_open && !closed) {  // crixus_d.cu:403-407
    if (lane == 0) a.vol[i] = 0.f;
    return false;
  }

  // ---------------- edge-normal pairing (crixus_d.cu:409-464): the <=2 segments that contain both ends of fan edge k.
  if (CSR) {
    // A segment with two corners in {t0, t1} contains t0, or (degenerate) t1 twice: the adjacency lists of t0 and t1 hold them
    // all.  Each segment is taken once (at its first corner equal to the list's vertex) and counted as often as the scan of the
    // 27 cells meets it (0: outside the neighbourhood).  The order of the two slots is irrelevant (their normals are added).
    __syncwarp();
    for (int k = 0; k < tris; k++) {
      const int t0 = w.tri[k][0], t1 = w.tri[k][1];
      int cnt = 0;   // warp-uniform
      for (int pass = 0; pass < (t0 == t1 ? 1 : 2); pass++) {
        const int v = pass == 0 ? t0 : t1;
        const int vb = __ldg(&a.adj_start[v]), vn = __ldg(&a.adj_start[v + 1]) - vb;
        for (int eb = 0; eb < vn; eb += 32) {
          const int e = eb + lane;
          bool hit = false;
          int sj = 0;
          if (e < vn) {
            const int ent = __ldg(&a.adj[vb + e]), c = ent & 3;
            sj = ent >> 2;
            const int4 E = __ldg(&a.ep[sj]);
            const bool first = (c == 0) || (c == 1 && E.x != v) || (c == 2 && E.x != v && E.y != v);
            const int vf = (E.x == t0 || E.x == t1) + (E.y == t0 || E.y == t1) + (E.z == t0 || E.z == t1);
            const bool has0 = E.x == t0 || E.y == t0 || E.z == t0;
            hit = first && vf == 2 && (pass == 0 || !has0);
          }
          unsigned hm = __ballot_sync(0xffffffffu, hit);
          while (hm) {
            const int src = __ffs(hm) - 1;
            hm &= hm - 1u;
            const int sh = __shfl_sync(0xffffffffu, sj, src);
            const int mult = __popc(__ballot_sync(0xffffffffu, lane < 27 && sh >= cb && sh < ce));
            for (int r = 0; r < mult; r++) {
              if (lane == 0) {
                if (cnt < 2) w.eseg[k][cnt] = sh;
                else atomicOr(&a.ctl->err, VVE_NONMANIFOLD);
              }
              cnt++;
            }
          }
        }
      }
      if (lane == 0) w.ecount[k] = cnt;
    }
  } else {
    // A 64-bit filter of the rim vertices (id mod 64) rejects almost every segment batch without entering the k loop.
    uint64_t rim = 0ull;
    for (int k = 0; k < tris; k++) rim |= (1ull << (w.tri[k][0] & 63)) | (1ull << (w.tri[k][1] & 63));
    for (int c27 = 0; c27 < 27; c27++) {
      const int b = __shfl_sync(0xffffffffu, cb, c27), e = __shfl_sync(0xffffffffu, ce, c27);
      for (int base = b; base < e; base += 32) {   // warp-uniform trip count: the warp must stay converged
        const int j = base + lane;
        int4 E = make_int4(-1, -1, -1, -1);
        bool cand = false;
        if (j < e) {
          E = __ldg(&a.ep[j]);
          cand = (int)((rim >> (E.x & 63)) & 1ull) + (int)((rim >> (E.y & 63)) & 1ull) + (int)((rim >> (E.z & 63)) & 1ull) >= 2;
        }
        if (!__any_sync(0xffffffffu, cand)) continue;
        for (int k = 0; k < tris; k++) {
          const int t0 = w.tri[k][0], t1 = w.tri[k][1];
          const int vf = (E.x == t0 || E.x == t1) + (E.y == t0 || E.y == t1) + (E.z == t0 || E.z == t1);
          if (vf == 2) {
            const int slot = atomicAdd(&w.ecount[k], 1);
            if (slot < 2) w.eseg[k][slot] = j;
            else atomicOr(&a.ctl->err, VVE_NONMANIFOLD);
          }
        }
      }
    }
    __syncwarp();
  }
  __syncwarp();

  // ---------------- avnorm (crixus_d.cu:524): float sum in fan order, every lane computes the same value
  f3 av = mk3(0.f, 0.f, 0.f);
  for (int j = 0; j < tris; j++) {
    const float4 n = __ldg(&a.norm[w.tri[j][2]]);
    av = mk3(fsub(av.x, n.x), fsub(av.y, n.y), fsub(av.z, n.z));
  }

  // ---------------- per-triangle plane vectors (crixus_d.cu:507-529 and :552-596), one lane per fan triangle
  for (int jb = 0; jb < tris; jb += 32) {
    const int j = jb + lane;
    if (j >= tris) continue;
    const f3 n = mk3(__ldg(&a.norm[w.tri[j][2]]));
    const f3 va = mk3(__ldg(&a.pos[w.tri[j][0]])), vb = mk3(__ldg(&a.pos[w.tri[j][1]]));
    const f3 e1 = unwrap3(p, sub3(va, pi), a.two_dr);  // cvec[5] (and cvec[0] before normalisation)
    const f3 e2 = unwrap3(p, sub3(vb, pi), a.two_dr);  // tvec[1] / cvec[7] (and cvec[3] before normalisation)
    const f3 c0 = normalise_ref(e1), c3 = normalise_ref(e2);
    const f3 c1 = cross_r2(c0, n), c4 = cross_r2(c3, n);
    const f3 c6 = mk3(halfway(pi.x, e1.x), halfway(pi.y, e1.y), halfway(pi.z, e1.z));
    const f3 c8 = mk3(halfway(pi.x, e2.x), halfway(pi.y, e2.y), halfway(pi.z, e2.z));
    f3 c9 = cross_r2(e1, e2), c10 = cross_r2(e2, av), c11 = cross_r2(av, e1);  // crixus_d.cu:570
    if (dot3_r1(n, c9) > 0.f) { c9 = neg3(c9); c10 = neg3(c10); c11 = neg3(c11); }
    // edge normal: sum of the (<=2) adjacent segment normals x edge (crixus_d.cu:447-473)
    f3 en = mk3(0.f, 0.f, 0.f);
    const int cnt = min(w.ecount[j], 2);
    if (cnt >= 1) {
      const f3 na = mk3(__ldg(&a.norm[w.eseg[j][0]]));
      en = add3(en, na);
      if (cnt == 2) en = add3(en, mk3(__ldg(&a.norm[w.eseg[j][1]])));
      f3 edge = sub3(va, vb);
      if (cnt == 2) edge = unwrap3(p, edge, a.two_dr);  // single-segment edges are not unwrapped (:465-473)
      en = cross_r2(en, edge);
    }
    if (dot3_r1(en, e1) < 0.f) en = neg3(en);  // crixus_d.cu:581-585
    float4 *P = w.plane[j];
    P[0] = make_float4(c0.x, c0.y, c0.z, c1.x);
    P[1] = make_float4(c1.y, c1.z, n.x, n.y);
    P[2] = make_float4(n.z, c3.x, c3.y, c3.z);
    P[3] = make_float4(c4.x, c4.y, c4.z, 0.f);
    P[4] = make_float4(e1.x, e1.y, e1.z, c6.x);
    P[5] = make_float4(c6.y, c6.z, e2.x, e2.y);
    P[6] = make_float4(e2.z, c8.x, c8.y, c8.z);
    P[7] = make_float4(c9.x, c9.y, c9.z, c10.x);
    P[8] = make_float4(c10.y, c10.z, c11.x, c11.y);
    P[9] = make_float4(c11.z, en.x, en.y, en.z);
  }
  __syncwarp();
  closed_out = closed;
  av_out = av;
  return true;
}

// Pass 1 of one vertex: rd[k] bit l = line (k, l) is removed whole by one triangle; rd[48 + k] bit l = some voronoi / edge plane
// touches line (k, l) (96 words of shared memory per warp, 2 x 41 used); result: the two line lists in `lines` (global memory)
template <int AX, class VVWarp>
__device__ __forceinline__ void vv_rows(const VVArgs &a, const VVWarp &w, const float *gzt, int tris, int closed, f3 pi, int lane,
                                        unsigned long long *rd, unsigned short *lines, int &n0_out, int &n1_out)
{
  for (int r = lane; r < 96; r += 32) rd[r] = 0ull;
  __syncwarp();
  if (closed) rows_dead<AX, false, VVWarp>(a, w, gzt, tris, pi, lane, rd);
  else rows_dead<AX, true, VVWarp>(a, w, gzt, tris, pi, lane, rd);
  __syncwarp();
  // The surviving lines as two lists in ONE array of VV_LSTRIDE entries (line = row * 41 + l): the lines no voronoi / edge plane
  // touches ascending from the front, the touched ones descending from the back.  Lane r expands rows r and r + 32: popcounts ->
  // warp prefix -> one store per set bit (~3x fewer instructions than 53 ballot steps over the 1681 lines, which were a quarter
  // of pass 2's instructions).
  const int rb = lane + 32;
  const uint64_t da = rd[lane], ta = rd[48 + lane];
  const uint64_t db = rb < VV_GS ? rd[rb] : VV_ALL, tb = rb < VV_GS ? rd[48 + rb] : 0ull;
  uint64_t u0a = ~da & ~ta & VV_ALL, u1a = ~da & ta & VV_ALL, u0b = ~db & ~tb & VV_ALL, u1b = ~db & tb & VV_ALL;
  const int c0a = __popcll(u0a), c1a = __popcll(u1a), c0b = __popcll(u0b), c1b = __popcll(u1b);
  // one scan for both lists and both row sets: four 11-bit counts cannot overflow 16-bit fields (<= 1681 in total)
  const unsigned long long packed = (unsigned long long)c0a | ((unsigned long long)c0b << 16) | ((unsigned long long)c1a << 32) |
                                    ((unsigned long long)c1b << 48);
  unsigned long long incl = packed;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  const unsigned long long tot = __shfl_sync(0xffffffffu, incl, 31), excl = incl - packed;
  const int t0a = (int)(tot & 0xffffu), t0b = (int)((tot >> 16) & 0xffffu), t1a = (int)((tot >> 32) & 0xffffu), t1b = (int)(tot >> 48);
  int o0a = (int)(excl & 0xffffu), o0b = t0a + (int)((excl >> 16) & 0xffffu);
  int o1a = (int)((excl >> 32) & 0xffffu), o1b = t1a + (int)(excl >> 48);
  auto expand = [&](uint64_t m, int row, int at, int step) {
    unsigned lo = (unsigned)m, hi = (unsigned)(m >> 32);
    const int base = row * VV_GS;
    while (lo) { lines[at] = (unsigned short)(base + __ffs(lo) - 1); at += step; lo &= lo - 1u; }
    while (hi) { lines[at] = (unsigned short)(base + 31 + __ffs(hi)); at += step; hi &= hi - 1u; }
  };
  expand(u0a, lane, o0a, 1);
  expand(u0b, rb, o0b, 1);
  expand(u1a, lane, VV_LSTRIDE - 1 - o1a, -1);
  expand(u1b, rb, VV_LSTRIDE - 1 - o1b, -1);
  n0_out = t0a + t0b;
  n1_out = t1a + t1b;
  __syncwarp();
}

// Pass 2 of one vertex: the surviving lines come as two lists (vv_rows), 32 per step, one per lane; the weights are summed in
// integers and converted once.  Lines that no voronoi / edge plane touches run line_weight without its first triangle loop: on a
// wall-like fan with the lines along the normal those planes are parallel to the lines, remove whole lines (pass 1) and touch
// the few that lie within eps of them, nothing else.
template <int AX, bool WIDE, class VVWarp>
__device__ __forceinline__ void vv_lines(const VVArgs &a, const VVWarp &w, const float *gzt, int tris, int closed, f3 pi, int lane,
                                         const unsigned short *lines, int n0, int n1, unsigned long long *xres, int64_t i)
{
  long long acc = 0;
  const int nb0 = (n0 + 31) >> 5, nb1 = (n1 + 31) >> 5;
  for (int b = 0; b < nb0 + nb1; b++) {   // one call site of line_weight for both lists (it is the bulk of the kernel's code)
    const bool ve = b >= nb0;
    const int at = (ve ? b - nb0 : b) * 32 + lane;
    const bool have = at < (ve ? n1 : n0);
    VV_STAT(12, 1);
    const int ln = have ? (int)lines[ve ? VV_LSTRIDE - 1 - at : at] : 0;
    const int kk = ln / VV_GS, ll = ln - kk * VV_GS;
    // gp = ((float)idx - 20.f) * gdr: identical expression for x, y and z, so one table serves all three
    const float h0 = gzt[kk], h1 = gzt[ll];
    acc += closed ? line_weight<AX, false, WIDE, VVWarp>(a, w, gzt, tris, pi, h0, h1, have, ve, xres, lane)
                  : line_weight<AX, true, WIDE, VVWarp>(a, w, gzt, tris, pi, h0, h1, have, ve, xres, lane);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    const float volf = (float)((double)acc * (1.0 / 2147483648.0));  // exact sum of 2^-k weights, rounded once
    a.vol[i] = (float)((double)volf * a.gd3);                          // crixus_d.cu:674
  }
}

// shared-memory tables of phase B: g[0..40] for lane-dependent lookups (padded to 64), then the two 64-entry uint64 mask tables
// used by mono_mask
__device__ __forceinline__ void vv_init_tables(const VVArgs &a, float *gzt)
{
  if (threadIdx.x < 64) {
    gzt[threadIdx.x] = threadIdx.x < VV_GS ? a.gz[threadIdx.x] : 0.f;
    uint64_t *lmt = reinterpret_cast<uint64_t *>(gzt + 64);
    const uint64_t lm = threadIdx.x <= VV_GS ? lowmask(threadIdx.x) : VV_ALL;
    lmt[threadIdx.x] = VV_ALL & ~lm;
    lmt[64 + threadIdx.x] = lm;
  }
  __syncthreads();
}

__device__ __forceinline__ int64_t vv_owned_vertex(const VVArgs &a, int64_t idx)
{
  const int64_t q = idx / a.chunk;
  return a.v_begin + (q * a.nparts + a.part) * a.chunk + (idx - q * a.chunk);
}

// The three phases run as three kernels over one chunk of the owned vertices (the instruction cache decides: the monolithic
// kernel's per-vertex path is ~60 KB of SASS against a 32 KB L1.5 instruction cache, and a third of its warp stall samples were
// "no instruction"), handing the fan's plane vectors (160 B per fan triangle) and the dead-line masks (41 words per vertex)
// through HBM scratch -- ~1.4 KB per vertex, noise against the arithmetic.
//   k_vv_fans         phase A; picks the line direction from the fan's own avnorm and appends the vertex to that work list
//   k_vv_rows<AX>     pass 1 (rows_dead)
//   k_vv_lines<AX>    pass 2 (ring compaction + line_weight)
// Fans with more than 16 triangles, and fans that find the chunk's plane store full, go to the monolithic k_vert_volume (list 3).
struct VVHead { int off, tris, closed, vertex, n0, n1; };

// vertex -> (segment, corner) adjacency: count, exclusive scan (cx_exclusive_scan_i32), fill.  The order inside a vertex's list
// is whatever the atomics give; vv_fan sorts the (<= trimax) entries it uses.
__global__ void __launch_bounds__(256) k_vv_adj_count(const int4 *__restrict__ ep, int64_t nbe, int64_t nvert, int *__restrict__ cnt)
{
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < nbe; s += (int64_t)gridDim.x * blockDim.x) {
    const int4 E = __ldg(&ep[s]);
    if ((uint64_t)E.x < (uint64_t)nvert) atomicAdd(&cnt[E.x], 1);
    if ((uint64_t)E.y < (uint64_t)nvert) atomicAdd(&cnt[E.y], 1);
    if ((uint64_t)E.z < (uint64_t)nvert) atomicAdd(&cnt[E.z], 1);
  }
}
__global__ void __launch_bounds__(256) k_vv_adj_fill(const int4 *__restrict__ ep, int64_t nbe, int64_t nvert, const int *__restrict__ start,
                                                     int *__restrict__ cursor, int *__restrict__ adj)
{
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < nbe; s += (int64_t)gridDim.x * blockDim.x) {
    const int4 E = __ldg(&ep[s]);
    const int e4 = (int)s * 4;
    if ((uint64_t)E.x < (uint64_t)nvert) adj[start[E.x] + atomicAdd(&cursor[E.x], 1)] = e4;
    if ((uint64_t)E.y < (uint64_t)nvert) adj[start[E.y] + atomicAdd(&cursor[E.y], 1)] = e4 + 1;
    if ((uint64_t)E.z < (uint64_t)nvert) adj[start[E.z] + atomicAdd(&cursor[E.z], 1)] = e4 + 2;
  }
}

template <bool CSR>
__global__ void __launch_bounds__(VV_WARPS * 32, VV_LB_FANS) k_vv_fans(const __grid_constant__ VVArgs a)
{
  typedef VVWarpT<VV_TCAP_SMALL> VVWarp;
  __shared__ VVWarp sm[VV_WARPS];
  const int lane = threadIdx.x & 31;
  VVWarp &w = sm[threadIdx.x >> 5];
  const int force = a.force_axis;
  for (;;) {
    unsigned long long widx = 0;
    if (lane == 0) widx = atomicAdd(&a.ctl->work[0], 1ull);
    widx = __shfl_sync(0xffffffffu, widx, 0);
    if ((int64_t)widx >= a.c_end - a.c_begin) break;
    const int64_t idx = a.c_begin + (int64_t)widx;
    const int64_t i = vv_owned_vertex(a, idx);
    __syncwarp();
    const float4 pi4 = __ldg(&a.pos[i]);
    if (pi4.x < -1e9f) continue;  // vertex deleted by the periodic pairing (crixus_d.cu:297)
    const f3 pi = mk3(pi4);
    const int tris = __ldg(&a.trisize[i]);
    if (tris > CX_TRIMAX) {  // crixus_d.cu:306: the reference thread gives up; here: flag + zero
      if (lane == 0) { atomicOr(&a.ctl->err, VVE_TRIMAX); a.vol[i] = 0.f; }
      continue;
    }
    if (tris <= 0) { if (lane == 0) a.vol[i] = 0.f; continue; }
    if (tris > VV_TCAP_SMALL) {  // fan too large for the shared-memory slot of the split kernels
      if (lane == 0) a.list[3 * a.nown + atomicAdd(&a.ctl->count[3], 1)] = (int)i;
      continue;
    }
    int closed;
    f3 av;
    if (!vv_fan<VV_TCAP_SMALL, CSR>(a, w, i, pi, tris, lane, closed, av)) continue;
    int off = 0;
    if (lane == 0) off = atomicAdd(&a.ctl->bump, tris);
    off = __shfl_sync(0xffffffffu, off, 0);
    if (off + tris > a.plane_cap) {   // plane store of this chunk is full: the monolithic kernel takes the vertex
      if (lane == 0) a.list[3 * a.nown + atomicAdd(&a.ctl->count[3], 1)] = (int)i;
      continue;
    }
    __syncwarp();
    const float4 *src = &w.plane[0][0];
    float4 *dst = a.planes + (int64_t)off * VV_PLANE_F4;
    for (int q = lane; q < tris * VV_PLANE_F4; q += 32) dst[q] = src[q];
    if (lane == 0) {
      // line direction = dominant axis of the fan normal; ties: z, then y (the cheaper evaluations)
      const float ax = fabsf(av.x), ay = fabsf(av.y), az = fabsf(av.z);
      int cls = (az >= ax && az >= ay) ? 2 : (ay >= ax ? 1 : 0);
      if (force >= 0) cls = force;
      VVHead h;
      h.off = off; h.tris = tris; h.closed = closed; h.vertex = (int)i; h.n0 = 0; h.n1 = 0;
      const int slot = (int)(idx - a.c_begin);
      a.heads[slot] = make_int4(h.off, h.tris, h.closed, h.vertex);
      a.list[(int64_t)cls * a.nown + atomicAdd(&a.ctl->count[cls], 1)] = slot;
    }
  }
}

// fetch of one work item of the split kernels: header + plane vectors into shared memory
template <class VVWarp>
__device__ __forceinline__ VVHead vv_fetch(const VVArgs &a, VVWarp &w, int slot, int lane)
{
  const int4 hv = a.heads[slot];
  VVHead h;
  h.off = hv.x; h.tris = hv.y; h.closed = hv.z & 1; h.vertex = hv.w; h.n0 = (hv.z >> 1) & 0x7ff; h.n1 = (hv.z >> 12) & 0x7ff;
  const float4 *src = a.planes + (int64_t)h.off * VV_PLANE_F4;
  float4 *dst = &w.plane[0][0];
  for (int q = lane; q < h.tris * VV_PLANE_F4; q += 32) dst[q] = src[q];
  __syncwarp();
  return h;
}

template <int AX>
__global__ void __launch_bounds__(VV_WARPS * 32, 6) k_vv_rows(const __grid_constant__ VVArgs a)
{
  typedef VVWarpT<VV_TCAP_SMALL> VVWarp;
  __shared__ VVWarp sm[VV_WARPS];
  __shared__ __align__(16) float gzt[64 + 256];
  __shared__ unsigned long long rds[VV_WARPS][96];
  const int lane = threadIdx.x & 31;
  VVWarp &w = sm[threadIdx.x >> 5];
  unsigned long long *rd = rds[threadIdx.x >> 5];
  vv_init_tables(a, gzt);
  const int *list = a.list + (int64_t)AX * a.nown;
  const int64_t nlist = a.ctl->count[AX];
  for (;;) {
    unsigned long long widx = 0;
    if (lane == 0) widx = atomicAdd(&a.ctl->work[1 + AX], 1ull);
    widx = __shfl_sync(0xffffffffu, widx, 0);
    if ((int64_t)widx >= nlist) break;
    const int slot = list[widx];
    __syncwarp();
    const VVHead h = vv_fetch(a, w, slot, lane);
    const f3 pi = mk3(__ldg(&a.pos[h.vertex]));
    int n0, n1;
    vv_rows<AX>(a, w, gzt, h.tris, h.closed, pi, lane, rd, a.lines + (int64_t)slot * VV_LSTRIDE, n0, n1);
    if (lane == 0) a.heads[slot].z = h.closed | (n0 << 1) | (n1 << 12);     // counts <= 1681: 11 bits each
    __syncwarp();
  }
}

template <int AX, bool WIDE>
__global__ void __launch_bounds__(VV_WARPS * 32, VV_LB_LINES) k_vv_lines(const __grid_constant__ VVArgs a)
{
  typedef VVWarpT<VV_TCAP_SMALL> VVWarp;
  __shared__ VVWarp sm[VV_WARPS];
  __shared__ __align__(16) float gzt[64 + 256];
  __shared__ unsigned long long xress[VV_WARPS][128]; // per-warp: results of the out-of-line exact path (exact_tri_*)
  const int lane = threadIdx.x & 31;
  VVWarp &w = sm[threadIdx.x >> 5];
  unsigned long long *xres = xress[threadIdx.x >> 5];
  vv_init_tables(a, gzt);
  const int *list = a.list + (int64_t)AX * a.nown;
  const int64_t nlist = a.ctl->count[AX];
  for (;;) {
    unsigned long long widx = 0;
    if (lane == 0) widx = atomicAdd(&a.ctl->work[4 + AX], 1ull);
    widx = __shfl_sync(0xffffffffu, widx, 0);
    if ((int64_t)widx >= nlist) break;
    const int slot = list[widx];
    __syncwarp();
    const VVHead h = vv_fetch(a, w, slot, lane);
    const f3 pi = mk3(__ldg(&a.pos[h.vertex]));
    vv_lines<AX, WIDE>(a, w, gzt, h.tris, h.closed, pi, lane, a.lines + (int64_t)slot * VV_LSTRIDE, h.n0, h.n1, xres, h.vertex);
    __syncwarp();
  }
}

// The monolithic form: all phases of a vertex in one kernel, fans of up to TCAP = trimax triangles; serves work list 3
template <int TCAP, bool WIDE, int AX, bool CSR>
__global__ void __launch_bounds__(VV_WARPS * 32, 4) k_vert_volume(const __grid_constant__ VVArgs a)
{
  typedef VVWarpT<TCAP> VVWarp;
  __shared__ VVWarp sm[VV_WARPS];
  __shared__ __align__(16) float gzt[64 + 256];
  __shared__ unsigned long long rds[VV_WARPS][96];
  __shared__ unsigned long long xress[VV_WARPS][128];
  const int lane = threadIdx.x & 31;
  VVWarp &w = sm[threadIdx.x >> 5];
  unsigned long long *rd = rds[threadIdx.x >> 5];
  unsigned long long *xres = xress[threadIdx.x >> 5];
  unsigned short *lines = a.lines_big + (int64_t)(blockIdx.x * VV_WARPS + (threadIdx.x >> 5)) * VV_LSTRIDE;   // this warp's line lists
  vv_init_tables(a, gzt);
  const int *list = a.list + 3 * a.nown;
  const int64_t nlist = a.ctl->count[3];     // complete: filled by the earlier launches
  for (;;) {
    unsigned long long widx = 0;
    if (lane == 0) widx = atomicAdd(&a.ctl->work[7], 1ull);
    widx = __shfl_sync(0xffffffffu, widx, 0);
    if ((int64_t)widx >= nlist) break;
    const int64_t i = list[widx];
    __syncwarp();
    const f3 pi = mk3(__ldg(&a.pos[i]));
    const int tris = __ldg(&a.trisize[i]);    // 1 .. trimax (k_vv_fans)
    int closed;
    f3 av;
    if (!vv_fan<TCAP, CSR>(a, w, i, pi, tris, lane, closed, av)) continue;
    int n0, n1;
    vv_rows<AX>(a, w, gzt, tris, closed, pi, lane, rd, lines, n0, n1);     // (global scratch of this warp; __syncwarp orders it)
    vv_lines<AX, WIDE>(a, w, gzt, tris, closed, pi, lane, lines, n0, n1, xres, i);
    __syncwarp();
  }
}

static int vert_volume_impl(cx_ctx *ctx, const cx_params *p, const float *pos_d, const float *norm_d, const int32_t *ep_d,
                            float *vol_d, const int32_t *trisize_d, const int32_t *cell_idx_d, int64_t nvert, int64_t nbe,
                            int zero_open, int64_t v_begin, int64_t v_end, int64_t chunk, int nparts, int part);

extern "C" int cx_calc_vert_volume(cx_ctx *ctx, const cx_params *p, const float *pos_d, const float *norm_d, const int32_t *ep_d,
                                   float *vol_d, const int32_t *trisize_d, const int32_t *cell_idx_d, int64_t nvert, int64_t nbe,
                                   int zero_open, int64_t v_begin, int64_t v_end)
{
  return vert_volume_impl(ctx, p, pos_d, norm_d, ep_d, vol_d, trisize_d, cell_idx_d, nvert, nbe, zero_open, v_begin, v_end,
                          v_end > v_begin ? v_end - v_begin : 1, 1, 0);
}

// Interleaved sharding for multi-GPU runs: the vertices are cut into chunks of `chunk`; this call computes chunks
// part, part + nparts, part + 2 nparts, ...  (the cost per vertex varies along the mesh, contiguous ranges balance badly).
// Only the owned entries of vol_d are written.
extern "C" int cx_calc_vert_volume_part(cx_ctx *ctx, const cx_params *p, const float *pos_d, const float *norm_d, const int32_t *ep_d,
                                        float *vol_d, const int32_t *trisize_d, const int32_t *cell_idx_d, int64_t nvert, int64_t nbe,
                                        int zero_open, int64_t chunk, int nparts, int part)
{
  if (chunk <= 0 || nparts <= 0 || part < 0 || part >= nparts) return CX_WRONG_INPUT;
  return vert_volume_impl(ctx, p, pos_d, norm_d, ep_d, vol_d, trisize_d, cell_idx_d, nvert, nbe, zero_open, 0, nvert, chunk, nparts, part);
}

template <class K>
static cudaError_t vv_launch(cx_ctx *ctx, K kernel, const VVArgs &a, int64_t items)
{
  int bps = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kernel, VV_WARPS * 32, 0);
  if (e != cudaSuccess) return e;
  if (bps < 1) bps = 1;
  int64_t grid = (int64_t)ctx->num_sms * bps;
  const int64_t want = (items + VV_WARPS - 1) / VV_WARPS;
  if (grid > want) grid = want;
  if (grid < 1) grid = 1;
  kernel<<<(unsigned)grid, VV_WARPS * 32, 0, ctx->stream>>>(a);
  ctx->launches++;
  return cudaGetLastError();
}

template <int AX>
static cudaError_t vv_launch_axis(cx_ctx *ctx, const VVArgs &a, bool wide, int64_t items)
{
  cudaError_t e = vv_launch(ctx, k_vv_rows<AX>, a, items);
  if (e != cudaSuccess) return e;
  return wide ? vv_launch(ctx, k_vv_lines<AX, true>, a, items) : vv_launch(ctx, k_vv_lines<AX, false>, a, items);
}

static int vert_volume_impl(cx_ctx *ctx, const cx_params *p, const float *pos_d, const float *norm_d, const int32_t *ep_d,
                            float *vol_d, const int32_t *trisize_d, const int32_t *cell_idx_d, int64_t nvert, int64_t nbe,
                            int zero_open, int64_t v_begin, int64_t v_end, int64_t chunk, int nparts, int part)
{
  if (!ctx || !p || !pos_d || !norm_d || !ep_d || !vol_d || !trisize_d || !cell_idx_d) return CX_WRONG_INPUT;
  if (v_begin < 0 || v_end > nvert || v_begin > v_end || nbe < 0) return CX_WRONG_INPUT;
  if (v_begin == v_end) return CX_OK;
  VVArgs a;
  a.p = *p;
  a.pos = (const float4 *)pos_d; a.norm = (const float4 *)norm_d; a.ep = (const int4 *)ep_d;
  a.vol = vol_d; a.trisize = trisize_d; a.cell_idx = cell_idx_d;
  a.v_begin = v_begin; a.v_end = v_end; a.zero_open = zero_open;
  a.chunk = chunk; a.nparts = nparts; a.part = part;
  {
    const int64_t n = v_end - v_begin, period = chunk * nparts;
    const int64_t full = n / period, rem = n - full * period;
    int64_t tail = rem - (int64_t)part * chunk;
    tail = tail < 0 ? 0 : (tail > chunk ? chunk : tail);
    a.nown = full * chunk + tail;
  }
  if (a.nown == 0) return CX_OK;
  const float dr = p->dr, eps = p->eps;
  const float gdr = (float)(dr / 2.0 / (float)CX_GRES);               // crixus_d.cu:281
  for (int m = 0; m < VV_GS; m++) a.gz[m] = ((float)m - (float)(VV_GS - 1) / 2) * gdr;  // :493-495
  const double thr_d = dr / 2. + eps;                                  // :535 (double)
  float thr = (float)thr_d;
  if ((double)thr > thr_d) thr = nextafterf(thr, -INFINITY);
  a.thr_cube = thr;
  a.eps = eps;
  a.two_dr = 2 * dr;
  const double gd = dr / 2.0 / (float)CX_GRES;
  a.gd3 = pow(gd, 3);
  int cmax = 0;
  while (cmax < 31 && !(ldexpf(1.0f, -(cmax + 1)) < eps)) cmax++;
  a.cmax = cmax;
  // CX_VV_AXIS=0/1/2 forces one line direction for every vertex (testing: every direction gives the same bits)
  const char *force_axis = getenv("CX_VV_AXIS");
  a.force_axis = (force_axis && force_axis[0] >= '0' && force_axis[0] <= '2') ? force_axis[0] - '0' : -1;
  // see plane_kill<WIDE>: below this dr several voxels of a cut line lie within eps of a plane as a rule
  const char *force_wide = getenv("CX_VV_WIDE");
  const bool wide = force_wide ? (force_wide[0] == '1') : (dr < 3.0e-3f);
  // The owned vertices are processed in chunks whose scratch (heads, line lists, plane store) stays within ~4 GB.  The plane
  // store holds 1.5 x the mean fan size per vertex (mean = 3 nbe / nvert: every segment sits in three fans); a chunk whose fans
  // are larger than that overflows into the monolithic kernel's list.
  const double mean_fan = nvert > 0 && nbe > 0 ? 3.0 * (double)nbe / (double)nvert : 6.0;
  const double tri_per_vertex = fmin(fmax(1.5 * mean_fan, 4.0), (double)VV_TCAP_SMALL);
  const double bytes_per_vertex = sizeof(int4) + VV_LSTRIDE * sizeof(unsigned short) + tri_per_vertex * VV_PLANE_F4 * sizeof(float4);
  static const char *chunk_env = getenv("CX_VV_CHUNK");      // vertices per chunk (testing)
  int64_t cmax_v = chunk_env ? atoll(chunk_env) : (int64_t)(4.0e9 / bytes_per_vertex);
  if (cmax_v < 1) cmax_v = 1;
  const int64_t cv = a.nown < cmax_v ? a.nown : cmax_v;
  a.plane_cap = (int)fmin(tri_per_vertex * (double)cv + 64.0, 2.0e9);
  if (const char *cap_env = getenv("CX_VV_PLANE_CAP")) a.plane_cap = atoi(cap_env) > 0 ? atoi(cap_env) : a.plane_cap;   // testing: overflow path
  auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t ctl_bytes = 256, head_bytes = up(sizeof(int4) * (size_t)cv), rd_bytes = up(sizeof(unsigned short) * VV_LSTRIDE * ((size_t)cv + VV_WARPS * (size_t)ctx->num_sms)),
               plane_bytes = up(sizeof(float4) * VV_PLANE_F4 * (size_t)a.plane_cap), list_bytes = up(sizeof(int) * 4 * (size_t)a.nown);
  // CX_VV_SCAN=1: fans by scanning the 27 cells (the form without the adjacency; also taken when segment * 4 + corner overflows)
  const char *scan_env = getenv("CX_VV_SCAN");
  const bool csr = !(scan_env && scan_env[0] == '1') && nbe > 0 && nbe < ((int64_t)1 << 29);
  const size_t start_bytes = csr ? up(sizeof(int) * ((size_t)nvert + 1)) : 0, adj_bytes = csr ? up(sizeof(int) * 3 * (size_t)nbe) : 0;
  char *tmp = nullptr;
  CX_CUDA(ctx, cudaMallocAsync(&tmp, ctl_bytes + head_bytes + rd_bytes + plane_bytes + list_bytes + 2 * start_bytes + adj_bytes, ctx->stream));
  a.ctl = (VVCtl *)tmp;
  a.heads = (int4 *)(tmp + ctl_bytes);
  a.lines = (unsigned short *)(tmp + ctl_bytes + head_bytes);
  a.lines_big = a.lines + (size_t)cv * VV_LSTRIDE;
  a.planes = (float4 *)(tmp + ctl_bytes + head_bytes + rd_bytes);
  a.list = (int *)(tmp + ctl_bytes + head_bytes + rd_bytes + plane_bytes);
  auto fail = [&](const char *what, cudaError_t e) { cudaFreeAsync(tmp, ctx->stream); return cx_fail(ctx, CX_CUDA_ERROR, what, e); };
  cudaError_t e = cudaMemsetAsync(tmp, 0, ctl_bytes, ctx->stream);
  if (e != cudaSuccess) return fail("calc_vert_volume: memset", e);
  a.adj_start = nullptr;
  a.adj = nullptr;
  if (csr) {
    int *start = (int *)(tmp + ctl_bytes + head_bytes + rd_bytes + plane_bytes + list_bytes);
    int *cursor = (int *)((char *)start + start_bytes), *adj = (int *)((char *)start + 2 * start_bytes);
    if ((e = cudaMemsetAsync(start, 0, 2 * start_bytes, ctx->stream)) != cudaSuccess) return fail("calc_vert_volume: memset", e);
    const int gb = cx_blocks(nbe, 256, (int64_t)ctx->num_sms * 16);
    k_vv_adj_count<<<gb, 256, 0, ctx->stream>>>(a.ep, nbe, nvert, start);
    ctx->launches++;
    if ((e = cudaGetLastError()) != cudaSuccess) return fail("k_vv_adj_count", e);
    const int rc = cx_exclusive_scan_i32(ctx, start, start, nvert + 1);   // in place; start[nvert] = number of entries
    if (rc != CX_OK) { cudaFreeAsync(tmp, ctx->stream); return rc; }
    k_vv_adj_fill<<<gb, 256, 0, ctx->stream>>>(a.ep, nbe, nvert, start, cursor, adj);
    ctx->launches++;
    if ((e = cudaGetLastError()) != cudaSuccess) return fail("k_vv_adj_fill", e);
    a.adj_start = start;
    a.adj = adj;
  }
  for (int64_t c0 = 0; c0 < a.nown; c0 += cv) {
    a.c_begin = c0;
    a.c_end = c0 + cv < a.nown ? c0 + cv : a.nown;
    if (c0 > 0) {   // new chunk: everything but the monolithic kernel's list length and the error bits starts over
      if ((e = cudaMemsetAsync(a.ctl->work, 0, sizeof(a.ctl->work), ctx->stream)) != cudaSuccess) return fail("calc_vert_volume: memset", e);
      if ((e = cudaMemsetAsync(a.ctl->count, 0, 3 * sizeof(int), ctx->stream)) != cudaSuccess) return fail("calc_vert_volume: memset", e);
      if ((e = cudaMemsetAsync(&a.ctl->bump, 0, sizeof(int), ctx->stream)) != cudaSuccess) return fail("calc_vert_volume: memset", e);
    }
    const int64_t items = a.c_end - a.c_begin;
    if ((e = csr ? vv_launch(ctx, k_vv_fans<true>, a, items) : vv_launch(ctx, k_vv_fans<false>, a, items)) != cudaSuccess) return fail("k_vv_fans", e);
    if ((e = vv_launch_axis<2>(ctx, a, wide, items)) != cudaSuccess) return fail("k_vv_rows/lines<z>", e);
    if ((e = vv_launch_axis<1>(ctx, a, wide, items)) != cudaSuccess) return fail("k_vv_rows/lines<y>", e);
    if ((e = vv_launch_axis<0>(ctx, a, wide, items)) != cudaSuccess) return fail("k_vv_rows/lines<x>", e);
  }
  // last launch: the (rare) vertices with more than 16 triangles, work list built by k_vv_fans
  if (wide && csr) k_vert_volume<CX_TRIMAX, true, 2, true><<<ctx->num_sms, VV_WARPS * 32, 0, ctx->stream>>>(a);
  else if (wide) k_vert_volume<CX_TRIMAX, true, 2, false><<<ctx->num_sms, VV_WARPS * 32, 0, ctx->stream>>>(a);
  else if (csr) k_vert_volume<CX_TRIMAX, false, 2, true><<<ctx->num_sms, VV_WARPS * 32, 0, ctx->stream>>>(a);
  else k_vert_volume<CX_TRIMAX, false, 2, false><<<ctx->num_sms, VV_WARPS * 32, 0, ctx->stream>>>(a);
  ctx->launches++;
  if ((e = cudaGetLastError()) != cudaSuccess) return fail("k_vert_volume", e);
  e = cudaMemcpyAsync(ctx->h_mail + 17, &a.ctl->err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
  if (e != cudaSuccess) return fail("calc_vert_volume: read-back", e);
  CX_CUDA(ctx, cudaFreeAsync(tmp, ctx->stream));
  CX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
#ifdef VV_STATS
  {
    unsigned long long st[16];
    cudaMemcpyFromSymbol(st, vv_stats, sizeof(st));
    fprintf(stderr, "VV_STATS verts=%lld pass1=%llu pass2=%llu searches=%llu cut_lanes=%llu exact=%llu exact_lanes=%llu alive_lines=%llu "
            "hw_inc=%llu hve_inc=%llu idle_it=%llu cube_it=%llu cube_search=%llu batches=%llu gV=%llu gE=%llu gW=%llu\n", (long long)(v_end - v_begin), st[0], st[1],
            st[2], st[3], st[4], st[5], st[6], st[7], st[8], st[9], st[10], st[11], st[12], st[13], st[14], st[15]);
    memset(st, 0, sizeof(st));
    cudaMemcpyToSymbol(vv_stats, st, sizeof(st));
  }
#endif
  const int err = *(int *)(ctx->h_mail + 17);
  if (err & VVE_TRIMAX) return cx_fail(ctx, CX_TRIMAX_EXCEEDED, "calc_vert_volume: a vertex has more than trimax triangles");
  if (err & VVE_FAN) return cx_fail(ctx, CX_FAN_INCOMPLETE, "calc_vert_volume: fan triangles missing from the 27-cell neighbourhood");
  if (err & VVE_NONMANIFOLD) ctx->last_error = "calc_vert_volume: warning: an edge is shared by more than two segments";
  return CX_OK;
}

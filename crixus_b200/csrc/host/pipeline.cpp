// crixus_b200 -- the whole hot path on host buffers: the sequence of crixus_main, /root/reference/src/crixus.cpp /
// src/crixus.cu:184-962, expressed through the C-ABI stage entry points.  This is what the CLI calls and what
// bench.py times end to end (host->device and device->host copies included).
#include <cuda_runtime.h>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <vector>
#include "internal.h"

namespace {
double now_s()
{
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
// bump allocation out of one arena block (256-byte aligned sub-buffers)
struct Bump {
  size_t off = 0;
  size_t take(size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; }
};
struct Buf {
  size_t off = 0;
  void *p = nullptr;
  template <class T> T *as() { return (T *)p; }
};
struct CopyStream {   // second stream + event for the overlapped download; released on every exit path
  cudaStream_t s = nullptr;
  cudaEvent_t e = nullptr;
  ~CopyStream() { if (s) cudaStreamDestroy(s); if (e) cudaEventDestroy(e); }
};
struct NormEvent {    // "the normals are on the device"; released on every exit path
  cudaEvent_t e = nullptr;
  ~NormEvent() { if (e) cudaEventDestroy(e); }
};
struct Events {
  cudaEvent_t e[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  ~Events() { for (int k = 0; k < 8; k++) if (e[k]) cudaEventDestroy(e[k]); }
};
struct DevTmp {       // stream-ordered temporary, released on every exit path
  void *p = nullptr;
  cudaStream_t st = nullptr;
  cudaError_t alloc(size_t bytes, cudaStream_t s) { st = s; return cudaMallocAsync(&p, bytes, s); }
  void *release() { void *q = p; p = nullptr; return q; }
  ~DevTmp() { if (p) cudaFreeAsync(p, st); }
};
#define RC(call) do { int rc__ = (call); if (rc__ != CX_OK) return rc__; } while (0)
#define CU(call) do { if ((call) != cudaSuccess) return CX_CUDA_ERROR; } while (0)
}  // namespace

// The result arrays live in a pinned arena owned by the context (valid until the next cx_preprocess_host on the same
// context or cx_destroy), so there is nothing to release here; the pointers are cleared.
extern "C" void cx_job_free(cx_job *job)
{
  if (!job) return;
  job->pos = job->norm = job->surf = job->vol = nullptr;
  job->ep = job->cell_idx = nullptr;
  job->fpos = nullptr;
  job->sbid = nullptr;
  job->d_pos = job->d_norm = job->d_surf = job->d_vol = nullptr;
  job->d_ep = job->d_sbid = nullptr;
  job->d_fpos = nullptr;
}

extern "C" int64_t cx_job_sizeof(void) { return (int64_t)sizeof(cx_job); }

// bounding box + weld of an auxiliary grid (special-boundary or free-surface STL) on the device
static int weld_grid(cx_ctx *ctx, cudaStream_t st, const float *corners, int64_t n, float dr, float **pos4_d, int32_t **ep4_d,
                     int64_t *nvert)
{
  DevTmp c, pos, ep;
  CU(c.alloc(36 * (size_t)n, st));
  CU(pos.alloc(48 * (size_t)n, st));
  CU(ep.alloc(16 * (size_t)n, st));
  float *c_d = (float *)c.p;
  CU(cudaMemcpyAsync(c_d, corners, 36 * (size_t)n, cudaMemcpyHostToDevice, st));
  float lo[3], hi[3];
  RC(cx_bbox_device(ctx, c_d, 3 * n, lo, hi));
  RC(cx_weld_device(ctx, c_d, 3 * n, dr, lo, hi, (float *)pos.p, (int32_t *)ep.p, nvert));
  *pos4_d = (float *)pos.release();   // handed to the caller, who frees them after the tagging kernel
  *ep4_d = (int32_t *)ep.release();
  return CX_OK;
}

static int preprocess_impl(cx_ctx *ctx, cx_job *job, bool dist)
{
  if (!ctx || !job || !job->normals || job->nbe <= 0 || !(job->dr > 0)) return CX_WRONG_INPUT;
  const int world = dist ? cx_dist_world(ctx) : 1, rank = dist ? cx_dist_rank(ctx) : 0;
  if (!job->corners && !(job->welded_pos && job->welded_ep && job->welded_nvert > 0)) return CX_WRONG_INPUT;
  const int64_t nbe = job->nbe;
  const float dr = job->dr;
  cudaStream_t st = (cudaStream_t)cx_get_stream(ctx);
  double t0 = now_s();
  float bbmin[3] = {1e10f, 1e10f, 1e10f}, bbmax[3] = {-1e10f, -1e10f, -1e10f};
  int64_t nvert;
  float *w_corners = nullptr, *w_pos4 = nullptr;   // device temporaries of the weld (stream-ordered pool)
  int32_t *w_ep4 = nullptr;
  // second stream: the normals follow the corners over PCIe while the bounding box and the weld run (into a staging buffer of
  // the context -- the working set is laid out only once the weld has counted the vertices), and the geometry arrays go back
  // while calc_vert_volume and the fills run
  CopyStream cs;
  CU(cudaStreamCreateWithFlags(&cs.s, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&cs.e, cudaEventDisableTiming));
  NormEvent ev_norm;
  float *w_norm3 = nullptr;
  const int64_t per_n = (nbe + world - 1) / world;   // normals: uploaded in N pieces like the corners
  if (job->welded_pos) {  // the mesh as crixus_main holds it right before the upload (crixus.cu:225-245)
    nvert = job->welded_nvert;
    for (int k = 0; k < 3; k++) { bbmin[k] = job->bbmin[k]; bbmax[k] = job->bbmax[k]; }
  } else {
    // raw corners -> device; bounding box (crixus.cu:185-208) and vertex weld (crixus.cu:214) run there
    // with N ranks every rank uploads 1/N of the triangles and the pieces are all-gathered over NVLink (equal pieces of `per`
    // triangles; the tail of the last piece is padding)
    const int64_t per = (nbe + world - 1) / world;
    const int64_t t_lo = std::min(nbe, per * rank), t_hi = std::min(nbe, per * (rank + 1));
    CU(cudaMallocAsync((void **)&w_corners, 36 * (size_t)(per * world), st));
    CU(cudaMallocAsync((void **)&w_pos4, 48 * (size_t)nbe, st));
    CU(cudaMallocAsync((void **)&w_ep4, 16 * (size_t)nbe, st));
    w_norm3 = (float *)cx_internal_norm_stage(ctx, 12 * (size_t)(per_n * world));
    if (!w_norm3) return CX_CUDA_ERROR;
    if (t_hi > t_lo)
      CU(cudaMemcpyAsync(w_corners + 9 * t_lo, job->corners + 9 * t_lo, 36 * (size_t)(t_hi - t_lo), cudaMemcpyHostToDevice, st));
    {
      CU(cudaEventRecord(cs.e, st));              // behind the corners, and behind whatever read the staging buffer last
      CU(cudaStreamWaitEvent(cs.s, cs.e, 0));
      const int64_t n_lo = std::min(nbe, per_n * rank), n_hi = std::min(nbe, per_n * (rank + 1));
      if (n_hi > n_lo)
        CU(cudaMemcpyAsync(w_norm3 + 3 * n_lo, job->normals + 3 * n_lo, 12 * (size_t)(n_hi - n_lo), cudaMemcpyHostToDevice, cs.s));
      CU(cudaEventCreateWithFlags(&ev_norm.e, cudaEventDisableTiming));
      CU(cudaEventRecord(ev_norm.e, cs.s));
    }
    if (world > 1) RC(cx_internal_allgather(ctx, w_corners, 36 * per));
    RC(cx_bbox_device(ctx, w_corners, 3 * nbe, bbmin, bbmax));
    RC(cx_check_triangle_size(ctx, w_corners, nbe, dr, &job->tri_failed, &job->tri_mind, &job->tri_maxd));
    RC(cx_weld_device(ctx, w_corners, 3 * nbe, dr, bbmin, bbmax, w_pos4, w_ep4, &nvert));
    CU(cudaFreeAsync(w_corners, st));
  }
  job->nvert = nvert;
  job->t_weld_s = now_s() - t0;

  cx_params p;
  cx_params_domain(&p, dr, bbmin, bbmax);  // crixus.cu:262-283
  const int64_t ncell = p.gridn[3];
  // the lattice of the fills is known up front (crixus.cu:463-467, 698-722 depend on the bounding box only)
  cx_params pf = p;
  for (int k = 0; k < 3; k++) pf.per[k] = job->periodic[k] ? 1 : 0;
  cx_params_fill_eps(&pf);
  cx_params_container(&pf, job->use_container, job->cmin, job->cmax);
  cx_fill_dims(&pf, job->dimg);
  const int64_t G = job->dimg[0] * job->dimg[1] * job->dimg[2];
  const int64_t nwords = (G + 31) / 32;
  const int64_t fsn = (job->fs_nbe > 0 && job->fs_corners && job->fs_normals) ? job->fs_nbe : 0;
  const int nsb = (job->nsb > 0 && job->sb_corners && job->sb_nbe) ? job->nsb : 0;

  // ---- one device block and one pinned host block, carved into the working set (no per-call cudaMalloc)
  t0 = now_s();
  Bump db, hb;
  Buf pre_pos, pos, pre_norm, norm, pre_ep, ep, pre_surf, surf, cell_idx, trisize, vol, newlink, norm3, fpos, fnorm, fep, fposv, sbid;
  pre_pos.off = db.take(16 * (size_t)(nvert + nbe)); pos.off = db.take(16 * (size_t)(nvert + nbe));
  pre_norm.off = db.take(16 * (size_t)nbe); norm.off = db.take(16 * (size_t)nbe);
  pre_ep.off = db.take(16 * (size_t)nbe); ep.off = db.take(16 * (size_t)nbe);
  pre_surf.off = db.take(4 * (size_t)nbe); surf.off = db.take(4 * (size_t)nbe);
  cell_idx.off = db.take(4 * (size_t)(ncell + 1)); trisize.off = db.take(4 * (size_t)nvert);
  vol.off = db.take(4 * (size_t)nvert); newlink.off = db.take(4 * (size_t)nvert);
  norm3.off = db.take(w_norm3 ? 0 : 12 * (size_t)(per_n * world)); fpos.off = db.take(4 * (size_t)nwords);
  if (fsn) {
    fnorm.off = db.take(16 * (size_t)(nbe + fsn)); fep.off = db.take(16 * (size_t)(nbe + fsn));
    fposv.off = db.take(16 * (size_t)(nvert + 3 * fsn));
  }
  if (nsb) sbid.off = db.take(4 * (size_t)(nvert + nbe));
  char *dbase = (char *)cx_internal_dev_arena(ctx, db.off + 256);
  if (!dbase) return CX_CUDA_ERROR;
  for (Buf *b : {&pre_pos, &pos, &pre_norm, &norm, &pre_ep, &ep, &pre_surf, &surf, &cell_idx, &trisize, &vol, &newlink, &norm3, &fpos,
                 &fnorm, &fep, &fposv, &sbid})
    b->p = dbase + b->off;
  const size_t o_pos = hb.take(16 * (size_t)(nvert + nbe)), o_norm = hb.take(16 * (size_t)nbe), o_ep = hb.take(16 * (size_t)nbe),
               o_surf = hb.take(4 * (size_t)nbe), o_vol = hb.take(4 * (size_t)nvert), o_cell = hb.take(4 * (size_t)(ncell + 1)),
               o_fpos = hb.take(4 * (size_t)nwords),
               o_sbid = hb.take(nsb ? 4 * (size_t)(nvert + nbe) : 0);
  char *hbase = (char *)cx_internal_host_arena(ctx, hb.off + 256);
  if (!hbase) return CX_CUDA_ERROR;
  if (job->welded_pos) {
    CU(cudaMemcpyAsync(pre_pos.p, job->welded_pos, 16 * (size_t)nvert, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(pre_ep.p, job->welded_ep, 16 * (size_t)nbe, cudaMemcpyHostToDevice, st));
  } else {
    CU(cudaMemcpyAsync(pre_pos.p, w_pos4, 16 * (size_t)nvert, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(pre_ep.p, w_ep4, 16 * (size_t)nbe, cudaMemcpyDeviceToDevice, st));
    CU(cudaFreeAsync(w_pos4, st));
    CU(cudaFreeAsync(w_ep4, st));
  }
  if (w_norm3) {   // uploaded behind the corners (above)
    CU(cudaStreamWaitEvent(st, ev_norm.e, 0));
    if (world > 1) RC(cx_internal_allgather(ctx, w_norm3, 12 * per_n));
    RC(cx_internal_expand3(ctx, w_norm3, pre_norm.as<float>(), nbe));
  } else {
    const int64_t n_lo = std::min(nbe, per_n * rank), n_hi = std::min(nbe, per_n * (rank + 1));
    if (n_hi > n_lo)
      CU(cudaMemcpyAsync(norm3.as<float>() + 3 * n_lo, job->normals + 3 * n_lo, 12 * (size_t)(n_hi - n_lo), cudaMemcpyHostToDevice, st));
    if (world > 1) RC(cx_internal_allgather(ctx, norm3.p, 12 * per_n));
    RC(cx_internal_expand3(ctx, norm3.as<float>(), pre_norm.as<float>(), nbe));
  }
  CU(cudaMemsetAsync(trisize.p, 0, 4 * (size_t)nvert, st));
  CU(cudaMemsetAsync(vol.p, 0, 4 * (size_t)nvert, st));
  CU(cudaMemsetAsync(fpos.p, 0, 4 * (size_t)nwords, st));
  RC(cx_synchronize(ctx));
  job->t_h2d_s = now_s() - t0;

  t0 = now_s();
  Events evs;   // released on every exit path
  CxStage stage;   // NVTX range of the running stage, closed on every exit path
  cudaEvent_t *ev = evs.e;
  for (int k = 0; k < 8; k++) CU(cudaEventCreate(&ev[k]));
  CU(cudaEventRecord(ev[0], st));
  stage.next("cx: swap_normals + set_bound_elem");
  if (job->swap_normals) RC(cx_swap_normals(ctx, pre_norm.as<float>(), nbe));  // crixus.cu:310-317
  RC(cx_set_bound_elem(ctx, pre_pos.as<float>(), pre_norm.as<float>(), pre_surf.as<float>(), pre_ep.as<int32_t>(), nbe, nvert, job->zstat));
  CU(cudaEventRecord(ev[1], st));
  stage.next("cx: cell binning");
  RC(cx_bin_segments(ctx, &p, pre_pos.as<float>(), pos.as<float>(), pre_norm.as<float>(), norm.as<float>(), pre_ep.as<int32_t>(),
                     ep.as<int32_t>(), pre_surf.as<float>(), surf.as<float>(), cell_idx.as<int32_t>(), nbe, nvert));
  CU(cudaEventRecord(ev[2], st));
  stage.next("cx: periodic links");
  for (int idim = 0; idim < 3; idim++) {  // crixus.cu:391-412
    if (!job->periodic[idim]) continue;
    CU(cudaMemsetAsync(newlink.p, 0xff, 4 * (size_t)nvert, st));
    RC(cx_find_links(ctx, &p, pos.as<float>(), nvert, newlink.as<int32_t>(), idim));
    RC(cx_periodicity_links(ctx, ep.as<int32_t>(), nbe, newlink.as<int32_t>()));
  }
  for (int k = 0; k < 3; k++) p.per[k] = job->periodic[k] ? 1 : 0;  // crixus.cu:428
  CU(cudaEventRecord(ev[3], st));
  stage.end();
  // pos / norm / ep / surf / cell_idx are final from here on (crixus.cu:363-365, 413): download them on a second stream
  // while calc_vert_volume and the fills run
  job->pos = (float *)(hbase + o_pos); job->norm = (float *)(hbase + o_norm); job->ep = (int32_t *)(hbase + o_ep);
  job->surf = (float *)(hbase + o_surf); job->cell_idx = (int32_t *)(hbase + o_cell);
  CU(cudaEventRecord(cs.e, st));
  CU(cudaStreamWaitEvent(cs.s, cs.e, 0));
  cudaStream_t cst = cs.s;
  // which part of the results this rank brings to the host: everything (single GPU, or rank 0 of a gathered run), nothing
  // (other ranks of a gathered run), or its 1/N share (out_sharded)
  const bool sharded = world > 1 && job->out_sharded;
  const bool all_here = world == 1 || (!sharded && rank == 0);
  auto share = [&](int64_t n, int64_t *lo, int64_t *hi) {
    if (sharded) { const int64_t per = (n + world - 1) / world; *lo = std::min(n, per * rank); *hi = std::min(n, per * (rank + 1)); }
    else { *lo = 0; *hi = all_here ? n : 0; }
  };
  share(nbe, &job->seg_lo, &job->seg_hi);
  share(nvert, &job->vert_lo, &job->vert_hi);
  share(nwords, &job->fword_lo, &job->fword_hi);
  const int64_t s_lo = job->seg_lo, s_n = job->seg_hi - job->seg_lo, v_lo = job->vert_lo, v_n = job->vert_hi - job->vert_lo;
  if (v_n > 0) CU(cudaMemcpyAsync(job->pos + 4 * v_lo, pos.as<float>() + 4 * v_lo, 16 * (size_t)v_n, cudaMemcpyDeviceToHost, cst));
  if (s_n > 0) {
    CU(cudaMemcpyAsync(job->pos + 4 * (nvert + s_lo), pos.as<float>() + 4 * (nvert + s_lo), 16 * (size_t)s_n, cudaMemcpyDeviceToHost, cst));
    CU(cudaMemcpyAsync(job->norm + 4 * s_lo, norm.as<float>() + 4 * s_lo, 16 * (size_t)s_n, cudaMemcpyDeviceToHost, cst));
    CU(cudaMemcpyAsync(job->ep + 4 * s_lo, ep.as<int32_t>() + 4 * s_lo, 16 * (size_t)s_n, cudaMemcpyDeviceToHost, cst));
    CU(cudaMemcpyAsync(job->surf + s_lo, surf.as<float>() + s_lo, 4 * (size_t)s_n, cudaMemcpyDeviceToHost, cst));
  }
  if (all_here || (sharded && rank == 0))
    CU(cudaMemcpyAsync(job->cell_idx, cell_idx.p, 4 * (size_t)(ncell + 1), cudaMemcpyDeviceToHost, cst));
  RC(cx_calc_trisize(ctx, ep.as<int32_t>(), trisize.as<int32_t>(), nbe));
  CU(cudaEventRecord(ev[4], st));
  stage.next("cx: calc_vert_volume");
  if (world > 1)
    RC(cx_dist_calc_vert_volume(ctx, &p, pos.as<float>(), norm.as<float>(), ep.as<int32_t>(), vol.as<float>(), trisize.as<int32_t>(),
                                cell_idx.as<int32_t>(), nvert, nbe, job->zero_open));
  else
    RC(cx_calc_vert_volume(ctx, &p, pos.as<float>(), norm.as<float>(), ep.as<int32_t>(), vol.as<float>(), trisize.as<int32_t>(),
                           cell_idx.as<int32_t>(), nvert, nbe, job->zero_open, 0, nvert));

  CU(cudaEventRecord(ev[5], st));
  stage.next("cx: special boundaries");
  cx_params_fill_eps(&p);   // crixus.cu:463-467
  // ---- special-boundary grids #1..nsb (crixus.cu:469-617); a later grid overwrites the ids of an earlier one
  job->sbid = nullptr;
  if (nsb) {
    CU(cudaMemsetAsync(sbid.p, 0, 4 * (size_t)(nvert + nbe), st));
    for (int k = 0; k < nsb; k++) {
      if (job->sb_nbe[k] <= 0 || !job->sb_corners[k]) continue;
      float *sp = nullptr;
      int32_t *se = nullptr;
      int64_t snv = 0;
      RC(weld_grid(ctx, st, job->sb_corners[k], job->sb_nbe[k], dr, &sp, &se, &snv));
      DevTmp gsp, gse;
      gsp.p = sp; gsp.st = st; gse.p = se; gse.st = st;
      RC(cx_identify_special_boundary(ctx, &p, pos.as<float>(), ep.as<int32_t>(), cell_idx.as<int32_t>(), nvert, nbe, sp, se,
                                      job->sb_nbe[k], sbid.as<int32_t>(), k + 1));
    }
    job->sbid = (int32_t *)(hbase + o_sbid);
    CU(cudaEventRecord(cs.e, st));
    CU(cudaStreamWaitEvent(cst, cs.e, 0));
    if (all_here || (sharded && rank == 0))
      CU(cudaMemcpyAsync(job->sbid, sbid.p, 4 * (size_t)(nvert + nbe), cudaMemcpyDeviceToHost, cst));
  }
  CU(cudaEventRecord(ev[6], st));
  stage.next("cx: fills");
  // ---- fluid (crixus.cu:698-961)
  cx_params_container(&p, job->use_container, job->cmin, job->cmax);
  job->nfluid_counted = 0;
  bool have_fs = false;
  int64_t fnbe = nbe, cnbe = 0;
  const float *g_norm = norm.as<float>(), *g_pos = pos.as<float>();
  const int32_t *g_ep = ep.as<int32_t>();
  for (int f = 0; f < job->nfills; f++) {
    const cx_fill_spec &fs = job->fills[f];
    if (fs.option == 2) {
      if (!have_fs && fsn > 0) {  // crixus.cu:809-930
        have_fs = true;
        // the free-surface STL is welded like the main mesh (crixus.cu:861-905) -- on the device -- and appended behind the
        // wall geometry: vertices behind the nvert wall vertices, segments behind the nbe wall segments
        float *fsp = nullptr;
        int32_t *fse = nullptr;
        int64_t fsnvert = 0;
        RC(weld_grid(ctx, st, job->fs_corners, fsn, dr, &fsp, &fse, &fsnvert));
        DevTmp gfp, gfe, gfn;
        gfp.p = fsp; gfp.st = st; gfe.p = fse; gfe.st = st;
        CU(gfn.alloc(12 * (size_t)fsn, st));
        CU(cudaMemcpyAsync(gfn.p, job->fs_normals, 12 * (size_t)fsn, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(fnorm.p, norm.p, 16 * (size_t)nbe, cudaMemcpyDeviceToDevice, st));
        CU(cudaMemcpyAsync(fep.p, ep.p, 16 * (size_t)nbe, cudaMemcpyDeviceToDevice, st));
        CU(cudaMemcpyAsync(fposv.p, pos.p, 16 * (size_t)nvert, cudaMemcpyDeviceToDevice, st));
        RC(cx_internal_expand3(ctx, (const float *)gfn.p, fnorm.as<float>() + 4 * nbe, fsn));
        CU(cudaMemcpyAsync(fposv.as<float>() + 4 * nvert, fsp, 16 * (size_t)fsnvert, cudaMemcpyDeviceToDevice, st));
        RC(cx_internal_offset_ep4(ctx, fse, fep.as<int32_t>() + 4 * nbe, fsn, (int32_t)nvert));
        g_norm = fnorm.as<float>(); g_ep = fep.as<int32_t>(); g_pos = fposv.as<float>();
        fnbe = nbe + fsn; cnbe = fsn;
      }
      p.dr_wall = fs.dr_wall > 0 ? fs.dr_wall : dr;  // crixus.cu:795
      const int64_t seed = cx_seed_index(&p, fs.seed);
      if (seed < 0) return CX_SEED_OUTSIDE;
      int64_t nfi = 0;
      int32_t rounds = 0;
      if (world > 1)
        RC(cx_dist_fill_geometry(ctx, &p, fpos.as<uint32_t>(), g_norm, g_ep, g_pos, fnbe, cnbe, cell_idx.as<int32_t>(), seed, &nfi, &rounds));
      else
        RC(cx_fill_geometry(ctx, &p, fpos.as<uint32_t>(), g_norm, g_ep, g_pos, fnbe, cnbe, cell_idx.as<int32_t>(), seed, &nfi, &rounds));
      job->nfluid_counted += nfi;
    } else {
      const float *lo = fs.lo, *hi = fs.hi;  // crixus.cu:759-773
      if (hi[0] - lo[0] < 1e-5 * dr || hi[1] - lo[1] < 1e-5 * dr || hi[2] - lo[2] < 1e-5 * dr) return CX_FLUID_NDEF;
      for (int k = 0; k < 3; k++)
        if (lo[k] + p.eps < p.fcmin[k] || hi[k] - p.eps > p.fcmax[k]) return CX_FLUID_BBOX_TOO_SMALL;
      int64_t nfi = 0;
      RC(cx_fill_box(ctx, &p, fpos.as<uint32_t>(), lo, hi, &nfi));
      job->nfluid_counted += nfi;
    }
  }
  CU(cudaEventRecord(ev[7], st));
  stage.end();
  RC(cx_synchronize(ctx));
  job->t_gpu_s = now_s() - t0;
  for (int k = 0; k < 5; k++) { CU(cudaEventElapsedTime(&job->stage_ms[k], ev[k], ev[k + 1])); }
  CU(cudaEventElapsedTime(&job->stage_ms[5], ev[6], ev[7]));
  CU(cudaEventElapsedTime(&job->stage_ms[6], ev[5], ev[6]));
  job->params_fill = p;
  job->d_pos = pos.as<float>(); job->d_norm = norm.as<float>(); job->d_ep = ep.as<int32_t>(); job->d_surf = surf.as<float>();
  job->d_vol = vol.as<float>(); job->d_fpos = fpos.as<uint32_t>(); job->d_sbid = nsb ? sbid.as<int32_t>() : nullptr;

  // ---- results to the host (crixus.cu:363-365, 413, 456, 962): pinned arena, asynchronous copies, one synchronise
  t0 = now_s();
  job->vol = (float *)(hbase + o_vol);
  job->fpos = (uint32_t *)(hbase + o_fpos);
  if (v_n > 0) CU(cudaMemcpyAsync(job->vol + v_lo, vol.as<float>() + v_lo, 4 * (size_t)v_n, cudaMemcpyDeviceToHost, st));
  if (job->fword_hi > job->fword_lo)
    CU(cudaMemcpyAsync(job->fpos + job->fword_lo, fpos.as<uint32_t>() + job->fword_lo, 4 * (size_t)(job->fword_hi - job->fword_lo),
                       cudaMemcpyDeviceToHost, st));
  uint64_t hp[2];
  RC(cx_hash_u32(ctx, fpos.p, nwords, hp));   // popcount on the device (every rank holds the complete bitfield); synchronises st
  job->nfluid = (int64_t)hp[1];
  CU(cudaStreamSynchronize(cst));
  job->t_d2h_s = now_s() - t0;
  return CX_OK;
}

extern "C" int cx_preprocess_host(cx_ctx *ctx, cx_job *job) { return preprocess_impl(ctx, job, false); }

// every rank of the context's communicator calls this with the same job (include/crixus_b200.h)
extern "C" int cx_preprocess_dist(cx_ctx *ctx, cx_job *job) { return preprocess_impl(ctx, job, true); }

/* crixus_b200 -- C-ABI of the B200-native Crixus preprocessing hot path.
 *
 * The reference (Azrael3000/Crixus) has no plugin/FFI surface: its hot path is the set of __global__ kernels in
 * src/crixus_d.cu, launched only from crixus_main (src/crixus.cu:31-1237), parameterised through __constant__
 * symbols (src/crixus_d.cu:13-32).  This header is the drop-in boundary for that path: one entry point per
 * reference kernel (same argument meaning, device pointers + sizes), the __constant__ symbols replaced by a
 * cx_params struct passed by pointer, silent device-side `return`s replaced by error codes in the style of
 * src/return.h:4-21.  No torch / C++ types cross this boundary.
 *
 * Array layouts are the reference's (src/crixus_d.cuh:8-16, src/crixus.cu:225-245):
 *   pos   float4[nvert+nbe]  vertices then segment barycentres (.w unused)
 *   norm  float4[nbe]        STL facet normals (.w unused)
 *   ep    int4[nbe]          the 3 vertex ids of a segment (.w unused)
 *   surf  float[nbe]         segment areas
 *   cell_idx int[ncell3+1]   exclusive prefix of segments per 3dr-cell, cell = gx*ny*nz + gy*nz + gz
 *   fpos  uint32[ceil(G/32)] fluid bitfield, LSB first, cell = i + j*idimg + k*idimg*jdimg
 * All pointers named *_d are DEVICE pointers on the context's device.  Calls are asynchronous on the context's
 * stream unless they return a host value (those synchronise the stream).
 */
#ifndef CRIXUS_B200_H
#define CRIXUS_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define CX_GRES 20   /* src/crixus.h:26  voxels per dr/2 */
#define CX_TRIMAX 50 /* src/crixus.h:27  max triangles per vertex */

/* error codes: src/return.h:4-21 where one exists, own codes below -100 */
enum {
  CX_OK = 0,
  CX_STL_NOT_BINARY = -1, CX_NO_FILE = -2, CX_FILE_NOT_OPEN = -3, CX_CANT_READ_CONFIG = -4, CX_READ_ERROR = -5,
  CX_FLUID_NDEF = -6, CX_WRITE_FAIL = -7, CX_NO_PER_VERT = -9, CX_WRONG_INPUT = -11, CX_NO_WRITE_FILE = -16,
  CX_FLUID_BBOX_TOO_SMALL = -17, CX_INTERNAL_ERROR = -18,
  CX_CUDA_ERROR = -101,      /* a CUDA runtime call failed; see cx_last_error */
  CX_TRIMAX_EXCEEDED = -102, /* a vertex has more than CX_TRIMAX triangles (reference: thread returns, crixus_d.cu:306) */
  CX_FAN_INCOMPLETE = -103,  /* a vertex' triangles were not all found in its 27 cells (mesh precondition violated) */
  CX_SEED_OUTSIDE = -104,    /* geometry-fill seed outside the fluid lattice (reference: out-of-bounds write) */
  CX_NO_DEVICE = -105,
  CX_WELD_CELL_TOO_FULL = -106, /* (no longer returned: the device weld handles cells of any size) */
  CX_NO_NCCL = -107            /* a multi-GPU communicator was requested but libnccl.so.2 cannot be loaded */
};

/* the __constant__ symbols of src/crixus_d.cu:13-32 */
typedef struct cx_params {
  float dr;       /* particle spacing */
  float dr_wall;  /* min distance fluid particle <-> segment (fill_N.dr_wall) */
  float eps;      /* NOTE: dr/gres*1e-4 up to calc_vert_volume, max(extent)*1e-5 for the fill (crixus.cu:265 vs :464-467) */
  float dmin[4], dmax[4]; /* domain bounding box (all raw STL corners) */
  float griddr[4];        /* 3dr-cell edge lengths */
  int32_t gridn[4];       /* 3dr-cell counts, [3] = product */
  float fcmin[4], fcmax[4]; /* fluid container */
  int32_t per[3];           /* periodicity flags */
} cx_params;

typedef struct cx_ctx cx_ctx;

/* ---- host-side glue of crixus_main, exact restatements (no device work) */
void cx_params_domain(cx_params *p, float dr, const float bbmin[3], const float bbmax[3]); /* crixus.cu:262-283 */
void cx_params_fill_eps(cx_params *p);                                                      /* crixus.cu:464-467 */
void cx_params_container(cx_params *p, int use, const float cmin[3], const float cmax[3]);  /* crixus.cu:698-722 */
void cx_fill_dims(const cx_params *p, int64_t dimg[3]);                 /* crixus_d.cu:985-987 / crixus.cu:802-803 */
int64_t cx_seed_index(const cx_params *p, const float spos[3]);         /* crixus.cu:799-804; -1 if outside */

/* ---- context */
int cx_create(int device, cx_ctx **ctx);   /* selects the device, creates a stream */
void cx_destroy(cx_ctx *ctx);
int cx_set_stream(cx_ctx *ctx, void *cuda_stream); /* run on a caller-owned cudaStream_t (e.g. torch's current stream) */
int cx_synchronize(cx_ctx *ctx);
void *cx_get_stream(cx_ctx *ctx);            /* the cudaStream_t the context launches on */
const char *cx_last_error(cx_ctx *ctx);
const char *cx_version(void);
int64_t cx_kernel_launches(cx_ctx *ctx);   /* number of kernels launched through this context so far */
int cx_measure_fp32_peak(cx_ctx *ctx, int reps, double *tflops); /* FFMA micro-benchmark: roofline denominator */

/* ---- stage entry points, one per reference kernel */

/* swap_normals, crixus_d.cu:203-211 */
int cx_swap_normals(cx_ctx *ctx, float *norm_d, int64_t nbe);

/* set_bound_elem, crixus_d.cu:105-201.  zstat_host[4] = {xminp, xminn, nminp, nminn} (may be NULL). */
int cx_set_bound_elem(cx_ctx *ctx, float *pos_d, const float *norm_d, float *surf_d, int32_t *ep_d, int64_t nbe,
                      int64_t nvert, float *zstat_host);

/* init_cell_idx + count_cell_bes + add_up_indices + sort_bes, crixus_d.cu:37-103, with a STABLE scatter
 * (original segment order inside a cell; the reference's order is run-dependent, crixus_d.cu:94). */
int cx_bin_segments(cx_ctx *ctx, const cx_params *p, const float *pre_pos_d, float *pos_d, const float *pre_norm_d,
                    float *norm_d, const int32_t *pre_ep_d, int32_t *ep_d, const float *pre_surf_d, float *surf_d,
                    int32_t *cell_idx_d, int64_t nbe, int64_t nvert);

/* find_links, crixus_d.cu:213-237: pairs max-face vertices with min-face vertices in dimension idim, writes newlink
 * (must be pre-set to -1), deletes paired vertices (pos = -1e10).  Returns CX_NO_PER_VERT if a partner is missing. */
int cx_find_links(cx_ctx *ctx, const cx_params *p, float *pos_d, int64_t nvert, int32_t *newlink_d, int idim);

/* periodicity_links, crixus_d.cu:240-256 */
int cx_periodicity_links(cx_ctx *ctx, int32_t *ep_d, int64_t nbe, const int32_t *newlink_d);

/* calc_trisize, crixus_d.cu:258-267 (trisize_d must be zeroed by the caller, as in crixus.cu:423-427) */
int cx_calc_trisize(cx_ctx *ctx, const int32_t *ep_d, int32_t *trisize_d, int64_t nbe);

/* calc_vert_volume, crixus_d.cu:271-678, for vertices [v_begin, v_end) (index-range sharding; whole mesh: 0, nvert) */
int cx_calc_vert_volume(cx_ctx *ctx, const cx_params *p, const float *pos_d, const float *norm_d, const int32_t *ep_d,
                        float *vol_d, const int32_t *trisize_d, const int32_t *cell_idx_d, int64_t nvert, int64_t nbe,
                        int zero_open, int64_t v_begin, int64_t v_end);

/* interleaved sharding of calc_vert_volume for multi-GPU runs: vertices are cut into chunks of `chunk`, this call
 * computes chunks part, part+nparts, ... and writes only those entries of vol_d (the others are left untouched) */
int cx_calc_vert_volume_part(cx_ctx *ctx, const cx_params *p, const float *pos_d, const float *norm_d, const int32_t *ep_d,
                             float *vol_d, const int32_t *trisize_d, const int32_t *cell_idx_d, int64_t nvert, int64_t nbe,
                             int zero_open, int64_t chunk, int nparts, int part);

/* fill_fluid, crixus_d.cu:680-744 (box).  *nfi_host = cells in the box (counted even when already set). */
int cx_fill_box(cx_ctx *ctx, const cx_params *p, uint32_t *fpos_d, const float lo[3], const float hi[3],
                int64_t *nfi_host);

/* fill_fluid_complex driven to its fixed point (crixus_d.cu:965-1058 + loop crixus.cu:934-947) by a frontier/tile
 * flood fill.  nbe counts wall + fshape segments, the last cnbe of them are the fshape ones (crixus_d.cu:951).
 * seed = cell index from cx_seed_index (or -1 for none).  *nfi_host = newly set cells, *rounds_host = global rounds. */
int cx_fill_geometry(cx_ctx *ctx, const cx_params *p, uint32_t *fpos_d, const float *norm_d, const int32_t *ep_d,
                     const float *pos_d, int64_t nbe, int64_t cnbe, const int32_t *cell_idx_d, int64_t seed,
                     int64_t *nfi_host, int32_t *rounds_host);

/* z-slab variant for multi-GPU runs: this rank owns lattice planes [k_begin, k_end); fpos_d is still the full packed
 * bitfield on every rank (<= 1.25 GB at 1e10 cells) but only owned planes are written.  One call = propagate inside
 * the slab until locally converged, given the current state of the two halo planes (k_begin-1, k_end) in fpos_d.
 * *changed_host = number of cells newly set in the slab.  The caller exchanges boundary planes between calls.
 * first_call: 1 = start of a fill (directed tests of the slab are resolved here), 0 = continuation, 2 = start of a fill whose
 * directed tests were resolved by cx_fill_resolve_part / _slab and merged by the caller (balanced; see below).  Only the
 * owned planes and one halo plane either side are touched, so a rank's cost does not grow with the whole lattice. */
int cx_fill_geometry_slab(cx_ctx *ctx, const cx_params *p, uint32_t *fpos_d, const float *norm_d, const int32_t *ep_d,
                          const float *pos_d, int64_t nbe, int64_t cnbe, const int32_t *cell_idx_d, int64_t seed,
                          int64_t k_begin, int64_t k_end, int first_call, int64_t *changed_host, int32_t *rounds_host);

/* Pipelining of the z-slab fill: bound the propagation rounds of one cx_fill_geometry_slab call (0 = run until the slab
 * is converged) so that boundary planes are exchanged while the flood front is still moving; cx_fill_unfinished tells
 * whether the last call stopped at the bound with active tiles left (the caller must then call again). */
int cx_fill_set_max_rounds(cx_ctx *ctx, int max_rounds);
int cx_fill_unfinished(cx_ctx *ctx);
/* how the directed tests of the geometry fill (checkCollision, crixus_d.cu:801-963) are resolved: 0 (default) = by dr -- at fine
 * dr, where eps >= dr_wall^2 kills the distance test (crixus_d.cu:940), the segments enumerate the cells they can block
 * (record-driven kernels); 1 = always the cell-driven kernel.  Same bits either way (tests/test_gpu_parity.py). */
int cx_fill_set_resolve_mode(cx_ctx *ctx, int mode);

/* Multi-GPU: the directed tests (pure functions of the geometry) are resolved for planes [k_begin, k_end) only;
 * *blocked_d points to the 6 "blocked" bit planes (*blocked_words uint32, zero outside this rank's share) which the caller
 * merges across ranks (every lattice cell is resolved by exactly one rank -- disjoint bits: SUM == OR).  Then either
 * cx_fill_propagate runs the flood on the whole lattice on every rank (small lattices: no halo exchange, no per-sweep
 * collective) or cx_fill_geometry_slab(first_call = 2) floods by z-slab (large lattices). */
int cx_fill_resolve_slab(cx_ctx *ctx, const cx_params *p, uint32_t *fpos_d, const float *norm_d, const int32_t *ep_d,
                         const float *pos_d, int64_t nbe, int64_t cnbe, const int32_t *cell_idx_d, int64_t k_begin,
                         int64_t k_end, uint32_t **blocked_d, int64_t *blocked_words);
/* same, but the 3dr-cells (and with them their lattice cells) are dealt to the ranks in interleaved chunks of 32 (balanced;
 * z-slabs are not: the floor slab holds a whole wall) */
int cx_fill_resolve_part(cx_ctx *ctx, const cx_params *p, uint32_t *fpos_d, const float *norm_d, const int32_t *ep_d,
                         const float *pos_d, int64_t nbe, int64_t cnbe, const int32_t *cell_idx_d, int nparts, int part,
                         uint32_t **blocked_d, int64_t *blocked_words);
int cx_fill_propagate(cx_ctx *ctx, const cx_params *p, uint32_t *fpos_d, const float *norm_d, const int32_t *ep_d,
                      const float *pos_d, int64_t nbe, int64_t cnbe, const int32_t *cell_idx_d, int64_t seed,
                      int64_t *nfi_host, int32_t *rounds_host);

/* ---- multi-GPU: one process (or host thread) per GPU.  The reference is single-GPU (src/crixus.cu:80-119 selects one device);
 * this is where a device list enters.  A context joins a communicator once; the cx_dist_* stage calls and cx_preprocess_dist then
 * run sharded over its ranks, all collectives on the context's stream:
 *   calc_vert_volume  interleaved chunks of CX_DIST_VV_CHUNK vertices per rank + one all-reduce of the volumes
 *   geometry fill     directed tests dealt to the ranks by 3dr-cell chunk, blocked planes reduced to the z-slab owners, slab
 *                     floods with boundary planes exchanged between z-neighbours every CX_DIST_FILL_ROUNDS rounds (NCCL
 *                     send/recv), termination by an all-reduce of two counters, owned planes all-gathered at the end
 *   cheap stages      (set_bound_elem, binning, links, trisize, device weld) replicated: recomputing them costs less HBM
 *                     traffic than gathering their outputs would cost NVLink traffic
 * Results are bit-identical for every rank count (the fill is a least fixed point; the volumes are per-vertex). */
#define CX_NCCL_ID_BYTES 128
#define CX_DIST_VV_CHUNK 1024
#define CX_DIST_FILL_ROUNDS 8
#define CX_DIST_SMALL_BYTES (64u << 20) /* six blocked planes below this size: all-reduce them and flood replicated */
int cx_dist_unique_id(void *id128);                                            /* ncclGetUniqueId; rank 0 calls it, the application ships the bytes */
int cx_dist_init(cx_ctx *ctx, const void *id128, int world, int rank);         /* ncclCommInitRank on the context's device */
int cx_dist_attach(cx_ctx *ctx, void *nccl_comm, int world, int rank);         /* use a ncclComm_t the caller owns */
int cx_dist_init_local(cx_ctx *const *ctxs, int n);  /* n contexts of ONE process form a group (one host thread per context must
                                                        then call the cx_dist_* functions concurrently); no NCCL involved */
int cx_dist_finalize(cx_ctx *ctx);
int cx_dist_world(cx_ctx *ctx);
int cx_dist_rank(cx_ctx *ctx);
/* same arguments as cx_calc_vert_volume / cx_fill_geometry, the (replicated) inputs resident on every rank; every rank ends with
 * all volumes / the complete bitfield.  *nfi_host = newly set cells over all ranks. */
int cx_dist_calc_vert_volume(cx_ctx *ctx, const cx_params *p, const float *pos_d, const float *norm_d, const int32_t *ep_d,
                             float *vol_d, const int32_t *trisize_d, const int32_t *cell_idx_d, int64_t nvert, int64_t nbe,
                             int zero_open);
int cx_dist_fill_geometry(cx_ctx *ctx, const cx_params *p, uint32_t *fpos_d, const float *norm_d, const int32_t *ep_d,
                          const float *pos_d, int64_t nbe, int64_t cnbe, const int32_t *cell_idx_d, int64_t seed,
                          int64_t *nfi_host, int32_t *rounds_host);
/* which form cx_dist_fill_geometry takes: 0 = by lattice size (CX_DIST_SMALL_BYTES), 1 = replicated flood, 2 = z-slab flood */
int cx_dist_set_fill_mode(cx_ctx *ctx, int mode);
/* order-independent checksum of n 32-bit words in device memory: out[0] = sum_i mix(i, word_i) mod 2^64, out[1] = popcount.
 * Used to compare results between rank counts and against the single-GPU pass without moving them. */
int cx_hash_u32(cx_ctx *ctx, const void *words_d, int64_t n, uint64_t out_host[2]);

/* identifySpecialBoundarySegments, crixus_d.cu:1085-1112 (+ segInTri :1060-1083): segments whose barycentre lies in a
 * triangle of the special-boundary grid (sbpos float4[sbnvert], sbep int4[sbnbe], welded like the main mesh,
 * crixus.cu:560) get sbid[nvert+seg] = sbi, and so do their three vertices.  sbid_d: int[nvert+nbe], zeroed by the caller
 * before the first grid (crixus.cu:590-597).  p->eps must be the fill eps (cx_params_fill_eps; crixus.cu:464-467 runs
 * first).  pos/ep/cell_idx are the cell-sorted arrays of cx_bin_segments: each grid triangle is only tested against the
 * 3dr-cells its grown bounding box touches instead of all nbe segments. */
int cx_identify_special_boundary(cx_ctx *ctx, const cx_params *p, const float *pos_d, const int32_t *ep_d,
                                 const int32_t *cell_idx_d, int64_t nvert, int64_t nbe, const float *sbpos_d,
                                 const int32_t *sbep_d, int64_t sbnbe, int32_t *sbid_d, int32_t sbi);

/* ---- whole hot path on HOST buffers (the call the CLI / a reference maintainer would make; includes H2D/D2H) */
typedef struct cx_fill_spec {
  int32_t option;      /* 1 = box, 2 = geometry (crixus.cu:737-741) */
  float lo[3], hi[3];  /* box */
  float seed[3];       /* geometry */
  float dr_wall;       /* geometry; <=0 -> dr */
} cx_fill_spec;

typedef struct cx_job {
  /* inputs */
  const float *corners;     /* nbe*9 floats: raw STL corner coordinates (may be NULL when the welded mesh is given) */
  const float *normals;     /* nbe*3 floats: raw STL facet normals */
  int64_t nbe;
  /* optional: the already welded mesh, i.e. exactly what crixus_main uploads at crixus.cu:295-297.  When welded_pos is
   * non-NULL the host weld is skipped; bbmin/bbmax (bounding box of the raw corners, crixus.cu:185-208) must be set. */
  const float *welded_pos;  /* welded_nvert*4 floats (float4) */
  const int32_t *welded_ep; /* nbe*4 ints (int4) */
  int64_t welded_nvert;
  float bbmin[3], bbmax[3];
  float dr;
  int32_t swap_normals, zero_open, periodic[3];
  int32_t use_container; float cmin[3], cmax[3];
  const float *fs_corners; const float *fs_normals; int64_t fs_nbe; /* optional fshape STL */
  const cx_fill_spec *fills; int32_t nfills;
  /* optional special-boundary grids #1..nsb (crixus.cu:469-617): sb_corners[k] = sb_nbe[k]*9 floats, raw STL corners */
  const float *const *sb_corners; const int64_t *sb_nbe; int32_t nsb;
  /* outputs: they live in a pinned host arena owned by the context and stay valid until the next cx_preprocess_host on
   * the same context (or cx_destroy); cx_job_free only clears the pointers */
  int64_t nvert, nfluid, nfluid_counted; /* popcount vs. the reference's running count (App. B-2 of SURVEY.md) */
  float *pos;  float *norm;  int32_t *ep;  float *surf;  float *vol;  int32_t *cell_idx; uint32_t *fpos;
  int32_t *sbid; /* int[nvert+nbe]: special-boundary id (KENT) of vertices, then segments; NULL when nsb == 0 */
  /* the same result arrays in HBM (context-owned, valid until the next cx_preprocess_host on this context): inputs of
   * cx_assemble_records_device, so that the particle file can be produced without a host round trip */
  float *d_pos; float *d_norm; int32_t *d_ep; float *d_surf; float *d_vol; uint32_t *d_fpos; int32_t *d_sbid;
  int64_t dimg[3];
  float zstat[4];
  cx_params params_fill;
  int64_t tri_failed; float tri_mind, tri_maxd; /* mesh precondition (cx_check_triangle_size), filled when raw corners are given */
  double t_weld_s, t_gpu_s, t_h2d_s, t_d2h_s; /* wall-clock breakdown */
  float stage_ms[8]; /* CUDA-event times: 0 swap+set_bound_elem, 1 binning, 2 periodic links, 3 trisize, 4 vert volume, 5 fills, 6 special-boundary tagging */
  /* cx_preprocess_dist: in: out_sharded; out: the ranges this rank downloaded (whole arrays on rank 0 when out_sharded = 0) */
  int32_t out_sharded;
  int64_t seg_lo, seg_hi, vert_lo, vert_hi, fword_lo, fword_hi;
} cx_job;

int cx_preprocess_host(cx_ctx *ctx, cx_job *job);
/* the same on every rank of the context's communicator (every rank passes the same job, i.e. can read the same input files):
 * each rank uploads 1/N of the raw corners and normals and the pieces are all-gathered over NVLink; calc_vert_volume and the
 * geometry fills run sharded (cx_dist_*).  job->out_sharded = 0: rank 0 downloads every result array (the other ranks none);
 * 1: every rank downloads its share -- segments [seg_lo, seg_hi) of pos(barycentres)/norm/ep/surf, vertices [vert_lo, vert_hi) of
 * pos/vol, bitfield words [fword_lo, fword_hi) -- into its own (full-size) host arrays, for a consumer that writes the particle
 * file in parallel.  nvert / nfluid / dimg are valid on every rank. */
int cx_preprocess_dist(cx_ctx *ctx, cx_job *job);
void cx_job_free(cx_job *job);
int64_t cx_job_sizeof(void); /* sizeof(cx_job) of the library, for bindings to check their mirror of the struct */

/* bounding box of the raw STL corners (src/crixus.cu:185-208) and vertex weld (remove_duplicate_vertices,
 * src/crixus.cpp:343-455) on the device: cell keys -> stable radix sort -> one thread per cell replays the claim loop ->
 * prefix sum = the reference's vertex numbering.  corners_d: ncorner*3 floats in STL order (ncorner = 3*nbe);
 * pos4_d: room for ncorner float4 (nvert are written); ep4_d: nbe int4.  Same outcome as cx_weld_host. */
int cx_bbox_device(cx_ctx *ctx, const float *corners_d, int64_t ncorner, float bbmin[3], float bbmax[3]);
int cx_weld_device(cx_ctx *ctx, const float *corners_d, int64_t ncorner, float dr, const float bbmin[3], const float bbmax[3],
                   float *pos4_d, int32_t *ep4_d, int64_t *nvert_host);

/* the mesh precondition of resources/test-triangle-size.c on the device: every vertex-barycentre distance must be
 * <= dr - 1e-4*dr.  *failed_host = number of (triangle, corner) pairs that violate it, *mind / *maxd = extreme distances */
int cx_check_triangle_size(cx_ctx *ctx, const float *corners_d, int64_t ntri, float dr, int64_t *failed_host, float *mind_host,
                           float *maxd_host);

/* host weld, src/crixus.cpp:343-455 (outcome-exact restatement).  out_verts: room for 3*ncorner floats,
 * out_ids: ncorner ints.  Returns the number of welded vertices. */
int64_t cx_weld_host(const float *corners, int64_t ncorner, float dr, float *out_verts, int32_t *out_ids);

/* ---- the reference's process surface (front end), usable without a GPU except cx_run_ini -------------------- */
/* .ini reader with the rules of the reference's vendored inih (src/ini/ini.c:60-167, src/ini/cpp/INIReader.cpp:15-75).
 * cx_ini_open returns INIReader::ParseError(): 0 ok, -1 cannot open, >0 first bad line; *out is valid in every case. */
typedef struct cx_ini cx_ini;
int cx_ini_open(const char *filename, cx_ini **out);
void cx_ini_close(cx_ini *ini);
int cx_ini_get(const cx_ini *ini, const char *section, const char *name, const char *def, char *buf, int buflen); /* -> length */
long cx_ini_get_integer(const cx_ini *ini, const char *section, const char *name, long def);
double cx_ini_get_real(const cx_ini *ini, const char *section, const char *name, double def);
int cx_ini_get_boolean(const cx_ini *ini, const char *section, const char *name, int def);

/* binary STL reader, src/crixus.cu:126-221: CX_OK / CX_FILE_NOT_OPEN / CX_STL_NOT_BINARY / CX_READ_ERROR.
 * corners: nbe*9 floats, normals: nbe*3 floats, both malloc'ed (release with cx_free). */
int cx_stl_read(const char *filename, float **corners, float **normals, int64_t *nbe);
void cx_free(void *p);

/* particle records, src/crixus.cu:973-1097: 64-byte OutBuf records (src/crixus.h:33-36) -- fluid particles in bit order,
 * alive vertices in index order, segments in sorted order, stably grouped by KENT when special boundaries are present
 * (crixus.cu:1088-1097).  Returns the record count (records may be NULL to size). */
int64_t cx_assemble_records(const cx_job *job, int64_t nfluid_for_indices, void *records, int64_t capacity);
int cx_write_vtu_records(const void *records, int64_t n, const char *filename);   /* vtk_output, src/crixus.cpp:62-259 */
int cx_write_h5sph_records(const void *records, int64_t n, const char *filename); /* hdf5_output, src/crixus.cpp:12-60 */
/* [output] split=true, src/crixus.cu:1137-1216: <outname>.fluid.<ext> holds the fluid particles, <outname>.boundary.kent<k>.<ext>
 * the vertex and segment records of special-boundary id k (k < num_kents, files without particles are skipped);
 * format 1 = vtu, 2 = h5sph (generic_output, src/crixus.cpp:261-289: "0." prefix and ".h5sph" suffix for h5sph). */
int cx_write_split_records(const void *records, int64_t n, int64_t nfluid_particles, int32_t num_kents, const char *outname, int format);

/* cx_assemble_records on the device (src/crixus.cu:973-1097): prefix sums over the bitfield popcounts / alive flags give every
 * particle its final index, one thread per bitfield word / vertex / segment writes the 64-byte records into records_d.
 * capacity: room in records_d (nfluid + nvert + nbe always suffices); records_d == NULL only counts.  *n_host = record count. */
int cx_assemble_records_device(cx_ctx *ctx, const cx_params *pfill, const int64_t dimg[3], const uint32_t *fpos_d,
                               const float *pos_d, const float *norm_d, const int32_t *ep_d, const float *surf_d,
                               const float *vol_d, const int32_t *sbid_d /* may be NULL */, int64_t nvert, int64_t nbe,
                               int64_t nfluid_for_indices, void *records_d, int64_t capacity, int64_t *n_host);
/* vtk_output / hdf5_output (src/crixus.cpp:62-259 / 12-60) from DEVICE records: the file body is laid out by a kernel and
 * streamed to the file through pinned staging buffers (copy of chunk c+1 overlaps the write of chunk c).  Same bytes as
 * cx_write_vtu_records / cx_write_h5sph_records. */
int cx_write_vtu_records_device(cx_ctx *ctx, const void *records_d, int64_t n, const char *filename);
int cx_write_h5sph_records_device(cx_ctx *ctx, const void *records_d, int64_t n, const char *filename);

/* The .h5sph file of a LARGE run without holding the records: cx_assemble_records_device needs all n 64-byte records resident (64 GB at
 * BASELINE configs[3]); here chunks of chunk_items bitfield words (32 chunk_items vertices / segments) are assembled into a bounded device
 * buffer and streamed to the file one after the other (src/crixus.cu:973-1097 + src/crixus.cpp:12-60, same bytes as the resident path).
 * chunk_items <= 0: 2^20.  Not with special boundaries (the KENT grouping permutes the segment records).  *n_host = records written. */
int cx_stream_h5sph_device(cx_ctx *ctx, const cx_params *pfill, const int64_t dimg[3], const uint32_t *fpos_d, const float *pos_d,
                           const float *norm_d, const int32_t *ep_d, const float *surf_d, const float *vol_d, int64_t nvert, int64_t nbe,
                           int64_t nfluid_for_indices, const char *filename, int64_t chunk_items, int64_t *n_host);

/* what `Crixus file.ini` does (crixus_main, src/crixus.cu:31-1237, hot path only): ini -> STL -> GPU -> output file */
int cx_run_ini(const char *ini_path);

#ifdef __cplusplus
}
#endif
#endif

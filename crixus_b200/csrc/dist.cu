// crixus_b200 -- multi-GPU plumbing behind the C-ABI: communicator handling and the collective back-ends of dist.h.
// The reference is single-GPU (src/crixus.cu:80-119 selects ONE device); this is where a device list enters.
//   cx_dist_unique_id / cx_dist_init   one process per GPU, NCCL over NVLink (ncclCommInitRank)
//   cx_dist_attach                      a communicator the caller already owns (e.g. an MPI application's)
//   cx_dist_init_local                  N contexts in one process (one host thread per rank): single-GPU emulation of N ranks
// NCCL is resolved with dlopen at that moment: single-GPU users need no NCCL, and a process that already loaded one (PyTorch)
// shares it.
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>
#include <condition_variable>
#include <mutex>
#include "dist.h"

// ------------------------------------------------------------------------------------------------ NCCL, late bound
namespace {
struct NcclApi {
  void *h = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclReduce) Reduce = nullptr;
  decltype(&ncclBroadcast) Broadcast = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  std::string err;
  bool load()
  {
    if (h) return true;
    const char *names[] = {getenv("CX_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the copy the process already uses, if any
    for (int i = 0; i < 3 && !h; i++)
      if (names[i]) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!h) { err = std::string("cannot load NCCL: ") + (dlerror() ? dlerror() : "?"); return false; }
#define CX_SYM(n) n = (decltype(n))dlsym(h, "nccl" #n); if (!n) { err = "NCCL symbol missing: nccl" #n; h = nullptr; return false; }
    CX_SYM(GetUniqueId) CX_SYM(CommInitRank) CX_SYM(CommDestroy) CX_SYM(AllReduce) CX_SYM(Reduce) CX_SYM(Broadcast)
    CX_SYM(AllGather) CX_SYM(Send) CX_SYM(Recv) CX_SYM(GroupStart) CX_SYM(GroupEnd) CX_SYM(GetErrorString)
#undef CX_SYM
    return true;
  }
};
NcclApi g_nccl;
std::mutex g_nccl_mutex;

#define CX_NCCL(ctx, call)                                                                         \
  do {                                                                                             \
    ncclResult_t r__ = (call);                                                                     \
    if (r__ != ncclSuccess) return cx_fail(ctx, CX_CUDA_ERROR, (std::string(#call ": ") + g_nccl.GetErrorString(r__)).c_str()); \
  } while (0)

struct NcclColl : CxColl {
  ncclComm_t comm = nullptr;
  bool own = false;
  ~NcclColl() override { if (own && comm) g_nccl.CommDestroy(comm); }
  int allreduce(cx_ctx *ctx, void *buf, int64_t count, ncclDataType_t t)
  {
    if (count <= 0) return CX_OK;
    CX_NCCL(ctx, g_nccl.AllReduce(buf, buf, (size_t)count, t, ncclSum, comm, ctx->stream));
    return CX_OK;
  }
  int allreduce_sum_u32(cx_ctx *ctx, uint32_t *buf, int64_t count) override { return allreduce(ctx, buf, count, ncclUint32); }
  int allreduce_sum_f32(cx_ctx *ctx, float *buf, int64_t count) override { return allreduce(ctx, buf, count, ncclFloat32); }
  int allreduce_sum_i64(cx_ctx *ctx, int64_t *buf, int64_t count) override { return allreduce(ctx, buf, count, ncclInt64); }
  int reduce_segments_u32(cx_ctx *ctx, uint32_t *base, const CxSeg *segs, int nseg) override
  {
    CX_NCCL(ctx, g_nccl.GroupStart());
    for (int i = 0; i < nseg; i++)
      if (segs[i].count > 0)
        CX_NCCL(ctx, g_nccl.Reduce(base + segs[i].offset, base + segs[i].offset, (size_t)segs[i].count, ncclUint32, ncclSum, segs[i].root,
                                   comm, ctx->stream));
    CX_NCCL(ctx, g_nccl.GroupEnd());
    return CX_OK;
  }
  int bcast_segments_u32(cx_ctx *ctx, uint32_t *base, const CxSeg *segs, int nseg) override
  {
    CX_NCCL(ctx, g_nccl.GroupStart());
    for (int i = 0; i < nseg; i++)
      if (segs[i].count > 0)
        CX_NCCL(ctx, g_nccl.Broadcast(base + segs[i].offset, base + segs[i].offset, (size_t)segs[i].count, ncclUint32, segs[i].root, comm,
                                      ctx->stream));
    CX_NCCL(ctx, g_nccl.GroupEnd());
    return CX_OK;
  }
  int halo_exchange_u32(cx_ctx *ctx, const uint32_t *up_send, uint32_t *up_recv, const uint32_t *dn_send, uint32_t *dn_recv,
                        int64_t count) override
  {
    if (count <= 0) return CX_OK;
    CX_NCCL(ctx, g_nccl.GroupStart());
    if (rank + 1 < world) {
      CX_NCCL(ctx, g_nccl.Send(up_send, (size_t)count, ncclUint32, rank + 1, comm, ctx->stream));
      CX_NCCL(ctx, g_nccl.Recv(up_recv, (size_t)count, ncclUint32, rank + 1, comm, ctx->stream));
    }
    if (rank > 0) {
      CX_NCCL(ctx, g_nccl.Send(dn_send, (size_t)count, ncclUint32, rank - 1, comm, ctx->stream));
      CX_NCCL(ctx, g_nccl.Recv(dn_recv, (size_t)count, ncclUint32, rank - 1, comm, ctx->stream));
    }
    CX_NCCL(ctx, g_nccl.GroupEnd());
    return CX_OK;
  }
  int allgather_bytes(cx_ctx *ctx, const void *send, void *recv, int64_t bytes) override
  {
    if (bytes <= 0) return CX_OK;
    CX_NCCL(ctx, g_nccl.AllGather(send, recv, (size_t)bytes, ncclUint8, comm, ctx->stream));
    return CX_OK;
  }
};

// ------------------------------------------------------------------------------------------------ in-process group
#define CX_LOCAL_MAX 16
struct LocalGroup {
  int n = 0, refs = 0;
  std::mutex m;
  std::condition_variable cv;
  int arrived = 0;
  unsigned gen = 0;
  const void *slot[CX_LOCAL_MAX][2];
  void barrier()
  {
    std::unique_lock<std::mutex> lk(m);
    const unsigned g = gen;
    if (++arrived == n) { arrived = 0; gen++; cv.notify_all(); }
    else cv.wait(lk, [&] { return gen != g; });
  }
};

struct PtrList { const void *p[CX_LOCAL_MAX]; int n; };

template <class T>
__global__ void __launch_bounds__(256) k_local_sum(PtrList L, T *__restrict__ out, int64_t offset, int64_t count)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    T s = 0;
    for (int q = 0; q < L.n; q++) s += ((const T *)L.p[q])[offset + i];   // rank order: the same sum on every rank
    out[i] = s;
  }
}

struct LocalColl : CxColl {
  LocalGroup *g = nullptr;
  ~LocalColl() override
  {
    bool last;
    { std::lock_guard<std::mutex> lk(g->m); last = --g->refs == 0; }
    if (last) delete g;
  }
  PtrList gather(const void *a)
  {
    PtrList L;
    L.n = world;
    for (int q = 0; q < world; q++) L.p[q] = g->slot[q][0];
    (void)a;
    return L;
  }
  template <class T> int allreduce(cx_ctx *ctx, T *buf, int64_t count)
  {
    if (count <= 0) return CX_OK;
    T *tmp = nullptr;
    CX_CUDA(ctx, cudaMallocAsync((void **)&tmp, sizeof(T) * (size_t)count, ctx->stream));
    CX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    g->slot[rank][0] = buf;
    g->barrier();
    k_local_sum<T><<<cx_blocks(count, 256, ctx->num_sms * 8), 256, 0, ctx->stream>>>(gather(buf), tmp, 0, count);
    CX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    g->barrier();
    CX_CUDA(ctx, cudaMemcpyAsync(buf, tmp, sizeof(T) * (size_t)count, cudaMemcpyDeviceToDevice, ctx->stream));
    CX_CUDA(ctx, cudaFreeAsync(tmp, ctx->stream));
    return CX_OK;
  }
  int allreduce_sum_u32(cx_ctx *ctx, uint32_t *buf, int64_t count) override { return allreduce(ctx, buf, count); }
  int allreduce_sum_f32(cx_ctx *ctx, float *buf, int64_t count) override { return allreduce(ctx, buf, count); }
  int allreduce_sum_i64(cx_ctx *ctx, int64_t *buf, int64_t count) override { return allreduce(ctx, (long long *)buf, count); }
  int reduce_segments_u32(cx_ctx *ctx, uint32_t *base, const CxSeg *segs, int nseg) override
  {
    CX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    g->slot[rank][0] = base;
    g->barrier();
    for (int i = 0; i < nseg; i++)   // the root's own segment is read and written element by element by the same thread
      if (segs[i].root == rank && segs[i].count > 0)
        k_local_sum<uint32_t><<<cx_blocks(segs[i].count, 256, ctx->num_sms * 8), 256, 0, ctx->stream>>>(gather(base), base + segs[i].offset,
                                                                                                         segs[i].offset, segs[i].count);
    CX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    g->barrier();
    return CX_OK;
  }
  int bcast_segments_u32(cx_ctx *ctx, uint32_t *base, const CxSeg *segs, int nseg) override
  {
    CX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    g->slot[rank][0] = base;
    g->barrier();
    for (int i = 0; i < nseg; i++)
      if (segs[i].root != rank && segs[i].count > 0)
        CX_CUDA(ctx, cudaMemcpyAsync(base + segs[i].offset, (const uint32_t *)g->slot[segs[i].root][0] + segs[i].offset,
                                     sizeof(uint32_t) * (size_t)segs[i].count, cudaMemcpyDeviceToDevice, ctx->stream));
    CX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    g->barrier();
    return CX_OK;
  }
  int halo_exchange_u32(cx_ctx *ctx, const uint32_t *up_send, uint32_t *up_recv, const uint32_t *dn_send, uint32_t *dn_recv,
                        int64_t count) override
  {
    CX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    g->slot[rank][0] = up_send;
    g->slot[rank][1] = dn_send;
    g->barrier();
    if (count > 0 && rank + 1 < world)
      CX_CUDA(ctx, cudaMemcpyAsync(up_recv, g->slot[rank + 1][1], sizeof(uint32_t) * (size_t)count, cudaMemcpyDeviceToDevice, ctx->stream));
    if (count > 0 && rank > 0)
      CX_CUDA(ctx, cudaMemcpyAsync(dn_recv, g->slot[rank - 1][0], sizeof(uint32_t) * (size_t)count, cudaMemcpyDeviceToDevice, ctx->stream));
    CX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    g->barrier();
    return CX_OK;
  }
  int allgather_bytes(cx_ctx *ctx, const void *send, void *recv, int64_t bytes) override
  {
    CX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    g->slot[rank][0] = send;
    g->barrier();
    for (int q = 0; q < world && bytes > 0; q++)
      if ((char *)recv + (size_t)q * bytes != (const char *)g->slot[q][0])   // in place: the own piece is already there
        CX_CUDA(ctx, cudaMemcpyAsync((char *)recv + (size_t)q * bytes, g->slot[q][0], (size_t)bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    CX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    g->barrier();
    return CX_OK;
  }
};
}  // namespace

void cx_dist_release(cx_ctx *ctx)
{
  if (!ctx || !ctx->dist) return;
  delete (CxColl *)ctx->dist;
  ctx->dist = nullptr;
}

extern "C" {

int cx_dist_unique_id(void *id128)
{
  if (!id128) return CX_WRONG_INPUT;
  std::lock_guard<std::mutex> lk(g_nccl_mutex);
  if (!g_nccl.load()) return CX_NO_NCCL;
  static_assert(sizeof(ncclUniqueId) == CX_NCCL_ID_BYTES, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return CX_CUDA_ERROR;
  memcpy(id128, &id, sizeof(id));
  return CX_OK;
}

int cx_dist_init(cx_ctx *ctx, const void *id128, int world, int rank)
{
  if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) return CX_WRONG_INPUT;
  {
    std::lock_guard<std::mutex> lk(g_nccl_mutex);
    if (!g_nccl.load()) return cx_fail(ctx, CX_NO_NCCL, g_nccl.err.c_str());
  }
  cx_dist_release(ctx);
  CX_CUDA(ctx, cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  NcclColl *c = new NcclColl();
  c->world = world; c->rank = rank; c->own = true;
  ncclResult_t r = g_nccl.CommInitRank(&c->comm, world, id, rank);
  if (r != ncclSuccess) {
    c->comm = nullptr;
    delete c;
    return cx_fail(ctx, CX_CUDA_ERROR, (std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r)).c_str());
  }
  ctx->dist = c;
  return CX_OK;
}

int cx_dist_attach(cx_ctx *ctx, void *nccl_comm, int world, int rank)
{
  if (!ctx || !nccl_comm || world < 1 || rank < 0 || rank >= world) return CX_WRONG_INPUT;
  {
    std::lock_guard<std::mutex> lk(g_nccl_mutex);
    if (!g_nccl.load()) return cx_fail(ctx, CX_NO_NCCL, g_nccl.err.c_str());
  }
  cx_dist_release(ctx);
  NcclColl *c = new NcclColl();
  c->world = world; c->rank = rank; c->own = false; c->comm = (ncclComm_t)nccl_comm;
  ctx->dist = c;
  return CX_OK;
}

int cx_dist_init_local(cx_ctx *const *ctxs, int n)
{
  if (!ctxs || n < 1 || n > CX_LOCAL_MAX) return CX_WRONG_INPUT;
  for (int r = 0; r < n; r++) if (!ctxs[r]) return CX_WRONG_INPUT;
  LocalGroup *g = new LocalGroup();
  g->n = n; g->refs = n;
  for (int r = 0; r < n; r++) {
    cx_dist_release(ctxs[r]);
    LocalColl *c = new LocalColl();
    c->world = n; c->rank = r; c->g = g;
    ctxs[r]->dist = c;
  }
  return CX_OK;
}

int cx_dist_finalize(cx_ctx *ctx)
{
  if (!ctx) return CX_WRONG_INPUT;
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  cx_dist_release(ctx);
  return CX_OK;
}

int cx_internal_allgather(cx_ctx *ctx, void *buf_d, int64_t bytes)
{
  if (!ctx || !buf_d) return CX_WRONG_INPUT;
  if (cx_world(ctx) == 1 || bytes <= 0) return CX_OK;
  return cx_coll(ctx)->allgather_bytes(ctx, (const char *)buf_d + (size_t)cx_rank(ctx) * bytes, buf_d, bytes);
}

int cx_dist_world(cx_ctx *ctx) { return ctx ? cx_world(ctx) : 0; }
int cx_dist_rank(cx_ctx *ctx) { return ctx ? cx_rank(ctx) : 0; }

// calc_vert_volume over all ranks: interleaved chunks of CX_DIST_VV_CHUNK vertices (the cost per vertex varies along the mesh,
// contiguous ranges balance badly), each rank fills its chunks of a zeroed array, one all-reduce(SUM) gathers them.
int cx_dist_calc_vert_volume(cx_ctx *ctx, const cx_params *p, const float *pos_d, const float *norm_d, const int32_t *ep_d, float *vol_d,
                             const int32_t *trisize_d, const int32_t *cell_idx_d, int64_t nvert, int64_t nbe, int zero_open)
{
  if (!ctx || !vol_d) return CX_WRONG_INPUT;
  const int world = cx_world(ctx);
  if (world == 1) return cx_calc_vert_volume(ctx, p, pos_d, norm_d, ep_d, vol_d, trisize_d, cell_idx_d, nvert, nbe, zero_open, 0, nvert);
  CX_CUDA(ctx, cudaMemsetAsync(vol_d, 0, sizeof(float) * (size_t)nvert, ctx->stream));
  int rc = cx_calc_vert_volume_part(ctx, p, pos_d, norm_d, ep_d, vol_d, trisize_d, cell_idx_d, nvert, nbe, zero_open, CX_DIST_VV_CHUNK, world,
                                    cx_rank(ctx));
  // every rank reaches the collective even when its own part failed (a rank that returned early would hang the others)
  const int rc2 = cx_coll(ctx)->allreduce_sum_f32(ctx, vol_d, nvert);
  return rc != CX_OK ? rc : rc2;
}

// ---- order-independent checksum of a device array: sum over i of mix(i, word[i]) mod 2^64, and the popcount.
__global__ void __launch_bounds__(256) k_hash_u32(const uint32_t *__restrict__ w, int64_t n, unsigned long long *out)
{
  unsigned long long h = 0, pc = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t v = w[i];
    unsigned long long x = (unsigned long long)v + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
    h += x;
    pc += __popc(v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { h += __shfl_xor_sync(0xffffffffu, h, o); pc += __shfl_xor_sync(0xffffffffu, pc, o); }
  if ((threadIdx.x & 31) == 0) { atomicAdd(out, h); atomicAdd(out + 1, pc); }
}

int cx_hash_u32(cx_ctx *ctx, const void *words_d, int64_t n, uint64_t out_host[2])
{
  if (!ctx || !out_host || (n > 0 && !words_d)) return CX_WRONG_INPUT;
  unsigned long long *d = (unsigned long long *)(ctx->d_mail + 32);
  CX_CUDA(ctx, cudaMemsetAsync(d, 0, 16, ctx->stream));
  if (n > 0) {
    k_hash_u32<<<cx_blocks(n, 256, ctx->num_sms * 8), 256, 0, ctx->stream>>>((const uint32_t *)words_d, n, d);
    CX_LAUNCH_CHECK(ctx, "k_hash_u32");
  }
  CX_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail + 32, d, 16, cudaMemcpyDeviceToHost, ctx->stream));
  CX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  out_host[0] = (uint64_t)ctx->h_mail[32];
  out_host[1] = (uint64_t)ctx->h_mail[33];
  return CX_OK;
}

}  // extern "C"

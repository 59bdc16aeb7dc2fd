// crixus_b200 -- collective layer of the multi-GPU path (library-internal).
// One rank = one cx_ctx = one GPU.  Two back-ends behind the same five operations:
//   * NCCL (one process per GPU, NVLink / NVSwitch): resolved with dlopen("libnccl.so.2") when a communicator is created, so
//     the library has no link-time NCCL dependency;
//   * local: N contexts inside ONE process, one host thread per rank, buffers exchanged by peer pointers (same device or
//     peer-enabled devices).  It exists so that every line of the sharded algorithms above this layer runs in the
//     single-GPU test suite; only the thin NCCL wrappers are specific to the real thing.
// All operations are issued on the context's stream and are in place.
#pragma once
#include "common.cuh"

struct CxSeg {
  int64_t offset, count;   // in elements of the call's type
  int root;
};

struct CxColl {
  int world = 1, rank = 0;
  virtual ~CxColl() {}
  // buf[i] = sum over ranks of buf[i]
  virtual int allreduce_sum_u32(cx_ctx *ctx, uint32_t *buf, int64_t count) = 0;
  virtual int allreduce_sum_f32(cx_ctx *ctx, float *buf, int64_t count) = 0;
  virtual int allreduce_sum_i64(cx_ctx *ctx, int64_t *buf, int64_t count) = 0;
  // for every segment: base[offset .. offset+count) of rank `root` = sum over ranks of the same range (other ranks keep theirs)
  virtual int reduce_segments_u32(cx_ctx *ctx, uint32_t *base, const CxSeg *segs, int nseg) = 0;
  // for every segment: base[offset .. offset+count) of every rank = the same range of rank `root` (an all-gather with unequal parts)
  virtual int bcast_segments_u32(cx_ctx *ctx, uint32_t *base, const CxSeg *segs, int nseg) = 0;
  // z-neighbour exchange: send `up_send` to rank+1 and `dn_send` to rank-1, receive their counterparts (count words each way;
  // ranks at the ends skip the missing neighbour)
  virtual int halo_exchange_u32(cx_ctx *ctx, const uint32_t *up_send, uint32_t *up_recv, const uint32_t *dn_send, uint32_t *dn_recv,
                                int64_t count) = 0;
  // recv[r*bytes .. (r+1)*bytes) = send of rank r
  virtual int allgather_bytes(cx_ctx *ctx, const void *send, void *recv, int64_t bytes) = 0;
};

inline CxColl *cx_coll(cx_ctx *ctx) { return (CxColl *)ctx->dist; }
inline int cx_world(cx_ctx *ctx) { return ctx->dist ? ((CxColl *)ctx->dist)->world : 1; }
inline int cx_rank(cx_ctx *ctx) { return ctx->dist ? ((CxColl *)ctx->dist)->rank : 0; }

// contiguous range [lo, hi) of rank r out of world; sizes differ by at most one
inline void cx_split_range(int64_t n, int world, int r, int64_t *lo, int64_t *hi)
{
  const int64_t base = n / world, rem = n % world;
  *lo = r * base + (r < rem ? r : rem);
  *hi = *lo + base + (r < rem ? 1 : 0);
}

// crixus_b200 -- context, scratch arena, and the host-side glue of crixus_main restated for the C-ABI.
#include <math.h>
#include <string.h>
#include "common.cuh"

int cx_fail(cx_ctx *ctx, int code, const char *what, cudaError_t e)
{
  if (ctx) {
    ctx->last_error = std::string(what ? what : "?");
    if (e != cudaSuccess) ctx->last_error += std::string(": ") + cudaGetErrorString(e);
  }
  return code;
}

void *cx_scratch_pinned(cx_ctx *ctx, size_t bytes)
{
  if (bytes > ctx->pin_stage_bytes) {
    cudaStreamSynchronize(ctx->stream);
    if (ctx->pin_stage) cudaFreeHost(ctx->pin_stage);
    ctx->pin_stage = nullptr; ctx->pin_stage_bytes = 0;
    if (cudaMallocHost(&ctx->pin_stage, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    ctx->pin_stage_bytes = bytes;
  }
  return ctx->pin_stage;
}

void *cx_scratch(cx_ctx *ctx, size_t bytes)
{
  if (bytes <= ctx->scratch_bytes) return ctx->scratch;
  if (ctx->scratch) {
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->scratch);
    ctx->scratch = nullptr;
    ctx->scratch_bytes = 0;
  }
  size_t want = bytes + bytes / 4 + (1 << 20);
  if (cudaMalloc(&ctx->scratch, want) != cudaSuccess) {
    cx_fail(ctx, CX_CUDA_ERROR, "cudaMalloc(scratch)", cudaGetLastError());
    return nullptr;
  }
  ctx->scratch_bytes = want;
  return ctx->scratch;
}

extern "C" {

const char *cx_version(void) { return "crixus_b200 0.1 (sm_100a)"; }

int cx_create(int device, cx_ctx **out)
{
  if (!out) return CX_WRONG_INPUT;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return CX_NO_DEVICE;  // no CPU fallback: fail loudly
  if (device < 0 || device >= n) return CX_NO_DEVICE;
  cx_ctx *ctx = new cx_ctx();
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return CX_NO_DEVICE; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->num_sms = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return CX_CUDA_ERROR; }
  ctx->own_stream = true;
  {  // stream-ordered temporaries (cudaMallocAsync) stay cached in the pool between calls
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      uint64_t thr = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
  }
  if (cudaMallocHost(&ctx->h_mail, 64 * sizeof(int64_t)) != cudaSuccess ||
      cudaMalloc(&ctx->d_mail, 64 * sizeof(int64_t)) != cudaSuccess) {
    cx_destroy(ctx);
    return CX_CUDA_ERROR;
  }
  *out = ctx;
  return CX_OK;
}

void cx_destroy(cx_ctx *ctx)
{
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  cx_dist_release(ctx);
  cx_fill_state_free(ctx);
  if (ctx->scratch) cudaFree(ctx->scratch);
  if (ctx->dev_arena) cudaFree(ctx->dev_arena);
  if (ctx->host_arena) cudaFreeHost(ctx->host_arena);
  if (ctx->h_mail) cudaFreeHost(ctx->h_mail);
  if (ctx->d_mail) cudaFree(ctx->d_mail);
  if (ctx->pin_stage) cudaFreeHost(ctx->pin_stage);
  if (ctx->norm_stage) cudaFree(ctx->norm_stage);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

void *cx_internal_norm_stage(cx_ctx *ctx, size_t bytes)
{
  if (bytes > ctx->norm_stage_bytes) {
    cudaStreamSynchronize(ctx->stream);
    if (ctx->norm_stage) cudaFree(ctx->norm_stage);
    ctx->norm_stage = nullptr; ctx->norm_stage_bytes = 0;
    if (cudaMalloc(&ctx->norm_stage, bytes + bytes / 8) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    ctx->norm_stage_bytes = bytes + bytes / 8;
  }
  return ctx->norm_stage;
}

void *cx_internal_dev_arena(cx_ctx *ctx, size_t bytes)
{
  if (bytes > ctx->dev_arena_bytes) {
    cudaStreamSynchronize(ctx->stream);
    if (ctx->dev_arena) cudaFree(ctx->dev_arena);
    ctx->dev_arena = nullptr; ctx->dev_arena_bytes = 0;
    const size_t want = bytes + bytes / 8;
    if (cudaMalloc(&ctx->dev_arena, want) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    ctx->dev_arena_bytes = want;
  }
  return ctx->dev_arena;
}

void *cx_internal_host_arena(cx_ctx *ctx, size_t bytes)
{
  if (bytes > ctx->host_arena_bytes) {
    cudaStreamSynchronize(ctx->stream);
    if (ctx->host_arena) cudaFreeHost(ctx->host_arena);
    ctx->host_arena = nullptr; ctx->host_arena_bytes = 0;
    const size_t want = bytes + bytes / 8;
    if (cudaMallocHost(&ctx->host_arena, want) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    ctx->host_arena_bytes = want;
  }
  return ctx->host_arena;
}

__global__ void __launch_bounds__(256) k_expand3(const float *__restrict__ s, float4 *__restrict__ d, int64_t n)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    d[i] = make_float4(s[3 * i], s[3 * i + 1], s[3 * i + 2], 0.f);
}

int cx_internal_expand3(cx_ctx *ctx, const float *src3_d, float *dst4_d, int64_t n)
{
  if (n <= 0) return CX_OK;
  k_expand3<<<cx_blocks(n, 256, ctx->num_sms * 16), 256, 0, ctx->stream>>>(src3_d, (float4 *)dst4_d, n);
  CX_LAUNCH_CHECK(ctx, "k_expand3");
  return CX_OK;
}

__global__ void __launch_bounds__(256) k_offset_ep4(const int4 *__restrict__ s, int4 *__restrict__ d, int64_t n, int32_t off)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int4 e = s[i];
    d[i] = make_int4(e.x + off, e.y + off, e.z + off, 0);
  }
}

int cx_internal_offset_ep4(cx_ctx *ctx, const int32_t *src_ep4_d, int32_t *dst_ep4_d, int64_t n, int32_t offset)
{
  if (n <= 0) return CX_OK;
  k_offset_ep4<<<cx_blocks(n, 256, ctx->num_sms * 16), 256, 0, ctx->stream>>>((const int4 *)src_ep4_d, (int4 *)dst_ep4_d, n, offset);
  CX_LAUNCH_CHECK(ctx, "k_offset_ep4");
  return CX_OK;
}

int cx_set_stream(cx_ctx *ctx, void *s)
{
  if (!ctx) return CX_WRONG_INPUT;
  if (ctx->own_stream && ctx->stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); }
  ctx->stream = (cudaStream_t)s;
  ctx->own_stream = false;
  return CX_OK;
}

int cx_synchronize(cx_ctx *ctx)
{
  CX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return CX_OK;
}

void *cx_get_stream(cx_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
const char *cx_last_error(cx_ctx *ctx) { return ctx ? ctx->last_error.c_str() : "no context"; }
int64_t cx_kernel_launches(cx_ctx *ctx) { return ctx ? ctx->launches : 0; }

// ---------------------------------------------------------------------------------------------- host glue
// /root/reference/src/crixus.cu:262-283
void cx_params_domain(cx_params *p, float dr, const float bbmin[3], const float bbmax[3])
{
  memset(p, 0, sizeof(*p));
  p->dr = dr;
  p->dr_wall = dr;
  for (int j = 0; j < 3; j++) { p->dmin[j] = bbmin[j]; p->dmax[j] = bbmax[j]; }
  p->eps = (float)(dr / (float)CX_GRES * 1e-4);
  for (int j = 0; j < 3; j++) {
    int g = (int)((bbmax[j] - bbmin[j]) / dr / 3.0);
    p->gridn[j] = g > 1 ? g : 1;
  }
  p->gridn[3] = p->gridn[0] * p->gridn[1] * p->gridn[2];
  for (int j = 0; j < 3; j++) p->griddr[j] = (p->dmax[j] - p->dmin[j] + p->eps) / (float)p->gridn[j];
  for (int j = 0; j < 3; j++) { p->fcmin[j] = p->dmin[j]; p->fcmax[j] = p->dmax[j]; }
}

// /root/reference/src/crixus.cu:464-467
void cx_params_fill_eps(cx_params *p)
{
  float eps = 1e-10f;
  for (int i = 0; i < 3; i++) {
    float e = (p->dmax[i] - p->dmin[i]) * 1e-5f;
    if (e > eps) eps = e;
  }
  p->eps = eps;
}

// /root/reference/src/crixus.cu:698-722
void cx_params_container(cx_params *p, int use, const float cmin[3], const float cmax[3])
{
  for (int i = 0; i < 3; i++) {
    if (use) {
      p->fcmin[i] = cmin[i];
      p->fcmax[i] = cmax[i];
    } else {
      p->fcmin[i] = p->dmin[i];
      p->fcmax[i] = p->dmax[i];
      if (p->per[i]) p->fcmax[i] -= p->dr;
    }
  }
}

// /root/reference/src/crixus_d.cu:985-987, src/crixus.cu:802-803
void cx_fill_dims(const cx_params *p, int64_t dimg[3])
{
  for (int j = 0; j < 3; j++) dimg[j] = (int)floorf((p->fcmax[j] - p->fcmin[j] + p->eps) / p->dr + 1);
}

// /root/reference/src/crixus.cu:799-804
int64_t cx_seed_index(const cx_params *p, const float spos[3])
{
  int64_t d[3], s[3];
  cx_fill_dims(p, d);
  for (int j = 0; j < 3; j++) {
    s[j] = (int)roundf((spos[j] - p->fcmin[j] + p->eps) / p->dr);
    if (s[j] < 0 || s[j] >= d[j]) return -1;
  }
  return s[0] + s[1] * d[0] + s[2] * d[0] * d[1];
}

}  // extern "C"

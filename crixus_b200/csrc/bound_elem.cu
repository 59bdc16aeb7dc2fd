// crixus_b200 -- swap_normals and set_bound_elem (per-segment barycentre, area, winding fix, z-min statistics).
// Reference: /root/reference/src/crixus_d.cu:105-211.  HBM-bound: 100 B / segment.
#include "common.cuh"

// ---------------------------------------------------------------------------------------------- swap_normals
// crixus_d.cu:203-211: norm[i].xyz *= -1.  (float4 read-modify-write, .w preserved)
__global__ void __launch_bounds__(256) k_swap_normals(float4 *__restrict__ norm, int64_t nbe)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nbe; i += (int64_t)gridDim.x * blockDim.x) {
    float4 n = norm[i];
    n.x = -n.x; n.y = -n.y; n.z = -n.z;
    norm[i] = n;
  }
}

// ---------------------------------------------------------------------------------------------- set_bound_elem
__device__ __forceinline__ unsigned int ordered_f32(float f)
{
  unsigned int b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float unordered_f32(unsigned int u)
{
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
#define CX_ZKEY_NONE 0xffffffffffffffffull

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w < v ? w : v;
  }
  return v;
}

// One thread per segment.  Reads ep (16 B), norm (16 B), gathers 3 vertices (3 x 16 B, read-only path), writes the
// barycentre (16 B), the area (4 B) and, if the winding was wrong, ep (16 B).
__global__ void __launch_bounds__(256) k_set_bound_elem(float4 *__restrict__ pos, const float4 *__restrict__ norm,
                                                        float *__restrict__ surf, int4 *__restrict__ ep, int64_t nbe,
                                                        int64_t nvert, unsigned long long *__restrict__ zkeys)
{
  unsigned long long kp = CX_ZKEY_NONE, kn = CX_ZKEY_NONE;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nbe; i += (int64_t)gridDim.x * blockDim.x) {
    int4 e = ep[i];
    const float4 n = norm[i];
    // vertices live in pos[0:nvert], barycentres are written to pos[nvert:], so the gathers never alias the stores
    const float4 q0 = __ldg(&pos[e.x]), q1 = __ldg(&pos[e.y]), q2 = __ldg(&pos[e.z]);
    const float v0[3] = {q0.x, q0.y, q0.z}, v1[3] = {q1.x, q1.y, q1.z}, v2[3] = {q2.x, q2.y, q2.z};
    float a2 = 0.f, b2 = 0.f, c2 = 0.f, dd[3], nd[3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
      // ddum[j] += pos/3. : double divide, double add, narrowed to float after every term (crixus_d.cu:132-134)
      float d = (float)(0.0 + (double)v0[j] / 3.);
      d = (float)((double)d + (double)v1[j] / 3.);
      d = (float)((double)d + (double)v2[j] / 3.);
      dd[j] = d;
      // a2 += pow(float diff, 2): the square of a float is exact in double (crixus_d.cu:135-137)
      double t;
      t = (double)fsub(v0[j], v1[j]); a2 = (float)((double)a2 + t * t);
      t = (double)fsub(v1[j], v2[j]); b2 = (float)((double)b2 + t * t);
      t = (double)fsub(v2[j], v0[j]); c2 = (float)((double)c2 + t * t);
      const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      nd[j] = det_r2(fsub(v1[j1], v0[j1]), fsub(v2[j2], v0[j2]), fsub(v1[j2], v0[j2]), fsub(v2[j1], v0[j1]));
    }
    if (dot_r3(n.x, n.y, n.z, nd[0], nd[1], nd[2]) < 0.0f) {  // crixus_d.cu:148: fix the winding
      int t = e.y; e.y = e.z; e.z = t;
      ep[i] = e;
    }
    const unsigned long long key = ((unsigned long long)ordered_f32(dd[2]) << 32) | (unsigned long long)(uint32_t)i;
    if ((double)n.z > 1e-5 && dd[2] < 1e10f) kp = key < kp ? key : kp;   // crixus_d.cu:153 (ties: lowest index)
    if ((double)n.z < -1e-5 && dd[2] < 1e10f) kn = key < kn ? key : kn;  // crixus_d.cu:157
    const double s = (double)fsub(fadd(a2, b2), c2);
    surf[i] = (float)(0.25 * sqrt(4. * (double)a2 * (double)b2 - s * s));  // crixus_d.cu:161
    pos[nvert + i] = make_float4(dd[0], dd[1], dd[2], 0.f);
  }
  kp = warp_min_u64(kp);
  kn = warp_min_u64(kn);
  if ((threadIdx.x & 31) == 0) {
    if (kp != CX_ZKEY_NONE) atomicMin(&zkeys[0], kp);
    if (kn != CX_ZKEY_NONE) atomicMin(&zkeys[1], kn);
  }
}

__global__ void k_zstat_init(unsigned long long *zkeys) { zkeys[0] = CX_ZKEY_NONE; zkeys[1] = CX_ZKEY_NONE; }

__global__ void k_zstat_finish(const unsigned long long *zkeys, const float4 *__restrict__ norm, float *out4)
{
  // {xminp, xminn, nminp, nminn}, initial values of crixus.cu:325-326 when no segment qualified
  out4[0] = 1e10f; out4[1] = 1e10f; out4[2] = 0.f; out4[3] = 0.f;
  for (int s = 0; s < 2; s++)
    if (zkeys[s] != CX_ZKEY_NONE) {
      out4[s] = unordered_f32((unsigned int)(zkeys[s] >> 32));
      out4[2 + s] = norm[(uint32_t)(zkeys[s] & 0xffffffffull)].z;
    }
}

extern "C" int cx_swap_normals(cx_ctx *ctx, float *norm_d, int64_t nbe)
{
  if (!ctx || !norm_d || nbe < 0) return CX_WRONG_INPUT;
  if (nbe == 0) return CX_OK;
  k_swap_normals<<<cx_blocks(nbe, 256, ctx->num_sms * 16), 256, 0, ctx->stream>>>((float4 *)norm_d, nbe);
  CX_LAUNCH_CHECK(ctx, "k_swap_normals");
  return CX_OK;
}

extern "C" int cx_set_bound_elem(cx_ctx *ctx, float *pos_d, const float *norm_d, float *surf_d, int32_t *ep_d,
                                 int64_t nbe, int64_t nvert, float *zstat_host)
{
  if (!ctx || !pos_d || !norm_d || !surf_d || !ep_d || nbe < 0 || nvert < 0) return CX_WRONG_INPUT;
  unsigned long long *zkeys = (unsigned long long *)ctx->d_mail;
  float *zout = (float *)(ctx->d_mail + 2);
  k_zstat_init<<<1, 1, 0, ctx->stream>>>(zkeys);
  CX_LAUNCH_CHECK(ctx, "k_zstat_init");
  if (nbe > 0) {
    k_set_bound_elem<<<cx_blocks(nbe, 256, ctx->num_sms * 16), 256, 0, ctx->stream>>>(
        (float4 *)pos_d, (const float4 *)norm_d, surf_d, (int4 *)ep_d, nbe, nvert, zkeys);
    CX_LAUNCH_CHECK(ctx, "k_set_bound_elem");
  }
  if (zstat_host) {
    k_zstat_finish<<<1, 1, 0, ctx->stream>>>(zkeys, (const float4 *)norm_d, zout);
    CX_LAUNCH_CHECK(ctx, "k_zstat_finish");
    CX_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail, zout, 4 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(zstat_host, ctx->h_mail, 4 * sizeof(float));
  }
  return CX_OK;
}

// crixus_b200 -- measured denominators for the roofline: an FFMA micro-kernel (non-tensor FP32 peak) and a
// device-to-device copy (HBM).  calc_vert_volume is FP32-ALU bound, which MEASURED_PEAKS.json does not cover.
#include "common.cuh"

__global__ void __launch_bounds__(256) k_ffma_peak(float *out, int iters, float a, float b)
{
  float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      x0 = __fmaf_rn(x0, a, b); x1 = __fmaf_rn(x1, a, b); x2 = __fmaf_rn(x2, a, b); x3 = __fmaf_rn(x3, a, b);
      x4 = __fmaf_rn(x4, a, b); x5 = __fmaf_rn(x5, a, b); x6 = __fmaf_rn(x6, a, b); x7 = __fmaf_rn(x7, a, b);
    }
  }
  const float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 12345.678f) out[0] = s;  // never true; keeps the chain alive
}

// returns achieved non-tensor FP32 throughput in TFLOP/s (2 flops per FFMA), best of `reps`
extern "C" int cx_measure_fp32_peak(cx_ctx *ctx, int reps, double *tflops)
{
  if (!ctx || !tflops) return CX_WRONG_INPUT;
  float *out = (float *)(ctx->d_mail + 40);
  cudaEvent_t e0, e1;
  CX_CUDA(ctx, cudaEventCreate(&e0));
  CX_CUDA(ctx, cudaEventCreate(&e1));
  const int iters = 2048, grid = ctx->num_sms * 8;
  double best = 0;
  for (int r = 0; r < reps + 1; r++) {
    CX_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
    k_ffma_peak<<<grid, 256, 0, ctx->stream>>>(out, iters, 0.999f, 0.001f);
    CX_LAUNCH_CHECK(ctx, "k_ffma_peak");
    CX_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
    CX_CUDA(ctx, cudaEventSynchronize(e1));
    float ms = 0;
    CX_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
    const double fl = 2.0 * 8 * 16 * (double)iters * 256.0 * grid;
    if (r > 0 && fl / (ms * 1e-3) / 1e12 > best) best = fl / (ms * 1e-3) / 1e12;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tflops = best;
  return CX_OK;
}

// crixus_b200 -- cell binning of segments into 3dr-cells: count -> scan -> STABLE scatter.
// Reference: /root/reference/src/crixus_d.cu:37-103 (init_cell_idx, count_cell_bes, add_up_indices<<<1,1>>>, sort_bes)
// and grid sizing /root/reference/src/crixus.cu:267-276.  The reference's scatter order inside a cell is
// run-dependent (atomicAdd cursor, crixus_d.cu:94); here the order is the original segment order, made
// deterministic by ranking the atomically scattered ids inside each cell.
// HBM-bound: ~140 B / segment + 12 B / 3dr-cell.
#include "common.cuh"

// pass 1: histogram (+ remember each segment's cell)
__global__ void __launch_bounds__(256) k_cell_count(const float4 *__restrict__ pre_pos, int32_t *__restrict__ count,
                                                    int32_t *__restrict__ cellid, cx_params p, int64_t nbe, int64_t nvert)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nbe; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 b = pre_pos[nvert + i];
    const int gx = cell_coord(b.x, p.dmin[0], p.griddr[0], p.gridn[0]);
    const int gy = cell_coord(b.y, p.dmin[1], p.griddr[1], p.gridn[1]);
    const int gz = cell_coord(b.z, p.dmin[2], p.griddr[2], p.gridn[2]);
    const int c = gx * p.gridn[1] * p.gridn[2] + gy * p.gridn[2] + gz;  // crixus_d.cu:56
    cellid[i] = c;
    atomicAdd(&count[c], 1);
  }
}

// pass 2: claim an (arbitrary) slot inside the cell's range
__global__ void __launch_bounds__(256) k_cell_claim(const int32_t *__restrict__ cellid, int32_t *__restrict__ cursor,
                                                    int32_t *__restrict__ order, int64_t nbe)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nbe; i += (int64_t)gridDim.x * blockDim.x) {
    const int slot = atomicAdd(&cursor[cellid[i]], 1);
    order[slot] = (int32_t)i;
  }
}

// pass 3: rank the claimed ids inside their cell (stable order) and move the 52-byte payload
__global__ void __launch_bounds__(256) k_cell_rank_move(const int32_t *__restrict__ order, const int32_t *__restrict__ cellid,
                                                        const int32_t *__restrict__ cell_idx, const float4 *__restrict__ pre_pos,
                                                        float4 *__restrict__ pos, const float4 *__restrict__ pre_norm,
                                                        float4 *__restrict__ norm, const int4 *__restrict__ pre_ep,
                                                        int4 *__restrict__ ep, const float *__restrict__ pre_surf,
                                                        float *__restrict__ surf, int64_t nbe, int64_t nvert)
{
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < nbe; s += (int64_t)gridDim.x * blockDim.x) {
    const int32_t x = order[s];
    const int c = cellid[x];
    const int b = cell_idx[c], e = cell_idx[c + 1];
    int r = 0;
    for (int t = b; t < e; t++) r += (order[t] < x);
    const int64_t j = b + r;
    pos[nvert + j] = pre_pos[nvert + x];
    norm[j] = pre_norm[x];
    ep[j] = pre_ep[x];
    surf[j] = pre_surf[x];
  }
}

__global__ void __launch_bounds__(256) k_copy_f4(const float4 *__restrict__ src, float4 *__restrict__ dst, int64_t n)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

extern "C" int cx_bin_segments(cx_ctx *ctx, const cx_params *p, const float *pre_pos_d, float *pos_d, const float *pre_norm_d,
                               float *norm_d, const int32_t *pre_ep_d, int32_t *ep_d, const float *pre_surf_d, float *surf_d,
                               int32_t *cell_idx_d, int64_t nbe, int64_t nvert)
{
  if (!ctx || !p || nbe < 0 || nvert < 0 || p->gridn[3] <= 0) return CX_WRONG_INPUT;
  const int64_t ncell = p->gridn[3];
  cudaStream_t st = ctx->stream;
  int32_t *count = nullptr, *cellid = nullptr, *order = nullptr, *cursor = nullptr;
  CX_CUDA(ctx, cudaMallocAsync(&count, sizeof(int32_t) * (size_t)(ncell + 1), st));
  CX_CUDA(ctx, cudaMallocAsync(&cursor, sizeof(int32_t) * (size_t)(ncell + 1), st));
  CX_CUDA(ctx, cudaMallocAsync(&cellid, sizeof(int32_t) * (size_t)(nbe + 1), st));
  CX_CUDA(ctx, cudaMallocAsync(&order, sizeof(int32_t) * (size_t)(nbe + 1), st));
  CX_CUDA(ctx, cudaMemsetAsync(count, 0, sizeof(int32_t) * (size_t)(ncell + 1), st));
  const int grid = cx_blocks(nbe, 256, ctx->num_sms * 16);
  if (nbe > 0) {
    k_cell_count<<<grid, 256, 0, st>>>((const float4 *)pre_pos_d, count, cellid, *p, nbe, nvert);
    CX_LAUNCH_CHECK(ctx, "k_cell_count");
  }
  int rc = cx_exclusive_scan_i32(ctx, count, cell_idx_d, ncell + 1);  // cell_idx[ncell] = nbe
  if (rc != CX_OK) return rc;
  CX_CUDA(ctx, cudaMemcpyAsync(cursor, cell_idx_d, sizeof(int32_t) * (size_t)(ncell + 1), cudaMemcpyDeviceToDevice, st));
  if (nbe > 0) {
    k_cell_claim<<<grid, 256, 0, st>>>(cellid, cursor, order, nbe);
    CX_LAUNCH_CHECK(ctx, "k_cell_claim");
    k_cell_rank_move<<<grid, 256, 0, st>>>(order, cellid, cell_idx_d, (const float4 *)pre_pos_d, (float4 *)pos_d,
                                          (const float4 *)pre_norm_d, (float4 *)norm_d, (const int4 *)pre_ep_d, (int4 *)ep_d,
                                          pre_surf_d, surf_d, nbe, nvert);
    CX_LAUNCH_CHECK(ctx, "k_cell_rank_move");
  }
  if (nvert > 0 && pos_d != pre_pos_d) {
    k_copy_f4<<<cx_blocks(nvert, 256, ctx->num_sms * 16), 256, 0, st>>>((const float4 *)pre_pos_d, (float4 *)pos_d, nvert);
    CX_LAUNCH_CHECK(ctx, "k_copy_f4");
  }
  CX_CUDA(ctx, cudaFreeAsync(count, st));
  CX_CUDA(ctx, cudaFreeAsync(cursor, st));
  CX_CUDA(ctx, cudaFreeAsync(cellid, st));
  CX_CUDA(ctx, cudaFreeAsync(order, st));
  return CX_OK;
}

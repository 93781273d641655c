// weights.cu -- per-step weight pipeline as two launches (instead of ~600 tiny framework kernels):
//
//   nunerf_weights_prepare : for every dense layer, in one launch: weight-norm (W = g v / |v|, nn.utils.weight_norm
//                            dim 0 as used by field.py:121-122, :386-394), optional row / column rotation and scaling,
//                            and conversion to the bf16 plane operands of the tensor-core kernels (K-major W and W^T).
//   nunerf_weights_backward: the adjoint: effective-weight gradients (in the prepared layout, produced by
//                            nunerf_linear_dw / nunerf_colsum) -> gradients of weight_v, weight_g, bias (or of the plain
//                            weight), accumulated straight into the parameters' .grad storage.
//
// One thread block per emitted weight row; descriptor tables live in device memory and are built once.
#include "common.cuh"

namespace nunerf {

constexpr int WTHREADS = 128;

__device__ __forceinline__ double block_sum_d(double x, double* red) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = x;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < WTHREADS / 32; ++w) t += red[w];
  return t;
}

__device__ __forceinline__ float block_sum(float x, float* red) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = x;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < WTHREADS / 32; ++w) t += red[w];
  return t;
}

__global__ void __launch_bounds__(WTHREADS)
weights_prepare_kernel(const nunerf_wdesc_t* __restrict__ descs, const int32_t* __restrict__ blk_desc,
                       const int32_t* __restrict__ blk_row) {
  __shared__ float red[WTHREADS / 32];
  const nunerf_wdesc_t d = descs[blk_desc[blockIdx.x]];
  const int r_dst = blk_row[blockIdx.x];
  const int src_row = (d.src_row0 + r_dst + d.row_rot) % d.N;
  const float* vrow = d.v + (long long)src_row * d.ld;
  float s = d.scale;
  if (d.g) {
    float ss = 0.f;
    for (int k = threadIdx.x; k < d.K; k += WTHREADS) { float x = vrow[k]; ss += x * x; }
    ss = block_sum(ss, red);
    const float inv = 1.0f / sqrtf(ss);
    if (threadIdx.x == 0 && d.inv_norm) d.inv_norm[src_row] = inv;
    s *= d.g[src_row] * inv;
  }
  __nv_bfloat16* wk = (__nv_bfloat16*)d.wk;
  __nv_bfloat16* wtk = (__nv_bfloat16*)d.wtk;
  for (int c = threadIdx.x; c < d.K; c += WTHREADS) {
    int sc = c + d.col_rot;
    if (sc >= d.K) sc -= d.K;
    const float w = vrow[sc] * s;
    if (wk) store_planes(wk, (long long)(d.wk_row_off + r_dst) * d.wk_ld + c, d.wk_lo, w);
    if (wtk) store_planes(wtk, (long long)c * d.wtk_ld + d.wtk_col_off + r_dst, d.wtk_lo, w);
    if (d.row_f32) d.row_f32[(long long)r_dst * d.K + c] = w;
  }
  if (threadIdx.x == 0 && d.bias_src && d.bias_dst) d.bias_dst[d.wk_row_off + r_dst] = d.bias_src[src_row];
}

__global__ void __launch_bounds__(WTHREADS)
weights_backward_kernel(const nunerf_wdesc_t* __restrict__ descs, const int32_t* __restrict__ blk_desc,
                        const int32_t* __restrict__ blk_row) {
  __shared__ double red[WTHREADS / 32];
  const nunerf_wdesc_t d = descs[blk_desc[blockIdx.x]];
  if (!d.dW) return;
  const int r_dst = blk_row[blockIdx.x];
  const int src_row = (d.src_row0 + r_dst + d.row_rot) % d.N;
  const float* vrow = d.v + (long long)src_row * d.ld;
  const float* grow = d.dW + (long long)(d.dw_row_off + r_dst) * d.lddw;
  float* dvrow = d.dv ? d.dv + (long long)src_row * d.ld : nullptr;   // null: the parameter is frozen (requires_grad=False)
  if (d.g) {
    // W = (g / |v|) v  ->  dg = (dW . v) / |v| ;  dv = (g / |v|) (dW - v (dW . v) / |v|^2)
    // (the two terms of dv nearly cancel for some layers: accumulate the row reductions in fp64)
    double dot = 0.0, ss = 0.0;
    for (int c = threadIdx.x; c < d.K; c += WTHREADS) {
      int sc = c + d.col_rot;
      if (sc >= d.K) sc -= d.K;
      dot += (double)grow[c] * (double)d.scale * (double)vrow[sc];
      ss += (double)vrow[sc] * (double)vrow[sc];
    }
    dot = block_sum_d(dot, red);
    ss = block_sum_d(ss, red);
    const double inv = 1.0 / sqrt(ss);
    const double gi = (double)d.g[src_row] * inv;
    if (threadIdx.x == 0 && d.dg) d.dg[src_row] += (float)(dot * inv);
    const double proj = dot / ss;
    for (int c = threadIdx.x; dvrow && c < d.K; c += WTHREADS) {
      int sc = c + d.col_rot;
      if (sc >= d.K) sc -= d.K;
      dvrow[sc] += (float)(gi * ((double)grow[c] * (double)d.scale - (double)vrow[sc] * proj));
    }
  } else {
    for (int c = threadIdx.x; dvrow && c < d.K; c += WTHREADS) {
      int sc = c + d.col_rot;
      if (sc >= d.K) sc -= d.K;
      dvrow[sc] += grow[c] * d.scale;
    }
  }
  if (threadIdx.x == 0 && d.db && d.dbias) d.dbias[src_row] += d.db[d.dw_row_off + r_dst];
}

}  // namespace nunerf

using namespace nunerf;

extern "C" int nunerf_weights_prepare(const nunerf_wdesc_t* descs, const int32_t* blk_desc, const int32_t* blk_row,
                                      int n_blocks, void* stream) {
  NUNERF_REQUIRE(descs && blk_desc && blk_row && n_blocks > 0, "weights_prepare: bad arguments");
  weights_prepare_kernel<<<n_blocks, WTHREADS, 0, (cudaStream_t)stream>>>(descs, blk_desc, blk_row);
  NUNERF_CHECK_LAUNCH("weights_prepare_kernel");
  return 0;
}

extern "C" int nunerf_weights_backward(const nunerf_wdesc_t* descs, const int32_t* blk_desc, const int32_t* blk_row,
                                       int n_blocks, void* stream) {
  NUNERF_REQUIRE(descs && blk_desc && blk_row && n_blocks > 0, "weights_backward: bad arguments");
  weights_backward_kernel<<<n_blocks, WTHREADS, 0, (cudaStream_t)stream>>>(descs, blk_desc, blk_row);
  NUNERF_CHECK_LAUNCH("weights_backward_kernel");
  return 0;
}

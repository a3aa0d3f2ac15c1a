// ffq_lpbq.cu -- LPBQ scale compression ("on-disk formats"): the arithmetic of
// export/_lpbq.py:131-160 (grouped_dynamic_quantize) in ONE pass over the per-block scales.
//
// A per-block quantized weight has one fp32 scale per (channel, block).  LPBQ stores it as a per-CHANNEL float scale
// and a small per-block integer:
//     f[c]    = max_b s[c, b] / 2^bw                       (aten: amax over the channel's blocks, then a division)
//     i[c, b] = clamp(rint(s[c, b] / f[c]), 1, 2^bw)       (IEEE division, half-to-even, then the uint32 cast)
// The reference runs amax, div, div, round, clamp, cast: six passes with five temporaries over a tensor that has
// 54.5 M entries for an 8B-parameter model at g = 128 (218 MB).  Here: read s once (twice through L1/L2 for long
// channels), write i and f.  HBM-bound, 8 B per scale algorithmic.
//
// Layouts: the scales of PerBlock(block_dims=1, per_channel_dims=0) are [channels, blocks] row-major (a channel is a
// ROW: one warp per row, coalesced reads, shuffle max); those of PerBlock(block_dims=0, per_channel_dims=1) are
// [blocks, channels] (a channel is a COLUMN: one thread per column, a warp reads 128 contiguous bytes of each row).
// NaN propagates through the max like aten's amax.  Every step is a separately rounded fp32 op (--fmad=false).
#include "ffq_common.cuh"

namespace ffq {

__device__ __forceinline__ int lpbq_code(float s, float f, float hi) {
  const float q = nan_clamp(rintf(__fdiv_rn(s, f)), 1.0f, hi);
  return (int)__float2uint_rz(q);                  // .to(torch.uint32): NaN -> 0
}

// a channel is a row of `cols` scales
__global__ void __launch_bounds__(256) lpbq_rows_kernel(const float* __restrict__ s, long long rows, long long cols,
                                                        float two_bw, int* __restrict__ iq, float* __restrict__ fs) {
  const int lane = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += warps) {
    const float* row = s + r * cols;
    float m = -INFINITY;
    for (long long c = lane; c < cols; c += 32) m = nan_max(m, row[c]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = nan_max(m, __shfl_xor_sync(0xffffffffu, m, d));
    const float f = __fdiv_rn(m, two_bw);
    if (lane == 0) fs[r] = f;
    int* out = iq + r * cols;
    for (long long c = lane; c < cols; c += 32) out[c] = lpbq_code(row[c], f, two_bw);
  }
}

// a channel is a column: [rows = blocks, cols = channels]
__global__ void __launch_bounds__(256) lpbq_cols_kernel(const float* __restrict__ s, long long rows, long long cols,
                                                        float two_bw, int* __restrict__ iq, float* __restrict__ fs) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float m = -INFINITY;
  for (long long r = 0; r < rows; ++r) m = nan_max(m, s[r * cols + c]);
  const float f = __fdiv_rn(m, two_bw);
  fs[c] = f;
  for (long long r = 0; r < rows; ++r) iq[r * cols + c] = lpbq_code(s[r * cols + c], f, two_bw);
}

}  // namespace ffq

extern "C" int ffq_lpbq_encode(const float* scale, int64_t rows, int64_t cols, int channel_axis, int bitwidth,
                               int32_t* int_scale, float* float_scale, void* stream) {
  using namespace ffq;
  if (scale == nullptr || int_scale == nullptr || float_scale == nullptr) { set_error("lpbq_encode: null pointer"); return FFQ_ERR_INVALID; }
  if (rows <= 0 || cols <= 0 || rows >= (1ll << 40) || cols >= (1ll << 40)) { set_error("lpbq_encode: bad shape"); return FFQ_ERR_INVALID; }
  if (bitwidth < 1 || bitwidth > 24) { set_error("lpbq_encode: bitwidth must be in [1, 24]"); return FFQ_ERR_BITWIDTH; }
  if (channel_axis != 0 && channel_axis != 1) { set_error("lpbq_encode: channel_axis must be 0 (rows) or 1 (columns)"); return FFQ_ERR_INVALID; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float two_bw = (float)(1u << bitwidth);
  if (channel_axis == 0) {
    const long long blocks = (rows + 7) / 8;
    const long long cap = (long long)sm_count() * 8;
    lpbq_rows_kernel<<<(unsigned int)(blocks < cap ? blocks : cap), 256, 0, st>>>(scale, rows, cols, two_bw, int_scale, float_scale);
  } else {
    lpbq_cols_kernel<<<(unsigned int)((cols + 255) / 256), 256, 0, st>>>(scale, rows, cols, two_bw, int_scale, float_scale);
  }
  FFQ_LAUNCH_CHECK();
  return FFQ_OK;
}

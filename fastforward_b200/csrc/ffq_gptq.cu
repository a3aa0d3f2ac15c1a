// ffq_gptq.cu -- the GPTQ inner block loop as ONE kernel per column block.
//
// Reference (quantization/gptq.py:100-132): for every column j of a block
//     q_j   = dequantize(quantize(w_j))                     per-row parameters (column_quantizer, :149-235)
//     e_j   = (w_j - q_j) / Hinv[j, j]
//     w_k  -= e_j * Hinv[j, k]            for the later columns k of the block
// i.e. per column one quantize, one dequantize, a subtraction, a division, a [R,1]x[1,n] matmul and an
// in-place subtraction: ~15 launches per column, thousands per layer.  Rows never interact (only through the
// shared Hinv block), so here one warp owns one row of the block: its <=128 columns live in registers (lane l
// holds columns l, l+32, l+64, l+96), the Hinv block sits in shared memory, and the 128 sequential steps are a
// shuffle broadcast of e_j followed by two FP32 instructions per owned column.
//
// Arithmetic is op-for-op the reference's: IEEE division by the scale, rint, clamp, (c + rint(o)) * s, then
// fl(fl(w - q) / d), fl(e * h) and fl(w - .) as separate roundings (a K=1 matmul is one rounded product), so the
// block's outputs are bit-identical to the reference given the same inputs (tests/test_gptq_gpu.py).
#include "ffq_common.cuh"

namespace ffq {

constexpr int GQ_MAXC = 128;            // columns per block handled in registers
constexpr int GQ_WARPS = 16;

struct GptqArgs {
  float* w; long long ldw;              // [R, ldw]: block columns, updated in place
  float* q; long long ldq;              // [R, ldq]: quantize-dequantized columns out
  float* err; long long lde;            // [R, lde]: errors out
  const float* hinv; long long ldh;     // [ncols, ldh]: Hinv[i:i+ncols, i:i+ncols] (upper triangle used)
  const void* scale; const void* offset; int s_dt, o_dt;
  const int* orig_col;                  // [ncols] column index in the ORIGINAL weight (activation ordering)
  int R, ncols;
  int rbs, cbs, ncb;                    // parameter index = (row / rbs) * ncb + orig_col / cbs
  float lo, hi;
  int code_is_int;                      // integer code dtype: -0 becomes +0 between quantize and dequantize
};

__global__ void __launch_bounds__(GQ_WARPS * 32) gptq_block_kernel(const GptqArgs a) {
  extern __shared__ float s_h[];        // [ncols][GQ_MAXC]
  __shared__ int s_col[GQ_MAXC];
  for (int i = threadIdx.x; i < a.ncols * GQ_MAXC; i += blockDim.x) {
    const int r = i / GQ_MAXC, c = i % GQ_MAXC;
    s_h[i] = (c < a.ncols) ? a.hinv[(long long)r * a.ldh + c] : 0.f;
  }
  for (int i = threadIdx.x; i < GQ_MAXC; i += blockDim.x) s_col[i] = i < a.ncols ? a.orig_col[i] : 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * GQ_WARPS + (threadIdx.x >> 5);
  if (row >= a.R) return;

  float w[4], s[4], o[4], qv[4], ev[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int c = lane + 32 * t;
    const bool in = c < a.ncols;
    w[t] = in ? a.w[row * a.ldw + c] : 0.f;
    const long long p = in ? (row / a.rbs) * (long long)a.ncb + s_col[c] / a.cbs : 0;
    s[t] = in ? load_as_float(a.scale, a.s_dt, p) : 1.f;
    o[t] = (in && a.offset) ? rintf(load_as_float(a.offset, a.o_dt, p)) : 0.f;
    qv[t] = 0.f; ev[t] = 0.f;
  }
#pragma unroll
  for (int t = 0; t < 4; ++t) {
#pragma unroll 1
    for (int jj = 0; jj < 32; ++jj) {
      const int j = t * 32 + jj;
      if (j >= a.ncols) break;
      // every lane runs the column arithmetic on its own slot-t value; lane jj's is the one that counts
      float c = __fsub_rn(__fdiv_rn(w[t], s[t]), o[t]);
      c = nan_clamp(rintf(c), a.lo, a.hi);
      if (a.code_is_int) c = __fadd_rn(c, 0.0f);
      const float qj = __fmul_rn(__fadd_rn(c, o[t]), s[t]);
      const float ej = __fdiv_rn(__fsub_rn(w[t], qj), s_h[j * GQ_MAXC + j]);
      if (lane == jj) { qv[t] = qj; ev[t] = ej; }
      const float e = __shfl_sync(0xffffffffu, ej, jj);
      const float* hrow = s_h + j * GQ_MAXC;
      if (lane > jj) w[t] = __fsub_rn(w[t], __fmul_rn(e, hrow[lane + 32 * t]));
#pragma unroll
      for (int t2 = t + 1; t2 < 4; ++t2) w[t2] = __fsub_rn(w[t2], __fmul_rn(e, hrow[lane + 32 * t2]));
    }
  }
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int c = lane + 32 * t;
    if (c < a.ncols) {
      a.w[row * a.ldw + c] = w[t];
      a.q[row * a.ldq + c] = qv[t];
      a.err[row * a.lde + c] = ev[t];
    }
  }
}

}  // namespace ffq

using namespace ffq;

extern "C" {

int ffq_gptq_block(float* w, int64_t ldw, float* q, int64_t ldq, float* err, int64_t lde, const float* hinv, int64_t ldh,
                   int64_t R, int64_t ncols, const void* scale, int scale_dtype, const void* offset, int offset_dtype,
                   const int32_t* orig_col, int64_t row_block, int64_t col_block, int64_t num_col_blocks, double num_bits,
                   int code_dtype, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (R <= 0 || ncols <= 0) return FFQ_OK;
  if (ncols > GQ_MAXC) { set_error("gptq_block: at most %d columns per block (got %lld)", GQ_MAXC, (long long)ncols); return FFQ_ERR_UNSUPPORTED; }
  if (!(scale_dtype == FFQ_F32 || scale_dtype == FFQ_F16 || scale_dtype == FFQ_BF16)) {
    set_error("gptq_block: unsupported scale dtype %s", dt_name(scale_dtype)); return FFQ_ERR_UNSUPPORTED;
  }
  if (offset == nullptr) offset_dtype = FFQ_NONE;
  else if (offset_dtype == FFQ_F64 || !(is_float_dt(offset_dtype) || is_int_dt(offset_dtype))) {
    set_error("gptq_block: unsupported offset dtype %s", dt_name(offset_dtype)); return FFQ_ERR_UNSUPPORTED;
  }
  if (row_block <= 0 || col_block <= 0 || num_col_blocks <= 0 || R > 0x7fffffffll) { set_error("gptq_block: bad parameter blocking"); return FFQ_ERR_INVALID; }
  const bool float_codes = code_dtype == FFQ_F32 || code_dtype == FFQ_F16 || code_dtype == FFQ_BF16;
  const int code_bits = float_codes ? 0 : 8 * dt_size(code_dtype);
  if (!float_codes && (!is_int_dt(code_dtype) || code_dtype == FFQ_U8 || code_bits < num_bits)) {
    set_error("gptq_block: code dtype %s cannot hold %g-bit signed codes", dt_name(code_dtype), num_bits); return FFQ_ERR_UNSUPPORTED;
  }
  if ((code_dtype == FFQ_BF16 && num_bits > 8) || (code_dtype == FFQ_F16 && num_bits > 11)) {
    set_error("gptq_block: %s codes are not exact for %g bits", dt_name(code_dtype), num_bits); return FFQ_ERR_UNSUPPORTED;
  }
  GptqArgs a{};
  a.w = w; a.ldw = ldw; a.q = q; a.ldq = ldq; a.err = err; a.lde = lde; a.hinv = hinv; a.ldh = ldh;
  a.scale = scale; a.offset = offset; a.s_dt = scale_dtype; a.o_dt = offset_dtype; a.orig_col = orig_col;
  a.R = (int)R; a.ncols = (int)ncols; a.rbs = (int)row_block; a.cbs = (int)col_block; a.ncb = (int)num_col_blocks;
  const double lo = -pow(2.0, num_bits - 1.0);
  a.lo = (float)lo; a.hi = (float)(-lo - 1.0);
  a.code_is_int = float_codes ? 0 : 1;
  const size_t smem = (size_t)ncols * GQ_MAXC * sizeof(float);
  static std::atomic<bool> attr_done{false};
  if (!attr_done.load()) {
    FFQ_CUDA_CHECK(cudaFuncSetAttribute(gptq_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GQ_MAXC * GQ_MAXC * 4));
    attr_done.store(true);
  }
  const unsigned int grid = (unsigned int)((R + GQ_WARPS - 1) / GQ_WARPS);
  gptq_block_kernel<<<grid, GQ_WARPS * 32, smem, st>>>(a);
  FFQ_LAUNCH_CHECK();
  return FFQ_OK;
}

}  // extern "C"

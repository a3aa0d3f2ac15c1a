/*
 * ffq_b200.h -- C ABI of the B200-native backend for FastForward's quantization hot path.
 *
 * This is the drop-in boundary.  Every entry point takes plain device
 * (or, for the *_host variants, host) pointers, dtype tags, a tile-layout descriptor and a CUDA
 * stream, and returns an int status.  No torch types appear here.  The Python host layer
 * (fastforward_b200/_cabi.py) binds these with ctypes; INTEGRATION.md shows the stub that
 * registers them under the reference's own op names.
 *
 * All "replaces:" citations are into the reference tree, /root/reference/src/fastforward/.
 *
 * Conventions
 *   - Tensors are dense, row-major ("contiguous" in torch terms).  `layout` gives the data
 *     shape and the tile shape; parameters (scale/offset/min/max/grads) are flat arrays of
 *     length num_tiles, tile index row-major over the block grid -- the ordering defined by
 *     quantization/tiled_tensor.py:71-98.
 *   - Arithmetic follows PyTorch eager semantics op by op: IEEE division, round-half-even,
 *     NaN-propagating clamp/min/max, one rounding to the promoted dtype after every op
 *     (SURVEY.md Appendix B).
 *   - Functions never allocate device memory and never synchronise the stream (except the
 *     *_host variants, which own their staging buffers and return after the result is on the
 *     host).  Scratch space is provided by the caller; ffq_workspace_bytes() sizes it.
 *   - Thread safety: no global mutable state except a per-thread last-error string and a
 *     per-device attribute cache.  One process per GPU is the intended deployment.
 *   - Return value: FFQ_OK or an ffq_status_t error; ffq_last_error() returns the message.
 */
#ifndef FFQ_B200_H
#define FFQ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(FFQ_BUILD) && defined(__GNUC__)
#define FFQ_API __attribute__((visibility("default")))
#else
#define FFQ_API
#endif

#define FFQ_ABI_VERSION 2
#define FFQ_MAX_RANK 8

typedef enum {
  FFQ_OK = 0,
  FFQ_ERR_INVALID = 1,     /* bad argument / tile does not divide shape  -> ValueError        */
  FFQ_ERR_UNSUPPORTED = 2, /* dtype/layout the backend does not implement -> NotImplementedError */
  FFQ_ERR_CUDA = 3,        /* CUDA runtime/driver failure                 -> RuntimeError      */
  FFQ_ERR_BITWIDTH = 4,    /* output dtype cannot hold num_bits           -> RuntimeError
                              (quantization/_quantizer_impl.py:165-167)                        */
  FFQ_ERR_WORKSPACE = 5    /* workspace too small                         -> RuntimeError      */
} ffq_status_t;

typedef enum {
  FFQ_F32 = 0,
  FFQ_F16 = 1,
  FFQ_BF16 = 2,
  FFQ_F64 = 3, /* accepted in the enum; kernels return FFQ_ERR_UNSUPPORTED for it */
  FFQ_I8 = 4,
  FFQ_I16 = 5,
  FFQ_I32 = 6,
  FFQ_U8 = 7,
  FFQ_I64 = 8,
  FFQ_NONE = 255 /* "this optional tensor is absent" */
} ffq_dtype_t;

/* Shape of the data tensor and of one tile.  tile[i] must divide dims[i]
 * (quantization/tiled_tensor.py:19-42). */
typedef struct {
  int32_t rank;
  int64_t dims[FFQ_MAX_RANK];
  int64_t tile[FFQ_MAX_RANK];
} ffq_layout_t;

/* Bits of the `allow_one_sided` argument of ffq_params_for_range, ffq_params_for_ranges_encode, ffq_dynamic_quantize,
 * ffq_calibrate_quantize and ffq_calibrate_fakequant (0 / 1 keep their plain meaning).
 * FFQ_FLAG_SCALAR_DIV_RECIPROCAL: do the three tensor / Python-scalar divisions of parameters_for_range
 * (affine/range.py:112-120) as aten's CUDA kernel does them, x * (1.0f / d), instead of the IEEE division of aten's CPU
 * kernel -- one ulp apart for divisors that are not powers of two.  The plugin sets it, so that a GPU run of the
 * unmodified reference gets bit-identical parameters with and without this backend; the default (CPU flavour) is what
 * the reference's golden vectors pin. */
#define FFQ_FLAG_ALLOW_ONE_SIDED 1
#define FFQ_FLAG_SCALAR_DIV_RECIPROCAL 2

/* Which scratch requirement ffq_workspace_bytes() reports. */
typedef enum {
  FFQ_WS_QUANTIZE_BWD = 0,
  FFQ_WS_MINMAX = 1,
  FFQ_WS_PARAMS_FOR_RANGE = 2,
  FFQ_WS_DYNAMIC_QUANTIZE = 3
} ffq_ws_kind_t;

/* ---- library -------------------------------------------------------------------------- */
FFQ_API int ffq_abi_version(void);
FFQ_API const char* ffq_last_error(void);
/* Number of kernel launches issued by this library since load (all threads).  bench.py reads
 * it around the timed region to report "gpu_launches". */
FFQ_API uint64_t ffq_launch_count(void);
FFQ_API size_t ffq_workspace_bytes(int kind, const ffq_layout_t* layout, int data_dtype);
/* Number of tiles / parameters implied by a layout; -1 if the tile does not divide the shape. */
FFQ_API int64_t ffq_num_tiles(const ffq_layout_t* layout);

/* ---- a1: quantize ------------------------------------------------------------------------
 * q = cast(clamp(rint(x / s_t - rint(o_t)), -2^(b-1), 2^(b-1)-1), q_dtype)
 * replaces: quantization/_quantizer_impl.py:144-169 (torch.ops.fastforward.quantize_by_tile)
 * `offset` may be NULL (offset_dtype FFQ_NONE): symmetric, not one-sided. */
FFQ_API int ffq_quantize(const void* x, int x_dtype, void* q, int q_dtype,
                 const void* scale, int scale_dtype, const void* offset, int offset_dtype,
                 const ffq_layout_t* layout, double num_bits, void* stream);

/* ---- a2: dequantize ----------------------------------------------------------------------
 * y = cast((q + rint(o_t)) * s_t, y_dtype)
 * replaces: quantization/_quantizer_impl.py:172-190 (torch.ops.fastforward.dequantize_by_tile),
 *           reached from QuantizedTensor.dequantize, quantized_tensor.py:384-388. */
FFQ_API int ffq_dequantize(const void* q, int q_dtype, void* y, int y_dtype,
                   const void* scale, int scale_dtype, const void* offset, int offset_dtype,
                   const ffq_layout_t* layout, void* stream);

/* ---- a1+a2 fused: fake-quantize ----------------------------------------------------------
 * y = dequantize(cast(quantize(x), q_dtype)) in ONE pass (x read once, y written once); if
 * `codes` is non-NULL the integer codes are stored too (dtype q_dtype).  Bit-identical to
 * ffq_quantize followed by ffq_dequantize.
 * replaces: the quantize->dequantize pair of affine/function.py:94-121 (export mode) and of
 *           quantization/fuse.py:91-121 (`weight.copy_(quantizer(weight).dequantize())`). */
FFQ_API int ffq_fakequant_fwd(const void* x, int x_dtype, void* y, int y_dtype, void* codes, int q_dtype,
                      const void* scale, int scale_dtype, const void* offset, int offset_dtype,
                      const ffq_layout_t* layout, double num_bits, void* stream);

/* ---- a3: straight-through backward ------------------------------------------------------
 * dx = clip ? 0 : g;  dscale_t = sum_tile g*(clip ? bound+rint(o) : q-pre);
 * doffset_t = sum_tile (clip ? s*g : 0)   (doffset may be NULL when offset is NULL)
 * dx has the dtype of g; dscale has scale_dtype; doffset has promote(scale_dtype, g_dtype).
 * Per-tile sums use a fixed reduction tree: deterministic run to run, no atomics.
 * replaces: quantization/_quantizer_impl.py:193-237 (quantize_by_tile_backward). */
FFQ_API int ffq_quantize_bwd(const void* x, int x_dtype, const void* g, int g_dtype, void* dx,
                     void* dscale, void* doffset,
                     const void* scale, int scale_dtype, const void* offset, int offset_dtype,
                     const ffq_layout_t* layout, double num_bits,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ---- a7: per-tile min/max, optionally merged into running ranges -------------------------
 * tile_min/tile_max (dtype of x, may be NULL) receive this batch's per-tile extrema.
 * If run_min/run_max are non-NULL (dtype run_dtype: the data dtype, or a wider float when an
 * existing fp32 range is being continued with bf16/fp16 data -- torch.min promotes) they are
 * updated in place: run_min = min(run_min, tile_min), run_max = max(run_max, tile_max) (NaN
 * propagates, as torch.min/torch.max do).
 * If `flags` (int32[1], device) is non-NULL, bit 0 is OR-ed in when any tile extremum of this
 * batch is +-inf -- the condition for which the reference raises NotImplementedError
 * (range_setting/minmax.py:233-234); the host checks it at its next sync point.
 * replaces: range_setting/minmax.py:226-237 (RunningMinMaxEstimator.estimate_step). */
FFQ_API int ffq_minmax(const void* x, int x_dtype, void* tile_min, void* tile_max,
               void* run_min, void* run_max, int run_dtype, int32_t* flags,
               const ffq_layout_t* layout, void* workspace, size_t workspace_bytes, void* stream);

/* ---- a6: range -> (scale, offset), sync-free ---------------------------------------------
 * fp32 arithmetic throughout (affine/range.py:89-90).  The global one-sided decision
 * `min.min() >= 0 and allow_one_sided` (range.py:100, a host sync in the reference) is taken
 * on the device.  offset_out may be NULL (symmetric and not allow_one_sided); when the
 * symmetric two-sided branch is taken and offset_out exists it is filled with 0
 * (nn/linear_quantizer.py:353-357).  round_offset != 0 stores rint(offset) (dynamic path,
 * _quantizer_impl.py:275).
 * replaces: quantization/affine/range.py:54-122 + nn/linear_quantizer.py:347-357. */
FFQ_API int ffq_params_for_range(const void* min_range, const void* max_range, int range_dtype, int64_t n,
                         double num_bits, int symmetric, int allow_one_sided, int round_offset,
                         void* scale_out, int scale_dtype, void* offset_out, int offset_dtype,
                         void* workspace, size_t workspace_bytes, void* stream);

/* Batched form for MANY quantizers whose ranges live in one contiguous buffer (the calibration block's arena): one
 * launch, one CTA per quantizer.  desc_dev: device int64 [num_quantizers][8] =
 * {start element in the range buffers, number of tiles, scale pointer (fp32), offset pointer (fp32, 0 = none),
 *  words[0], words[1], words[2] of ffq_params_for_ranges_encode(num_bits, symmetric, allow_one_sided), 0}.
 * Results equal ffq_params_for_range called per quantizer (round_offset = 0).  Used when the data-parallel range
 * exchange has merged the ranges of all quantizers at the end of an estimate_ranges block. */
FFQ_API int ffq_params_for_ranges_batched(const void* min_base, const void* max_base, int range_dtype,
                                  const int64_t* desc_dev, int64_t num_quantizers, void* stream);
FFQ_API void ffq_params_for_ranges_encode(double num_bits, int symmetric, int allow_one_sided, int64_t words[3]);

/* ---- a4: dynamic quantize -----------------------------------------------------------------
 * per-tile min/max -> params -> quantize, three launches, no host sync.
 * scale_out/offset_out are fp32[num_tiles]; offset_out is the rounded offset (zeros when the
 * symmetric two-sided branch is taken).
 * replaces: quantization/_quantizer_impl.py:243-285 (quantize_dynamic_by_tile). */
FFQ_API int ffq_dynamic_quantize(const void* x, int x_dtype, void* q, int q_dtype,
                         float* scale_out, float* offset_out,
                         const ffq_layout_t* layout, double num_bits, int symmetric, int allow_one_sided,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ---- a7 + a6 + a1 fused: one RunningMinMax calibration step of one quantizer -----------------
 * tile min/max of x -> run_min/run_max updated in place (NaN propagates; bit 0 of *flags is OR-ed in
 * when a tile extremum is +-inf) -> (scale, offset) for the UPDATED running range -> int8 codes of x
 * under those parameters [-> rowsum[r] = sum of the codes of row r, rows of rowsum_row_len elements],
 * with x read from HBM once.  Bit-identical to ffq_minmax + ffq_params_for_range + ffq_quantize
 * (+ ffq_rowsum_i8).  x: f32/f16/bf16; q: int8, num_bits <= 8; scale/offset: fp32 [num_tiles]
 * (offset may be NULL only for symmetric && !allow_one_sided; it is filled with 0 on the symmetric
 * two-sided branch); run_dtype must hold x_dtype.
 * Layouts: contiguous tiles of 64..4096 16-byte vectors (per-channel weight rows; rowsum_row_len must
 * equal the tile length), ONE tile (per-tensor; rowsum_row_len divides numel), or contiguous tiles of
 * 1, 2, 4, ... 32 vectors (per-group weights, e.g. g = 128; no row sums).  Anything else
 * returns FFQ_ERR_UNSUPPORTED; ffq_calibrate_quantize_mode() tells in advance (0 unsupported, 1 rows,
 * 2 per-tensor, 3 per-group).  The per-tensor kernel is a cooperative launch with a grid barrier: `workspace`
 * (ffq_calibrate_quantize_workspace_bytes(), 16-byte aligned) must be ZERO before its first use and
 * be used by one stream at a time; bit 1 of *flags reports a barrier time-out (never expected).
 * Rows path, symmetric && allow_one_sided: the global one-sided decision (range.py:100) is taken by a second,
 * tiny fix-up launch that finishes exactly the rows whose running min is >= 0 (normally none).  `settled`
 * (optional int32[1], zero-initialised, one per running range) is set by that launch once every running min
 * is negative -- permanent, because running mins only decrease; a caller that has OBSERVED *settled != 0 may
 * pass run_fixup = 0 from then on and save the launch.  run_fixup must be non-zero otherwise.
 * replaces: range_setting/minmax.py:215-239 + nn/linear_quantizer.py:347-357 +
 *           quantization/affine/range.py:54-122 + quantization/_quantizer_impl.py:144-169. */
FFQ_API int ffq_calibrate_quantize(const void* x, int x_dtype, int8_t* q,
                           void* run_min, void* run_max, int run_dtype,
                           float* scale, float* offset, int32_t* rowsum, int64_t rowsum_row_len,
                           int32_t* flags, int32_t* settled, int run_fixup,
                           const ffq_layout_t* layout, double num_bits, int symmetric, int allow_one_sided,
                           void* workspace, size_t workspace_bytes, void* stream);
FFQ_API int ffq_calibrate_quantize_mode(const ffq_layout_t* layout, int x_dtype);
FFQ_API size_t ffq_calibrate_quantize_workspace_bytes(void);

/* ---- a7 + a6 + (a1 + a2) fused: calibrate a quantizer on a tensor and snap the tensor to its grid ------------
 * tile min/max of x [merged into run_min/run_max when given] -> (scale, offset) -> y = dequantize(quantize(x)) under
 * those parameters, x read once; y may alias x (in place).  This is `quantizer.quantization_range = (tile min, tile
 * max)` followed by `w.copy_(quantizer(w).dequantize())` -- the unit of work of whole-model weight quantization
 * (BASELINE config 3) -- in 2s bytes per element instead of 3s and one launch (+ the one-sided fix-up launch for
 * symmetric && allow_one_sided quantizers, which returns after one load unless a tile is entirely non-negative).
 * Layouts: modes 1 (rows) and 3 (per-group) of ffq_calibrate_quantize_mode; x, y: f32/f16/bf16 of the same dtype;
 * scale/offset fp32; code_dtype = the quantizer's quantized_dtype (or the data dtype): integer codes turn -0 into +0.
 * `workspace`: >= 16 bytes, shared with ffq_calibrate_quantize.  Bit-identical to ffq_minmax + ffq_params_for_range +
 * ffq_fakequant_fwd.
 * replaces: quantization/fuse.py:91-121 after range_setting/minmax.py:215-239 + nn/linear_quantizer.py:347-357. */
FFQ_API int ffq_calibrate_fakequant(const void* x, int x_dtype, void* y, void* run_min, void* run_max, int run_dtype,
                            float* scale, float* offset, int32_t* flags, const ffq_layout_t* layout, double num_bits,
                            int symmetric, int allow_one_sided, int code_dtype,
                            void* workspace, size_t workspace_bytes, void* stream);

/* The same for MANY tensors in one launch (+ one fix-up launch): whole-model weight fake-quant, the loop of
 * quantization/fuse.py:199-241 over every quantized linear, without one launch (and one launch tail) per weight.
 * items_dev: device array of descriptors; block_start_dev: device uint32[num_items + 1], prefix sums of
 * ceil(num_tiles_i / 256) (CTAs per tensor), total_blocks = block_start[num_items].  All tensors: dtype x_dtype
 * (bf16 / f16), 32-byte aligned, contiguous tiles of tile_len (64 or 128) elements, one quantizer configuration;
 * y may alias x.  workspace: >= 8 * num_items bytes (zeroed by the call).  Per tensor bit-identical to
 * ffq_calibrate_fakequant without a running range.  Anything else returns FFQ_ERR_UNSUPPORTED (use the per-tensor call). */
typedef struct {
  const void* x;
  void* y;
  float* scale;   /* fp32 [numel / tile_len] */
  float* offset;  /* fp32 [numel / tile_len] or NULL */
  int64_t numel;
} ffq_fq_item_t;
FFQ_API int ffq_calibrate_fakequant_batched(const ffq_fq_item_t* items_dev, const uint32_t* block_start_dev,
                                    int64_t num_items, int64_t total_blocks, int x_dtype, int64_t tile_len,
                                    double num_bits, int symmetric, int allow_one_sided, int code_dtype,
                                    void* workspace, size_t workspace_bytes, void* stream);

/* ---- a12: quantized linear (new kernel behind ff.dispatcher "linear") ---------------------
 * y[m,n] = sx * sw[n] * ( sum_k qx[m,k] qw[n,k] + ox*rowsum_w[n] + ow[n]*rowsum_x[m] + K*ox*ow[n] )
 *          + bias[n]
 * qx: int8 [M,K] per-tensor (sx, ox scalars on the device, fp32); qw: int8 [N,K] per-channel
 * (sw, ow fp32[N]; ow may be NULL).  int32 accumulation on tcgen05 (kind::i8) tensor cores,
 * dequantisation fused into the epilogue.  y_dtype in {F32, BF16, F16}.
 * rowsum_w: int32[N] precomputed with ffq_rowsum_i8 (weights are static); rowsum_x may be NULL
 * when ow is NULL.  K must be a multiple of 16 (TMA row pitch); M and N are arbitrary.
 * The offset corrections are added in int32 (exact) whenever they fit; for offsets so far from zero that
 * K*ox*ow or ox*rowsum_w could leave int32 they are added in float, like the reference's float path.
 * `requant` (may be NULL): the linear's output quantizer fused into the epilogue -- see ffq_requant_t.
 * replaces: _gen/fallback.py:77-112 (dequantize x2 + torch.nn.functional.linear [+ output_quantizer,
 *           fallback.py:109-110]), selected via dispatcher.py:268-283. */
/* Static per-tensor output quantizer applied in the GEMM epilogue:
 *   codes[m,n] = int8(clamp(rint(cast_y_dtype(y[m,n]) / scale - rint(offset)), -2^(b-1), 2^(b-1)-1))
 * i.e. quantize_by_tile (_quantizer_impl.py:144-169) of the output tensor the fallback would have written, without
 * that tensor making a round trip through HBM.  scale/offset: device fp32[1] (offset may be NULL); num_bits <= 8;
 * codes: int8 [M,N]; rowsum (may be NULL): int32[M], ZERO-initialised by the caller, receives the row sums of the
 * codes (what the next W8A8 linear needs).  `y` of ffq_qlinear_w8a8 may be NULL when only the codes are wanted. */
typedef struct {
  const float* scale;
  const float* offset;
  double num_bits;
  int8_t* codes;
  int32_t* rowsum;
} ffq_requant_t;
FFQ_API int ffq_qlinear_w8a8(const int8_t* qx, const int8_t* qw, void* y, int y_dtype,
                     int64_t M, int64_t N, int64_t K,
                     const float* sx, const float* ox, const float* sw, const float* ow,
                     const int32_t* rowsum_w, const int32_t* rowsum_x,
                     const void* bias, int bias_dtype,
                     const ffq_requant_t* requant, void* stream);

/* rowsum[r] = sum_k q[r,k]  (int8 [R,K] -> int32[R]) */
FFQ_API int ffq_rowsum_i8(const int8_t* q, int32_t* rowsum, int64_t R, int64_t K, void* stream);

/* ---- a12: weight-only quantized linear (W4A16; any integer weight codes of <= 8 bits) --------
 * y[m,n] = sum_k x[m,k] * w[n,k] + bias[n],  w[n,k] = cast_x_dtype((qw[n,k] + rint(ow[n,k/group])) * sw[n,k/group])
 * x, y: bf16 or f16 [M,K] / [M,N]; qw: int8 codes [N,K]; sw, ow: fp32 [N, K/group] (ow may be NULL);
 * group: elements of K that share one parameter (a multiple of 64 that divides K; group == K is
 * per-channel).  The codes are dequantized inside the k-loop with dequantize_by_tile's arithmetic
 * and fed to tcgen05.mma.kind::f16 (fp32 accumulation), so the result equals the reference's
 * fallback up to the GEMM's accumulation order, without the dequantized weight ever touching HBM.
 * replaces: _gen/fallback.py:77-112 (dequantize weight + torch.nn.functional.linear), selected via
 *           dispatcher.py:268-283. */
FFQ_API int ffq_qlinear_w4a16(const void* x, int x_dtype, const int8_t* qw, void* y,
                      int64_t M, int64_t N, int64_t K,
                      const float* sw, const float* ow, int64_t group,
                      const void* bias, int bias_dtype, void* stream);

/* ---- host-buffer convenience (end-to-end path; copies inside) -----------------------------
 * Fake-quant forward + STE backward of one tensor whose data lives in HOST memory:
 * H2D(x, g) -> ffq_fakequant_fwd -> ffq_quantize_bwd -> D2H(y, dx, dscale, doffset).
 * scale/offset are host arrays too.  All dtypes as in the device entry points.
 * Returns after the results are in the host buffers. */
FFQ_API int ffq_fakequant_fwd_bwd_host(const void* x_host, const void* g_host, int dtype,
                               void* y_host, void* dx_host, float* dscale_host, float* doffset_host,
                               const float* scale_host, const float* offset_host,
                               const ffq_layout_t* layout, double num_bits, int device);

/* ---- f2: fused MSE grid search ---------------------------------------------------------------
 * err_accum[c][t] += mean over tile t of (dequantize_c(quantize_c(x)) - x)^2 for all C candidate
 * (scale, offset) sets in ONE pass over x.  cand_scale / cand_offset: fp32 [C][num_tiles]
 * (cand_offset may be NULL); err_accum: fp32 [C][num_tiles].  Contiguous-tile layouts only
 * (FFQ_ERR_UNSUPPORTED otherwise: the caller evaluates candidates one by one).
 * replaces: range_setting/min_error.py:206-221 (_MinAvgErrorGridEstimator.estimate_step). */
FFQ_API int ffq_grid_mse(const void* x, int x_dtype, const float* cand_scale, const float* cand_offset,
                 int num_candidates, float* err_accum, const ffq_layout_t* layout, double num_bits,
                 void* workspace, size_t workspace_bytes, void* stream);
FFQ_API size_t ffq_grid_mse_workspace_bytes(const ffq_layout_t* layout, int x_dtype, int num_candidates);

/* ---- f3: GPTQ inner block loop ----------------------------------------------------------------
 * For the `ncols` (<= 128) columns of one block, in order j = 0..ncols-1 and for every row r independently:
 *     q[r,j]   = dequantize(quantize(w[r,j]))          parameters of (r, orig_col[j]), see below
 *     err[r,j] = (w[r,j] - q[r,j]) / hinv[j,j]
 *     w[r,k]  -= err[r,j] * hinv[j,k]                   for k = j+1 .. ncols-1
 * w: fp32 [R, ldw] (the block's columns, updated in place -- the reference's `weights_block`); q, err: fp32
 * outputs with their own leading dimensions; hinv: fp32 [ncols, ldh], the diagonal block of the upper Cholesky
 * factor of the inverse Hessian.  The parameter of element (r, c) is scale[(r / row_block) * num_col_blocks +
 * c / col_block] with c = orig_col[j] the column index in the un-permuted weight: per-tensor (row_block = R,
 * col_block = C), per-channel(0) (1, C), per-channel(1) (R, 1), per-element (1, 1), per-block/tile.
 * code_dtype: dtype the codes take between quantize and dequantize (the quantizer's quantized_dtype or f32).
 * Every step is rounded like the reference's separate aten ops (no FMA contraction), so the results are
 * bit-identical to quantization/gptq.py:100-132 + column_quantizer (:149-235) on the same inputs.
 * replaces: quantization/gptq.py:106-130 (the `for j in range(block_end - i)` loop). */
FFQ_API int ffq_gptq_block(float* w, int64_t ldw, float* q, int64_t ldq, float* err, int64_t lde,
                   const float* hinv, int64_t ldh, int64_t R, int64_t ncols,
                   const void* scale, int scale_dtype, const void* offset, int offset_dtype,
                   const int32_t* orig_col, int64_t row_block, int64_t col_block, int64_t num_col_blocks,
                   double num_bits, int code_dtype, void* stream);

/* ---- f4: LPBQ scale compression (on-disk encodings of per-block weights) --------------------
 * scale: fp32 [rows, cols] per-block scales of ONE weight.  channel_axis = 0: a channel is a row (the scales of
 * PerBlock(block_dims=1, per_channel_dims=0), [out_channels, blocks]); 1: a channel is a column
 * (PerBlock(block_dims=0, per_channel_dims=1), [blocks, in_channels]).  For every channel c
 *     float_scale[c]  = max over the channel's blocks of scale / 2^bitwidth
 *     int_scale[., .] = clamp(rint(scale / float_scale[c]), 1, 2^bitwidth)      (same shape as scale, int32)
 * with aten's roundings (IEEE fp32 division, half-to-even), bit-identical to the reference on the same input.
 * replaces: export/_lpbq.py:131-160 (LPBQProcessor.grouped_dynamic_quantize: amax, 2x div, round, clamp, cast). */
FFQ_API int ffq_lpbq_encode(const float* scale, int64_t rows, int64_t cols, int channel_axis, int bitwidth,
                            int32_t* int_scale, float* float_scale, void* stream);

/* ---- test hook ---------------------------------------------------------------------------
 * Sweeps the kernels' shared-reciprocal division against __fdiv_rn over n pseudo-random
 * (dividend, scale) pairs.  counts_dev: uint64[4], zero-initialised by the caller:
 * [0] accepted quotients (>= 2^-50 in magnitude) that differ from __fdiv_rn  -- must stay 0
 * [1] quotients accepted by the magnitude guard   [2] accepted by the strict guard
 * [3] strict-accepted quotients that differ from __fdiv_rn                   -- must stay 0 */
FFQ_API int ffq_selftest_shared_div(unsigned long long n, unsigned int seed, unsigned long long* counts_dev,
                                    void* stream);

/* Test hook: while `counters_dev` is non-NULL, the CTA-pair kernel of ffq_qlinear_w8a8 records per CTA sixteen uint64 of
 * SM clocks (see GemmArgs::prof in csrc/ffq_qlinear.cu): how long the TMA producer waited for a free stage, the MMA
 * issuer for operands and for a free accumulator, the epilogue for a finished tile, each role's total, and the
 * epilogue's phases (column parameters, tcgen05.ld, staging box, arithmetic, store issue).
 * counters_dev: uint64[16 * grid] (grid <= 8 * SM count).  NULL switches it off.  Used by tools/prof_gemm_roles.py. */
FFQ_API void ffq_debug_gemm_profile(unsigned long long* counters_dev);

#ifdef __cplusplus
}
#endif
#endif /* FFQ_B200_H */

// ffq_gridsearch.cu -- fused MSE grid search.
//
// The reference's MinErrorGridRangeEstimator evaluates C candidate ranges per batch by running
// the whole quantize -> dequantize -> (y - x)^2 -> mean chain C times over the tensor
// (range_setting/min_error.py:206-221): C x ~30 full-tensor passes.  Here the tensor is read ONCE:
// a warp keeps a segment of its tile in registers and loops over the candidates, so the kernel is
// bound by the FP32 pipes (about 14 instructions per element per candidate), not by HBM.
//   err[c][t] += mean over tile t of (dequant_c(quant_c(x)) - x)^2
// Partial sums of multi-segment tiles go through a workspace and a fixed-order second stage:
// deterministic, no floating-point atomics.
#include "ffq_common.cuh"

namespace ffq {

constexpr int GS_THREADS = 256;

struct GsArgs {
  const void* x; int x_dt;
  const float* cscale; const float* coffset;   // [C][num_tiles]
  int C;
  float* err;                                   // [C][num_tiles], accumulated
  float* part;                                  // [units][C] when segs_per_tile > 1
  unsigned long long tile_numel, num_tiles, total_units;
  unsigned int segs_per_tile, seg_vecs;
  float lo, hi, inv_tile;
  int m_out;                                    // rounding of the data dtype (y, y - x and its square are data-dtype ops)
};

// G lanes share one unit (a tile, or a 32*U-vector segment of a long tile); 32/G units per warp.
// Short tiles (per-group quantization, g = 128) take G < 32 so that a lane still holds U vectors
// and the per-candidate overhead (parameter loads, reciprocal, G-lane sum) is amortised.
template <typename T, int U, int G>
__global__ void __launch_bounds__(GS_THREADS) grid_mse_kernel(const GsArgs a) {
  constexpr int EPT = 16 / sizeof(T);
  // y, y - x and its square are ops of the data dtype: round through it (compile-time mode)
  constexpr int RM = sizeof(T) == 4 ? RM_F32 : (Elem<T>::dt == FFQ_BF16 ? RM_BF16 : RM_F16);
  constexpr int UPW = 32 / G;                    // units per warp
  const unsigned int lane = threadIdx.x & 31;
  const unsigned int gl = lane % G;
  const unsigned long long warp = (unsigned long long)blockIdx.x * (GS_THREADS / 32) + (threadIdx.x >> 5);
  const unsigned long long unit = warp * UPW + lane / G;
  const bool unit_ok = unit < a.total_units;     // whole groups go idle together; shuffles stay warp-wide
  const unsigned long long tile = unit_ok ? unit / a.segs_per_tile : 0;
  const unsigned int seg = unit_ok ? (unsigned int)(unit - tile * a.segs_per_tile) : 0;
  const unsigned int tvec = (unsigned int)(a.tile_numel / EPT);
  const unsigned int vec0 = seg * a.seg_vecs;
  const unsigned int nv = unit_ok ? ((tvec - vec0) < a.seg_vecs ? (tvec - vec0) : a.seg_vecs) : 0;
  const T* __restrict__ in = static_cast<const T*>(a.x) + tile * a.tile_numel + (unsigned long long)vec0 * EPT;

  float x[U][EPT];
  bool live[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const unsigned int j = gl + u * G;
    live[u] = j < nv;
    if (live[u]) {
      const Vec<T, EPT> v = ld_stream<T, EPT>(in + (size_t)j * EPT);
#pragma unroll
      for (int i = 0; i < EPT; ++i) x[u][i] = Elem<T>::to_f(v.v[i]);
    } else {
#pragma unroll
      for (int i = 0; i < EPT; ++i) x[u][i] = 0.f;
    }
  }
  for (int c = 0; c < a.C; ++c) {
    const float s = a.cscale[(size_t)c * a.num_tiles + tile];
    const float o = a.coffset ? rintf(a.coffset[(size_t)c * a.num_tiles + tile]) : 0.f;
    const SharedRcp k = make_shared_rcp(s);
    float err = 0.f;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!live[u]) continue;
      float t[EPT];
      float amax = 0.f;
#pragma unroll
      for (int i = 0; i < EPT; ++i) {
        const float q0 = __fmul_rn(x[u][i], k.r);
        const float e = __fmaf_rn(-k.s, q0, x[u][i]);
        t[i] = __fmaf_rn(k.r, e, q0);
        amax = nan_max(amax, fabsf(t[i]));
      }
      if (!(k.ok && amax <= 0x1p60f)) {          // outside the proven box: plain IEEE division
#pragma unroll
        for (int i = 0; i < EPT; ++i) t[i] = __fdiv_rn(x[u][i], s);
      }
#pragma unroll
      for (int i = 0; i < EPT; ++i) {
        const float q = nan_clamp(rintf(__fsub_rn(t[i], o)), a.lo, a.hi);
        const float y = rnd(__fmul_rn(__fadd_rn(q, o), s), RM);
        const float d = rnd(__fsub_rn(y, x[u][i]), RM);
        err = __fadd_rn(err, rnd(__fmul_rn(d, d), RM));
      }
    }
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) err = __fadd_rn(err, __shfl_xor_sync(0xffffffffu, err, off));
    if (gl == 0 && unit_ok) {
      if (a.segs_per_tile == 1) a.err[(size_t)c * a.num_tiles + tile] += err * a.inv_tile;
      else a.part[unit * a.C + c] = err;
    }
  }
}

__global__ void __launch_bounds__(GS_THREADS) grid_mse_finalize_kernel(const GsArgs a) {
  const unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;   // (tile, c)
  if (idx >= a.num_tiles * (unsigned long long)a.C) return;
  const unsigned long long tile = idx / a.C;
  const int c = (int)(idx - tile * a.C);
  float sum = 0.f;
  for (unsigned int sgi = 0; sgi < a.segs_per_tile; ++sgi) sum += a.part[(tile * a.segs_per_tile + sgi) * a.C + c];
  a.err[(size_t)c * a.num_tiles + tile] += sum * a.inv_tile;
}

}  // namespace ffq

using namespace ffq;

static int gs_lanes(unsigned long long tvec, int U) {   // lanes per unit: pow2, each lane ~U vectors
  int g = 1;
  while (g < 32 && (unsigned long long)g * U < tvec) g <<= 1;
  return g;
}

static int gs_geometry(const Plan& plan, int x_dtype, unsigned int* seg_vecs, unsigned int* segs_per_tile, int* lanes) {
  const int sz = dt_size(x_dtype);
  const int ept = 16 / sz;
  const int U = sz == 4 ? 8 : 4;
  if (!plan.row || plan.tile_numel % ept != 0 || plan.tile_numel / ept >= (1ll << 31)) return 0;
  *lanes = gs_lanes((unsigned long long)(plan.tile_numel / ept), U);
  *seg_vecs = (unsigned int)(*lanes) * U;
  *segs_per_tile = (unsigned int)((plan.tile_numel / ept + *seg_vecs - 1) / *seg_vecs);
  return 1;
}

template <typename T, int U>
static void gs_launch(const GsArgs& a, int lanes, cudaStream_t st) {
  const unsigned long long warps = (a.total_units + (32 / lanes) - 1) / (32 / lanes);
  const unsigned int blocks = (unsigned int)((warps + GS_THREADS / 32 - 1) / (GS_THREADS / 32));
  switch (lanes) {
    case 1: grid_mse_kernel<T, U, 1><<<blocks, GS_THREADS, 0, st>>>(a); break;
    case 2: grid_mse_kernel<T, U, 2><<<blocks, GS_THREADS, 0, st>>>(a); break;
    case 4: grid_mse_kernel<T, U, 4><<<blocks, GS_THREADS, 0, st>>>(a); break;
    case 8: grid_mse_kernel<T, U, 8><<<blocks, GS_THREADS, 0, st>>>(a); break;
    case 16: grid_mse_kernel<T, U, 16><<<blocks, GS_THREADS, 0, st>>>(a); break;
    default: grid_mse_kernel<T, U, 32><<<blocks, GS_THREADS, 0, st>>>(a); break;
  }
}

extern "C" {

size_t ffq_grid_mse_workspace_bytes(const ffq_layout_t* layout, int x_dtype, int num_candidates) {
  Plan plan;
  if (make_plan(layout, &plan) != FFQ_OK || plan.numel == 0) return 0;
  if (!(x_dtype == FFQ_F32 || x_dtype == FFQ_F16 || x_dtype == FFQ_BF16)) return 0;
  unsigned int sv, spt;
  int lanes;
  if (!gs_geometry(plan, x_dtype, &sv, &spt, &lanes) || spt <= 1) return 0;
  return (size_t)plan.num_tiles * spt * (size_t)num_candidates * sizeof(float);
}

int ffq_grid_mse(const void* x, int x_dtype, const float* cand_scale, const float* cand_offset, int num_candidates,
                 float* err_accum, const ffq_layout_t* layout, double num_bits, void* workspace,
                 size_t workspace_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!(x_dtype == FFQ_F32 || x_dtype == FFQ_F16 || x_dtype == FFQ_BF16)) {
    set_error("grid_mse: data must be float32/float16/bfloat16"); return FFQ_ERR_UNSUPPORTED;
  }
  if (num_candidates <= 0) return FFQ_OK;
  Plan plan;
  int rc = make_plan(layout, &plan);
  if (rc != FFQ_OK) return rc;
  if (plan.numel == 0) { set_error("grid_mse: empty tensor"); return FFQ_ERR_INVALID; }
  GsArgs a{};
  int lanes = 32;
  if (!gs_geometry(plan, x_dtype, &a.seg_vecs, &a.segs_per_tile, &lanes) || (reinterpret_cast<uintptr_t>(x) & 15u)) {
    set_error("grid_mse: only contiguous-tile layouts with vector-aligned tiles are fused; use the per-candidate path");
    return FFQ_ERR_UNSUPPORTED;
  }
  a.x = x; a.x_dt = x_dtype; a.cscale = cand_scale; a.coffset = cand_offset; a.C = num_candidates; a.err = err_accum;
  a.tile_numel = plan.tile_numel; a.num_tiles = plan.num_tiles;
  a.total_units = (unsigned long long)plan.num_tiles * a.segs_per_tile;
  const double lo = -pow(2.0, num_bits - 1.0);
  a.lo = (float)lo; a.hi = (float)(-lo - 1.0); a.inv_tile = (float)(1.0 / (double)plan.tile_numel);
  a.m_out = round_mode_of(x_dtype);
  a.part = nullptr;
  if (a.segs_per_tile > 1) {
    const size_t need = (size_t)a.total_units * num_candidates * sizeof(float);
    if (workspace == nullptr || workspace_bytes < need) {
      set_error("grid_mse: workspace of %zu bytes required, %zu given", need, workspace_bytes); return FFQ_ERR_WORKSPACE;
    }
    a.part = static_cast<float*>(workspace);
  }
  if (a.total_units / (32 / lanes) / (GS_THREADS / 32) > 0x7ffffff0ull) { set_error("grid_mse: tensor too large"); return FFQ_ERR_UNSUPPORTED; }
  switch (x_dtype) {
    case FFQ_F32: gs_launch<float, 8>(a, lanes, st); break;
    case FFQ_BF16: gs_launch<__nv_bfloat16, 4>(a, lanes, st); break;
    default: gs_launch<__half, 4>(a, lanes, st); break;
  }
  FFQ_LAUNCH_CHECK();
  if (a.segs_per_tile > 1) {
    const unsigned long long n = a.num_tiles * (unsigned long long)a.C;
    grid_mse_finalize_kernel<<<(unsigned int)((n + GS_THREADS - 1) / GS_THREADS), GS_THREADS, 0, st>>>(a);
    FFQ_LAUNCH_CHECK();
  }
  return FFQ_OK;
}

}  // extern "C"

// ffq_calibrate.cu -- one RunningMinMax calibration step of one quantizer in ONE pass over HBM
// (SURVEY.md section 8a rows a7 + a6 + a1, plus the code row sums the W8A8 linear of a12 needs):
//
//     tile min/max -> running range update (+-inf flag) -> range -> (scale, offset) -> int8 codes
//     [-> int32 row sums of the codes]
//
// The reference does this as RunningMinMaxEstimator.estimate_step (range_setting/minmax.py:215-239:
// two full reads + a host sync), the quantization_range setter (nn/linear_quantizer.py:347-357 ->
// affine/range.py:54-122: a second host sync) and quantize_by_tile (_quantizer_impl.py:144-169: a third
// read) per quantizer per forward.  The unfused B200 path (ffq_minmax + ffq_params_for_range +
// ffq_quantize + ffq_rowsum_i8) already has no syncs but still reads the tensor three times and the
// codes once; here the data crosses HBM once (s + c bytes per element) in a single launch:
//
//   * calq_rows_kernel  : per-channel weights (tile = one contiguous row).  One CTA holds a row in
//     registers (<= 8 x 16 B per thread), reduces it, updates the running range, derives the
//     parameters and quantizes from the registers.  (A persistent variant streaming rows through a
//     3-deep shared-memory ring with 1-D TMA bulk copies measured 15-20 % slower on B200: the kernel is
//     bound by instruction issue -- exact IEEE division -- not by bytes in flight.)
//   * calq_tensor_kernel: per-tensor activations.  A co-resident (cooperative) grid; every CTA parks
//     its first 32 KB chunk in shared memory (148 SMs x 4 CTAs x 32 KB = 19 MB of the tensor never
//     gets re-read), partial extrema meet at a grid barrier, then every CTA derives the same
//     parameters and quantizes (remaining chunks are re-read through L2).
//
// The symmetric "one-sided" decision of parameters_for_range (`min.min() >= 0`, range.py:100) is global
// over all tiles.  A row whose running min is negative (or NaN) already proves the answer is "two-sided",
// so it is finished immediately; rows with a non-negative running min are left to calq_rows_fixup_kernel,
// which runs right behind, takes the global decision and finishes exactly those rows (normally none).
//
// Arithmetic is the same op-by-op sequence as the unfused kernels (ffq_common.cuh), so the results are
// bit-identical to them and to the reference (tests/test_calibrate_gpu.py).
#include <cstdlib>

#include "ffq_common.cuh"

namespace ffq {

struct CalqArgs {
  const void* x;
  int8_t* q;
  void* run_min; void* run_max; int run_dt;
  float* scale; float* offset;          // fp32 [tiles]; offset may be null (symmetric, one-sided not allowed)
  int32_t* rowsum;                      // optional: int32 per code row
  int32_t* flags;                       // optional: bit 0 = +-inf seen, bit 1 = grid barrier timed out
  int32_t* settled;                     // optional, rows: set once every running min is negative (sticky)
  unsigned long long rows;              // rows kernels: number of tiles.  tensor kernel: rows of the rowsum
  unsigned int row_len;                 // elements per row
  unsigned long long numel;
  float int_min_abs, int_max_abs, neg_int_min, steps, lo, hi;
  int symmetric, allow_one_sided, sat8;
  int rcp_div;                          // scalar divisions of parameters_for_range as aten's CUDA kernel does them
  // tensor kernel
  float* part;                          // [2 * gridDim.x]
  unsigned int* bar;                    // [0] root arrivals (wraps to 0), [1] generation; zero before first use
  unsigned int* bar_groups;             // first-level arrival counters, one per 32 bytes (wrap to 0)
  unsigned int nchunks;
  FastDiv rdiv;                         // division by row_len (tensor kernel row sums)
  // fake-quant output mode (calq_group_kernel / calq_rows_fq_kernel): y = dequantize(quantize(x)), may alias x
  void* y;
  int code_is_int;                      // integer code dtype between quantize and dequantize: -0 becomes +0
  unsigned int* ws_flags;               // [0] some tile was deferred, [1] some (running) tile min is negative/NaN
  unsigned int lanes;                   // group kernel: 16-byte vectors per tile
  int prof;                             // test hook (FFQ_CALQ_PROF=1): the tensor kernel leaves %globaltimer stamps in the workspace
};

constexpr int CQ_CHUNK_VECS = 2048;     // tensor kernel: 32 KB of input per CTA chunk
constexpr int CQ_T = 256;
constexpr int CQ_U = 4;               // 16-byte loads in flight per lane
constexpr unsigned int CQ_BAR_GROUPS = 16;   // first-level groups of the tensor kernel's grid barrier

// parameters_for_range for one tile (affine/range.py:89-122), fp32; `one_sided` is the global decision
__device__ __forceinline__ void calq_params(const CalqArgs& a, float mn, float mx, bool one_sided, float& sc, float& off) {
  if (a.symmetric && !one_sided) {
    const float neg = scalar_div(fabsf(mn), a.int_min_abs, a.rcp_div != 0);
    const float pos = scalar_div(fabsf(mx), a.int_max_abs, a.rcp_div != 0);
    sc = nan_max(neg, pos);
    off = 0.f;
    return;
  }
  if (a.symmetric) mn = 0.f;
  sc = scalar_div(__fsub_rn(mx, mn), a.steps, a.rcp_div != 0);
  const float eps = 1.1920928955078125e-07f;
  sc = (sc != sc) ? sc : fmaxf(sc, eps);
  off = __fadd_rn(__fdiv_rn(mn, sc), a.neg_int_min);
}

// quantize one 16-byte vector to int8 codes (quantize_by_tile with an fp32 chain); returns the codes
// packed little-endian and adds them to `sum`
template <typename XT, int EPT>
__device__ __forceinline__ void calq_vec(const Vec<XT, EPT>& xin, const SharedRcp& k, float o, float lo, float hi,
                                         bool sat8, uint32_t (&packed)[EPT / 4], int& sum) {
  float x[EPT], t[EPT];
  int c[EPT];
#pragma unroll
  for (int i = 0; i < EPT; ++i) x[i] = Elem<XT>::to_f(xin.v[i]);
  float amax = 0.f;
#pragma unroll
  for (int i = 0; i < EPT; ++i) {
    const float q0 = __fmul_rn(x[i], k.r);
    const float e = __fmaf_rn(-k.s, q0, x[i]);
    const float q = __fmaf_rn(k.r, e, q0);
    amax = nan_max(amax, fabsf(q));
    t[i] = __fsub_rn(q, o);
  }
  const bool ok = k.ok && (amax <= 0x1p60f);
  if (ok && sat8) {
#pragma unroll
    for (int i = 0; i < EPT; ++i) asm("cvt.rni.sat.s8.f32 %0, %1;" : "=r"(c[i]) : "f"(t[i]));
  } else {
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
      const float tt = ok ? t[i] : __fsub_rn(__fdiv_rn(x[i], k.s), o);
      c[i] = (int)(int8_t)__float2int_rz(nan_clamp(rintf(tt), lo, hi));
    }
  }
#pragma unroll
  for (int w = 0; w < EPT / 4; ++w) {
    packed[w] = (uint32_t)(c[4 * w] & 0xff) | ((uint32_t)(c[4 * w + 1] & 0xff) << 8) |
                ((uint32_t)(c[4 * w + 2] & 0xff) << 16) | ((uint32_t)(c[4 * w + 3] & 0xff) << 24);
    sum += c[4 * w] + c[4 * w + 1] + c[4 * w + 2] + c[4 * w + 3];
  }
}

// Fast variant for tiles whose guard was settled ONCE from the running range (which contains every element of
// the tile): the scale is inside the shared-reciprocal box and max(|min|, |max|) / s <= 2^59, so every quotient
// passes the magnitude guard and there is no NaN (a NaN element would have made the range NaN).  Per element:
// unpack, FMUL + 2 FFMA (exact quotient), FADD, F2I; packing saturates to int8 two codes at a time and the row
// sum is one dp4a per four codes.
template <typename XT, int EPT, bool SAT8>
__device__ __forceinline__ void calq_vec_fast(const Vec<XT, EPT>& xin, const SharedRcp& k, float o, int lo, int hi,
                                              uint32_t (&packed)[EPT / 4], int& sum) {
  int c[EPT];
#pragma unroll
  for (int i = 0; i < EPT; ++i) {
    const float x = Elem<XT>::to_f(xin.v[i]);
    const float q0 = __fmul_rn(x, k.r);
    const float e = __fmaf_rn(-k.s, q0, x);
    const float q = __fmaf_rn(k.r, e, q0);
    c[i] = __float2int_rn(__fsub_rn(q, o));          // saturates to int32; rint-then-clamp == clamp of this
    if constexpr (!SAT8) c[i] = min(max(c[i], lo), hi);
  }
#pragma unroll
  for (int w = 0; w < EPT / 4; ++w) {
    uint32_t hi16, word;
    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(hi16) : "r"(c[4 * w + 3]), "r"(c[4 * w + 2]), "r"(0));
    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(word) : "r"(c[4 * w + 1]), "r"(c[4 * w]), "r"(hi16));
    packed[w] = word;
    sum = __dp4a((int)word, 0x01010101, sum);
  }
}

// guard for calq_vec_fast from the tile's running range
__device__ __forceinline__ bool calq_fast_ok(const SharedRcp& k, float rmn, float rmx) {
  const float bound = __fmul_rn(nan_max(fabsf(rmn), fabsf(rmx)), fabsf(k.r));
  return k.ok && (bound <= 0x1p59f);                 // NaN compares false
}

template <typename XT, int EPT>
__device__ __forceinline__ void calq_vec_any(const Vec<XT, EPT>& xin, const SharedRcp& k, float o, const CalqArgs& a,
                                             bool fast, uint32_t (&packed)[EPT / 4], int& sum) {
  if (fast) {
    if (a.sat8) calq_vec_fast<XT, EPT, true>(xin, k, o, 0, 0, packed, sum);
    else calq_vec_fast<XT, EPT, false>(xin, k, o, (int)a.lo, (int)a.hi, packed, sum);
  } else {
    calq_vec<XT, EPT>(xin, k, o, a.lo, a.hi, a.sat8 != 0, packed, sum);
  }
}

template <int EPT>
__device__ __forceinline__ void calq_store(int8_t* q, const uint32_t (&packed)[EPT / 4]) {
  if constexpr (EPT == 8) *reinterpret_cast<uint2*>(q) = make_uint2(packed[0], packed[1]);
  else *reinterpret_cast<uint32_t*>(q) = packed[0];
}

// CTA-wide integer sum (result valid in thread 0); smem: 32 ints
__device__ __forceinline__ int block_isum(int v, int* smem) {
  v = __reduce_add_sync(0xffffffffu, v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) smem[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  if (w == 0) {
    v = lane < nw ? smem[lane] : 0;
    v = __reduce_add_sync(0xffffffffu, v);
  }
  return v;
}

// ------------------------------------------------------------------------------------------
// rows: one CTA per tile, the tile lives in registers
// ------------------------------------------------------------------------------------------
// FULL: the row is exactly blockDim.x * VPT vectors (the usual case: the host sizes the CTA to the row), so no
// per-vector bounds predicates are needed.
template <typename XT, int VPT, bool FULL>
__global__ void __launch_bounds__(512, 2) calq_rows_kernel(const CalqArgs a) {
  constexpr int EPT = 16 / sizeof(XT);
  __shared__ float s_mn[16], s_mx[16];
  __shared__ int s_i[32];
  const unsigned long long row = blockIdx.x;
  const unsigned int nvec = a.row_len / EPT;
  const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  const XT* __restrict__ x = static_cast<const XT*>(a.x) + row * a.row_len;
  pdl_wait();                    // programmatic dependent launch (ffq_common.cuh): no-ops under FFQ_PDL=0
  pdl_trigger();

  Vec<XT, EPT> xin[VPT];
#pragma unroll
  for (int u = 0; u < VPT; ++u) {
    const unsigned int j = threadIdx.x + u * blockDim.x;
    if (FULL || j < nvec) xin[u] = ld_stream<XT, EPT>(x + (size_t)j * EPT);
  }
  // the old running range is requested behind the data so that its latency hides under the stream; every thread
  // reads it before the barrier below, thread 0 overwrites it after the barrier
  const float old_mn = load_as_float(a.run_min, a.run_dt, row);
  const float old_mx = load_as_float(a.run_max, a.run_dt, row);
  float mn = INFINITY, mx = -INFINITY;
#pragma unroll
  for (int u = 0; u < VPT; ++u) {
    const unsigned int j = threadIdx.x + u * blockDim.x;
    if (FULL || j < nvec) {
      float vmn, vmx;
      vec_minmax<XT, EPT>(xin[u], vmn, vmx);
      mn = nan_min(mn, vmn);
      mx = nan_max(mx, vmx);
    }
  }
  mn = group_min<32>(mn);
  mx = group_max<32>(mx);
  if (lane == 0) { s_mn[warp] = mn; s_mx[warp] = mx; }
  __syncthreads();
  // every thread finishes the reduction and derives the parameters itself: no second barrier, no broadcast
  mn = s_mn[0]; mx = s_mx[0];
  for (unsigned int w = 1; w < nw; ++w) { mn = nan_min(mn, s_mn[w]); mx = nan_max(mx, s_mx[w]); }
  // running update: torch.min(self.min, data_min) / torch.max(...) in the range's dtype (minmax.py:235-237)
  const float rmn = nan_min(old_mn, mn);
  const float rmx = nan_max(old_mx, mx);
  // a negative (or NaN) running min settles the global one-sided question: two-sided
  const bool deferred = a.symmetric && a.allow_one_sided && (rmn >= 0.f);
  float sc = 1.f, off = 0.f;
  if (!deferred) calq_params(a, rmn, rmx, false, sc, off);
  if (threadIdx.x == 0) {
    store_from_float(a.run_min, a.run_dt, row, rmn);
    store_from_float(a.run_max, a.run_dt, row, rmx);
    if (a.flags && (isinf(mn) || isinf(mx))) atomicOr(a.flags, 1);
    if (!deferred) {
      a.scale[row] = sc;
      if (a.offset) a.offset[row] = off;
    }
  }
  if (deferred) return;                    // finished by calq_rows_fixup_kernel
  const float o = rintf(off);
  const SharedRcp k = make_shared_rcp(sc);
  const bool fast = calq_fast_ok(k, rmn, rmx);
  int8_t* __restrict__ q = a.q + row * a.row_len;
  int sum = 0;
  // the path is chosen once per row, not once per vector
  auto run = [&](auto kind) {
#pragma unroll
    for (int u = 0; u < VPT; ++u) {
      const unsigned int j = threadIdx.x + u * blockDim.x;
      if (FULL || j < nvec) {
        uint32_t packed[EPT / 4];
        if constexpr (decltype(kind)::value == 0) calq_vec_fast<XT, EPT, true>(xin[u], k, o, 0, 0, packed, sum);
        else if constexpr (decltype(kind)::value == 1) calq_vec_fast<XT, EPT, false>(xin[u], k, o, (int)a.lo, (int)a.hi, packed, sum);
        else calq_vec<XT, EPT>(xin[u], k, o, a.lo, a.hi, a.sat8 != 0, packed, sum);
        calq_store<EPT>(q + (size_t)j * EPT, packed);
      }
    }
  };
  if (fast && a.sat8) run(std::integral_constant<int, 0>{});
  else if (fast) run(std::integral_constant<int, 1>{});
  else run(std::integral_constant<int, 2>{});
  if (a.rowsum) {
    sum = block_isum(sum, s_i);
    if (threadIdx.x == 0) a.rowsum[row] = sum;
  }
}

// Finishes the rows calq_rows_kernel left open (running min >= 0): takes the global one-sided decision
// over ALL running mins and quantizes those rows, streaming them from HBM/L2.  Grid: a few CTAs per SM.
template <typename XT>
__global__ void __launch_bounds__(CQ_T) calq_rows_fixup_kernel(const CalqArgs a) {
  pdl_wait();                    // programmatic dependent launch: no-ops unless launched that way
  pdl_trigger();
  constexpr int EPT = 16 / sizeof(XT);
  __shared__ float s_f[64];
  __shared__ int s_i[32];
  __shared__ float s_g[2];
  // running mins only ever decrease: once every row has a negative one, no row is deferred again
  if (a.settled && *reinterpret_cast<const volatile int32_t*>(a.settled) != 0) return;
  float gmn = INFINITY, any = -INFINITY;
  for (unsigned long long r = threadIdx.x; r < a.rows; r += blockDim.x) {
    const float v = load_as_float(a.run_min, a.run_dt, r);
    gmn = nan_min(gmn, v);
    any = fmaxf(any, (v >= 0.f) ? 1.f : 0.f);
  }
  block_minmax(gmn, any, s_f);
  if (threadIdx.x == 0) { s_g[0] = gmn; s_g[1] = any; }
  __syncthreads();
  if (!(s_g[1] > 0.f)) {                   // nothing was deferred
    if (a.settled && blockIdx.x == 0 && threadIdx.x == 0 && !(s_g[0] != s_g[0])) *a.settled = 1;
    return;
  }
  const bool one_sided = s_g[0] >= 0.f;    // NaN >= 0 is false, as in Python
  const unsigned int nvec = a.row_len / EPT;
  for (unsigned long long row = blockIdx.x; row < a.rows; row += gridDim.x) {
    const float rmn = load_as_float(a.run_min, a.run_dt, row);
    if (!(rmn >= 0.f)) continue;
    const float rmx = load_as_float(a.run_max, a.run_dt, row);
    float sc, off;
    calq_params(a, rmn, rmx, one_sided, sc, off);
    if (threadIdx.x == 0) {
      a.scale[row] = sc;
      if (a.offset) a.offset[row] = off;
    }
    const float o = rintf(off);
    const SharedRcp k = make_shared_rcp(sc);
    const bool fast = calq_fast_ok(k, rmn, rmx);
    const XT* __restrict__ x = static_cast<const XT*>(a.x) + row * a.row_len;
    int8_t* __restrict__ q = a.q + row * a.row_len;
    int sum = 0;
    for (unsigned int j = threadIdx.x; j < nvec; j += blockDim.x) {
      const Vec<XT, EPT> xv = ld_stream<XT, EPT>(x + (size_t)j * EPT);
      uint32_t packed[EPT / 4];
      calq_vec_any<XT, EPT>(xv, k, o, a, fast, packed, sum);
      calq_store<EPT>(q + (size_t)j * EPT, packed);
    }
    if (a.rowsum) {
      __syncthreads();
      sum = block_isum(sum, s_i);
      if (threadIdx.x == 0) a.rowsum[row] = sum;
    }
  }
}

// ------------------------------------------------------------------------------------------
// fake-quant output (quantize + dequantize, fp32 chain): the arithmetic of ew_tile_kernel<OP_FAKEQUANT>
// ------------------------------------------------------------------------------------------
struct FqConst { float lo, hi; bool int_zero; };    // read once per thread: no per-element constant loads / branches

template <typename XT, int EPT, bool FAST, bool INT_ZERO>
__device__ __forceinline__ Vec<XT, EPT> calq_vec_fq_impl(const Vec<XT, EPT>& xin, const SharedRcp& k, float o, const FqConst c) {
  float y[EPT];
#pragma unroll
  for (int i = 0; i < EPT; ++i) {
    const float x = Elem<XT>::to_f(xin.v[i]);
    float t;
    if constexpr (FAST) {
      const float q0 = __fmul_rn(x, k.r);
      const float e = __fmaf_rn(-k.s, q0, x);
      t = __fsub_rn(__fmaf_rn(k.r, e, q0), o);
    } else {
      t = __fsub_rn(__fdiv_rn(x, k.s), o);
    }
    float q = nan_clamp(rintf(t), c.lo, c.hi);
    // integer code dtype: the cast to it turns -0 into +0 and NaN into 0 (what aten's float -> int conversion gives)
    if constexpr (INT_ZERO) q = (q == q) ? __fadd_rn(q, 0.0f) : 0.0f;
    y[i] = __fmul_rn(__fadd_rn(q, o), k.s);
  }
  Vec<XT, EPT> out;
  if constexpr (std::is_same<XT, __nv_bfloat16>::value) {
#pragma unroll
    for (int i = 0; i < EPT; i += 2) *reinterpret_cast<__nv_bfloat162*>(&out.v[i]) = __floats2bfloat162_rn(y[i], y[i + 1]);
  } else if constexpr (std::is_same<XT, __half>::value) {
#pragma unroll
    for (int i = 0; i < EPT; i += 2) *reinterpret_cast<__half2*>(&out.v[i]) = __floats2half2_rn(y[i], y[i + 1]);
  } else {
#pragma unroll
    for (int i = 0; i < EPT; ++i) out.v[i] = Elem<XT>::from_f(y[i]);
  }
  return out;
}

template <typename XT, int EPT>
__device__ __forceinline__ Vec<XT, EPT> calq_vec_fq(const Vec<XT, EPT>& xin, const SharedRcp& k, float o, const FqConst c,
                                                    bool fast) {
  if (c.int_zero) return fast ? calq_vec_fq_impl<XT, EPT, true, true>(xin, k, o, c) : calq_vec_fq_impl<XT, EPT, false, true>(xin, k, o, c);
  return fast ? calq_vec_fq_impl<XT, EPT, true, false>(xin, k, o, c) : calq_vec_fq_impl<XT, EPT, false, false>(xin, k, o, c);
}

constexpr float CQ_DEFERRED = -1.0f;    // scale sentinel: a legitimate scale is never negative

// Short tiles (per-group weights): LANES 16-byte vectors per tile, one vector per lane, the tile's extrema by
// sub-warp shuffles; four tiles' worth of loads in flight per thread.  FQ: fake-quant output (weight QDQ, may run
// in place); otherwise int8 codes.  The running range is optional (null: the tile's own min/max is the range).
// Deferred tiles (symmetric one-sided candidates) get the sentinel scale and are finished by calq_sentinel_fixup.
template <typename XT, typename RT, int LANES, bool FQ>
__global__ void __launch_bounds__(256) calq_group_kernel(const CalqArgs a) {
  pdl_wait();                    // programmatic dependent launch: no-ops unless launched that way
  pdl_trigger();
  constexpr int EPT = 16 / sizeof(XT);
  constexpr int U = 4;
  const unsigned long long nvec = a.numel / EPT;
  const unsigned long long vbase = (unsigned long long)blockIdx.x * (256 * U) + threadIdx.x;
  // everything below indexes relative to this thread's first vector with compile-time offsets (u * 256 vectors)
  const XT* __restrict__ x = static_cast<const XT*>(a.x) + vbase * EPT;
  const unsigned int remaining = (unsigned int)(nvec > vbase ? (nvec - vbase < 0x7fffffffull ? nvec - vbase : 0x7fffffffull) : 0ull);
  const unsigned long long tile0 = vbase / LANES;               // tile of vector u: tile0 + u * (256 / LANES)
  const unsigned int sel = threadIdx.x & (LANES - 1);          // position inside the tile's lane group
  const bool lead = sel == 0;
  const bool decide = a.symmetric && a.allow_one_sided;
  const FqConst fqc{a.lo, a.hi, a.code_is_int != 0};
  Vec<XT, EPT> xv[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if ((unsigned int)(u * 256) < remaining) xv[u] = ld_stream<XT, EPT>(x + u * 256 * EPT);
  }
  // ---- extrema and (running) range of the U tiles this lane group touches ----
  float rmn[U], rmx[U];
  bool neg_any = false, def_any = false;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const bool live = (unsigned int)(u * 256) < remaining;   // whole groups are live or dead together (nvec % LANES == 0)
    float mn = INFINITY, mx = -INFINITY;
    if (live) vec_minmax<XT, EPT>(xv[u], mn, mx);
    mn = group_min<LANES>(mn);
    mx = group_max<LANES>(mx);
    rmn[u] = mn; rmx[u] = mx;
    if (!live) continue;
    const unsigned long long tile = tile0 + u * (256 / LANES);
    if (a.run_min) {              // RT: the running range's dtype (the data dtype or fp32), typed accesses
      RT* __restrict__ run_mn = static_cast<RT*>(a.run_min);
      RT* __restrict__ run_mx = static_cast<RT*>(a.run_max);
      rmn[u] = nan_min(Elem<RT>::to_f(run_mn[tile]), mn);
      rmx[u] = nan_max(Elem<RT>::to_f(run_mx[tile]), mx);
      if (lead) {
        run_mn[tile] = Elem<RT>::from_f(rmn[u]);
        run_mx[tile] = Elem<RT>::from_f(rmx[u]);
      }
    }
    if (lead && a.flags && (isinf(mn) || isinf(mx))) atomicOr(a.flags, 1);
    neg_any = neg_any || !(rmn[u] >= 0.f);
    def_any = def_any || (decide && rmn[u] >= 0.f);
  }
  // ---- parameters: lane `sel` of a group derives them for tile u = sel (once per thread instead of once per
  // vector), the other lanes fetch them by shuffle.  Groups narrower than U fall back to one derivation per vector. ----
  constexpr bool SPREAD = LANES >= U;
  float sc_m = 1.f, off_m = 0.f, r_m = 1.f;
  int fast_m = 0;
  if constexpr (SPREAD) {
    float mn_m = rmn[0], mx_m = rmx[0];
#pragma unroll
    for (int u = 1; u < U; ++u)
      if (sel == (unsigned int)u) { mn_m = rmn[u]; mx_m = rmx[u]; }
    calq_params(a, mn_m, mx_m, false, sc_m, off_m);
    const SharedRcp km = make_shared_rcp(sc_m);
    r_m = km.r;
    fast_m = (calq_fast_ok(km, mn_m, mx_m) ? 1 : 0) | (km.ok ? 2 : 0);
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    float sc, off;
    SharedRcp k;
    bool fast;
    if constexpr (SPREAD) {
      const int src = (int)((threadIdx.x & 31u) - sel) + u;      // lane u of this group
      sc = __shfl_sync(0xffffffffu, sc_m, src);
      off = __shfl_sync(0xffffffffu, off_m, src);
      const int fl = __shfl_sync(0xffffffffu, fast_m, src);
      k.s = sc; k.r = __shfl_sync(0xffffffffu, r_m, src); k.ok = (fl & 2) != 0;
      fast = (fl & 1) != 0;
    }
    if ((unsigned int)(u * 256) >= remaining) continue;
    const unsigned long long tile = tile0 + u * (256 / LANES);
    if (decide && rmn[u] >= 0.f) {         // finished by calq_sentinel_fixup_kernel
      if (lead) a.scale[tile] = CQ_DEFERRED;
      continue;
    }
    if constexpr (!SPREAD) {
      calq_params(a, rmn[u], rmx[u], false, sc, off);
      k = make_shared_rcp(sc);
      fast = calq_fast_ok(k, rmn[u], rmx[u]);
    }
    if (lead) {
      a.scale[tile] = sc;
      if (a.offset) a.offset[tile] = off;
    }
    const float o = rintf(off);
    if constexpr (FQ) {
      const Vec<XT, EPT> yv = calq_vec_fq<XT, EPT>(xv[u], k, o, fqc, fast);
      st_vec<XT, EPT>(static_cast<XT*>(a.y) + vbase * EPT + u * 256 * EPT, yv);
    } else {
      uint32_t packed[EPT / 4];
      int sum = 0;
      calq_vec_any<XT, EPT>(xv[u], k, o, a, fast, packed, sum);
      calq_store<EPT>(a.q + vbase * EPT + u * 256 * EPT, packed);
    }
  }
  if (decide) {
    // one word each, written at most once per CTA and only while still unset
    const int neg = __syncthreads_or(neg_any ? 1 : 0);
    const int def = __syncthreads_or(def_any ? 1 : 0);
    if (threadIdx.x == 0) {
      if (neg && *reinterpret_cast<volatile unsigned int*>(a.ws_flags + 1) == 0) a.ws_flags[1] = 1;
      if (def && *reinterpret_cast<volatile unsigned int*>(a.ws_flags) == 0) a.ws_flags[0] = 1;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Short tiles of 16-bit data, ONE THREAD PER TILE (g = 64 / 128 elements = 128 / 256 bytes): the tile is read with
// 32-byte loads (LDG.256: one whole sector per lane per instruction; a warp's eight loads sweep 8 KB of consecutive
// memory), reduced, parameterised and quantized entirely in the thread's registers -- no shuffles for the extrema, no
// parameter broadcast, parameters derived once per 128 elements.  That takes the per-element instruction count of the
// sub-warp kernel above (28: 2 for the shuffle trees, 2 for the parameter broadcast, the rest arithmetic) to ~14, which
// is what an instruction-issue-bound kernel needs (at 6.4 TB/s a bf16 in/out stream leaves ~22 issue slots per element).
// ZOFF: symmetric quantizer -- every tile finished here has offset 0 (one-sided candidates are deferred), so the
// subtraction of the rounded offset disappears (x - 0 == x bit for bit; the +0 of the dequantize step stays: it turns
// -0 codes into +0 as the reference's addition does).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void ld256_stream(const void* p, uint4& a, uint4& b) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}
__device__ __forceinline__ void st256(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
               "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
template <typename XT> struct Packed2;
template <> struct Packed2<__nv_bfloat16> {
  using T = __nv_bfloat162;
  static __device__ __forceinline__ T mn(T a, T b) { return __hmin2_nan(a, b); }
  static __device__ __forceinline__ T mx(T a, T b) { return __hmax2_nan(a, b); }
};
template <> struct Packed2<__half> {
  using T = __half2;
  static __device__ __forceinline__ T mn(T a, T b) { return __hmin2_nan(a, b); }
  static __device__ __forceinline__ T mx(T a, T b) { return __hmax2_nan(a, b); }
};

// The whole tile through ONE variant of the per-element code.  The variant (exact-division fast path or not, integer
// codes or not, saturating 8-bit pack or not) is chosen once per thread OUTSIDE the unrolled loops: with the choice
// inside, the variants of all 16 vectors interleave in one 10 000-instruction body of which a warp executes 1 800
// scattered over 160 KB, and a third of its stall samples are instruction-cache misses (`no_instruction`, ncu).
template <typename XT, int NV, bool FAST, bool IZ>
__device__ __forceinline__ void tt_fq_all(const uint4 (&w)[NV], XT* __restrict__ y, const SharedRcp& k, float o, const FqConst& fqc) {
#pragma unroll
  for (int i = 0; i < NV; i += 2) {
    const Vec<XT, 8> y0 = calq_vec_fq_impl<XT, 8, FAST, IZ>(*reinterpret_cast<const Vec<XT, 8>*>(&w[i]), k, o, fqc);
    const Vec<XT, 8> y1 = calq_vec_fq_impl<XT, 8, FAST, IZ>(*reinterpret_cast<const Vec<XT, 8>*>(&w[i + 1]), k, o, fqc);
    st256(y + i * 8, *reinterpret_cast<const uint4*>(&y0), *reinterpret_cast<const uint4*>(&y1));
  }
}
template <typename XT, int NV, int MODE>     // 0: fast + saturating 8-bit pack, 1: fast, 2: guarded division
__device__ __forceinline__ void tt_codes_all(const uint4 (&w)[NV], int8_t* __restrict__ q, const SharedRcp& k, float o, const CalqArgs& a) {
  int sum = 0;
#pragma unroll
  for (int i = 0; i < NV; i += 4) {       // four 16-byte vectors -> 32 bytes of codes
    uint32_t p[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const Vec<XT, 8>& xv = *reinterpret_cast<const Vec<XT, 8>*>(&w[i + j]);
      if constexpr (MODE == 0) calq_vec_fast<XT, 8, true>(xv, k, o, 0, 0, p[j], sum);
      else if constexpr (MODE == 1) calq_vec_fast<XT, 8, false>(xv, k, o, (int)a.lo, (int)a.hi, p[j], sum);
      else calq_vec<XT, 8>(xv, k, o, a.lo, a.hi, a.sat8 != 0, p[j], sum);
    }
    st256(q + i * 8, make_uint4(p[0][0], p[0][1], p[1][0], p[1][1]), make_uint4(p[2][0], p[2][1], p[3][0], p[3][1]));
  }
}

template <typename XT, typename RT, int NV, bool FQ, bool ZOFF>
__device__ __forceinline__ void calq_tile_thread_body(const CalqArgs& a, unsigned long long tile) {
  static_assert(sizeof(XT) == 2 && NV % 2 == 0, "16-bit data, whole 32-byte accesses");
  constexpr int EPT = 8;
  using P2 = Packed2<XT>;
  const unsigned long long ntiles = a.numel / (unsigned long long)(NV * EPT);
  const bool decide = a.symmetric && a.allow_one_sided;
  bool neg = false, def = false;
  if (tile < ntiles) {
    const XT* __restrict__ x = static_cast<const XT*>(a.x) + tile * (unsigned long long)(NV * EPT);
    uint4 w[NV];
#pragma unroll
    for (int i = 0; i < NV; i += 2) ld256_stream(x + i * EPT, w[i], w[i + 1]);
    // extrema in the packed domain (NaN-propagating HMNMX2), one unpack at the end
    typename P2::T lo2, hi2;
    {
      // four independent chains (min / max are exact: any association gives the same bits)
      const typename P2::T* h = reinterpret_cast<const typename P2::T*>(&w[0]);
      typename P2::T l4[4] = {h[0], h[1], h[2], h[3]}, h4[4] = {h[0], h[1], h[2], h[3]};
#pragma unroll
      for (int i = 4; i < NV * 4; ++i) { l4[i & 3] = P2::mn(l4[i & 3], h[i]); h4[i & 3] = P2::mx(h4[i & 3], h[i]); }
      lo2 = P2::mn(P2::mn(l4[0], l4[1]), P2::mn(l4[2], l4[3]));
      hi2 = P2::mx(P2::mx(h4[0], h4[1]), P2::mx(h4[2], h4[3]));
    }
    const float mn = nan_min(__low2float(lo2), __high2float(lo2));
    const float mx = nan_max(__low2float(hi2), __high2float(hi2));
    float rmn = mn, rmx = mx;
    if (a.run_min) {
      RT* __restrict__ run_mn = static_cast<RT*>(a.run_min);
      RT* __restrict__ run_mx = static_cast<RT*>(a.run_max);
      rmn = nan_min(Elem<RT>::to_f(run_mn[tile]), mn);
      rmx = nan_max(Elem<RT>::to_f(run_mx[tile]), mx);
      run_mn[tile] = Elem<RT>::from_f(rmn);
      run_mx[tile] = Elem<RT>::from_f(rmx);
    }
    if (a.flags && (isinf(mn) || isinf(mx))) atomicOr(a.flags, 1);
    neg = !(rmn >= 0.f);
    if (decide && rmn >= 0.f) {            // finished by calq_sentinel_fixup_kernel
      a.scale[tile] = CQ_DEFERRED;
      def = true;
    } else {
      float sc, off;
      calq_params(a, rmn, rmx, false, sc, off);
      a.scale[tile] = sc;
      if (a.offset) a.offset[tile] = off;
      const SharedRcp k = make_shared_rcp(sc);
      const bool fast = calq_fast_ok(k, rmn, rmx);
      const float o = ZOFF ? 0.0f : rintf(off);
      if constexpr (FQ) {
        const FqConst fqc{a.lo, a.hi, a.code_is_int != 0};
        XT* __restrict__ y = static_cast<XT*>(a.y) + tile * (unsigned long long)(NV * EPT);
        if (fast) {
          if (fqc.int_zero) tt_fq_all<XT, NV, true, true>(w, y, k, o, fqc);
          else tt_fq_all<XT, NV, true, false>(w, y, k, o, fqc);
        } else {
          if (fqc.int_zero) tt_fq_all<XT, NV, false, true>(w, y, k, o, fqc);
          else tt_fq_all<XT, NV, false, false>(w, y, k, o, fqc);
        }
      } else {
        int8_t* __restrict__ q = a.q + tile * (unsigned long long)(NV * EPT);
        if (fast && a.sat8) tt_codes_all<XT, NV, 0>(w, q, k, o, a);
        else if (fast) tt_codes_all<XT, NV, 1>(w, q, k, o, a);
        else tt_codes_all<XT, NV, 2>(w, q, k, o, a);
      }
    }
  }
  if (decide) {
    // one word each, written at most once per CTA and only while still unset
    const int n = __syncthreads_or(neg ? 1 : 0);
    const int d = __syncthreads_or(def ? 1 : 0);
    if (threadIdx.x == 0) {
      if (n && *reinterpret_cast<volatile unsigned int*>(a.ws_flags + 1) == 0) a.ws_flags[1] = 1;
      if (d && *reinterpret_cast<volatile unsigned int*>(a.ws_flags) == 0) a.ws_flags[0] = 1;
    }
  }
}

template <typename XT, typename RT, int NV, bool FQ, bool ZOFF>
__global__ void __launch_bounds__(256) calq_tile_thread_kernel(const CalqArgs a) {
  pdl_wait();                    // programmatic dependent launch: no-ops unless launched that way
  pdl_trigger();
  calq_tile_thread_body<XT, RT, NV, FQ, ZOFF>(a, (unsigned long long)blockIdx.x * 256ull + threadIdx.x);
}

// ---- many tensors in ONE launch (whole-model weight fake-quant: quantization/fuse.py over every linear) -----------
// items[t] = {x, y, scale, offset, numel}; block_start[t] = first CTA of tensor t (prefix sums of ceil(tiles_t / 256),
// n + 1 entries); flags_base + 2t = the tensor's {deferred, negative} words.  Every CTA belongs to one tensor.
struct FqItem { const void* x; void* y; float* scale; float* offset; long long numel; };

__device__ __forceinline__ int find_item(const unsigned int* __restrict__ block_start, int n, unsigned int block) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(block_start + mid) <= block) lo = mid; else hi = mid - 1;
  }
  return lo;
}

template <typename XT, int NV, bool ZOFF>
__global__ void __launch_bounds__(256) calq_tile_thread_batched_kernel(CalqArgs a, const FqItem* __restrict__ items,
                                                                       const unsigned int* __restrict__ block_start, int n,
                                                                       unsigned int* flags_base) {
  pdl_wait();                    // programmatic dependent launch: no-ops unless launched that way
  pdl_trigger();
  const int t = find_item(block_start, n, blockIdx.x);
  const FqItem it = items[t];
  a.x = it.x; a.y = it.y; a.scale = it.scale; a.offset = it.offset; a.numel = (unsigned long long)it.numel;
  a.ws_flags = flags_base + 2 * t;
  calq_tile_thread_body<XT, XT, NV, true, ZOFF>(a, (unsigned long long)(blockIdx.x - __ldg(block_start + t)) * 256ull + threadIdx.x);
}

// Long rows with fake-quant output: calq_rows_kernel's structure (row in registers), the sentinel protocol for
// deferred rows, optional running range.
template <typename XT, int VPT>
__global__ void __launch_bounds__(512) calq_rows_fq_kernel(const CalqArgs a) {
  pdl_wait();                    // programmatic dependent launch: no-ops unless launched that way
  pdl_trigger();
  constexpr int EPT = 16 / sizeof(XT);
  __shared__ float s_mn[16], s_mx[16];
  const unsigned long long row = blockIdx.x;
  const unsigned int nvec = a.row_len / EPT;
  const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  const XT* __restrict__ x = static_cast<const XT*>(a.x) + row * a.row_len;
  const FqConst fqc{a.lo, a.hi, a.code_is_int != 0};
  Vec<XT, EPT> xin[VPT];
#pragma unroll
  for (int u = 0; u < VPT; ++u) {
    const unsigned int j = threadIdx.x + u * blockDim.x;
    if (j < nvec) xin[u] = ld_stream<XT, EPT>(x + (size_t)j * EPT);
  }
  float old_mn = INFINITY, old_mx = -INFINITY;
  if (a.run_min) {
    old_mn = load_as_float(a.run_min, a.run_dt, row);
    old_mx = load_as_float(a.run_max, a.run_dt, row);
  }
  float mn = INFINITY, mx = -INFINITY;
#pragma unroll
  for (int u = 0; u < VPT; ++u) {
    const unsigned int j = threadIdx.x + u * blockDim.x;
    if (j < nvec) {
      float vmn, vmx;
      vec_minmax<XT, EPT>(xin[u], vmn, vmx);
      mn = nan_min(mn, vmn);
      mx = nan_max(mx, vmx);
    }
  }
  mn = group_min<32>(mn);
  mx = group_max<32>(mx);
  if (lane == 0) { s_mn[warp] = mn; s_mx[warp] = mx; }
  __syncthreads();
  mn = s_mn[0]; mx = s_mx[0];
  for (unsigned int w = 1; w < nw; ++w) { mn = nan_min(mn, s_mn[w]); mx = nan_max(mx, s_mx[w]); }
  const float rmn = a.run_min ? nan_min(old_mn, mn) : mn;
  const float rmx = a.run_min ? nan_max(old_mx, mx) : mx;
  const bool live_decision = a.symmetric && a.allow_one_sided;
  const bool deferred = live_decision && (rmn >= 0.f);
  float sc = CQ_DEFERRED, off = 0.f;
  if (!deferred) calq_params(a, rmn, rmx, false, sc, off);
  if (threadIdx.x == 0) {
    if (a.run_min) {
      store_from_float(a.run_min, a.run_dt, row, rmn);
      store_from_float(a.run_max, a.run_dt, row, rmx);
    }
    if (a.flags && (isinf(mn) || isinf(mx))) atomicOr(a.flags, 1);
    a.scale[row] = sc;
    if (!deferred && a.offset) a.offset[row] = off;
    if (live_decision) {
      if (!(rmn >= 0.f) && *reinterpret_cast<volatile unsigned int*>(a.ws_flags + 1) == 0) a.ws_flags[1] = 1;
      if (deferred && *reinterpret_cast<volatile unsigned int*>(a.ws_flags) == 0) a.ws_flags[0] = 1;
    }
  }
  if (deferred) return;
  const float o = rintf(off);
  const SharedRcp k = make_shared_rcp(sc);
  const bool fast = calq_fast_ok(k, rmn, rmx);
  XT* __restrict__ y = static_cast<XT*>(a.y) + row * a.row_len;
#pragma unroll
  for (int u = 0; u < VPT; ++u) {
    const unsigned int j = threadIdx.x + u * blockDim.x;
    if (j < nvec) st_vec<XT, EPT>(y + (size_t)j * EPT, calq_vec_fq<XT, EPT>(xin[u], k, o, fqc, fast));
  }
}

// Finishes the tiles marked with the sentinel scale: a warp per tile.  Runs right behind calq_group_kernel /
// calq_rows_fq_kernel and returns after one load unless something was deferred (weights whose tiles are all
// non-negative are the only case).  The global decision: one-sided unless some tile min was negative or NaN.
template <typename XT, bool FQ>
__device__ __forceinline__ void calq_sentinel_fixup_body(const CalqArgs& a, unsigned long long warp0, unsigned long long nwarps) {
  constexpr int EPT = 16 / sizeof(XT);
  if (*reinterpret_cast<volatile unsigned int*>(a.ws_flags) == 0) return;
  const bool one_sided = *reinterpret_cast<volatile unsigned int*>(a.ws_flags + 1) == 0;
  const unsigned int lane = threadIdx.x & 31;
  const unsigned int tvec = a.row_len / EPT;
  const FqConst fqc{a.lo, a.hi, a.code_is_int != 0};
  for (unsigned long long tile = warp0; tile < a.rows; tile += nwarps) {
    if (a.scale[tile] != CQ_DEFERRED) continue;
    const XT* __restrict__ x = static_cast<const XT*>(a.x) + tile * a.row_len;
    float rmn, rmx;
    if (a.run_min) {
      rmn = load_as_float(a.run_min, a.run_dt, tile);
      rmx = load_as_float(a.run_max, a.run_dt, tile);
    } else {
      float mn = INFINITY, mx = -INFINITY;
      for (unsigned int j = lane; j < tvec; j += 32) {
        float vmn, vmx;
        vec_minmax<XT, EPT>(ld_stream<XT, EPT>(x + (size_t)j * EPT), vmn, vmx);
        mn = nan_min(mn, vmn);
        mx = nan_max(mx, vmx);
      }
      rmn = group_min<32>(mn);
      rmx = group_max<32>(mx);
    }
    float sc, off;
    calq_params(a, rmn, rmx, one_sided, sc, off);
    const float o = rintf(off);
    const SharedRcp k = make_shared_rcp(sc);
    const bool fast = calq_fast_ok(k, rmn, rmx);
    for (unsigned int j = lane; j < tvec; j += 32) {
      const Vec<XT, EPT> xv = ld_stream<XT, EPT>(x + (size_t)j * EPT);
      if constexpr (FQ) {
        st_vec<XT, EPT>(static_cast<XT*>(a.y) + tile * a.row_len + (size_t)j * EPT, calq_vec_fq<XT, EPT>(xv, k, o, fqc, fast));
      } else {
        uint32_t packed[EPT / 4];
        int sum = 0;
        calq_vec_any<XT, EPT>(xv, k, o, a, fast, packed, sum);
        calq_store<EPT>(a.q + tile * a.row_len + (size_t)j * EPT, packed);
      }
    }
    __syncwarp();
    if (lane == 0) {          // parameters last: the sentinel is this tile's "still to do" mark
      a.scale[tile] = sc;
      if (a.offset) a.offset[tile] = off;
    }
  }
}

template <typename XT, bool FQ>
__global__ void __launch_bounds__(256) calq_sentinel_fixup_kernel(const CalqArgs a) {
  pdl_wait();                    // programmatic dependent launch: no-ops unless launched that way
  pdl_trigger();
  calq_sentinel_fixup_body<XT, FQ>(a, ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5,
                                   ((unsigned long long)gridDim.x * blockDim.x) >> 5);
}

constexpr int FQ_FIXUP_BLOCKS = 8;      // CTAs per tensor of the batched fix-up (they return at once unless a tile was deferred)
template <typename XT>
__global__ void __launch_bounds__(256) calq_sentinel_fixup_batched_kernel(CalqArgs a, const FqItem* __restrict__ items,
                                                                          unsigned int* flags_base) {
  pdl_wait();                    // programmatic dependent launch: no-ops unless launched that way
  pdl_trigger();
  const int t = blockIdx.x / FQ_FIXUP_BLOCKS;
  const FqItem it = items[t];
  a.x = it.x; a.y = it.y; a.scale = it.scale; a.offset = it.offset; a.numel = (unsigned long long)it.numel;
  a.rows = (unsigned long long)it.numel / a.row_len;
  a.ws_flags = flags_base + 2 * t;
  calq_sentinel_fixup_body<XT, true>(a, (unsigned long long)(blockIdx.x % FQ_FIXUP_BLOCKS) * 8 + (threadIdx.x >> 5),
                                     (unsigned long long)FQ_FIXUP_BLOCKS * 8);
}

// ------------------------------------------------------------------------------------------
// per-tensor: co-resident grid, first chunk of every CTA parked in shared memory
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Phase 3 of calq_tensor_kernel for one (MODE, ROWS) variant.  MODE 0: fast exact division + saturating 8-bit pack,
// 1: fast, 2: guarded division.  ROWS 0: no row sums, 1: a warp's span lies in one row (one atomic per chunk),
// 2: rows straddle warp spans (one atomic per (warp, row), sums carried while the whole warp stays in one row).
template <typename XT, int T, int CV, int MODE, int ROWS>
__device__ __forceinline__ void calq_tensor_pass3(const CalqArgs& a, const uint4* s_chunk, const XT* __restrict__ x,
                                                  unsigned long long nvec, const SharedRcp& k, float o) {
  constexpr int EPT = 16 / sizeof(XT);
  constexpr int VPW = CV / (T / 32);
  const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned int G = gridDim.x;
  for (unsigned int c = blockIdx.x; c < a.nchunks; c += G) {
    const unsigned long long w0 = (unsigned long long)c * CV + warp * VPW;   // warp's first vector
    const bool keep = c == blockIdx.x;
    unsigned long long row0 = 0;
    unsigned int rem0 = 0;
    if (ROWS != 0) { row0 = (w0 * EPT) / a.row_len; rem0 = (unsigned int)(w0 * EPT - row0 * a.row_len); }
    int sum = 0;
    unsigned int cur = ~0u;                  // row (relative to row0) the running sum belongs to
#pragma unroll
    for (int ub = 0; ub < VPW / 32; ub += CQ_U) {
      Vec<XT, EPT> xv[CQ_U];
#pragma unroll
      for (int u = 0; u < CQ_U; ++u) {
        const unsigned long long v = w0 + (ub + u) * 32 + lane;
        if (v < nvec) {
          if (keep) *reinterpret_cast<uint4*>(&xv[u]) = s_chunk[warp * VPW + (ub + u) * 32 + lane];
          else xv[u] = ld_stream<XT, EPT>(x + v * EPT);
        }
      }
#pragma unroll
      for (int u = 0; u < CQ_U; ++u) {
        const unsigned long long v = w0 + (ub + u) * 32 + lane;
        if (w0 + (ub + u) * 32 >= nvec) break;           // the whole warp is past the end
        const bool live = v < nvec;
        int vs = 0;
        if (live) {
          uint32_t packed[EPT / 4];
          if constexpr (MODE == 0) calq_vec_fast<XT, EPT, true>(xv[u], k, o, 0, 0, packed, vs);
          else if constexpr (MODE == 1) calq_vec_fast<XT, EPT, false>(xv[u], k, o, (int)a.lo, (int)a.hi, packed, vs);
          else calq_vec<XT, EPT>(xv[u], k, o, a.lo, a.hi, a.sat8 != 0, packed, vs);
          calq_store<EPT>(a.q + v * EPT, packed);
        }
        if constexpr (ROWS == 1) {
          sum += vs;                       // the warp's whole span lies in one row: one atomic per chunk
        } else if constexpr (ROWS == 2) {
          const unsigned int r = fast_div(rem0 + ((ub + u) * 32 + lane) * EPT, a.rdiv);
          const unsigned int r_first = __shfl_sync(0xffffffffu, r, 0);
          const unsigned int r_last = __shfl_sync(0xffffffffu, r, 31);
          if (r_first != cur || r_last != cur) {
            if (cur != ~0u) {
              const int tot = __reduce_add_sync(0xffffffffu, sum);
              if (lane == 0) atomicAdd(&a.rowsum[row0 + cur], tot);
            }
            sum = 0;
            cur = (r_first == r_last) ? r_first : ~0u;
          }
          if (cur != ~0u) sum += vs;
          else if (live) atomicAdd(&a.rowsum[row0 + r], vs);
        }
      }
    }
    if constexpr (ROWS == 1) {
      if (w0 < nvec) {
        const int tot = __reduce_add_sync(0xffffffffu, sum);
        if (lane == 0) atomicAdd(&a.rowsum[row0], tot);
      }
    } else if constexpr (ROWS == 2) {
      if (cur != ~0u) {
        const int tot = __reduce_add_sync(0xffffffffu, sum);
        if (lane == 0) atomicAdd(&a.rowsum[row0 + cur], tot);
      }
    }
  }
}

__device__ __forceinline__ unsigned int ld_relaxed_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// T threads per CTA, CV 16-byte vectors per chunk.  (256, 2048): four 32 KB CTAs per SM; (1024, 8192): ONE 128 KB CTA per SM --
// a quarter of the arrivals at the grid barrier and of the partials every CTA re-reduces, which is what a 16 MB
// activation tensor (one chunk per CTA either way) spends its time on.
template <typename XT, int T, int CV>
__global__ void __launch_bounds__(T, (T >= 1024 ? 1 : 4)) calq_tensor_kernel(const CalqArgs a) {
  constexpr int EPT = 16 / sizeof(XT);
  constexpr int VPW = CV / (T / 32);       // vectors of a chunk owned by one warp (contiguous)
  extern __shared__ uint4 s_chunk[];                      // CV vectors
  __shared__ float s_f[64];
  __shared__ float s_par[4];
  const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned int G = gridDim.x;
  const unsigned long long nvec = a.numel / EPT;
  const XT* __restrict__ x = static_cast<const XT*>(a.x);

  // test hook: thread 0 of every CTA stamps the nanosecond timer at start / extrema done / barrier passed / parameters
  // known / end, 8 u64 per CTA behind the partial extrema (tools/prof_calq_phases.py); grids of <= 448 CTAs only
  unsigned long long* stamps = (a.prof && G <= 448 && threadIdx.x == 0)
                                   ? reinterpret_cast<unsigned long long*>(a.part + 1024) + (size_t)blockIdx.x * 8 : nullptr;
  auto stamp = [&](int i) {
    if (stamps) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); stamps[i] = t; }
  };
  stamp(0);
  float old_mn = 0.f, old_mx = 0.f;
  if (threadIdx.x == 0) {
    old_mn = load_as_float(a.run_min, a.run_dt, 0);
    old_mx = load_as_float(a.run_max, a.run_dt, 0);
  }
  if (a.rowsum)
    for (unsigned long long r = (unsigned long long)blockIdx.x * T + threadIdx.x; r < a.rows; r += (unsigned long long)G * T)
      a.rowsum[r] = 0;

  // ---- phase 1: extrema of this CTA's chunks; the first chunk stays in shared memory ----
  float mn = INFINITY, mx = -INFINITY;
  for (unsigned int c = blockIdx.x; c < a.nchunks; c += G) {
    const unsigned long long v0 = (unsigned long long)c * CV + warp * VPW + lane;
    const bool keep = c == blockIdx.x;
#pragma unroll
    for (int ub = 0; ub < VPW / 32; ub += CQ_U) {
      Vec<XT, EPT> xv[CQ_U];
#pragma unroll
      for (int u = 0; u < CQ_U; ++u) {
        const unsigned long long v = v0 + (ub + u) * 32;
        if (v < nvec) xv[u] = ld_stream<XT, EPT>(x + v * EPT);
      }
#pragma unroll
      for (int u = 0; u < CQ_U; ++u) {
        const unsigned long long v = v0 + (ub + u) * 32;
        if (v < nvec) {
          if (keep) s_chunk[warp * VPW + (ub + u) * 32 + lane] = *reinterpret_cast<const uint4*>(&xv[u]);
          float vmn, vmx;
          vec_minmax<XT, EPT>(xv[u], vmn, vmx);
          mn = nan_min(mn, vmn);
          mx = nan_max(mx, vmx);
        }
      }
    }
  }
  block_minmax(mn, mx, s_f);
  stamp(1);
  // ---- grid barrier (all CTAs are co-resident: cooperative launch, grid <= occupancy).  Sense-reversing:
  // bar[0] counts arrivals and wraps to 0 with the last one, which then bumps the generation bar[1]; nothing to
  // reset between launches, any grid size ----
  if (threadIdx.x == 0) {
    a.part[blockIdx.x] = mn;
    a.part[G + blockIdx.x] = mx;
    // the generation is polled with RELAXED loads (an acquire load per poll drags a gpu-scope fence through every
    // iteration: the barrier took 4.9 us of a 12.8 us launch that way); one fence after the loop orders the reads of
    // the partials behind the last arrival's release
    const unsigned int gen0 = ld_relaxed_u32(&a.bar[1]);
    // two-level arrival: ~150 atomics on ONE address serialise in its L2 slice (~27 clocks each: the last arrival
    // waited 3-4 us); CTAs first meet in CQ_BAR_GROUPS groups on separate words, the last of a group arrives at the root
    const unsigned int ngroups = G < CQ_BAR_GROUPS ? G : CQ_BAR_GROUPS;
    const unsigned int grp = blockIdx.x % ngroups;
    const unsigned int members = G / ngroups + (grp < G % ngroups ? 1u : 0u);
    unsigned int old;
    asm volatile("atom.inc.acq_rel.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(a.bar_groups + grp * 8), "r"(members - 1) : "memory");
    bool last = false;
    if (old == members - 1) {
      asm volatile("atom.inc.acq_rel.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(a.bar), "r"(ngroups - 1) : "memory");
      last = old == ngroups - 1;
    }
    if (last) {
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(a.bar + 1) : "memory");
    } else {
      unsigned int spins = 0;
      while (ld_relaxed_u32(&a.bar[1]) == gen0) {
        if (++spins > (1u << 21)) {           // ~1 s of L2 round trips: never hang the GPU; the host raises on bit 1
          if (a.flags) atomicOr(a.flags, 2);
          break;
        }
      }
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
  }
  stamp(2);
  __syncthreads();
  // ---- phase 2: warp 0 of every CTA reduces the partials to the same (min, max) and derives the parameters ----
  if (warp == 0) {
    mn = INFINITY; mx = -INFINITY;
    for (unsigned int i = lane; i < G; i += 32) {
      mn = nan_min(mn, __ldcg(&a.part[i]));
      mx = nan_max(mx, __ldcg(&a.part[G + i]));
    }
    mn = group_min<32>(mn);
    mx = group_max<32>(mx);
    if (lane == 0) {
      const float rmn = nan_min(old_mn, mn);
      const float rmx = nan_max(old_mx, mx);
      float sc, off;
      calq_params(a, rmn, rmx, a.symmetric && a.allow_one_sided && (rmn >= 0.f), sc, off);
      if (blockIdx.x == 0) {
        store_from_float(a.run_min, a.run_dt, 0, rmn);
        store_from_float(a.run_max, a.run_dt, 0, rmx);
        if (a.flags && (isinf(mn) || isinf(mx))) atomicOr(a.flags, 1);
        a.scale[0] = sc;
        if (a.offset) a.offset[0] = off;
      }
      s_par[0] = sc; s_par[1] = off; s_par[2] = rmn; s_par[3] = rmx;
    }
  }
  __syncthreads();
  stamp(3);
  const float s = s_par[0], o = rintf(s_par[1]);
  const SharedRcp k = make_shared_rcp(s);
  const bool fast = calq_fast_ok(k, s_par[2], s_par[3]);
  // ---- phase 3: quantize.  The per-element variant (fast / saturating pack) and the row-sum bookkeeping are chosen
  // ONCE, outside the unrolled loops: interleaved inside them they made a 5 000-instruction body whose executed part
  // was scattered (ncu: 4.4 warps per issue slot waiting for instructions) ----
  // rows that are a whole number of warp spans (VPW vectors): a warp never straddles two rows
  const bool simple_rows = a.rowsum != nullptr && (a.row_len % (VPW * EPT)) == 0;
  const int mode = fast ? (a.sat8 ? 0 : 1) : 2;
  const int rows = a.rowsum == nullptr ? 0 : (simple_rows ? 1 : 2);
#define FFQ_PASS3(M, R) calq_tensor_pass3<XT, T, CV, M, R>(a, s_chunk, x, nvec, k, o)
  switch (mode * 3 + rows) {
    case 0: FFQ_PASS3(0, 0); break;
    case 1: FFQ_PASS3(0, 1); break;
    case 2: FFQ_PASS3(0, 2); break;
    case 3: FFQ_PASS3(1, 0); break;
    case 4: FFQ_PASS3(1, 1); break;
    case 5: FFQ_PASS3(1, 2); break;
    case 6: FFQ_PASS3(2, 0); break;
    case 7: FFQ_PASS3(2, 1); break;
    default: FFQ_PASS3(2, 2); break;
  }
#undef FFQ_PASS3
  __syncthreads();
  stamp(4);
}

constexpr int CQ_FAT_T = 1024, CQ_FAT_CHUNK_VECS = 8192;     // one 128 KB CTA per SM

template <typename XT, int T, int CV>
static int tensor_grid_cap() {
  static std::atomic<uint64_t> done{0};
  static int occ_dev[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  once_per_device(done, [&]() -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(calq_tensor_kernel<XT, T, CV>, cudaFuncAttributeMaxDynamicSharedMemorySize, CV * 16);
    if (e != cudaSuccess) return e;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, calq_tensor_kernel<XT, T, CV>, T, CV * 16) != cudaSuccess) occ = 1;
    occ_dev[dev] = occ < 1 ? 1 : occ;
    return cudaSuccess;
  });
  return (occ_dev[dev] < 1 ? 1 : occ_dev[dev]) * sm_count();
}

template <typename XT, int T, int CV>
static cudaError_t launch_tensor(const CalqArgs& a, unsigned int grid, cudaStream_t st) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(T);
  cfg.dynamicSmemBytes = CV * 16; cfg.stream = st;
  // cooperative: the driver guarantees that the whole grid is co-resident (the barrier cannot starve behind
  // another kernel).  FFQ_CALQ_COOP=0 launches it as an ordinary kernel (same grid, sized to the occupancy).
  static const bool coop = []() { const char* e = getenv("FFQ_CALQ_COOP"); return !(e && e[0] == '0'); }();
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = coop ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, calq_tensor_kernel<XT, T, CV>, a);
}

// chunks, grid and launch of the per-tensor kernel in one of its two shapes
template <typename XT, int T, int CV>
static int run_tensor(CalqArgs& a, unsigned long long nvec, cudaStream_t st) {
  const unsigned long long nchunks = (nvec + CV - 1) / CV;
  if (nchunks >= (1ull << 31)) { set_error("calibrate_quantize: tensor too large"); return FFQ_ERR_UNSUPPORTED; }
  a.nchunks = (unsigned int)nchunks;
  int cap = tensor_grid_cap<XT, T, CV>();
  if (cap > 4096) cap = 4096;              // partial extrema of at most 4096 CTAs fit the workspace
  const unsigned int grid = (unsigned int)(nchunks < (unsigned long long)cap ? nchunks : (unsigned long long)cap);
  const cudaError_t e = launch_tensor<XT, T, CV>(a, grid, st);
  count_launch();
  if (e != cudaSuccess) { cudaGetLastError(); set_error("calibrate_quantize: per-tensor kernel launch failed: %s", cudaGetErrorString(e)); return FFQ_ERR_CUDA; }
  return FFQ_OK;
}

template <typename XT>
static int run_tensor_any(CalqArgs& a, unsigned long long nvec, cudaStream_t st) {
  // the fat shape from ~2 MB per SM-wave on; FFQ_CALQ_TENSOR_FAT=0 keeps the four-small-CTAs shape (A/B switch)
  static const bool no_fat = []() { const char* e = getenv("FFQ_CALQ_TENSOR_FAT"); return e && e[0] == '0'; }();
  if (!no_fat && nvec >= (unsigned long long)CQ_FAT_CHUNK_VECS * 32)
    return run_tensor<XT, CQ_FAT_T, CQ_FAT_CHUNK_VECS>(a, nvec, st);
  return run_tensor<XT, CQ_T, CQ_CHUNK_VECS>(a, nvec, st);
}

// ------------------------------------------------------------------------------------------
// per-tensor, optimistic: ONE pass in the common case that the batch does not move the running range
// ------------------------------------------------------------------------------------------
// The two-pass structure of the kernel above (extrema -> grid barrier -> quantize) is the reference's; a 16 MB
// activation tensor spends 3-5 of its 16 us in the barrier and runs its two passes one after the other.  But a running
// range only ever widens, and after the first few batches most batches do not widen it (the chance that batch t sets
// a new extreme is ~2/t).  So:
//   kernel A (calq_tensor_opt_kernel<.., false>): every CTA derives (scale, offset) from the OLD running range and, in
//     one pass, takes the batch's extrema AND writes codes + row sums for those parameters.  No barrier: the CTAs leave
//     their partial extrema in the workspace and take a ticket; the LAST CTA merges them into the running range,
//     writes the parameters of the merged range and records whether the merged range differs from the old one
//     in any bit (`redo`);
//   kernel B (<.., true>): returns after one load unless `redo` is set; then it writes codes + row sums again with the
//     parameters of the merged range (the tensor is still in L2).
// Result: codes = quantize(x, parameters(merged range)) in every case, exactly what the two-pass kernel produces (when
// the merged range equals the old one bit for bit, so do the parameters and therefore the codes).  The first batch of a
// block (no range yet) and every batch that widens the range cost two passes, as before, without the barrier.
// Work is assigned by rows (a warp per row of the row-sum layout, or per 512 vectors when no row sums are wanted), so
// a row's sum is one plain store: no atomics, no zero-initialisation.
constexpr int CQO_T = 256;
constexpr unsigned int CQO_SPAN = 512;                 // vectors per warp task when there are no rows

// MODE 0: fast exact division + saturating 8-bit pack, 1: fast, 2: guarded division, 3: no codes (extrema only)
template <typename XT, int MODE, bool MINMAX, int U>
__device__ __forceinline__ void calq_opt_pass(const CalqArgs& a, const XT* __restrict__ x, unsigned long long nvec,
                                              unsigned int span, unsigned long long ntasks, const SharedRcp& k, float o,
                                              float& mn, float& mx) {
  constexpr int EPT = 16 / sizeof(XT);
  const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned long long nwarps = (unsigned long long)gridDim.x * (CQO_T / 32);
  for (unsigned long long task = (unsigned long long)blockIdx.x * (CQO_T / 32) + warp; task < ntasks; task += nwarps) {
    const unsigned long long v0 = task * span;
    const unsigned int n = (unsigned int)((nvec - v0) < span ? (nvec - v0) : span);
    int sum = 0;
    for (unsigned int b = 0; b < n; b += 32 * U) {
      Vec<XT, EPT> xv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const unsigned int i = b + u * 32 + lane;
        if (i < n) xv[u] = ld_stream<XT, EPT>(x + (v0 + i) * EPT);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const unsigned int i = b + u * 32 + lane;
        if (i < n) {
          if constexpr (MINMAX) {
            float vmn, vmx;
            vec_minmax<XT, EPT>(xv[u], vmn, vmx);
            mn = nan_min(mn, vmn);
            mx = nan_max(mx, vmx);
          }
          if constexpr (MODE != 3) {
            uint32_t packed[EPT / 4];
            if constexpr (MODE == 0) calq_vec_fast<XT, EPT, true>(xv[u], k, o, 0, 0, packed, sum);
            else if constexpr (MODE == 1) calq_vec_fast<XT, EPT, false>(xv[u], k, o, (int)a.lo, (int)a.hi, packed, sum);
            else calq_vec<XT, EPT>(xv[u], k, o, a.lo, a.hi, a.sat8 != 0, packed, sum);
            calq_store<EPT>(a.q + (v0 + i) * EPT, packed);
          }
        }
      }
    }
    if constexpr (MODE != 3) {
      if (a.rowsum) {
        const int tot = __reduce_add_sync(0xffffffffu, sum);
        if (lane == 0) a.rowsum[task] = tot;
      }
    }
  }
}

template <typename XT, bool REDO, int U>
__global__ void __launch_bounds__(CQO_T, (U > 4 ? 2 : 4)) calq_tensor_opt_kernel(const CalqArgs a) {
  constexpr int EPT = 16 / sizeof(XT);
  __shared__ float s_f[64];
  __shared__ float s_par[4];
  __shared__ int s_flag;
  const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned int G = gridDim.x;
  const unsigned long long nvec = a.numel / EPT;
  const unsigned int span = a.rowsum ? a.row_len / EPT : CQO_SPAN;
  const unsigned long long ntasks = (nvec + span - 1) / span;
  const XT* __restrict__ x = static_cast<const XT*>(a.x);
  unsigned int* redo = a.bar + 2;                        // workspace word 2: written by A's last CTA, read by B
  pdl_wait();
  pdl_trigger();

  if (threadIdx.x == 0) {
    float rmn, rmx;
    int go;
    if constexpr (REDO) {
      go = (int)__ldcg(redo);
      rmn = load_as_float(a.run_min, a.run_dt, 0);       // the merged range and its parameters, written by kernel A
      rmx = load_as_float(a.run_max, a.run_dt, 0);
      s_par[0] = go ? __ldcg(a.scale) : 1.f;
      s_par[1] = (go && a.offset) ? __ldcg(a.offset) : 0.f;
    } else {
      rmn = load_as_float(a.run_min, a.run_dt, 0);       // the OLD range: never trust the stored parameters to match it
      rmx = load_as_float(a.run_max, a.run_dt, 0);
      go = (rmn <= rmx) && !isinf(rmn) && !isinf(rmx);   // a usable range (false for the +-inf start and for NaN)
      float sc = 1.f, off = 0.f;
      if (go) calq_params(a, rmn, rmx, a.symmetric && a.allow_one_sided && (rmn >= 0.f), sc, off);
      s_par[0] = sc; s_par[1] = off;
    }
    s_par[2] = rmn; s_par[3] = rmx;
    s_flag = go;
  }
  __syncthreads();
  const bool go = s_flag != 0;
  if (REDO && !go) return;
  const float s = s_par[0], o = rintf(s_par[1]);
  const float old_mn = s_par[2], old_mx = s_par[3];
  const SharedRcp k = make_shared_rcp(s);
  // the fast variants are exact only for elements inside the range the guard was settled from; a speculative pass may
  // meet elements outside it -- then the range widens and kernel B rewrites everything
  const bool fast = calq_fast_ok(k, old_mn, old_mx);
  const int mode = !go ? 3 : (fast ? (a.sat8 ? 0 : 1) : 2);
  float mn = INFINITY, mx = -INFINITY;
  switch (mode) {
    case 0: calq_opt_pass<XT, 0, !REDO, U>(a, x, nvec, span, ntasks, k, o, mn, mx); break;
    case 1: calq_opt_pass<XT, 1, !REDO, U>(a, x, nvec, span, ntasks, k, o, mn, mx); break;
    case 2: calq_opt_pass<XT, 2, !REDO, U>(a, x, nvec, span, ntasks, k, o, mn, mx); break;
    default:
      if constexpr (!REDO) calq_opt_pass<XT, 3, true, U>(a, x, nvec, span, ntasks, k, o, mn, mx);
      break;
  }
  if constexpr (!REDO) {
    block_minmax(mn, mx, s_f);
    if (threadIdx.x == 0) {
      a.part[blockIdx.x] = mn;
      a.part[G + blockIdx.x] = mx;
      __threadfence();
      const unsigned int ticket = atomicInc(a.bar, G - 1);     // wraps to 0 with the last arrival: nothing to reset
      s_flag = ticket == G - 1;
    }
    __syncthreads();
    if (s_flag && warp == 0) {
      __threadfence();
      mn = INFINITY; mx = -INFINITY;
      for (unsigned int i = lane; i < G; i += 32) {
        mn = nan_min(mn, __ldcg(&a.part[i]));
        mx = nan_max(mx, __ldcg(&a.part[G + i]));
      }
      mn = group_min<32>(mn);
      mx = group_max<32>(mx);
      if (lane == 0) {
        const float rmn = nan_min(old_mn, mn);
        const float rmx = nan_max(old_mx, mx);
        float sc, off;
        calq_params(a, rmn, rmx, a.symmetric && a.allow_one_sided && (rmn >= 0.f), sc, off);
        store_from_float(a.run_min, a.run_dt, 0, rmn);
        store_from_float(a.run_max, a.run_dt, 0, rmx);
        if (a.flags && (isinf(mn) || isinf(mx))) atomicOr(a.flags, 1);
        a.scale[0] = sc;
        if (a.offset) a.offset[0] = off;
        const bool same = go && __float_as_uint(rmn) == __float_as_uint(old_mn) && __float_as_uint(rmx) == __float_as_uint(old_mx);
        *redo = same ? 0u : 1u;
      }
    }
  }
}

template <typename XT>
static int run_tensor_opt(CalqArgs& a, unsigned long long nvec, cudaStream_t st) {
  constexpr int EPT = 16 / sizeof(XT);
  const unsigned int span = a.rowsum ? a.row_len / EPT : CQO_SPAN;
  const unsigned long long ntasks = (nvec + span - 1) / span;
  const unsigned long long want = (ntasks + CQO_T / 32 - 1) / (CQO_T / 32);
  unsigned long long cap = (unsigned long long)sm_count() * 8;
  if (cap > 4096) cap = 4096;              // partial extrema of at most 4096 CTAs fit the workspace
  const unsigned int grid = (unsigned int)(want < cap ? want : cap);
  // 16-byte loads in flight per lane: 8 while the whole tensor is one wave at two CTAs per SM (a 4096-element bf16 row
  // is two L2 round trips per lane: 2048x4096 12.8 -> 12.1 us), else 4 (four CTAs per SM: 8192x4096 26.3 us against
  // 34.3 with 8); FFQ_CALQ_OPT_U=4 forces the latter (A/B)
  static const bool u4 = []() { const char* e = getenv("FFQ_CALQ_OPT_U"); return e && e[0] == '4'; }();
  if (u4 || want > (unsigned long long)sm_count() * 2) {
    launch_pdl(calq_tensor_opt_kernel<XT, false, 4>, dim3(grid), dim3(CQO_T), 0, st, a);
    FFQ_LAUNCH_CHECK();
    launch_pdl(calq_tensor_opt_kernel<XT, true, 4>, dim3(grid), dim3(CQO_T), 0, st, a);
  } else {
    launch_pdl(calq_tensor_opt_kernel<XT, false, 8>, dim3(grid), dim3(CQO_T), 0, st, a);
    FFQ_LAUNCH_CHECK();
    launch_pdl(calq_tensor_opt_kernel<XT, true, 8>, dim3(grid), dim3(CQO_T), 0, st, a);
  }
  FFQ_LAUNCH_CHECK();
  return FFQ_OK;
}

// the optimistic pair serves tensors of at least 256 KB whose row-sum rows (if any) are 32 .. 8192 vectors long;
// FFQ_CALQ_OPTIMISTIC=0 keeps the two-pass kernel everywhere (A/B switch)
static bool tensor_opt_applies(const CalqArgs& a, unsigned long long nvec, int ept) {
  static const bool off = []() { const char* e = getenv("FFQ_CALQ_OPTIMISTIC"); return e && e[0] == '0'; }();
  if (off || a.prof || nvec < 16384) return false;
  if (a.rowsum) {
    const unsigned int rv = a.row_len / (unsigned int)ept;
    if (a.row_len % (unsigned int)ept != 0 || rv < 32 || rv > 8192) return false;
  }
  return true;
}

template <typename XT>
static cudaError_t launch_rows(const CalqArgs& a, unsigned int nvec, cudaStream_t st) {
  // one CTA per row, sized to the row (no idle slots beyond the last warp); 8 vectors (128 B) per thread in
  // flight for rows of >= 512 vectors: half as many warps share the per-row work (reduction, parameters)
  static const int force = []() { const char* e = getenv("FFQ_CALQ_VPT"); return e ? atoi(e) : 0; }();
  int vpt = (nvec >= 512 && a.rows >= 2048) ? 8 : (nvec >= 256 ? 4 : (nvec >= 128 ? 2 : 1));
  if (nvec > 2048) vpt = 8;
  if ((force == 4 || force == 8) && nvec >= 256 && (unsigned long long)force * 512 >= nvec) vpt = force;
  const unsigned int threads = ((nvec + vpt - 1) / vpt + 31) / 32 * 32;
  const unsigned int grid = (unsigned int)a.rows;
  const bool full = threads * (unsigned int)vpt == nvec;
#define FFQ_ROWS_LAUNCH(V)                                                                    \
  do {                                                                                        \
    if (full) launch_pdl(calq_rows_kernel<XT, V, true>, dim3(grid), dim3(threads), 0, st, a);  \
    else launch_pdl(calq_rows_kernel<XT, V, false>, dim3(grid), dim3(threads), 0, st, a);      \
  } while (0)
  switch (vpt) {
    case 1: FFQ_ROWS_LAUNCH(1); break;
    case 2: FFQ_ROWS_LAUNCH(2); break;
    case 4: FFQ_ROWS_LAUNCH(4); break;
    default: FFQ_ROWS_LAUNCH(8); break;
  }
#undef FFQ_ROWS_LAUNCH
  return cudaSuccess;
}

template <typename XT, typename RT, bool FQ>
static void launch_group_rt(const CalqArgs& a, cudaStream_t st) {
  constexpr int EPT = 16 / sizeof(XT);
  const unsigned long long nvec = a.numel / EPT;
  if constexpr (sizeof(XT) == 2) {
    // 16-bit tiles of 128 / 256 bytes whose data (and output) sit on 32-byte boundaries: one thread per tile
    static const bool off = getenv("FFQ_CALQ_GROUP_SUBWARP") != nullptr;      // A/B switch: keep the sub-warp kernel
    const void* out = FQ ? a.y : static_cast<const void*>(a.q);
    const bool al = (reinterpret_cast<uintptr_t>(a.x) & 31u) == 0 && (reinterpret_cast<uintptr_t>(out) & 31u) == 0;
    if (!off && al && (a.lanes == 8 || a.lanes == 16)) {
      const unsigned long long ntiles = nvec / a.lanes;
      const unsigned int grid = (unsigned int)((ntiles + 255) / 256);
      const bool z = a.symmetric != 0;
#define FFQ_TT(NV)                                                                                              \
      do {                                                                                                        \
        if (z) launch_pdl(calq_tile_thread_kernel<XT, RT, NV, FQ, true>, dim3(grid), dim3(256), 0, st, a);                            \
        else launch_pdl(calq_tile_thread_kernel<XT, RT, NV, FQ, false>, dim3(grid), dim3(256), 0, st, a);                             \
      } while (0)
      if (a.lanes == 8) FFQ_TT(8); else FFQ_TT(16);
#undef FFQ_TT
      return;
    }
  }
  const unsigned int grid = (unsigned int)((nvec + 256 * 4 - 1) / (256 * 4));
  switch (a.lanes) {
    case 1: launch_pdl(calq_group_kernel<XT, RT, 1, FQ>, dim3(grid), dim3(256), 0, st, a); break;
    case 2: launch_pdl(calq_group_kernel<XT, RT, 2, FQ>, dim3(grid), dim3(256), 0, st, a); break;
    case 4: launch_pdl(calq_group_kernel<XT, RT, 4, FQ>, dim3(grid), dim3(256), 0, st, a); break;
    case 8: launch_pdl(calq_group_kernel<XT, RT, 8, FQ>, dim3(grid), dim3(256), 0, st, a); break;
    case 16: launch_pdl(calq_group_kernel<XT, RT, 16, FQ>, dim3(grid), dim3(256), 0, st, a); break;
    default: launch_pdl(calq_group_kernel<XT, RT, 32, FQ>, dim3(grid), dim3(256), 0, st, a); break;
  }
}

template <typename XT, bool FQ>
static void launch_group(const CalqArgs& a, cudaStream_t st) {
  // the running range (when there is one) is stored in the data dtype or in fp32 (torch.min promotes)
  if (a.run_min != nullptr && a.run_dt == FFQ_F32 && !std::is_same<XT, float>::value) launch_group_rt<XT, float, FQ>(a, st);
  else launch_group_rt<XT, XT, FQ>(a, st);
}

template <typename XT>
static void launch_rows_fq(const CalqArgs& a, unsigned int nvec, cudaStream_t st) {
  int vpt = (nvec >= 512 && a.rows >= 2048) ? 8 : (nvec >= 256 ? 4 : (nvec >= 128 ? 2 : 1));
  if (nvec > 2048) vpt = 8;
  const unsigned int threads = ((nvec + vpt - 1) / vpt + 31) / 32 * 32;
  const unsigned int grid = (unsigned int)a.rows;
  switch (vpt) {
    case 1: launch_pdl(calq_rows_fq_kernel<XT, 1>, dim3(grid), dim3(threads), 0, st, a); break;
    case 2: launch_pdl(calq_rows_fq_kernel<XT, 2>, dim3(grid), dim3(threads), 0, st, a); break;
    case 4: launch_pdl(calq_rows_fq_kernel<XT, 4>, dim3(grid), dim3(threads), 0, st, a); break;
    default: launch_pdl(calq_rows_fq_kernel<XT, 8>, dim3(grid), dim3(threads), 0, st, a); break;
  }
}

template <typename XT, bool FQ>
static void launch_sentinel_fixup(const CalqArgs& a, cudaStream_t st) {
  const unsigned long long warps_wanted = a.rows;
  unsigned long long blocks = (warps_wanted + 7) / 8;
  const unsigned long long cap = (unsigned long long)4 * sm_count();
  if (blocks > cap) blocks = cap;
  launch_pdl(calq_sentinel_fixup_kernel<XT, FQ>, dim3((unsigned int)blocks), dim3(256), 0, st, a);
}

}  // namespace ffq

using namespace ffq;

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// 0: not supported (use the unfused sequence), 1: rows kernel, 2: per-tensor kernel, 3: sub-warp group kernel
static int calq_mode(const Plan& plan, int x_dtype) {
  if (!(x_dtype == FFQ_F32 || x_dtype == FFQ_F16 || x_dtype == FFQ_BF16)) return 0;
  if (!plan.row || plan.numel == 0) return 0;
  const int ept = 16 / dt_size(x_dtype);
  if (plan.tile_numel % ept != 0) return 0;
  const long long tvec = plan.tile_numel / ept;
  if (plan.num_tiles == 1) return (plan.numel < (1ll << 40) && tvec >= 64) ? 2 : 0;
  if (tvec >= 64 && tvec <= 4096 && plan.num_tiles < (1ll << 31) && plan.tile_numel < (1ll << 31)) return 1;
  if (tvec <= 32 && (tvec & (tvec - 1)) == 0 && plan.num_tiles < (1ll << 40)) return 3;   // sub-warp groups
  return 0;
}

extern "C" {

int ffq_calibrate_quantize_mode(const ffq_layout_t* layout, int x_dtype) {
  Plan plan;
  if (make_plan(layout, &plan) != FFQ_OK) return 0;
  return calq_mode(plan, x_dtype);
}

size_t ffq_calibrate_quantize_workspace_bytes(void) {
  // [bar0, bar1, pad x2 | partial mins and maxes of up to 4096 CTAs | first-level barrier counters, 32 B apart]
  return 16 + (size_t)2 * 4096 * sizeof(float) + (size_t)CQ_BAR_GROUPS * 32;
}

int ffq_calibrate_quantize(const void* x, int x_dtype, int8_t* q, void* run_min, void* run_max, int run_dtype,
                           float* scale, float* offset, int32_t* rowsum, int64_t rowsum_row_len, int32_t* flags,
                           int32_t* settled, int run_fixup, const ffq_layout_t* layout, double num_bits, int symmetric,
                           int allow_one_sided, void* workspace, size_t workspace_bytes, void* stream) {
  const int rcp_div = (allow_one_sided & FFQ_FLAG_SCALAR_DIV_RECIPROCAL) ? 1 : 0;
  allow_one_sided &= FFQ_FLAG_ALLOW_ONE_SIDED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Plan plan;
  int rc = make_plan(layout, &plan);
  if (rc != FFQ_OK) return rc;
  const int mode = calq_mode(plan, x_dtype);
  if (mode == 0 || !aligned16(x) || (reinterpret_cast<uintptr_t>(q) & 7u) != 0) {
    set_error("calibrate_quantize: layout/dtype/alignment not handled by the fused kernels");
    return FFQ_ERR_UNSUPPORTED;
  }
  if (num_bits > 8 || num_bits < 1) { set_error("calibrate_quantize: int8 codes hold at most 8 bits"); return FFQ_ERR_BITWIDTH; }
  if (!(run_dtype == FFQ_F32 || run_dtype == FFQ_F16 || run_dtype == FFQ_BF16) || promote(run_dtype, x_dtype) != run_dtype) {
    set_error("calibrate_quantize: running-range dtype %s cannot hold %s data", dt_name(run_dtype), dt_name(x_dtype));
    return FFQ_ERR_UNSUPPORTED;
  }
  if (run_min == nullptr || run_max == nullptr || scale == nullptr || (offset == nullptr && !(symmetric && !allow_one_sided))) {
    set_error("calibrate_quantize: run_min, run_max, scale (and offset unless symmetric two-sided only) are required");
    return FFQ_ERR_INVALID;
  }
  CalqArgs a{};
  a.x = x; a.q = q; a.run_min = run_min; a.run_max = run_max; a.run_dt = run_dtype;
  a.scale = scale; a.offset = offset; a.rowsum = rowsum; a.flags = flags; a.settled = settled;
  a.numel = (unsigned long long)plan.numel;
  const double lo = -pow(2.0, num_bits - 1.0);
  a.int_min_abs = (float)fabs(lo);
  a.int_max_abs = (float)fabs(-lo - 1.0);
  a.neg_int_min = (float)(-lo);
  a.steps = (float)(pow(2.0, num_bits) - 1.0);
  a.lo = (float)lo; a.hi = (float)(-lo - 1.0);
  a.symmetric = symmetric; a.allow_one_sided = allow_one_sided; a.rcp_div = rcp_div;
  a.sat8 = (a.lo == -128.f && a.hi == 127.f) ? 1 : 0;
  const int ept = 16 / dt_size(x_dtype);
  if (mode == 3) {
    if (rowsum) { set_error("calibrate_quantize: row sums are not produced for per-group tiles"); return FFQ_ERR_UNSUPPORTED; }
    if (workspace == nullptr || workspace_bytes < 16) { set_error("calibrate_quantize: workspace required"); return FFQ_ERR_WORKSPACE; }
    a.rows = (unsigned long long)plan.num_tiles;
    a.row_len = (unsigned int)plan.tile_numel;
    a.lanes = (unsigned int)(plan.tile_numel / ept);
    a.ws_flags = reinterpret_cast<unsigned int*>(static_cast<char*>(workspace) + 8);
    const bool decide = symmetric && allow_one_sided;
    if (decide) FFQ_CUDA_CHECK(cudaMemsetAsync(a.ws_flags, 0, 8, st));
    switch (x_dtype) {
      case FFQ_F32: launch_group<float, false>(a, st); break;
      case FFQ_BF16: launch_group<__nv_bfloat16, false>(a, st); break;
      default: launch_group<__half, false>(a, st); break;
    }
    FFQ_LAUNCH_CHECK();
    if (decide) {
      switch (x_dtype) {
        case FFQ_F32: launch_sentinel_fixup<float, false>(a, st); break;
        case FFQ_BF16: launch_sentinel_fixup<__nv_bfloat16, false>(a, st); break;
        default: launch_sentinel_fixup<__half, false>(a, st); break;
      }
      FFQ_LAUNCH_CHECK();
    }
    return FFQ_OK;
  }
  if (mode == 1) {
    a.rows = (unsigned long long)plan.num_tiles;
    a.row_len = (unsigned int)plan.tile_numel;
    if (rowsum && rowsum_row_len != plan.tile_numel) {
      set_error("calibrate_quantize: per-channel row sums need rowsum_row_len == tile length");
      return FFQ_ERR_INVALID;
    }
    const unsigned int nvec = a.row_len / ept;
    cudaError_t le;
    switch (x_dtype) {
      case FFQ_F32: le = launch_rows<float>(a, nvec, st); break;
      case FFQ_BF16: le = launch_rows<__nv_bfloat16>(a, nvec, st); break;
      default: le = launch_rows<__half>(a, nvec, st); break;
    }
    if (le != cudaSuccess) { set_error("calibrate_quantize: %s", cudaGetErrorString(le)); return FFQ_ERR_CUDA; }
    FFQ_LAUNCH_CHECK();
    if (symmetric && allow_one_sided && run_fixup) {
      unsigned int grid = (unsigned int)(a.rows < (unsigned long long)sm_count() ? a.rows : sm_count());
      switch (x_dtype) {
        case FFQ_F32: launch_pdl(calq_rows_fixup_kernel<float>, dim3(grid), dim3(CQ_T), 0, st, a); break;
        case FFQ_BF16: launch_pdl(calq_rows_fixup_kernel<__nv_bfloat16>, dim3(grid), dim3(CQ_T), 0, st, a); break;
        default: launch_pdl(calq_rows_fixup_kernel<__half>, dim3(grid), dim3(CQ_T), 0, st, a); break;
      }
      FFQ_LAUNCH_CHECK();
    }
    return FFQ_OK;
  }
  // per-tensor
  if (workspace == nullptr || workspace_bytes < ffq_calibrate_quantize_workspace_bytes() ||
      (reinterpret_cast<uintptr_t>(workspace) & 15u) != 0) {
    set_error("calibrate_quantize: a 16-byte aligned workspace of %zu bytes is required", ffq_calibrate_quantize_workspace_bytes());
    return FFQ_ERR_WORKSPACE;
  }
  if (rowsum) {
    if (rowsum_row_len <= 0 || plan.numel % rowsum_row_len != 0 || rowsum_row_len % ept != 0 || rowsum_row_len >= (1ll << 31)) {
      set_error("calibrate_quantize: rowsum_row_len must divide the tensor and be a multiple of the vector width");
      return FFQ_ERR_INVALID;
    }
    a.rows = (unsigned long long)(plan.numel / rowsum_row_len);
    a.row_len = (unsigned int)rowsum_row_len;
  } else {
    a.rows = 0; a.row_len = (unsigned int)ept;
  }
  a.rdiv = make_fast_div(a.row_len);
  { static const bool prof = []() { const char* e = getenv("FFQ_CALQ_PROF"); return e && e[0] == '1'; }(); a.prof = prof ? 1 : 0; }
  a.bar = static_cast<unsigned int*>(workspace);
  a.part = reinterpret_cast<float*>(static_cast<char*>(workspace) + 16);
  a.bar_groups = reinterpret_cast<unsigned int*>(static_cast<char*>(workspace) + 16 + (size_t)2 * 4096 * sizeof(float));
  const unsigned long long nvec = a.numel / ept;
  if (tensor_opt_applies(a, nvec, ept)) {
    switch (x_dtype) {
      case FFQ_F32: return run_tensor_opt<float>(a, nvec, st);
      case FFQ_BF16: return run_tensor_opt<__nv_bfloat16>(a, nvec, st);
      default: return run_tensor_opt<__half>(a, nvec, st);
    }
  }
  switch (x_dtype) {
    case FFQ_F32: return run_tensor_any<float>(a, nvec, st);
    case FFQ_BF16: return run_tensor_any<__nv_bfloat16>(a, nvec, st);
    default: return run_tensor_any<__half>(a, nvec, st);
  }
}

int ffq_calibrate_fakequant(const void* x, int x_dtype, void* y, void* run_min, void* run_max, int run_dtype,
                            float* scale, float* offset, int32_t* flags, const ffq_layout_t* layout, double num_bits,
                            int symmetric, int allow_one_sided, int code_dtype, void* workspace, size_t workspace_bytes,
                            void* stream) {
  const int rcp_div = (allow_one_sided & FFQ_FLAG_SCALAR_DIV_RECIPROCAL) ? 1 : 0;
  allow_one_sided &= FFQ_FLAG_ALLOW_ONE_SIDED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Plan plan;
  int rc = make_plan(layout, &plan);
  if (rc != FFQ_OK) return rc;
  const int mode = calq_mode(plan, x_dtype);
  if (!(mode == 1 || mode == 3) || !aligned16(x) || !aligned16(y)) {
    set_error("calibrate_fakequant: layout/dtype/alignment not handled by the fused kernels");
    return FFQ_ERR_UNSUPPORTED;
  }
  if ((run_min == nullptr) != (run_max == nullptr)) { set_error("calibrate_fakequant: run_min and run_max go together"); return FFQ_ERR_INVALID; }
  if (run_min && (!(run_dtype == FFQ_F32 || run_dtype == FFQ_F16 || run_dtype == FFQ_BF16) || promote(run_dtype, x_dtype) != run_dtype)) {
    set_error("calibrate_fakequant: running-range dtype %s cannot hold %s data", dt_name(run_dtype), dt_name(x_dtype));
    return FFQ_ERR_UNSUPPORTED;
  }
  if (scale == nullptr || (offset == nullptr && !(symmetric && !allow_one_sided))) {
    set_error("calibrate_fakequant: scale (and offset unless symmetric two-sided only) are required");
    return FFQ_ERR_INVALID;
  }
  const bool float_codes = code_dtype == FFQ_F32 || code_dtype == FFQ_F16 || code_dtype == FFQ_BF16;
  const int mant = code_dtype == FFQ_F32 ? 23 : (code_dtype == FFQ_F16 ? 10 : 7);
  if (float_codes ? (mant + 2 < num_bits) : (!is_int_dt(code_dtype) || code_dtype == FFQ_U8 || 8 * dt_size(code_dtype) < num_bits)) {
    set_error("calibrate_fakequant: code dtype %s cannot hold %g-bit signed codes exactly", dt_name(code_dtype), num_bits);
    return FFQ_ERR_BITWIDTH;
  }
  if (float_codes && mant + 1 < num_bits) {   // representable by the reference's rule but not exactly: keep the separate kernels
    set_error("calibrate_fakequant: %s codes round %g-bit values", dt_name(code_dtype), num_bits);
    return FFQ_ERR_UNSUPPORTED;
  }
  if (workspace == nullptr || workspace_bytes < 16) { set_error("calibrate_fakequant: workspace required"); return FFQ_ERR_WORKSPACE; }
  CalqArgs a{};
  a.x = x; a.y = y; a.run_min = run_min; a.run_max = run_max; a.run_dt = run_dtype;
  a.scale = scale; a.offset = offset; a.flags = flags;
  a.numel = (unsigned long long)plan.numel;
  const double lo = -pow(2.0, num_bits - 1.0);
  a.int_min_abs = (float)fabs(lo);
  a.int_max_abs = (float)fabs(-lo - 1.0);
  a.neg_int_min = (float)(-lo);
  a.steps = (float)(pow(2.0, num_bits) - 1.0);
  a.lo = (float)lo; a.hi = (float)(-lo - 1.0);
  a.symmetric = symmetric; a.allow_one_sided = allow_one_sided; a.rcp_div = rcp_div;
  a.sat8 = 0;
  a.code_is_int = float_codes ? 0 : 1;
  const int ept = 16 / dt_size(x_dtype);
  a.rows = (unsigned long long)plan.num_tiles;
  a.row_len = (unsigned int)plan.tile_numel;
  a.lanes = (unsigned int)(plan.tile_numel / ept);
  a.ws_flags = reinterpret_cast<unsigned int*>(static_cast<char*>(workspace) + 8);
  const bool decide = symmetric && allow_one_sided;
  if (decide) FFQ_CUDA_CHECK(cudaMemsetAsync(a.ws_flags, 0, 8, st));
  if (mode == 3) {
    switch (x_dtype) {
      case FFQ_F32: launch_group<float, true>(a, st); break;
      case FFQ_BF16: launch_group<__nv_bfloat16, true>(a, st); break;
      default: launch_group<__half, true>(a, st); break;
    }
  } else {
    const unsigned int nvec = a.row_len / ept;
    switch (x_dtype) {
      case FFQ_F32: launch_rows_fq<float>(a, nvec, st); break;
      case FFQ_BF16: launch_rows_fq<__nv_bfloat16>(a, nvec, st); break;
      default: launch_rows_fq<__half>(a, nvec, st); break;
    }
  }
  FFQ_LAUNCH_CHECK();
  if (decide) {
    switch (x_dtype) {
      case FFQ_F32: launch_sentinel_fixup<float, true>(a, st); break;
      case FFQ_BF16: launch_sentinel_fixup<__nv_bfloat16, true>(a, st); break;
      default: launch_sentinel_fixup<__half, true>(a, st); break;
    }
    FFQ_LAUNCH_CHECK();
  }
  return FFQ_OK;
}

int ffq_calibrate_fakequant_batched(const ffq_fq_item_t* items_dev, const uint32_t* block_start_dev, int64_t num_items,
                                    int64_t total_blocks, int x_dtype, int64_t tile_len, double num_bits, int symmetric,
                                    int allow_one_sided, int code_dtype, void* workspace, size_t workspace_bytes, void* stream) {
  static_assert(sizeof(ffq_fq_item_t) == sizeof(FqItem), "descriptor layout");
  const int rcp_div = (allow_one_sided & FFQ_FLAG_SCALAR_DIV_RECIPROCAL) ? 1 : 0;
  allow_one_sided &= FFQ_FLAG_ALLOW_ONE_SIDED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (num_items <= 0 || total_blocks <= 0) return FFQ_OK;
  if (!(x_dtype == FFQ_BF16 || x_dtype == FFQ_F16) || !(tile_len == 64 || tile_len == 128)) {
    set_error("calibrate_fakequant_batched: 16-bit data in tiles of 64 or 128 elements only");
    return FFQ_ERR_UNSUPPORTED;
  }
  if (items_dev == nullptr || block_start_dev == nullptr || num_items > 0x7fffffffll / (2 * FQ_FIXUP_BLOCKS) || total_blocks > 0x7fffffffll) {
    set_error("calibrate_fakequant_batched: bad descriptor table");
    return FFQ_ERR_INVALID;
  }
  const bool float_codes = code_dtype == FFQ_F32 || code_dtype == FFQ_F16 || code_dtype == FFQ_BF16;
  if (!float_codes && !is_int_dt(code_dtype)) { set_error("calibrate_fakequant_batched: bad code dtype"); return FFQ_ERR_INVALID; }
  const int mant = code_dtype == FFQ_F32 ? 23 : (code_dtype == FFQ_F16 ? 10 : 7);
  if (float_codes && mant + 1 < num_bits) { set_error("calibrate_fakequant_batched: %s codes round %g-bit values", dt_name(code_dtype), num_bits); return FFQ_ERR_UNSUPPORTED; }
  if (!float_codes && num_bits > dt_size(code_dtype) * 8) { set_error("calibrate_fakequant_batched: code dtype too narrow"); return FFQ_ERR_UNSUPPORTED; }
  const size_t need = (size_t)num_items * 8;
  if (workspace == nullptr || workspace_bytes < need) { set_error("calibrate_fakequant_batched: workspace of %zu bytes required", need); return FFQ_ERR_WORKSPACE; }
  CalqArgs a{};
  const double lo = -pow(2.0, num_bits - 1.0);
  a.int_min_abs = (float)fabs(lo);
  a.int_max_abs = (float)fabs(-lo - 1.0);
  a.neg_int_min = (float)(-lo);
  a.steps = (float)(pow(2.0, num_bits) - 1.0);
  a.lo = (float)lo; a.hi = (float)(-lo - 1.0);
  a.symmetric = symmetric; a.allow_one_sided = allow_one_sided; a.rcp_div = rcp_div;
  a.code_is_int = float_codes ? 0 : 1;
  a.row_len = (unsigned int)tile_len;
  a.lanes = (unsigned int)(tile_len / 8);
  unsigned int* flags_base = static_cast<unsigned int*>(workspace);
  const bool decide = symmetric && allow_one_sided;
  if (decide) FFQ_CUDA_CHECK(cudaMemsetAsync(flags_base, 0, need, st));
  const FqItem* items = reinterpret_cast<const FqItem*>(items_dev);
  const unsigned int grid = (unsigned int)total_blocks;
  const int n = (int)num_items;
#define FFQ_TTB(XT, NV)                                                                                                 \
  do {                                                                                                                \
    if (symmetric) launch_pdl(calq_tile_thread_batched_kernel<XT, NV, true>, dim3(grid), dim3(256), 0, st, a, items, block_start_dev, n, flags_base);   \
    else launch_pdl(calq_tile_thread_batched_kernel<XT, NV, false>, dim3(grid), dim3(256), 0, st, a, items, block_start_dev, n, flags_base);            \
  } while (0)
  if (x_dtype == FFQ_BF16) { if (tile_len == 64) FFQ_TTB(__nv_bfloat16, 8); else FFQ_TTB(__nv_bfloat16, 16); }
  else { if (tile_len == 64) FFQ_TTB(__half, 8); else FFQ_TTB(__half, 16); }
#undef FFQ_TTB
  FFQ_LAUNCH_CHECK();
  if (decide) {
    const unsigned int fgrid = (unsigned int)(num_items * FQ_FIXUP_BLOCKS);
    if (x_dtype == FFQ_BF16) launch_pdl(calq_sentinel_fixup_batched_kernel<__nv_bfloat16>, dim3(fgrid), dim3(256), 0, st, a, items, flags_base);
    else launch_pdl(calq_sentinel_fixup_batched_kernel<__half>, dim3(fgrid), dim3(256), 0, st, a, items, flags_base);
    FFQ_LAUNCH_CHECK();
  }
  return FFQ_OK;
}

}  // extern "C"

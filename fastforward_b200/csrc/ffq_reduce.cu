// ffq_reduce.cu -- the kernels with a per-tile reduction (a3, a4, a6, a7):
//   * straight-through backward: dx + per-tile dscale / doffset sums
//   * per-tile min/max, merged into running ranges (RunningMinMax calibration step)
//   * range -> (scale, offset) with the global one-sided decision taken on the device
//   * dynamic quantize = min/max -> params -> quantize, no host sync
//
// All are HBM-bound.  "Row" layouts (contiguous tiles) get three mappings:
//   - small tiles (tile_numel == LANES*EPT, LANES in 1..32): a LANES-wide sub-warp owns a tile,
//     the CTA streams a contiguous 16 KB chunk per unrolled step, reduction by shuffles only;
//   - medium tiles (a weight row, up to 64 Ki elements) when there are enough of them: one warp
//     per tile, shuffle-only reduction, no shared memory and no block barrier;
//   - large / few tiles: a 256-thread CTA owns a tile segment; few-and-huge tiles (per-tensor)
//     are split into S segments whose partials are combined by a second tiny kernel.
// Reductions are fixed trees (thread-serial -> shuffle -> shared memory -> optional second
// stage): deterministic run to run, no atomics on floating-point data.
// Algorithmic traffic per element: backward 3s (x, g in; dx out), min/max s.
#include <type_traits>

#include <cstring>

#include "ffq_common.cuh"

namespace ffq {

constexpr int RD_THREADS = 256;
constexpr int RD_UNROLL = 4;

// ------------------------------------------------------------------------------------------
// backward element math                              quantization/_quantizer_impl.py:203-237
// ------------------------------------------------------------------------------------------
struct BParams {
  float lo, hi;
  float lo_s, hi_s;   // the bounds as stored in a scale-dtype tensor (scale.new_tensor([lo]))
  int m_div, m_sub;   // as in the forward
  int m_s;            // rounding of the scale dtype (dscale chain runs in scale.dtype)
  int m_sg;           // rounding of promote(scale, grad)
  int has_offset;
};

__device__ __forceinline__ void bwd_terms(float x, float g, float s, float o, const BParams& p,
                                          float& dx, float& dsc, float& doff) {
  float pre = rnd(__fdiv_rn(x, s), p.m_div);
  pre = rnd(__fsub_rn(pre, o), p.m_sub);
  const float q = rintf(pre);
  const bool below = q < p.lo, above = q > p.hi;
  const bool clip = below || above;
  dx = clip ? 0.f : g;
  doff = clip ? rnd(__fmul_rn(s, g), p.m_sg) : 0.f;
  const float bound = rnd(__fadd_rn(below ? p.lo_s : p.hi_s, rnd(o, p.m_s)), p.m_s);
  const float resid = rnd(rnd(__fsub_rn(q, pre), p.m_sub), p.m_s);
  const float v = clip ? bound : resid;
  dsc = rnd(rnd(__fmul_rn(v, g), p.m_sg), p.m_s);
}

// fast path: all EPT elements share a tile and the whole chain has one promoted dtype RM
// per-tile constants of the fast path: formed once per warp / CTA (once per vector for small tiles)
struct BwdTile {
  SharedRcp k;
  float o, bound_lo, bound_hi;
};
template <int RM>
__device__ __forceinline__ BwdTile make_bwd_tile(float s, float o, const BParams& p) {
  BwdTile t;
  t.k = make_shared_rcp(s);
  t.o = o;
  const float o_s = rndc<RM>(o);
  t.bound_lo = rndc<RM>(__fadd_rn(p.lo_s, o_s));
  t.bound_hi = rndc<RM>(__fadd_rn(p.hi_s, o_s));
  return t;
}

// exact element (plain IEEE division), single promoted dtype RM
template <int RM>
__device__ __forceinline__ void bwd_elem_exact(float x, float g, const BwdTile& t, const BParams& p, float& dx,
                                               float& dsc, float& doff) {
  const float s = t.k.s;
  float pre = rndc<RM>(__fdiv_rn(x, s));
  pre = rndc<RM>(__fsub_rn(pre, t.o));
  const float q = rintf(pre);
  const bool below = q < p.lo, above = q > p.hi, clip = below || above;
  dx = clip ? 0.f : g;
  doff = clip ? rndc<RM>(__fmul_rn(s, g)) : 0.f;
  const float v = clip ? (below ? t.bound_lo : t.bound_hi) : rndc<RM>(__fsub_rn(q, pre));
  dsc = rndc<RM>(__fmul_rn(v, g));
}

// fast path: all EPT elements share a tile and the whole chain has one promoted dtype RM.
// The quotient comes from the shared reciprocal; the vector is redone exactly when the scale or any
// quotient leaves the box [2^-50, 2^60] in which that quotient is proven to equal __fdiv_rn (this
// includes vectors containing exact zeros -- rare in dense weights/activations).
// Returns true when NO element of the vector clips: then dx equals g bit for bit (the caller stores the raw
// gradient vector, no select / re-pack) and only the q - pre residuals feed dscale.  Calibrated ranges clip few
// elements, so this is the common case; a sub-vector with a clipped element redoes its sums with the selections
// from the quotients kept in registers.  The vector is processed in sub-vectors of at most 4 elements (fewer live
// registers for 16-bit data); the sums run sequentially over the elements in every path, so the association --
// and with it the result -- does not depend on which path a sub-vector took.
template <int RM, int EPT>
__device__ __forceinline__ bool bwd_vector(const float (&x)[EPT], const float (&g)[EPT], float (&dx)[EPT],
                                           const BwdTile& t, const BParams& p, float& sum_sc, float& sum_off) {
  constexpr int SUB = EPT < 4 ? EPT : 4;
  const float s = t.k.s, r = t.k.r;
  float sc = 0.f, off = 0.f;
  bool clean = true;
#pragma unroll
  for (int h = 0; h < EPT; h += SUB) {
    const float sc0 = sc;
    float amax = 0.f, amin = INFINITY, qmin = INFINITY, qmax = -INFINITY;
    float qv[SUB], rv[SUB];
#pragma unroll
    for (int j = 0; j < SUB; ++j) {
      const int i = h + j;
      const float q0 = __fmul_rn(x[i], r);
      const float e = __fmaf_rn(-s, q0, x[i]);
      const float quo = __fmaf_rn(r, e, q0);
      amax = nan_max(amax, fabsf(quo));
      amin = fminf(amin, fabsf(quo));
      const float pre = rndc<RM>(__fsub_rn(rndc<RM>(quo), t.o));
      const float q = rintf(pre);
      qv[j] = q;
      rv[j] = rndc<RM>(__fsub_rn(q, pre));
      qmin = fminf(qmin, q);                     // NaN-ignoring: a NaN code does not clip (q < lo, q > hi are false)
      qmax = fmaxf(qmax, q);
      sc += rndc<RM>(__fmul_rn(rv[j], g[i]));
      dx[i] = g[i];
    }
    if (!(t.k.ok && amax <= 0x1p60f && amin >= 0x1p-50f)) {     // outside the proven box: plain IEEE division
      sc = sc0;
      clean = false;
#pragma unroll
      for (int j = 0; j < SUB; ++j) {
        float dsc, doff;
        bwd_elem_exact<RM>(x[h + j], g[h + j], t, p, dx[h + j], dsc, doff);
        sc += dsc; off += doff;
      }
    } else if (!(qmin >= p.lo && qmax <= p.hi)) {               // something clips: redo the sums with the selections
      sc = sc0;
      clean = false;
#pragma unroll
      for (int j = 0; j < SUB; ++j) {
        const int i = h + j;
        const bool below = qv[j] < p.lo, above = qv[j] > p.hi, clip = below || above;
        dx[i] = clip ? 0.f : g[i];
        if (p.has_offset) off += clip ? rndc<RM>(__fmul_rn(s, g[i])) : 0.f;
        const float v = clip ? (below ? t.bound_lo : t.bound_hi) : rv[j];
        sc += rndc<RM>(__fmul_rn(v, g[i]));
      }
    }
  }
  sum_sc += sc;
  sum_off += off;
  return clean;
}

template <typename T, int N>
__device__ __forceinline__ void unpack(const Vec<T, N>& v, float (&f)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) f[i] = Elem<T>::to_f(v.v[i]);
}
template <typename T, int N>
__device__ __forceinline__ void pack(const float (&f)[N], Vec<T, N>& v) {
  if constexpr (std::is_same<T, __nv_bfloat16>::value && (N % 2 == 0)) {
#pragma unroll
    for (int i = 0; i < N; i += 2)
      *reinterpret_cast<__nv_bfloat162*>(&v.v[i]) = __floats2bfloat162_rn(f[i], f[i + 1]);
  } else if constexpr (std::is_same<T, __half>::value && (N % 2 == 0)) {
#pragma unroll
    for (int i = 0; i < N; i += 2) *reinterpret_cast<__half2*>(&v.v[i]) = __floats2half2_rn(f[i], f[i + 1]);
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) v.v[i] = Elem<T>::from_f(f[i]);
  }
}

struct BwdArgs {
  const void* x; const void* g; void* dx;
  int x_dt, g_dt;
  const void* scale; const void* offset; int s_dt, o_dt;
  void* dscale; void* doffset; int dsc_dt, doff_dt;   // final outputs (S == 1) ...
  float* part;                                        // ... or partials [2][num_tiles*S]
  unsigned long long numel, tile_numel, num_tiles;
  unsigned long long seg_len; unsigned int S;
  BParams bp;
  GenericLayout gl;
};

__device__ __forceinline__ float block_sum(float v, float* smem) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) smem[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float r = (threadIdx.x < nw) ? smem[threadIdx.x] : 0.f;
  if (w == 0) r = warp_sum(r);
  return r;  // valid in warp 0
}

// per-type unroll: the same bytes in flight per thread, fewer live registers for 16-bit data
template <typename T> struct BwdUnroll { static constexpr int value = sizeof(T) >= 4 ? 4 : 2; };

// --- small tiles: one LANES-wide group per tile -------------------------------------------
// T is the dtype of x, g and dx (autograd hands back the gradient in the dtype of the output);
// mixed x/g dtypes take the generic kernel.
template <int RM> struct RParamT { using type = float; };
template <> struct RParamT<RM_BF16> { using type = __nv_bfloat16; };
template <> struct RParamT<RM_F16> { using type = __half; };

template <typename T, int LANES, int RM>
__global__ void __launch_bounds__(RD_THREADS, 3) bwd_row_group_kernel(const BwdArgs a) {
  pdl_wait();                    // programmatic dependent launch: no-ops unless launched that way
  pdl_trigger();
  constexpr int EPT = 16 / sizeof(T);
  constexpr int U = BwdUnroll<T>::value;
  using PT = typename RParamT<RM>::type;      // the fast path requires scale/offset stored in the chain's dtype
  const PT* __restrict__ scale = static_cast<const PT*>(a.scale);
  const PT* __restrict__ offset = static_cast<const PT*>(a.offset);
  const T* __restrict__ x = static_cast<const T*>(a.x);
  const T* __restrict__ g = static_cast<const T*>(a.g);
  T* __restrict__ dx = static_cast<T*>(a.dx);
  const unsigned long long nvec = a.numel / EPT;
  const unsigned long long vbase = (unsigned long long)blockIdx.x * (RD_THREADS * U) + threadIdx.x;

  Vec<T, EPT> xv[U], gv[U];
  float sv[U], ov[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const unsigned long long v = vbase + (unsigned long long)u * RD_THREADS;
    sv[u] = 1.f; ov[u] = 0.f;
    if (v < nvec) {
      xv[u] = ld_stream<T, EPT>(x + v * EPT);
      gv[u] = ld_stream<T, EPT>(g + v * EPT);
      sv[u] = Elem<PT>::to_f(scale[v / LANES]);          // requested with the data: latency overlaps the stream
      if (offset) ov[u] = Elem<PT>::to_f(offset[v / LANES]);
    }
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const unsigned long long v = vbase + (unsigned long long)u * RD_THREADS;
    const bool live = v < nvec;   // whole groups are live or dead together (nvec % LANES == 0)
    const unsigned long long tile = v / LANES;
    float sum_sc = 0.f, sum_off = 0.f;
    if (live) {
      const float s = sv[u];
      const float o = rintf(ov[u]);
      float xf[EPT], gf[EPT], df[EPT];
      unpack<T, EPT>(xv[u], xf);
      unpack<T, EPT>(gv[u], gf);
      const BwdTile bt = make_bwd_tile<RM>(s, o, a.bp);
      if (bwd_vector<RM, EPT>(xf, gf, df, bt, a.bp, sum_sc, sum_off)) {
        st_vec<T, EPT>(dx + v * EPT, gv[u]);               // no element clipped: dx is the gradient itself
      } else {
        Vec<T, EPT> d;
        pack<T, EPT>(df, d);
        st_vec<T, EPT>(dx + v * EPT, d);
      }
    }
    sum_sc = group_sum<LANES>(sum_sc);
    if (a.bp.has_offset) sum_off = group_sum<LANES>(sum_off);
    if (live && (threadIdx.x & (LANES - 1)) == 0) {
      store_from_float(a.dscale, a.dsc_dt, tile, sum_sc);
      if (a.bp.has_offset) store_from_float(a.doffset, a.doff_dt, tile, sum_off);
    }
  }
}

// Shared body: `nthreads` cooperating threads (a warp or a CTA), thread `tid` of them, stream
// elements [0, len) of x/g/dx (already offset to the tile segment) and accumulate the sums.
template <typename T, int EPT, int RM>
__device__ __forceinline__ void bwd_stream(const T* __restrict__ x, const T* __restrict__ g, T* __restrict__ dx,
                                           unsigned int len, unsigned int tid, unsigned int nthreads, float s,
                                           float o, const BParams& bp, float& sum_sc, float& sum_off) {
  constexpr int U = BwdUnroll<T>::value;
  const unsigned int step = nthreads * EPT;
  const BwdTile bt = make_bwd_tile<RM>(s, o, bp);
  for (unsigned int i0 = tid * EPT; i0 < len; i0 += step * U) {
    Vec<T, EPT> xv[U], gv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned int i = i0 + u * step;
      if (i < len) {
        if constexpr (EPT == 1) { xv[u].v[0] = x[i]; gv[u].v[0] = g[i]; }
        else { xv[u] = ld_stream<T, EPT>(x + i); gv[u] = ld_stream<T, EPT>(g + i); }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned int i = i0 + u * step;
      if (i < len) {
        float xf[EPT], gf[EPT], df[EPT];
        unpack<T, EPT>(xv[u], xf);
        unpack<T, EPT>(gv[u], gf);
        if (bwd_vector<RM, EPT>(xf, gf, df, bt, bp, sum_sc, sum_off)) {   // no element clipped: dx is the gradient itself
          if constexpr (EPT == 1) dx[i] = gv[u].v[0];
          else st_vec<T, EPT>(dx + i, gv[u]);
        } else {
          Vec<T, EPT> d;
          pack<T, EPT>(df, d);
          if constexpr (EPT == 1) dx[i] = d.v[0];
          else st_vec<T, EPT>(dx + i, d);
        }
      }
    }
  }
}

// --- medium tiles: one warp per tile ---------------------------------------------------------
// fp32 data fits 64 registers without spills (4 CTAs per SM: cfg1 55 -> 50 us); the 16-bit variants would spill
template <typename T, int EPT, int RM>
__global__ void __launch_bounds__(RD_THREADS, 4) bwd_row_warp_kernel(const BwdArgs a) {
  pdl_wait();                    // programmatic dependent launch: no-ops unless launched that way
  pdl_trigger();
  const unsigned long long tile = (unsigned long long)blockIdx.x * (RD_THREADS / 32) + (threadIdx.x >> 5);
  if (tile >= a.num_tiles) return;
  const unsigned int lane = threadIdx.x & 31;
  const unsigned long long base = tile * a.tile_numel;
  const float s = load_as_float(a.scale, a.s_dt, tile);
  const float o = load_offset(a.offset, a.o_dt, tile);
  float sum_sc = 0.f, sum_off = 0.f;
  bwd_stream<T, EPT, RM>(static_cast<const T*>(a.x) + base, static_cast<const T*>(a.g) + base,
                         static_cast<T*>(a.dx) + base, (unsigned int)a.tile_numel, lane, 32u, s, o, a.bp, sum_sc,
                         sum_off);
  sum_sc = warp_sum(sum_sc);
  if (a.bp.has_offset) sum_off = warp_sum(sum_off);
  if (lane == 0) {
    store_from_float(a.dscale, a.dsc_dt, tile, sum_sc);
    if (a.bp.has_offset) store_from_float(a.doffset, a.doff_dt, tile, sum_off);
  }
}

// --- large / few tiles: one CTA per tile segment ---------------------------------------------
template <typename T, int EPT, int RM>
__global__ void __launch_bounds__(RD_THREADS, 4) bwd_row_cta_kernel(const BwdArgs a) {
  pdl_wait();                    // programmatic dependent launch: no-ops unless launched that way
  pdl_trigger();
  __shared__ float smem[32];
  const unsigned long long tile = blockIdx.x / a.S;
  const unsigned int seg = blockIdx.x % a.S;
  const unsigned long long begin = (unsigned long long)seg * a.seg_len;
  const unsigned long long remain = a.tile_numel - begin;
  const unsigned int len = (unsigned int)(remain < a.seg_len ? remain : a.seg_len);
  const unsigned long long base = tile * a.tile_numel + begin;
  const float s = load_as_float(a.scale, a.s_dt, tile);
  const float o = load_offset(a.offset, a.o_dt, tile);
  float sum_sc = 0.f, sum_off = 0.f;
  bwd_stream<T, EPT, RM>(static_cast<const T*>(a.x) + base, static_cast<const T*>(a.g) + base,
                         static_cast<T*>(a.dx) + base, len, threadIdx.x, blockDim.x, s, o, a.bp, sum_sc, sum_off);
  sum_sc = block_sum(sum_sc, smem);
  if (a.bp.has_offset) sum_off = block_sum(sum_off, smem);
  if (threadIdx.x == 0) {
    if (a.S == 1) {
      store_from_float(a.dscale, a.dsc_dt, tile, sum_sc);
      if (a.bp.has_offset) store_from_float(a.doffset, a.doff_dt, tile, sum_off);
    } else {
      a.part[blockIdx.x] = sum_sc;
      if (a.bp.has_offset) a.part[a.num_tiles * a.S + blockIdx.x] = sum_off;
    }
  }
}

// second stage: one warp per tile folds the S partials in a fixed order
__global__ void __launch_bounds__(RD_THREADS) bwd_finalize_kernel(const BwdArgs a) {
  pdl_wait();                    // programmatic dependent launch: no-ops unless launched that way
  pdl_trigger();
  const unsigned long long tile = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (tile >= a.num_tiles) return;
  float sc = 0.f, off = 0.f;
  for (unsigned int k = lane; k < a.S; k += 32) {
    sc += a.part[tile * a.S + k];
    if (a.bp.has_offset) off += a.part[a.num_tiles * a.S + tile * a.S + k];
  }
  sc = warp_sum(sc);
  off = warp_sum(off);
  if (lane == 0) {
    store_from_float(a.dscale, a.dsc_dt, tile, sc);
    if (a.bp.has_offset) store_from_float(a.doffset, a.doff_dt, tile, off);
  }
}

// element offset of the k-th element (row-major inside the tile) of tile `t`
__device__ __forceinline__ unsigned long long tile_elem_offset(unsigned long long t, unsigned long long k,
                                                               const GenericLayout& gl) {
  unsigned long long off = 0;
#pragma unroll 1
  for (int d = gl.rank - 1; d >= 0; --d) {
    const unsigned long long tc = t % gl.grid[d]; t /= gl.grid[d];
    const unsigned long long kc = k % gl.tile[d]; k /= gl.tile[d];
    off += (tc * gl.tile[d] + kc) * gl.stride[d];
  }
  return off;
}

// any layout, any dtype: one CTA per tile, scalar index math.  Correctness path.
__global__ void __launch_bounds__(128) bwd_generic_kernel(const BwdArgs a) {
  __shared__ float smem[32];
  const unsigned long long tile = blockIdx.x;
  const float s = load_as_float(a.scale, a.s_dt, tile);
  const float o = load_offset(a.offset, a.o_dt, tile);
  float sum_sc = 0.f, sum_off = 0.f;
  for (unsigned long long k = threadIdx.x; k < a.tile_numel; k += blockDim.x) {
    const unsigned long long e = a.gl.rank <= 1 ? tile * a.tile_numel + k : tile_elem_offset(tile, k, a.gl);
    float dxi, dsc, doff;
    bwd_terms(load_as_float(a.x, a.x_dt, e), load_as_float(a.g, a.g_dt, e), s, o, a.bp, dxi, dsc, doff);
    store_from_float(a.dx, a.g_dt, e, dxi);
    sum_sc += dsc;
    sum_off += doff;
  }
  sum_sc = block_sum(sum_sc, smem);
  sum_off = block_sum(sum_off, smem);
  if (threadIdx.x == 0) {
    store_from_float(a.dscale, a.dsc_dt, tile, sum_sc);
    if (a.bp.has_offset) store_from_float(a.doffset, a.doff_dt, tile, sum_off);
  }
}

// ------------------------------------------------------------------------------------------
// segmentation shared by backward and min/max
// ------------------------------------------------------------------------------------------
struct Segmentation { unsigned int S; unsigned long long seg_len; };

constexpr long long WARP_TILE_MAX = 1ll << 16;   // largest tile a single warp streams

// One warp per tile pays off when there are enough tiles to fill the machine with warps.
static bool use_warp_per_tile(const Plan& plan, int ept) {
  return plan.tile_numel > 32ll * ept && plan.tile_numel <= WARP_TILE_MAX &&
         plan.num_tiles >= 16ll * sm_count();
}

static Segmentation choose_segments(const Plan& plan, int ept) {
  Segmentation sg{1, (unsigned long long)plan.tile_numel};
  if (use_warp_per_tile(plan, ept)) return sg;
  const long long target = 6ll * sm_count();
  const unsigned long long quantum = (unsigned long long)RD_THREADS * ept * RD_UNROLL;
  const unsigned long long max_seg = 1ull << 30;   // in-segment indices are 32-bit
  if ((plan.num_tiles >= target || (unsigned long long)plan.tile_numel <= quantum) &&
      (unsigned long long)plan.tile_numel <= max_seg)
    return sg;
  unsigned long long want = (unsigned long long)((target + plan.num_tiles - 1) / plan.num_tiles);
  if (want < 1) want = 1;
  unsigned long long seg = ((unsigned long long)plan.tile_numel + want - 1) / want;
  seg = (seg + quantum - 1) / quantum * quantum;
  if (seg > max_seg) seg = max_seg;
  sg.seg_len = seg;
  sg.S = (unsigned int)(((unsigned long long)plan.tile_numel + seg - 1) / seg);
  if (sg.S <= 1) { sg.S = 1; sg.seg_len = (unsigned long long)plan.tile_numel; }
  return sg;
}

static int group_lanes(const Plan& plan, int ept) {
  // tile_numel == LANES*ept with LANES a power of two <= 32
  if (plan.tile_numel % ept) return 0;
  const long long l = plan.tile_numel / ept;
  if (l >= 1 && l <= 32 && (l & (l - 1)) == 0) return (int)l;
  return 0;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename T, int RM>
static void launch_bwd_row(const BwdArgs& a, const Plan& plan, bool vec_ok, cudaStream_t st) {
  constexpr int EPT = 16 / sizeof(T);
  const int lanes = vec_ok ? group_lanes(plan, EPT) : 0;
  if (lanes) {
    const unsigned long long nvec = a.numel / EPT;
    constexpr int U = BwdUnroll<T>::value;
    const unsigned int blocks = (unsigned int)((nvec + RD_THREADS * U - 1) / (RD_THREADS * U));
    switch (lanes) {
      case 1: launch_pdl(bwd_row_group_kernel<T, 1, RM>, dim3(blocks), dim3(RD_THREADS), 0, st, a); break;
      case 2: launch_pdl(bwd_row_group_kernel<T, 2, RM>, dim3(blocks), dim3(RD_THREADS), 0, st, a); break;
      case 4: launch_pdl(bwd_row_group_kernel<T, 4, RM>, dim3(blocks), dim3(RD_THREADS), 0, st, a); break;
      case 8: launch_pdl(bwd_row_group_kernel<T, 8, RM>, dim3(blocks), dim3(RD_THREADS), 0, st, a); break;
      case 16: launch_pdl(bwd_row_group_kernel<T, 16, RM>, dim3(blocks), dim3(RD_THREADS), 0, st, a); break;
      default: launch_pdl(bwd_row_group_kernel<T, 32, RM>, dim3(blocks), dim3(RD_THREADS), 0, st, a); break;
    }
    return;
  }
  const int ept = (vec_ok && plan.tile_numel % EPT == 0) ? EPT : 1;
  if (a.S == 1 && use_warp_per_tile(plan, EPT)) {
    const unsigned int blocks = (unsigned int)((a.num_tiles + RD_THREADS / 32 - 1) / (RD_THREADS / 32));
    if (ept == EPT) launch_pdl(bwd_row_warp_kernel<T, EPT, RM>, dim3(blocks), dim3(RD_THREADS), 0, st, a);
    else launch_pdl(bwd_row_warp_kernel<T, 1, RM>, dim3(blocks), dim3(RD_THREADS), 0, st, a);
    return;
  }
  const unsigned int blocks = (unsigned int)(a.num_tiles * a.S);
  // shrink the CTA for short tiles so that lanes are not idle
  unsigned long long per_seg = (a.seg_len + ept - 1) / ept;
  int threads = RD_THREADS;
  while (threads > 32 && (unsigned long long)threads / 2 >= per_seg) threads /= 2;
  if (ept == EPT) launch_pdl(bwd_row_cta_kernel<T, EPT, RM>, dim3(blocks), dim3(threads), 0, st, a);
  else launch_pdl(bwd_row_cta_kernel<T, 1, RM>, dim3(blocks), dim3(threads), 0, st, a);
}

// fast kernels exist for x.dtype == g.dtype with a single promoted dtype for the whole chain
static bool dispatch_bwd(const BwdArgs& a, const Plan& plan, bool vec_ok, cudaStream_t st) {
  if (a.x_dt != a.g_dt) return false;
  const BParams& p = a.bp;
  if (!(p.m_div == p.m_sub && p.m_sub == p.m_s && p.m_s == p.m_sg)) return false;
  const int rm = p.m_div;
  switch (a.x_dt) {
    case FFQ_F32: if (rm == RM_F32) { launch_bwd_row<float, RM_F32>(a, plan, vec_ok, st); return true; } break;
    case FFQ_BF16:
      if (rm == RM_F32) { launch_bwd_row<__nv_bfloat16, RM_F32>(a, plan, vec_ok, st); return true; }
      if (rm == RM_BF16) { launch_bwd_row<__nv_bfloat16, RM_BF16>(a, plan, vec_ok, st); return true; }
      break;
    case FFQ_F16:
      if (rm == RM_F32) { launch_bwd_row<__half, RM_F32>(a, plan, vec_ok, st); return true; }
      if (rm == RM_F16) { launch_bwd_row<__half, RM_F16>(a, plan, vec_ok, st); return true; }
      break;
  }
  return false;
}


static int ept_of(int dt) {
  const int sz = dt_size(dt);
  return (sz > 0 && sz <= 4) ? 16 / sz : 4;
}

// ------------------------------------------------------------------------------------------
// min / max                                                range_setting/minmax.py:226-237
// ------------------------------------------------------------------------------------------
struct MmArgs {
  const void* x; int x_dt;
  void* tile_min; void* tile_max;     // optional, dtype x_dt
  void* run_min; void* run_max;       // optional, dtype run_dt, updated in place
  int run_dt;
  int32_t* flags;                     // optional
  float* part;                        // [2][num_tiles*S] when S > 1
  unsigned long long numel, tile_numel, num_tiles, seg_len; unsigned int S;
  GenericLayout gl;
};

__device__ __forceinline__ void mm_emit(const MmArgs& a, unsigned long long tile, float mn, float mx) {
  if (a.tile_min) store_from_float(a.tile_min, a.x_dt, tile, mn);
  if (a.tile_max) store_from_float(a.tile_max, a.x_dt, tile, mx);
  if (a.run_min) store_from_float(a.run_min, a.run_dt, tile, nan_min(load_as_float(a.run_min, a.run_dt, tile), mn));
  if (a.run_max) store_from_float(a.run_max, a.run_dt, tile, nan_max(load_as_float(a.run_max, a.run_dt, tile), mx));
  if (a.flags && (isinf(mn) || isinf(mx))) atomicOr(a.flags, 1);
}

template <typename XT, int LANES>
__global__ void __launch_bounds__(RD_THREADS) mm_row_group_kernel(const MmArgs a) {
  pdl_wait();                    // programmatic dependent launch: no-ops unless launched that way
  pdl_trigger();
  constexpr int EPT = 16 / sizeof(XT);
  const XT* __restrict__ x = static_cast<const XT*>(a.x);
  const unsigned long long nvec = a.numel / EPT;
  const unsigned long long vbase = (unsigned long long)blockIdx.x * (RD_THREADS * RD_UNROLL) + threadIdx.x;
  Vec<XT, EPT> xv[RD_UNROLL];
#pragma unroll
  for (int u = 0; u < RD_UNROLL; ++u) {
    const unsigned long long v = vbase + (unsigned long long)u * RD_THREADS;
    if (v < nvec) xv[u] = ld_stream<XT, EPT>(x + v * EPT);
  }
#pragma unroll
  for (int u = 0; u < RD_UNROLL; ++u) {
    const unsigned long long v = vbase + (unsigned long long)u * RD_THREADS;
    const bool live = v < nvec;   // whole groups are live or dead together (nvec % LANES == 0)
    float mn = INFINITY, mx = -INFINITY;
    if (live) vec_minmax<XT, EPT>(xv[u], mn, mx);
    mn = group_min<LANES>(mn);
    mx = group_max<LANES>(mx);
    if (live && (threadIdx.x & (LANES - 1)) == 0) mm_emit(a, v / LANES, mn, mx);
  }
}

template <typename XT, int EPT>
__device__ __forceinline__ void mm_stream(const XT* __restrict__ x, unsigned int len, unsigned int tid,
                                          unsigned int nthreads, float& mn, float& mx) {
  constexpr int U = RD_UNROLL;
  const unsigned int step = nthreads * EPT;
  for (unsigned int i0 = tid * EPT; i0 < len; i0 += step * U) {
    Vec<XT, EPT> xv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned int i = i0 + u * step;
      if (i < len) {
        if constexpr (EPT == 1) xv[u].v[0] = x[i];
        else xv[u] = ld_stream<XT, EPT>(x + i);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned int i = i0 + u * step;
      if (i < len) {
        float vmn, vmx;
        vec_minmax<XT, EPT>(xv[u], vmn, vmx);
        mn = nan_min(mn, vmn);
        mx = nan_max(mx, vmx);
      }
    }
  }
}

template <typename XT, int EPT>
__global__ void __launch_bounds__(RD_THREADS) mm_row_warp_kernel(const MmArgs a) {
  pdl_wait();                    // programmatic dependent launch: no-ops unless launched that way
  pdl_trigger();
  const unsigned long long tile = (unsigned long long)blockIdx.x * (RD_THREADS / 32) + (threadIdx.x >> 5);
  if (tile >= a.num_tiles) return;
  const unsigned int lane = threadIdx.x & 31;
  const XT* __restrict__ x = static_cast<const XT*>(a.x) + tile * a.tile_numel;
  // a tile is never empty, so its first element is a valid identity for every lane
  float mn = Elem<XT>::to_f(x[0]), mx = mn;
  mm_stream<XT, EPT>(x, (unsigned int)a.tile_numel, lane, 32u, mn, mx);
  mn = group_min<32>(mn);
  mx = group_max<32>(mx);
  if (lane == 0) mm_emit(a, tile, mn, mx);
}

template <typename XT, int EPT>
__global__ void __launch_bounds__(RD_THREADS) mm_row_cta_kernel(const MmArgs a) {
  pdl_wait();                    // programmatic dependent launch: no-ops unless launched that way
  pdl_trigger();
  __shared__ float smem[64];
  const unsigned long long tile = blockIdx.x / a.S;
  const unsigned int seg = blockIdx.x % a.S;
  const unsigned long long begin = (unsigned long long)seg * a.seg_len;
  const unsigned long long remain = a.tile_numel - begin;
  const unsigned int len = (unsigned int)(remain < a.seg_len ? remain : a.seg_len);
  const XT* __restrict__ x = static_cast<const XT*>(a.x) + tile * a.tile_numel + begin;
  float mn = Elem<XT>::to_f(x[0]), mx = mn;
  mm_stream<XT, EPT>(x, len, threadIdx.x, blockDim.x, mn, mx);
  block_minmax(mn, mx, smem);
  if (threadIdx.x == 0) {
    if (a.S == 1) mm_emit(a, tile, mn, mx);
    else { a.part[blockIdx.x] = mn; a.part[a.num_tiles * a.S + blockIdx.x] = mx; }
  }
}

__global__ void __launch_bounds__(RD_THREADS) mm_finalize_kernel(const MmArgs a) {
  pdl_wait();                    // programmatic dependent launch: no-ops unless launched that way
  pdl_trigger();
  const unsigned long long tile = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (tile >= a.num_tiles) return;
  float mn = a.part[tile * a.S], mx = a.part[a.num_tiles * a.S + tile * a.S];
  for (unsigned int k = lane; k < a.S; k += 32) {
    mn = nan_min(mn, a.part[tile * a.S + k]);
    mx = nan_max(mx, a.part[a.num_tiles * a.S + tile * a.S + k]);
  }
  mn = group_min<32>(mn);
  mx = group_max<32>(mx);
  if (lane == 0) mm_emit(a, tile, mn, mx);
}

__global__ void __launch_bounds__(128) mm_generic_kernel(const MmArgs a) {
  __shared__ float smem[64];
  const unsigned long long tile = blockIdx.x;
  const unsigned long long e0 = a.gl.rank <= 1 ? tile * a.tile_numel : tile_elem_offset(tile, 0, a.gl);
  float mn = load_as_float(a.x, a.x_dt, e0), mx = mn;
  for (unsigned long long k = threadIdx.x; k < a.tile_numel; k += blockDim.x) {
    const unsigned long long e = a.gl.rank <= 1 ? tile * a.tile_numel + k : tile_elem_offset(tile, k, a.gl);
    const float f = load_as_float(a.x, a.x_dt, e);
    mn = nan_min(mn, f);
    mx = nan_max(mx, f);
  }
  block_minmax(mn, mx, smem);
  if (threadIdx.x == 0) mm_emit(a, tile, mn, mx);
}

template <typename XT>
static void launch_mm_row(const MmArgs& a, const Plan& plan, bool vec_ok, cudaStream_t st) {
  constexpr int EPT = 16 / sizeof(XT);
  const int lanes = vec_ok ? group_lanes(plan, EPT) : 0;
  if (lanes) {
    const unsigned long long nvec = a.numel / EPT;
    const unsigned int blocks = (unsigned int)((nvec + RD_THREADS * RD_UNROLL - 1) / (RD_THREADS * RD_UNROLL));
    switch (lanes) {
      case 1: launch_pdl(mm_row_group_kernel<XT, 1>, dim3(blocks), dim3(RD_THREADS), 0, st, a); break;
      case 2: launch_pdl(mm_row_group_kernel<XT, 2>, dim3(blocks), dim3(RD_THREADS), 0, st, a); break;
      case 4: launch_pdl(mm_row_group_kernel<XT, 4>, dim3(blocks), dim3(RD_THREADS), 0, st, a); break;
      case 8: launch_pdl(mm_row_group_kernel<XT, 8>, dim3(blocks), dim3(RD_THREADS), 0, st, a); break;
      case 16: launch_pdl(mm_row_group_kernel<XT, 16>, dim3(blocks), dim3(RD_THREADS), 0, st, a); break;
      default: launch_pdl(mm_row_group_kernel<XT, 32>, dim3(blocks), dim3(RD_THREADS), 0, st, a); break;
    }
    return;
  }
  const int ept = (vec_ok && plan.tile_numel % EPT == 0) ? EPT : 1;
  if (a.S == 1 && use_warp_per_tile(plan, EPT)) {
    const unsigned int blocks = (unsigned int)((a.num_tiles + RD_THREADS / 32 - 1) / (RD_THREADS / 32));
    if (ept == EPT) launch_pdl(mm_row_warp_kernel<XT, EPT>, dim3(blocks), dim3(RD_THREADS), 0, st, a);
    else launch_pdl(mm_row_warp_kernel<XT, 1>, dim3(blocks), dim3(RD_THREADS), 0, st, a);
    return;
  }
  const unsigned int blocks = (unsigned int)(a.num_tiles * a.S);
  unsigned long long per_seg = (a.seg_len + ept - 1) / ept;
  int threads = RD_THREADS;
  while (threads > 32 && (unsigned long long)threads / 2 >= per_seg) threads /= 2;
  if (ept == EPT) launch_pdl(mm_row_cta_kernel<XT, EPT>, dim3(blocks), dim3(threads), 0, st, a);
  else launch_pdl(mm_row_cta_kernel<XT, 1>, dim3(blocks), dim3(threads), 0, st, a);
}

// ------------------------------------------------------------------------------------------
// range -> (scale, offset)                                 quantization/affine/range.py:54-122
// ------------------------------------------------------------------------------------------
struct PrArgs {
  const void* mn; const void* mx; int r_dt;
  unsigned long long n;
  float int_min_abs, int_max_abs, neg_int_min, steps;
  int symmetric, allow_one_sided, round_offset, rcp;
  void* scale; int s_dt; void* offset; int o_dt;
  float* part; unsigned int nparts;
};

// stage A (only when the one-sided decision is live): per-block min of min_range -> part[]
__global__ void __launch_bounds__(RD_THREADS) pr_min_kernel(const PrArgs a) {
  pdl_wait();                    // programmatic dependent launch: no-ops unless launched that way
  pdl_trigger();
  __shared__ float smem[64];
  float mn = INFINITY, dummy = 0.f;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n;
       i += (unsigned long long)gridDim.x * blockDim.x)
    mn = nan_min(mn, load_as_float(a.mn, a.r_dt, i));
  block_minmax(mn, dummy, smem);
  if (threadIdx.x == 0) a.part[blockIdx.x] = mn;
}

__global__ void __launch_bounds__(RD_THREADS) pr_apply_kernel(const PrArgs a) {
  pdl_wait();                    // programmatic dependent launch: no-ops unless launched that way
  pdl_trigger();
  __shared__ float smem[64];
  __shared__ int s_one_sided;
  bool one_sided = false;
  if (a.allow_one_sided && a.symmetric) {   // for asymmetric quantizers the flag changes nothing
    float mn = INFINITY, dummy = 0.f;
    for (unsigned int i = threadIdx.x; i < a.nparts; i += blockDim.x) mn = nan_min(mn, a.part[i]);
    block_minmax(mn, dummy, smem);
    if (threadIdx.x == 0) s_one_sided = (mn >= 0.f) ? 1 : 0;   // NaN >= 0 is false, as in Python
    __syncthreads();
    one_sided = s_one_sided != 0;
  }
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  float mn = load_as_float(a.mn, a.r_dt, i);
  const float mx = load_as_float(a.mx, a.r_dt, i);
  if (a.symmetric && !one_sided) {
    const float neg = scalar_div(fabsf(mn), a.int_min_abs, a.rcp != 0);
    const float pos = scalar_div(fabsf(mx), a.int_max_abs, a.rcp != 0);
    store_from_float(a.scale, a.s_dt, i, nan_max(neg, pos));
    if (a.offset) store_from_float(a.offset, a.o_dt, i, 0.f);
    return;
  }
  if (a.symmetric) mn = 0.f;
  float sc = scalar_div(__fsub_rn(mx, mn), a.steps, a.rcp != 0);
  const float eps = 1.1920928955078125e-07f;
  sc = (sc != sc) ? sc : fmaxf(sc, eps);
  float off = __fadd_rn(__fdiv_rn(mn, sc), a.neg_int_min);   // min/scale - int_min
  if (a.round_offset) off = rintf(off);
  store_from_float(a.scale, a.s_dt, i, sc);
  if (a.offset) store_from_float(a.offset, a.o_dt, i, off);
}

// Batched variant: one CTA per quantizer, all running ranges living in one contiguous buffer (the estimator's
// arena).  desc[q] = {start, len, scale_ptr, offset_ptr (0: none), c0, c1, cfg, 0} with c0/c1/cfg the words
// ffq_params_for_ranges_encode() produces (the integer-grid constants as float bits; bit 0 symmetric, bit 1
// allow_one_sided); scale/offset are fp32.  Same arithmetic as pr_min_kernel + pr_apply_kernel.
__global__ void __launch_bounds__(RD_THREADS) pr_batched_kernel(const void* mn_base, const void* mx_base, int r_dt,
                                                                const long long* __restrict__ desc) {
  __shared__ float smem[64];
  __shared__ int s_one_sided;
  const long long* d = desc + (long long)blockIdx.x * 8;
  const unsigned long long start = (unsigned long long)d[0], len = (unsigned long long)d[1];
  float* scale = reinterpret_cast<float*>(d[2]);
  float* offset = reinterpret_cast<float*>(d[3]);
  const float int_min_abs = __int_as_float((int)(d[4] & 0xffffffffll)), int_max_abs = __int_as_float((int)(d[4] >> 32));
  const float neg_int_min = __int_as_float((int)(d[5] & 0xffffffffll)), steps = __int_as_float((int)(d[5] >> 32));
  const bool symmetric = (d[6] & 1) != 0, allow_one_sided = (d[6] & 2) != 0, rcp = (d[6] & 4) != 0;
  bool one_sided = false;
  if (symmetric && allow_one_sided) {
    float mn = INFINITY, dummy = 0.f;
    for (unsigned long long i = threadIdx.x; i < len; i += blockDim.x) mn = nan_min(mn, load_as_float(mn_base, r_dt, start + i));
    block_minmax(mn, dummy, smem);
    if (threadIdx.x == 0) s_one_sided = (mn >= 0.f) ? 1 : 0;
    __syncthreads();
    one_sided = s_one_sided != 0;
  }
  for (unsigned long long i = threadIdx.x; i < len; i += blockDim.x) {
    float mn = load_as_float(mn_base, r_dt, start + i);
    const float mx = load_as_float(mx_base, r_dt, start + i);
    if (symmetric && !one_sided) {
      scale[i] = nan_max(scalar_div(fabsf(mn), int_min_abs, rcp), scalar_div(fabsf(mx), int_max_abs, rcp));
      if (offset) offset[i] = 0.f;
      continue;
    }
    if (symmetric) mn = 0.f;
    float sc = scalar_div(__fsub_rn(mx, mn), steps, rcp);
    const float eps = 1.1920928955078125e-07f;
    sc = (sc != sc) ? sc : fmaxf(sc, eps);
    scale[i] = sc;
    if (offset) offset[i] = __fadd_rn(__fdiv_rn(mn, sc), neg_int_min);
  }
}

}  // namespace ffq

using namespace ffq;

static size_t seg_workspace(const Plan& plan, int ept) {
  if (!plan.row) return 0;
  const Segmentation sg = choose_segments(plan, ept);
  return sg.S > 1 ? (size_t)2 * plan.num_tiles * sg.S * sizeof(float) : 0;
}

static const unsigned int PR_MAX_PARTS = 1024;

extern "C" {

size_t ffq_workspace_bytes(int kind, const ffq_layout_t* layout, int data_dtype) {
  Plan plan;
  if (make_plan(layout, &plan) != FFQ_OK || plan.numel == 0) return 0;
  const int ept = ept_of(data_dtype);
  switch (kind) {
    case FFQ_WS_QUANTIZE_BWD:
    case FFQ_WS_MINMAX:
      return seg_workspace(plan, ept);
    case FFQ_WS_PARAMS_FOR_RANGE:
      return PR_MAX_PARTS * sizeof(float);
    case FFQ_WS_DYNAMIC_QUANTIZE:
      // [partials for min/max | param partials | tile_min | tile_max] (fp32 holds every data dtype exactly)
      return seg_workspace(plan, ept) + PR_MAX_PARTS * sizeof(float) + (size_t)2 * plan.num_tiles * 8 + 64;
  }
  return 0;
}

int ffq_quantize_bwd(const void* x, int x_dtype, const void* g, int g_dtype, void* dx, void* dscale,
                     void* doffset, const void* scale, int scale_dtype, const void* offset, int offset_dtype,
                     const ffq_layout_t* layout, double num_bits, void* workspace, size_t workspace_bytes,
                     void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!(x_dtype == FFQ_F32 || x_dtype == FFQ_F16 || x_dtype == FFQ_BF16) ||
      !(g_dtype == FFQ_F32 || g_dtype == FFQ_F16 || g_dtype == FFQ_BF16) ||
      !(scale_dtype == FFQ_F32 || scale_dtype == FFQ_F16 || scale_dtype == FFQ_BF16)) {
    set_error("quantize_bwd: data, grad and scale must be float32/float16/bfloat16 (got %s, %s, %s)",
              dt_name(x_dtype), dt_name(g_dtype), dt_name(scale_dtype));
    return FFQ_ERR_UNSUPPORTED;
  }
  if (offset == nullptr) offset_dtype = FFQ_NONE;
  else if (offset_dtype == FFQ_F64 || !(is_float_dt(offset_dtype) || is_int_dt(offset_dtype))) {
    set_error("quantize_bwd: unsupported offset dtype %s", dt_name(offset_dtype));
    return FFQ_ERR_UNSUPPORTED;
  }
  Plan plan;
  int rc = make_plan(layout, &plan);
  if (rc != FFQ_OK) return rc;
  if (plan.numel == 0) return FFQ_OK;

  BwdArgs a{};
  a.x = x; a.g = g; a.dx = dx; a.x_dt = x_dtype; a.g_dt = g_dtype;
  a.scale = scale; a.offset = offset; a.s_dt = scale_dtype; a.o_dt = offset_dtype;
  a.dscale = dscale; a.doffset = doffset;
  a.dsc_dt = scale_dtype; a.doff_dt = promote(scale_dtype, g_dtype);
  a.numel = plan.numel; a.tile_numel = plan.tile_numel; a.num_tiles = plan.num_tiles;
  const QParams qp = make_qparams(x_dtype, scale_dtype, offset_dtype, num_bits);
  a.bp.lo = qp.lo; a.bp.hi = qp.hi; a.bp.m_div = qp.m_div; a.bp.m_sub = qp.m_sub;
  a.bp.m_s = round_mode_of(scale_dtype);
  a.bp.m_sg = round_mode_of(promote(scale_dtype, g_dtype));
  // scale.new_tensor([lo]) : the bound rounded to the scale dtype
  auto to_s = [&](float v) {
    if (scale_dtype == FFQ_BF16) return __bfloat162float(__float2bfloat16_rn(v));
    if (scale_dtype == FFQ_F16) return __half2float(__float2half_rn(v));
    return v;
  };
  a.bp.lo_s = to_s(qp.lo); a.bp.hi_s = to_s(qp.hi);
  a.bp.has_offset = (offset != nullptr && doffset != nullptr) ? 1 : 0;
  a.gl = make_generic_layout(plan);
  a.S = 1; a.seg_len = plan.tile_numel;

  const bool fast = plan.row && x_dtype == g_dtype && a.bp.m_div == a.bp.m_sub && a.bp.m_sub == a.bp.m_s &&
                    a.bp.m_s == a.bp.m_sg && (offset == nullptr || offset_dtype == scale_dtype) &&
                    round_mode_of(scale_dtype) == a.bp.m_div &&
                    (scale_dtype == FFQ_F32 || scale_dtype == x_dtype);
  if (fast) {
    const bool vec_ok = aligned16(x) && aligned16(g) && aligned16(dx);
    const Segmentation sg = choose_segments(plan, ept_of(x_dtype));
    const bool group = vec_ok && group_lanes(plan, ept_of(x_dtype)) != 0;
    if (!group && sg.S > 1) {
      const size_t need = (size_t)2 * plan.num_tiles * sg.S * sizeof(float);
      if (workspace == nullptr || workspace_bytes < need) {
        set_error("quantize_bwd: workspace of %zu bytes required, %zu given", need, workspace_bytes);
        return FFQ_ERR_WORKSPACE;
      }
      a.S = sg.S; a.seg_len = sg.seg_len; a.part = static_cast<float*>(workspace);
    }
    if (!dispatch_bwd(a, plan, vec_ok, st)) { set_error("quantize_bwd: dtype dispatch failed"); return FFQ_ERR_UNSUPPORTED; }
    FFQ_LAUNCH_CHECK();
    if (a.S > 1) {
      const unsigned long long threads = a.num_tiles * 32;
      launch_pdl(bwd_finalize_kernel, dim3((unsigned int)((threads + RD_THREADS - 1) / RD_THREADS)), dim3(RD_THREADS), 0, st, a);
      FFQ_LAUNCH_CHECK();
    }
    return FFQ_OK;
  }
  if (plan.num_tiles > 0x7fffffffll) { set_error("quantize_bwd: too many tiles for the generic kernel"); return FFQ_ERR_UNSUPPORTED; }
  bwd_generic_kernel<<<(unsigned int)plan.num_tiles, 128, 0, st>>>(a);
  FFQ_LAUNCH_CHECK();
  return FFQ_OK;
}

int ffq_minmax(const void* x, int x_dtype, void* tile_min, void* tile_max, void* run_min, void* run_max,
               int run_dtype, int32_t* flags, const ffq_layout_t* layout, void* workspace, size_t workspace_bytes,
               void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (x_dtype == FFQ_F64 || !(is_float_dt(x_dtype) || is_int_dt(x_dtype))) {
    set_error("minmax: unsupported data dtype %s", dt_name(x_dtype));
    return FFQ_ERR_UNSUPPORTED;
  }
  Plan plan;
  int rc = make_plan(layout, &plan);
  if (rc != FFQ_OK) return rc;
  if (plan.numel == 0) { set_error("minmax: empty tensor"); return FFQ_ERR_INVALID; }
  MmArgs a{};
  a.x = x; a.x_dt = x_dtype; a.tile_min = tile_min; a.tile_max = tile_max;
  a.run_min = run_min; a.run_max = run_max; a.flags = flags;
  a.run_dt = (run_min || run_max) ? run_dtype : x_dtype;
  if ((run_min || run_max) && (run_dtype == FFQ_F64 || !(is_float_dt(run_dtype) || is_int_dt(run_dtype)))) {
    set_error("minmax: unsupported running-range dtype %s", dt_name(run_dtype));
    return FFQ_ERR_UNSUPPORTED;
  }
  a.numel = plan.numel; a.tile_numel = plan.tile_numel; a.num_tiles = plan.num_tiles;
  a.gl = make_generic_layout(plan);
  a.S = 1; a.seg_len = plan.tile_numel;
  const bool typed = x_dtype == FFQ_F32 || x_dtype == FFQ_F16 || x_dtype == FFQ_BF16;
  if (plan.row && typed) {
    const bool vec_ok = aligned16(x);
    const Segmentation sg = choose_segments(plan, ept_of(x_dtype));
    const bool group = vec_ok && group_lanes(plan, ept_of(x_dtype)) != 0;
    if (!group && sg.S > 1) {
      const size_t need = (size_t)2 * plan.num_tiles * sg.S * sizeof(float);
      if (workspace == nullptr || workspace_bytes < need) {
        set_error("minmax: workspace of %zu bytes required, %zu given", need, workspace_bytes);
        return FFQ_ERR_WORKSPACE;
      }
      a.S = sg.S; a.seg_len = sg.seg_len; a.part = static_cast<float*>(workspace);
    }
    switch (x_dtype) {
      case FFQ_F32: launch_mm_row<float>(a, plan, vec_ok, st); break;
      case FFQ_BF16: launch_mm_row<__nv_bfloat16>(a, plan, vec_ok, st); break;
      default: launch_mm_row<__half>(a, plan, vec_ok, st); break;
    }
    FFQ_LAUNCH_CHECK();
    if (a.S > 1) {
      const unsigned long long threads = a.num_tiles * 32;
      launch_pdl(mm_finalize_kernel, dim3((unsigned int)((threads + RD_THREADS - 1) / RD_THREADS)), dim3(RD_THREADS), 0, st, a);
      FFQ_LAUNCH_CHECK();
    }
    return FFQ_OK;
  }
  if (plan.num_tiles > 0x7fffffffll) { set_error("minmax: too many tiles for the generic kernel"); return FFQ_ERR_UNSUPPORTED; }
  mm_generic_kernel<<<(unsigned int)plan.num_tiles, 128, 0, st>>>(a);
  FFQ_LAUNCH_CHECK();
  return FFQ_OK;
}

int ffq_params_for_range(const void* min_range, const void* max_range, int range_dtype, int64_t n,
                         double num_bits, int symmetric, int allow_one_sided, int round_offset,
                         void* scale_out, int scale_dtype, void* offset_out, int offset_dtype,
                         void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n <= 0) return FFQ_OK;
  if (range_dtype == FFQ_F64 || !(is_float_dt(range_dtype) || is_int_dt(range_dtype)) ||
      !(scale_dtype == FFQ_F32 || scale_dtype == FFQ_F16 || scale_dtype == FFQ_BF16)) {
    set_error("params_for_range: unsupported dtypes (range %s, scale %s)", dt_name(range_dtype), dt_name(scale_dtype));
    return FFQ_ERR_UNSUPPORTED;
  }
  PrArgs a{};
  a.mn = min_range; a.mx = max_range; a.r_dt = range_dtype; a.n = (unsigned long long)n;
  const double lo = -pow(2.0, num_bits - 1.0);
  a.int_min_abs = (float)fabs(lo);
  a.int_max_abs = (float)fabs(-lo - 1.0);
  a.neg_int_min = (float)(-lo);
  a.steps = (float)(pow(2.0, num_bits) - 1.0);
  a.symmetric = symmetric; a.allow_one_sided = allow_one_sided & FFQ_FLAG_ALLOW_ONE_SIDED; a.round_offset = round_offset;
  a.rcp = (allow_one_sided & FFQ_FLAG_SCALAR_DIV_RECIPROCAL) ? 1 : 0;
  allow_one_sided &= FFQ_FLAG_ALLOW_ONE_SIDED;
  a.scale = scale_out; a.s_dt = scale_dtype; a.offset = offset_out; a.o_dt = offset_out ? offset_dtype : FFQ_NONE;
  a.part = static_cast<float*>(workspace);
  a.nparts = 0;
  if (symmetric && allow_one_sided) {
    unsigned long long nb = ((unsigned long long)n + RD_THREADS * 8 - 1) / (RD_THREADS * 8);
    if (nb > PR_MAX_PARTS) nb = PR_MAX_PARTS;
    if (nb < 1) nb = 1;
    if (workspace == nullptr || workspace_bytes < nb * sizeof(float)) {
      set_error("params_for_range: workspace of %zu bytes required", (size_t)(PR_MAX_PARTS * sizeof(float)));
      return FFQ_ERR_WORKSPACE;
    }
    a.nparts = (unsigned int)nb;
    launch_pdl(pr_min_kernel, dim3((unsigned int)nb), dim3(RD_THREADS), 0, st, a);
    FFQ_LAUNCH_CHECK();
  }
  launch_pdl(pr_apply_kernel, dim3((unsigned int)(((unsigned long long)n + RD_THREADS - 1) / RD_THREADS)), dim3(RD_THREADS), 0, st, a);
  FFQ_LAUNCH_CHECK();
  return FFQ_OK;
}

int ffq_dynamic_quantize(const void* x, int x_dtype, void* q, int q_dtype, float* scale_out, float* offset_out,
                         const ffq_layout_t* layout, double num_bits, int symmetric, int allow_one_sided,
                         void* workspace, size_t workspace_bytes, void* stream) {
  Plan plan;
  int rc = make_plan(layout, &plan);
  if (rc != FFQ_OK) return rc;
  if (plan.numel == 0) { set_error("Cannot dynamically quantize an empty tensor"); return FFQ_ERR_INVALID; }
  const size_t need = ffq_workspace_bytes(FFQ_WS_DYNAMIC_QUANTIZE, layout, x_dtype);
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("dynamic_quantize: workspace of %zu bytes required, %zu given", need, workspace_bytes);
    return FFQ_ERR_WORKSPACE;
  }
  const int ept = ept_of(x_dtype);
  char* ws = static_cast<char*>(workspace);
  const size_t seg_bytes = (seg_workspace(plan, ept) + 15) / 16 * 16;
  float* pr_part = reinterpret_cast<float*>(ws + seg_bytes);
  // min/max are kept in the data dtype by the reference and only then cast to fp32
  // (range.py:90); the data dtype embeds exactly in fp32, so fp32 scratch is equivalent
  // only if we store through the data dtype -- ffq_minmax does (tile_min has dtype x_dtype).
  char* tmin = ws + seg_bytes + PR_MAX_PARTS * sizeof(float);
  char* tmax = tmin + (size_t)plan.num_tiles * 8;
  rc = ffq_minmax(x, x_dtype, tmin, tmax, nullptr, nullptr, FFQ_NONE, nullptr, layout, ws, seg_bytes, stream);
  if (rc != FFQ_OK) return rc;
  rc = ffq_params_for_range(tmin, tmax, x_dtype, plan.num_tiles, num_bits, symmetric, allow_one_sided, 1,
                            scale_out, FFQ_F32, offset_out, FFQ_F32, pr_part, PR_MAX_PARTS * sizeof(float), stream);
  if (rc != FFQ_OK) return rc;
  return ffq_quantize(x, x_dtype, q, q_dtype, scale_out, FFQ_F32, offset_out, FFQ_F32, layout, num_bits, stream);
}

void ffq_params_for_ranges_encode(double num_bits, int symmetric, int allow_one_sided, int64_t words[3]) {
  const double lo = -pow(2.0, num_bits - 1.0);
  const float f[4] = {(float)fabs(lo), (float)fabs(-lo - 1.0), (float)(-lo), (float)(pow(2.0, num_bits) - 1.0)};
  uint32_t u[4];
  memcpy(u, f, sizeof(u));
  words[0] = (int64_t)(((uint64_t)u[1] << 32) | u[0]);
  words[1] = (int64_t)(((uint64_t)u[3] << 32) | u[2]);
  words[2] = (symmetric ? 1 : 0) | ((allow_one_sided & FFQ_FLAG_ALLOW_ONE_SIDED) ? 2 : 0) |
             ((allow_one_sided & FFQ_FLAG_SCALAR_DIV_RECIPROCAL) ? 4 : 0);
}

int ffq_params_for_ranges_batched(const void* min_base, const void* max_base, int range_dtype, const int64_t* desc_dev,
                                  int64_t num_quantizers, void* stream) {
  if (num_quantizers <= 0) return FFQ_OK;
  if (range_dtype == FFQ_F64 || !(is_float_dt(range_dtype) || is_int_dt(range_dtype))) {
    set_error("params_for_ranges_batched: unsupported range dtype %s", dt_name(range_dtype));
    return FFQ_ERR_UNSUPPORTED;
  }
  if (num_quantizers > 0x7fffffffll || desc_dev == nullptr) { set_error("params_for_ranges_batched: bad descriptor table"); return FFQ_ERR_INVALID; }
  pr_batched_kernel<<<(unsigned int)num_quantizers, RD_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      min_base, max_base, range_dtype, reinterpret_cast<const long long*>(desc_dev));
  FFQ_LAUNCH_CHECK();
  return FFQ_OK;
}

}  // extern "C"

// ffq_elementwise.cu -- quantize / dequantize / fused fake-quantize (a1, a2).
//
// HBM-bound streaming kernels.  Fast path ("row" layouts: every tile is one contiguous run,
// which covers per-tensor, per-channel(0) and per-group weights): 16-byte vector loads with
// L1 no-allocate, 4 vectors in flight per thread, one parameter fetch and ONE reciprocal per
// vector; the IEEE-exact quotient costs 3 FMA-pipe instructions per element (shared_div in
// ffq_common.cuh).  Algorithmic traffic per element: quantize s+c, dequantize c+s,
// fake-quantize 2s bytes (s = data bytes, c = code bytes).  Everything else (strided tiles,
// integer inputs, mixed promotion chains, unaligned pointers) runs on a scalar generic kernel:
// correct for any rank<=8 tiling and any dtype, not tuned.
#include <type_traits>

#include "ffq_common.cuh"

namespace ffq {

enum EwOp : int { OP_QUANT = 0, OP_DEQUANT = 1, OP_FAKEQUANT = 2 };

struct EwArgs {
  const void* in;
  void* out;
  void* codes;      // fake-quant only, optional
  int in_dt, out_dt, codes_dt;
  const void* scale;
  const void* offset;
  int s_dt, o_dt;
  unsigned long long numel;
  unsigned long long tile_numel;
  FastDiv tdiv;     // division by tile_numel (fast kernel: numel < 2^31)
  QParams qp;
  DParams dp;
  int q_rt_dt;      // fake-quant: dtype the codes take between quantize and dequantize
  int rt_needed;    // fake-quant: that round trip can change the value (integer wrap)
  int q_is_int;     // fake-quant: integer code dtype (drops the sign of a zero)
  int tiles_aligned; // tile_numel is a multiple of the vector width: no vector straddles tiles
  int sat8;          // bounds are exactly [-128, 127]: integer codes use the saturating conversion
  unsigned long long total_segs;   // tile-streaming kernel: number of (tile, segment) work units
  unsigned int segs_per_tile;
  unsigned int seg_vecs;           // vectors per segment (multiple of 32 * EW_UNROLL)
  GenericLayout gl; // generic kernel only
};

__device__ __forceinline__ float code_roundtrip(float q, int dt) {
  switch (dt) {
    case FFQ_BF16: return __bfloat162float(__float2bfloat16_rn(q));
    case FFQ_F16: return __half2float(__float2half_rn(q));
    case FFQ_I8: return (float)(int8_t)__float2int_rz(q);
    case FFQ_U8: return (float)(uint8_t)__float2int_rz(q);
    case FFQ_I16: return (float)(int16_t)__float2int_rz(q);
    case FFQ_I32: return __int2float_rn(__float2int_rz(q));
    default: return q;
  }
}

// runtime-mode scalar arithmetic (generic kernel, and the slow path of the fast kernel)
template <int OP>
__device__ __forceinline__ float ew_apply(float v, float s, float o, const EwArgs& a, float* code_out) {
  if constexpr (OP == OP_QUANT) {
    return quantize_value(v, s, o, a.qp);
  } else if constexpr (OP == OP_DEQUANT) {
    return dequantize_value(v, s, o, a.dp);
  } else {
    float q = quantize_value(v, s, o, a.qp);
    q = code_roundtrip(q, a.q_rt_dt);
    *code_out = q;
    return dequantize_value(q, s, o, a.dp);
  }
}

constexpr int EW_THREADS = 256;
constexpr int EW_UNROLL = 4;

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

template <int RM> struct ParamT { using type = float; };
template <> struct ParamT<RM_BF16> { using type = __nv_bfloat16; };
template <> struct ParamT<RM_F16> { using type = __half; };

// Fast kernel.  Preconditions (checked by the host): row layout, numel < 2^31, 16-byte aligned
// pointers, one promoted dtype RM for the whole chain, scale/offset stored in that dtype.
template <typename InT, typename OutT> struct EwEpt {
  static constexpr int value = 16 / (sizeof(InT) > sizeof(OutT) ? sizeof(InT) : sizeof(OutT));
};

template <int OP, typename InT, typename OutT, int RM>
__global__ void __launch_bounds__(EW_THREADS, 4) ew_row_kernel(const EwArgs a) {
  pdl_wait();                    // programmatic dependent launch: no-ops unless launched that way
  pdl_trigger();
  constexpr int EPT = EwEpt<InT, OutT>::value;
  constexpr bool FLOAT_OUT = std::is_same<OutT, float>::value || std::is_same<OutT, __half>::value ||
                             std::is_same<OutT, __nv_bfloat16>::value;
  using PT = typename ParamT<RM>::type;
  const InT* __restrict__ in = static_cast<const InT*>(a.in);
  OutT* __restrict__ out = static_cast<OutT*>(a.out);
  const PT* __restrict__ scale = static_cast<const PT*>(a.scale);
  const PT* __restrict__ offset = static_cast<const PT*>(a.offset);
  const unsigned int numel = (unsigned int)a.numel;
  const unsigned int nvec = numel / EPT;
  const unsigned int vbase = blockIdx.x * (EW_THREADS * EW_UNROLL) + threadIdx.x;
  const float lo = a.qp.lo, hi = a.qp.hi;

  Vec<InT, EPT> xin[EW_UNROLL];
#pragma unroll
  for (int u = 0; u < EW_UNROLL; ++u) {
    const unsigned int v = vbase + u * EW_THREADS;
    if (v < nvec) xin[u] = ld_stream<InT, EPT>(in + (size_t)v * EPT);
  }
  // parameters of every vector are requested right behind the data loads, so their latency
  // overlaps the stream instead of stalling the arithmetic
  unsigned int pv[EW_UNROLL];
  float sv[EW_UNROLL], ov[EW_UNROLL];
#pragma unroll
  for (int u = 0; u < EW_UNROLL; ++u) {
    const unsigned int v = vbase + u * EW_THREADS;
    pv[u] = fast_div((v < nvec ? v : 0u) * EPT, a.tdiv);
    sv[u] = Elem<PT>::to_f(scale[pv[u]]);
    ov[u] = offset ? Elem<PT>::to_f(offset[pv[u]]) : 0.f;
  }
  float s = 1.f, o = 0.f;
  SharedRcp k{};
#pragma unroll
  for (int u = 0; u < EW_UNROLL; ++u) {
    const unsigned int v = vbase + u * EW_THREADS;
    if (v >= nvec) continue;
    const unsigned int e0 = v * EPT;
    const unsigned int p0 = pv[u];
    if (u == 0 || pv[u] != pv[u > 0 ? u - 1 : 0]) {   // the reciprocal is formed once per tile change
      s = sv[u];
      o = rintf(ov[u]);
      if constexpr (OP != OP_DEQUANT) k = make_shared_rcp(s);
    }
    float x[EPT], y[EPT], c[EPT];
    unpack<InT, EPT>(xin[u], x);
    const bool straddle = !a.tiles_aligned && fast_div(e0 + (EPT - 1), a.tdiv) != p0;
    if (!straddle) {
      if constexpr (OP == OP_DEQUANT) {
#pragma unroll
        for (int i = 0; i < EPT; ++i) y[i] = rndc<RM>(__fmul_rn(rndc<RM>(__fadd_rn(x[i], o)), s));
      } else {
        bool ok = k.ok;
        if (a.sat8 && !FLOAT_OUT) {
          // 8-bit codes into an integer container: rint + clamp(-128, 127) is ONE saturating
          // conversion (cvt.rni.sat.s8.f32).  NaN would convert to 0, so NaN inputs fail the
          // magnitude guard below and take the exact path.
#pragma unroll
          for (int i = 0; i < EPT; ++i) {
            float t = rndc<RM>(shared_div<false>(x[i], k, ok));
            t = rndc<RM>(__fsub_rn(t, o));
            int ci;
            asm("cvt.rni.sat.s8.f32 %0, %1;" : "=r"(ci) : "f"(t));
            c[i] = (float)ci;
          }
        } else {
#pragma unroll
          for (int i = 0; i < EPT; ++i) {
            // float codes expose the sign of a zero quotient: take the strict guard there
            float t = rndc<RM>(shared_div<(OP == OP_QUANT) && FLOAT_OUT>(x[i], k, ok));
            t = rndc<RM>(__fsub_rn(t, o));
            c[i] = nan_clamp(rintf(t), lo, hi);
          }
        }
        if (!ok) {   // rare: scale or quotient outside the proven box -> plain IEEE division
#pragma unroll
          for (int i = 0; i < EPT; ++i) {
            float t = rndc<RM>(__fdiv_rn(x[i], s));
            t = rndc<RM>(__fsub_rn(t, o));
            c[i] = nan_clamp(rintf(t), lo, hi);
          }
        }
        if constexpr (OP == OP_FAKEQUANT) {
          if (a.rt_needed) {          // integer code dtype narrower than num_bits: wraps
#pragma unroll
            for (int i = 0; i < EPT; ++i) c[i] = code_roundtrip(c[i], a.q_rt_dt);
          } else if (a.q_is_int) {    // integer codes cannot carry -0
#pragma unroll
            for (int i = 0; i < EPT; ++i) c[i] = __fadd_rn(c[i], 0.0f);
          }
#pragma unroll
          for (int i = 0; i < EPT; ++i) y[i] = rndc<RM>(__fmul_rn(rndc<RM>(__fadd_rn(c[i], o)), s));
        } else {
#pragma unroll
          for (int i = 0; i < EPT; ++i) y[i] = c[i];
        }
      }
    } else {   // a vector straddling tiles (tile_numel not a multiple of the vector width)
      unsigned int pp = p0;
      float ss = s, oo = o;
#pragma unroll
      for (int i = 0; i < EPT; ++i) {
        const unsigned int p = fast_div(e0 + i, a.tdiv);
        if (p != pp) { pp = p; ss = Elem<PT>::to_f(scale[p]); oo = offset ? rintf(Elem<PT>::to_f(offset[p])) : 0.f; }
        y[i] = ew_apply<OP>(x[i], ss, oo, a, &c[i]);
      }
    }
    Vec<OutT, EPT> yo;
    pack<OutT, EPT>(y, yo);
    st_vec<OutT, EPT>(out + e0, yo);
    if constexpr (OP == OP_FAKEQUANT) {
      if (a.codes) {
        const bool float_codes = a.codes_dt == FFQ_F32 || a.codes_dt == FFQ_F16 || a.codes_dt == FFQ_BF16;
#pragma unroll
        for (int i = 0; i < EPT; ++i) {
          float ci = c[i];
          // float codes expose the sign of a zero quotient, which the shared-reciprocal path does
          // not guarantee for zero / sub-2^-90 dividends: recompute those few exactly
          if (float_codes && !straddle && !(fabsf(x[i]) >= 0x1p-90f)) ew_apply<OP_FAKEQUANT>(x[i], s, o, a, &ci);
          store_from_float(a.codes, a.codes_dt, e0 + i, ci);
        }
      }
    }
  }
  // scalar tail (numel % EPT elements), done by the last block's first threads
  if (blockIdx.x == gridDim.x - 1) {
    const unsigned int e = nvec * EPT + threadIdx.x;
    if (e < numel) {
      const unsigned int p = fast_div(e, a.tdiv);
      const float ss = Elem<PT>::to_f(scale[p]);
      const float oo = offset ? rintf(Elem<PT>::to_f(offset[p])) : 0.f;
      float c;
      const float r = ew_apply<OP>(Elem<InT>::to_f(in[e]), ss, oo, a, &c);
      out[e] = Elem<OutT>::from_f(r);
      if constexpr (OP == OP_FAKEQUANT) {
        if (a.codes) store_from_float(a.codes, a.codes_dt, e, c);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Tile-streaming kernel: the tuned path for tiles of at least 32 vectors whose length is a multiple
// of the vector width (weight rows, whole activations).  One warp owns one (tile, segment): the
// parameters are fetched and the reciprocal is formed ONCE per warp, then the warp streams up to
// SEG_VECS 16-byte vectors with UNROLL loads in flight per lane.  No per-vector index division, no
// parameter reload inside the stream, 64-bit addressing only in the prologue.
// ------------------------------------------------------------------------------------------------
constexpr int SEG_VECS_MAX = 1024;   // per-warp segment: up to 16 KB of input, shrunk for small tensors

// exact recomputation of one vector (scale or quotient outside the guard of shared_div)
template <int OP, int RM, int EPT>
__device__ __forceinline__ void ew_vector_exact(const float (&x)[EPT], float (&y)[EPT], float s, float o, float lo, float hi) {
#pragma unroll
  for (int i = 0; i < EPT; ++i) {
    float t = rndc<RM>(__fdiv_rn(x[i], s));
    t = rndc<RM>(__fsub_rn(t, o));
    const float c = nan_clamp(rintf(t), lo, hi);
    y[i] = (OP == OP_FAKEQUANT) ? rndc<RM>(__fmul_rn(rndc<RM>(__fadd_rn(c, o)), s)) : c;
  }
}

template <int OP, typename InT, typename OutT, int RM>
__global__ void __launch_bounds__(EW_THREADS, 4) ew_tile_kernel(const EwArgs a) {
  pdl_wait();                    // programmatic dependent launch: no-ops unless launched that way
  pdl_trigger();
  constexpr int EPT = EwEpt<InT, OutT>::value;
  constexpr int U = EW_UNROLL;
  constexpr bool FLOAT_OUT = std::is_same<OutT, float>::value || std::is_same<OutT, __half>::value ||
                             std::is_same<OutT, __nv_bfloat16>::value;
  using PT = typename ParamT<RM>::type;
  const unsigned int lane = threadIdx.x & 31;
  const unsigned long long seg_id = (unsigned long long)blockIdx.x * (EW_THREADS / 32) + (threadIdx.x >> 5);
  if (seg_id >= a.total_segs) return;
  const unsigned long long tile = seg_id / a.segs_per_tile;
  const unsigned int seg = (unsigned int)(seg_id - tile * a.segs_per_tile);
  const unsigned int tvec = (unsigned int)(a.tile_numel / EPT);
  const unsigned int vec0 = seg * a.seg_vecs;
  const unsigned int nv = (tvec - vec0) < a.seg_vecs ? (tvec - vec0) : a.seg_vecs;
  const unsigned long long base = tile * a.tile_numel + (unsigned long long)vec0 * EPT;
  const InT* __restrict__ in = static_cast<const InT*>(a.in) + base;
  OutT* __restrict__ out = static_cast<OutT*>(a.out) + base;

  const float s = Elem<PT>::to_f(static_cast<const PT*>(a.scale)[tile]);
  const float o = a.offset ? rintf(Elem<PT>::to_f(static_cast<const PT*>(a.offset)[tile])) : 0.f;
  const float lo = a.qp.lo, hi = a.qp.hi;
  SharedRcp k{};
  if constexpr (OP != OP_DEQUANT) k = make_shared_rcp(s);
  // integer codes of exactly 8 bits: rint + clamp is one saturating conversion (NaN fails the guard)
  const bool sat8 = a.sat8 != 0;
  const bool int_zero = a.q_is_int != 0;       // fake-quant through integer codes: -0 becomes +0

  for (unsigned int j0 = lane; j0 < nv; j0 += 32 * U) {
    Vec<InT, EPT> xin[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned int j = j0 + u * 32;
      if (j < nv) xin[u] = ld_stream<InT, EPT>(in + (size_t)j * EPT);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned int j = j0 + u * 32;
      if (j >= nv) continue;
      float x[EPT], y[EPT];
      unpack<InT, EPT>(xin[u], x);
      if constexpr (OP == OP_DEQUANT) {
#pragma unroll
        for (int i = 0; i < EPT; ++i) y[i] = rndc<RM>(__fmul_rn(rndc<RM>(__fadd_rn(x[i], o)), s));
      } else {
        float t[EPT];
        float amax = 0.f, amin = INFINITY;
#pragma unroll
        for (int i = 0; i < EPT; ++i) {
          const float q0 = __fmul_rn(x[i], k.r);
          const float e = __fmaf_rn(-k.s, q0, x[i]);
          const float q = __fmaf_rn(k.r, e, q0);
          amax = nan_max(amax, fabsf(q));                    // NaN-propagating: a NaN quotient fails the guard
          if constexpr (OP == OP_QUANT && FLOAT_OUT) amin = fminf(amin, fabsf(q));
          t[i] = rndc<RM>(__fsub_rn(rndc<RM>(q), o));
        }
        bool ok = k.ok && (amax <= 0x1p60f);
        if constexpr (OP == OP_QUANT && FLOAT_OUT) ok = ok && (amin >= 0x1p-50f);   // float codes expose the sign of zero
        if (ok) {
          if (sat8 && (OP == OP_FAKEQUANT || !FLOAT_OUT)) {
#pragma unroll
            for (int i = 0; i < EPT; ++i) {
              int ci;
              asm("cvt.rni.sat.s8.f32 %0, %1;" : "=r"(ci) : "f"(t[i]));
              const float c = (float)ci;
              y[i] = (OP == OP_FAKEQUANT) ? rndc<RM>(__fmul_rn(rndc<RM>(__fadd_rn(c, o)), s)) : c;
            }
          } else {
#pragma unroll
            for (int i = 0; i < EPT; ++i) {
              float c = nan_clamp(rintf(t[i]), lo, hi);
              if constexpr (OP == OP_FAKEQUANT) {
                if (int_zero) c = __fadd_rn(c, 0.0f);
                y[i] = rndc<RM>(__fmul_rn(rndc<RM>(__fadd_rn(c, o)), s));
              } else {
                y[i] = c;
              }
            }
          }
        } else {
          ew_vector_exact<OP, RM, EPT>(x, y, s, o, lo, hi);
        }
      }
      Vec<OutT, EPT> yo;
      pack<OutT, EPT>(y, yo);
      st_vec<OutT, EPT>(out + (size_t)j * EPT, yo);
    }
  }
}

// tile index of a linear element index under a collapsed layout of any rank
__device__ __forceinline__ unsigned long long tile_of(unsigned long long e, const GenericLayout& g) {
  unsigned long long p = 0;
#pragma unroll 1
  for (int d = 0; d < g.rank; ++d) {
    const unsigned long long c = e / g.stride[d];
    e -= c * g.stride[d];
    p = p * g.grid[d] + c / g.tile[d];
  }
  return p;
}

template <int OP>
__global__ void __launch_bounds__(256) ew_generic_kernel(const EwArgs a) {
  const unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.numel) return;
  const unsigned long long p = a.gl.rank <= 1 ? e / a.tile_numel : tile_of(e, a.gl);
  const float s = load_as_float(a.scale, a.s_dt, p);
  const float o = load_offset(a.offset, a.o_dt, p);
  float c;
  const float r = ew_apply<OP>(load_as_float(a.in, a.in_dt, e), s, o, a, &c);
  store_from_float(a.out, a.out_dt, e, r);
  if constexpr (OP == OP_FAKEQUANT) {
    if (a.codes) store_from_float(a.codes, a.codes_dt, e, c);
  }
}

template <int OP, typename InT, typename OutT, int RM>
static void launch_tile(const EwArgs& a, cudaStream_t st) {
  const unsigned long long blocks = (a.total_segs + EW_THREADS / 32 - 1) / (EW_THREADS / 32);
  launch_pdl(ew_tile_kernel<OP, InT, OutT, RM>, dim3((unsigned int)blocks), dim3(EW_THREADS), 0, st, a);
  count_launch();
}

template <int OP, typename InT, typename OutT, int RM>
static void launch_row(const EwArgs& a, cudaStream_t st) {
  constexpr int EPT = EwEpt<InT, OutT>::value;
  if (a.total_segs) { launch_tile<OP, InT, OutT, RM>(a, st); return; }
  const unsigned long long nvec = a.numel / EPT;
  unsigned long long blocks = (nvec + EW_THREADS * EW_UNROLL - 1) / (EW_THREADS * EW_UNROLL);
  if (blocks == 0) blocks = 1;
  launch_pdl(ew_row_kernel<OP, InT, OutT, RM>, dim3((unsigned int)blocks), dim3(EW_THREADS), 0, st, a);
  count_launch();
}

// Instantiated combinations: every float in/out pair with an fp32 chain; bf16->bf16 with a bf16
// chain and f16->f16 with an f16 chain (model.to(bfloat16) after the quantizers were created);
// integer code outputs (quantize) / inputs (dequantize) with an fp32 chain.
template <int OP, typename InT>
static bool dispatch_out(const EwArgs& a, int rm, cudaStream_t st) {
  if (rm == RM_F32) {
    switch (a.out_dt) {
      case FFQ_F32: launch_row<OP, InT, float, RM_F32>(a, st); return true;
      case FFQ_BF16: launch_row<OP, InT, __nv_bfloat16, RM_F32>(a, st); return true;
      case FFQ_F16: launch_row<OP, InT, __half, RM_F32>(a, st); return true;
      case FFQ_I8: if constexpr (OP == OP_QUANT) { launch_row<OP, InT, int8_t, RM_F32>(a, st); return true; } break;
      case FFQ_I16: if constexpr (OP == OP_QUANT) { launch_row<OP, InT, int16_t, RM_F32>(a, st); return true; } break;
      case FFQ_I32: if constexpr (OP == OP_QUANT) { launch_row<OP, InT, int32_t, RM_F32>(a, st); return true; } break;
    }
    return false;
  }
  if constexpr (std::is_same<InT, __nv_bfloat16>::value) {
    if (rm == RM_BF16 && a.out_dt == FFQ_BF16) { launch_row<OP, InT, __nv_bfloat16, RM_BF16>(a, st); return true; }
  }
  if constexpr (std::is_same<InT, __half>::value) {
    if (rm == RM_F16 && a.out_dt == FFQ_F16) { launch_row<OP, InT, __half, RM_F16>(a, st); return true; }
  }
  return false;
}

template <int OP>
static bool dispatch_row(const EwArgs& a, int rm, cudaStream_t st) {
  switch (a.in_dt) {
    case FFQ_F32: return dispatch_out<OP, float>(a, rm, st);
    case FFQ_BF16: return dispatch_out<OP, __nv_bfloat16>(a, rm, st);
    case FFQ_F16: return dispatch_out<OP, __half>(a, rm, st);
    case FFQ_I8: if constexpr (OP == OP_DEQUANT) return dispatch_out<OP, int8_t>(a, rm, st); break;
    case FFQ_I16: if constexpr (OP == OP_DEQUANT) return dispatch_out<OP, int16_t>(a, rm, st); break;
    case FFQ_I32: if constexpr (OP == OP_DEQUANT) return dispatch_out<OP, int32_t>(a, rm, st); break;
  }
  return false;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int check_dt(int dt, const char* what) {
  if (dt == FFQ_F64) {
    set_error("%s: float64 is not implemented by the B200 backend", what);
    return FFQ_ERR_UNSUPPORTED;
  }
  if (!(is_float_dt(dt) || is_int_dt(dt))) {
    set_error("%s: unknown dtype tag %d", what, dt);
    return FFQ_ERR_INVALID;
  }
  return FFQ_OK;
}

QParams make_qparams(int x_dt, int s_dt, int o_dt, double num_bits) {
  QParams qp;
  // num_bits arrives as a float (8.0); fractional widths follow the same formula
  const double lo_exact = -pow(2.0, num_bits - 1.0);
  qp.lo = (float)lo_exact;
  qp.hi = (float)(-lo_exact - 1.0);
  const int p_div = promote(x_dt, s_dt);
  const int p_sub = (o_dt == FFQ_NONE) ? p_div : promote(p_div, o_dt);  // zeros_like(scale) when absent
  qp.m_div = round_mode_of(p_div);
  qp.m_sub = round_mode_of(p_sub);
  return qp;
}

DParams make_dparams(int q_dt, int s_dt, int o_dt) {
  DParams dp;
  const int off_dt = (o_dt == FFQ_NONE) ? s_dt : o_dt;  // zeros_like(scale)
  const int p_add = promote(q_dt, off_dt);
  const int p_mul = promote(p_add, s_dt);
  dp.m_add = round_mode_of(p_add);
  dp.m_mul = round_mode_of(p_mul);
  dp.int_add_bits = is_int_dt(p_add) ? dt_size(p_add) * 8 : 0;
  return dp;
}

static int param_rm(int dt) { return dt == FFQ_F32 ? RM_F32 : dt == FFQ_BF16 ? RM_BF16 : dt == FFQ_F16 ? RM_F16 : -1; }

static int run_elementwise(int op, EwArgs& a, const ffq_layout_t* layout, double num_bits, cudaStream_t st) {
  Plan plan;
  int rc = make_plan(layout, &plan);
  if (rc != FFQ_OK) return rc;
  if (plan.numel == 0) return FFQ_OK;
  a.numel = (unsigned long long)plan.numel;
  a.tile_numel = (unsigned long long)plan.tile_numel;
  a.gl = make_generic_layout(plan);
  // an integer code dtype narrower than num_bits wraps on the round trip (the reference's
  // can_support_bitwidth admits iinfo.bits + 2)
  a.rt_needed = (op == OP_FAKEQUANT && is_int_dt(a.q_rt_dt) && num_bits > dt_size(a.q_rt_dt) * 8) ? 1 : 0;
  a.q_is_int = (op == OP_FAKEQUANT && is_int_dt(a.q_rt_dt)) ? 1 : 0;
  a.sat8 = (op != OP_DEQUANT && a.qp.lo == -128.f && a.qp.hi == 127.f &&
            (op == OP_QUANT ? is_int_dt(a.out_dt) : (is_int_dt(a.q_rt_dt) && !a.rt_needed))) ? 1 : 0;

  // the fast kernel handles chains with a single promoted dtype whose parameters are stored in it
  int rm = -1;
  if (op == OP_QUANT) {
    if (a.qp.m_div == a.qp.m_sub) rm = a.qp.m_div;
  } else if (op == OP_DEQUANT) {
    if (a.dp.m_add == a.dp.m_mul && a.dp.int_add_bits == 0) rm = a.dp.m_add;
  } else {
    if (a.qp.m_div == a.qp.m_sub && a.dp.m_add == a.dp.m_mul && a.qp.m_div == a.dp.m_add && a.dp.int_add_bits == 0)
      rm = a.qp.m_div;
  }
  if (rm >= 0 && (param_rm(a.s_dt) != rm || (a.offset && param_rm(a.o_dt) != rm))) rm = -1;
  // dequantize into an integer dtype (codes of integer data: dequantize_dtype = the data dtype, affine/function.py:137):
  // the generic kernel casts like aten does, (q + o) * s truncated toward zero
  if (op == OP_DEQUANT && !is_float_dt(a.out_dt)) rm = -1;

  const int in_sz = dt_size(a.in_dt), out_sz = dt_size(a.out_dt), code_sz = a.codes ? dt_size(a.codes_dt) : 0;
  const int p_sz = dt_size(a.s_dt);
  a.total_segs = 0;
  if (rm >= 0 && plan.row && aligned16(a.in) && aligned16(a.out) && in_sz > 0 && in_sz <= 4 &&
      !(op == OP_FAKEQUANT && a.codes != nullptr) && !a.rt_needed) {
    // tile-streaming kernel: tiles of >= 128 whole vectors (64-bit addressing in its prologue: any size)
    const int ept = 16 / (in_sz > out_sz ? in_sz : out_sz);
    const unsigned long long T = a.tile_numel;
    if (T % ept == 0 && T / ept >= 128 && T / ept < (1ull << 31)) {
      EwArgs c = a;
      // enough warps to fill the machine (>= 32 per SM) while keeping >= one full unrolled step per warp
      const unsigned long long total_vecs = a.numel / ept, quantum = 32ull * EW_UNROLL;
      unsigned long long sv = total_vecs / (32ull * sm_count());
      sv = sv / quantum * quantum;
      if (sv < quantum) sv = quantum;
      if (sv > (unsigned long long)SEG_VECS_MAX) sv = SEG_VECS_MAX;
      // equal segments within a tile (a 1024-vector row is 2 x 512, not 768 + 256)
      const unsigned long long tvec = T / ept;
      const unsigned long long nseg = (tvec + sv - 1) / sv;
      sv = ((tvec + nseg - 1) / nseg + quantum - 1) / quantum * quantum;
      c.seg_vecs = (unsigned int)sv;
      c.segs_per_tile = (unsigned int)((tvec + sv - 1) / sv);
      c.total_segs = (unsigned long long)plan.num_tiles * c.segs_per_tile;
      if (c.total_segs / (EW_THREADS / 32) < 0x7fffffffull) {
        bool ok;
        if (op == OP_QUANT) ok = dispatch_row<OP_QUANT>(c, rm, st);
        else if (op == OP_DEQUANT) ok = dispatch_row<OP_DEQUANT>(c, rm, st);
        else ok = dispatch_row<OP_FAKEQUANT>(c, rm, st);
        if (ok) {
          cudaError_t e = cudaPeekAtLastError();
          if (e != cudaSuccess) { cudaGetLastError(); set_error("kernel launch failed: %s", cudaGetErrorString(e)); return FFQ_ERR_CUDA; }
          return FFQ_OK;
        }
      }
    }
  }
  if (rm >= 0 && plan.row && aligned16(a.in) && aligned16(a.out) && in_sz > 0 && in_sz <= 4) {
    // launches of < 2^31 elements each, cut at tile boundaries (a single launch in practice)
    const int ept = 16 / (in_sz > out_sz ? in_sz : out_sz);
    const unsigned long long T = a.tile_numel, LIM = 1ull << 31;
    unsigned long long chunk;        // elements per launch
    unsigned long long chunk_tiles;  // parameter advance per launch (0: all launches inside one tile)
    if (a.numel < LIM) { chunk = a.numel; chunk_tiles = 0; }
    else if (T * 16 <= LIM) { chunk = (LIM / (T * 16)) * (T * 16); chunk_tiles = chunk / T; }
    else { chunk = 0; chunk_tiles = 0; }   // giant tiles in a giant tensor: handled below
    if (chunk) {
      bool ok = true;
      const EwArgs base = a;
      for (unsigned long long off = 0; off < base.numel && ok; off += chunk) {
        EwArgs c = base;
        c.numel = (base.numel - off < chunk) ? base.numel - off : chunk;
        c.in = static_cast<const char*>(base.in) + off * in_sz;
        c.out = static_cast<char*>(base.out) + off * out_sz;
        if (base.codes) c.codes = static_cast<char*>(base.codes) + off * code_sz;
        const unsigned long long tile0 = chunk_tiles ? (off / chunk) * chunk_tiles : 0;
        c.scale = static_cast<const char*>(base.scale) + tile0 * p_sz;
        if (base.offset) c.offset = static_cast<const char*>(base.offset) + tile0 * p_sz;
        c.tdiv = make_fast_div((unsigned int)(T < LIM ? T : LIM - 1));
        if (T >= LIM) c.tdiv = make_fast_div(0x80000000u);   // one tile: every index maps to parameter 0
        c.tiles_aligned = (T % ept == 0) ? 1 : 0;
        if (op == OP_QUANT) ok = dispatch_row<OP_QUANT>(c, rm, st);
        else if (op == OP_DEQUANT) ok = dispatch_row<OP_DEQUANT>(c, rm, st);
        else ok = dispatch_row<OP_FAKEQUANT>(c, rm, st);
        if (!ok && off != 0) { set_error("elementwise: dtype dispatch changed between chunks"); return FFQ_ERR_CUDA; }
      }
      if (ok) {
        cudaError_t e = cudaPeekAtLastError();
        if (e != cudaSuccess) { cudaGetLastError(); set_error("kernel launch failed: %s", cudaGetErrorString(e)); return FFQ_ERR_CUDA; }
        return FFQ_OK;
      }
    }
  }
  const unsigned long long blocks = (a.numel + 255) / 256;
  if (blocks > 0x7fffffffull) {
    set_error("tensor too large for the generic elementwise kernel");
    return FFQ_ERR_UNSUPPORTED;
  }
  if (op == OP_QUANT) ew_generic_kernel<OP_QUANT><<<(unsigned int)blocks, 256, 0, st>>>(a);
  else if (op == OP_DEQUANT) ew_generic_kernel<OP_DEQUANT><<<(unsigned int)blocks, 256, 0, st>>>(a);
  else ew_generic_kernel<OP_FAKEQUANT><<<(unsigned int)blocks, 256, 0, st>>>(a);
  FFQ_LAUNCH_CHECK();
  return FFQ_OK;
}

}  // namespace ffq

using namespace ffq;

extern "C" {

int ffq_quantize(const void* x, int x_dtype, void* q, int q_dtype, const void* scale, int scale_dtype,
                 const void* offset, int offset_dtype, const ffq_layout_t* layout, double num_bits,
                 void* stream) {
  int rc;
  if ((rc = check_dt(x_dtype, "quantize: data")) || (rc = check_dt(q_dtype, "quantize: output")) ||
      (rc = check_dt(scale_dtype, "quantize: scale")))
    return rc;
  if (!is_float_dt(scale_dtype)) { set_error("quantize: scale must be floating point"); return FFQ_ERR_INVALID; }
  if (offset == nullptr) offset_dtype = FFQ_NONE;
  else if ((rc = check_dt(offset_dtype, "quantize: offset"))) return rc;
  EwArgs a{};
  a.in = x; a.out = q; a.codes = nullptr;
  a.in_dt = x_dtype; a.out_dt = q_dtype; a.codes_dt = FFQ_NONE;
  a.scale = scale; a.offset = offset; a.s_dt = scale_dtype; a.o_dt = offset_dtype;
  a.qp = make_qparams(x_dtype, scale_dtype, offset_dtype, num_bits);
  return run_elementwise(OP_QUANT, a, layout, num_bits, static_cast<cudaStream_t>(stream));
}

int ffq_dequantize(const void* q, int q_dtype, void* y, int y_dtype, const void* scale, int scale_dtype,
                   const void* offset, int offset_dtype, const ffq_layout_t* layout, void* stream) {
  int rc;
  if ((rc = check_dt(q_dtype, "dequantize: codes")) || (rc = check_dt(y_dtype, "dequantize: output")) ||
      (rc = check_dt(scale_dtype, "dequantize: scale")))
    return rc;
  if (offset == nullptr) offset_dtype = FFQ_NONE;
  else if ((rc = check_dt(offset_dtype, "dequantize: offset"))) return rc;
  EwArgs a{};
  a.in = q; a.out = y; a.codes = nullptr;
  a.in_dt = q_dtype; a.out_dt = y_dtype; a.codes_dt = FFQ_NONE;
  a.scale = scale; a.offset = offset; a.s_dt = scale_dtype; a.o_dt = offset_dtype;
  a.dp = make_dparams(q_dtype, scale_dtype, offset_dtype);
  return run_elementwise(OP_DEQUANT, a, layout, 0.0, static_cast<cudaStream_t>(stream));
}

int ffq_fakequant_fwd(const void* x, int x_dtype, void* y, int y_dtype, void* codes, int q_dtype,
                      const void* scale, int scale_dtype, const void* offset, int offset_dtype,
                      const ffq_layout_t* layout, double num_bits, void* stream) {
  int rc;
  if ((rc = check_dt(x_dtype, "fakequant: data")) || (rc = check_dt(y_dtype, "fakequant: output")) ||
      (rc = check_dt(q_dtype, "fakequant: codes")) || (rc = check_dt(scale_dtype, "fakequant: scale")))
    return rc;
  if (!is_float_dt(y_dtype)) { set_error("fakequant: output dtype must be floating point"); return FFQ_ERR_UNSUPPORTED; }
  if (offset == nullptr) offset_dtype = FFQ_NONE;
  else if ((rc = check_dt(offset_dtype, "fakequant: offset"))) return rc;
  EwArgs a{};
  a.in = x; a.out = y; a.codes = codes;
  a.in_dt = x_dtype; a.out_dt = y_dtype; a.codes_dt = q_dtype;
  a.scale = scale; a.offset = offset; a.s_dt = scale_dtype; a.o_dt = offset_dtype;
  a.qp = make_qparams(x_dtype, scale_dtype, offset_dtype, num_bits);
  a.dp = make_dparams(q_dtype, scale_dtype, offset_dtype);
  a.q_rt_dt = q_dtype;
  return run_elementwise(OP_FAKEQUANT, a, layout, num_bits, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

// ffq_elementwise.cu -- quantize / dequantize / fused fake-quantize (a1, a2).
//
// HBM-bound streaming kernels.  Fast path ("row" layouts: every tile is one contiguous run,
// which covers per-tensor, per-channel(0) and per-group weights): 16-byte vector loads with
// L1 no-allocate, 4 vectors in flight per thread, one parameter fetch per vector.  Algorithmic
// traffic per element: quantize s+c, dequantize c+s, fake-quantize 2s bytes (s = data bytes,
// c = code bytes).  Everything else (strided tiles, integer inputs, unaligned pointers) runs on
// a scalar generic kernel: correct for any rank<=8 tiling and any dtype, not tuned.
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
  FastDiv tdiv;     // valid when numel < 2^32
  int big;          // numel >= 2^32: 64-bit index math
  QParams qp;
  DParams dp;
  int q_rt_dt;      // fake-quant: dtype the codes take between quantize and dequantize
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

template <int OP, typename InT, typename OutT>
__global__ void __launch_bounds__(EW_THREADS) ew_row_kernel(const EwArgs a) {
  constexpr int EPT = 16 / sizeof(InT);
  const InT* __restrict__ in = static_cast<const InT*>(a.in);
  OutT* __restrict__ out = static_cast<OutT*>(a.out);
  const unsigned long long nvec = a.numel / EPT;
  const unsigned long long vbase = (unsigned long long)blockIdx.x * (EW_THREADS * EW_UNROLL) + threadIdx.x;

  Vec<InT, EPT> xin[EW_UNROLL];
#pragma unroll
  for (int u = 0; u < EW_UNROLL; ++u) {
    const unsigned long long v = vbase + (unsigned long long)u * EW_THREADS;
    if (v < nvec) xin[u] = ld_stream<InT, EPT>(in + v * EPT);
  }
#pragma unroll
  for (int u = 0; u < EW_UNROLL; ++u) {
    const unsigned long long v = vbase + (unsigned long long)u * EW_THREADS;
    if (v >= nvec) continue;
    const unsigned long long e0 = v * EPT;
    unsigned long long p0, p1;
    if (!a.big) {
      p0 = fast_div((unsigned int)e0, a.tdiv);
      p1 = fast_div((unsigned int)e0 + (EPT - 1), a.tdiv);
    } else {
      p0 = e0 / a.tile_numel;
      p1 = (e0 + (EPT - 1)) / a.tile_numel;
    }
    Vec<OutT, EPT> y;
    float s = load_as_float(a.scale, a.s_dt, p0);
    float o = load_offset(a.offset, a.o_dt, p0);
    if (p0 == p1) {
#pragma unroll
      for (int i = 0; i < EPT; ++i) {
        float c;
        const float r = ew_apply<OP>(Elem<InT>::to_f(xin[u].v[i]), s, o, a, &c);
        y.v[i] = Elem<OutT>::from_f(r);
        if constexpr (OP == OP_FAKEQUANT) {
          if (a.codes) store_from_float(a.codes, a.codes_dt, e0 + i, c);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < EPT; ++i) {
        const unsigned long long p = a.big ? (e0 + i) / a.tile_numel
                                           : (unsigned long long)fast_div((unsigned int)e0 + i, a.tdiv);
        if (p != p0) { p0 = p; s = load_as_float(a.scale, a.s_dt, p); o = load_offset(a.offset, a.o_dt, p); }
        float c;
        const float r = ew_apply<OP>(Elem<InT>::to_f(xin[u].v[i]), s, o, a, &c);
        y.v[i] = Elem<OutT>::from_f(r);
        if constexpr (OP == OP_FAKEQUANT) {
          if (a.codes) store_from_float(a.codes, a.codes_dt, e0 + i, c);
        }
      }
    }
    st_vec<OutT, EPT>(out + e0, y);
  }
  // scalar tail (numel % EPT elements), done by the last block's first threads
  if (blockIdx.x == gridDim.x - 1) {
    const unsigned long long e = nvec * EPT + threadIdx.x;
    if (e < a.numel) {
      const unsigned long long p = e / a.tile_numel;
      const float s = load_as_float(a.scale, a.s_dt, p);
      const float o = load_offset(a.offset, a.o_dt, p);
      float c;
      const float r = ew_apply<OP>(Elem<InT>::to_f(in[e]), s, o, a, &c);
      out[e] = Elem<OutT>::from_f(r);
      if constexpr (OP == OP_FAKEQUANT) {
        if (a.codes) store_from_float(a.codes, a.codes_dt, e, c);
      }
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

template <int OP, typename InT, typename OutT>
static void launch_row(const EwArgs& a, cudaStream_t st) {
  constexpr int EPT = 16 / sizeof(InT);
  const unsigned long long nvec = a.numel / EPT;
  unsigned long long blocks = (nvec + EW_THREADS * EW_UNROLL - 1) / (EW_THREADS * EW_UNROLL);
  if (blocks == 0) blocks = 1;
  ew_row_kernel<OP, InT, OutT><<<(unsigned int)blocks, EW_THREADS, 0, st>>>(a);
}

template <int OP, typename InT>
static bool dispatch_out(const EwArgs& a, cudaStream_t st) {
  switch (a.out_dt) {
    case FFQ_F32: launch_row<OP, InT, float>(a, st); return true;
    case FFQ_BF16: launch_row<OP, InT, __nv_bfloat16>(a, st); return true;
    case FFQ_F16: launch_row<OP, InT, __half>(a, st); return true;
    case FFQ_I8: if constexpr (OP == OP_QUANT) { launch_row<OP, InT, int8_t>(a, st); return true; } break;
    case FFQ_I16: if constexpr (OP == OP_QUANT) { launch_row<OP, InT, int16_t>(a, st); return true; } break;
    case FFQ_I32: if constexpr (OP == OP_QUANT) { launch_row<OP, InT, int32_t>(a, st); return true; } break;
  }
  return false;
}

template <int OP>
static bool dispatch_row(const EwArgs& a, cudaStream_t st) {
  switch (a.in_dt) {
    case FFQ_F32: return dispatch_out<OP, float>(a, st);
    case FFQ_BF16: return dispatch_out<OP, __nv_bfloat16>(a, st);
    case FFQ_F16: return dispatch_out<OP, __half>(a, st);
    case FFQ_I8: if constexpr (OP == OP_DEQUANT) return dispatch_out<OP, int8_t>(a, st); break;
    case FFQ_I16: if constexpr (OP == OP_DEQUANT) return dispatch_out<OP, int16_t>(a, st); break;
    case FFQ_I32: if constexpr (OP == OP_DEQUANT) return dispatch_out<OP, int32_t>(a, st); break;
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

static int run_elementwise(int op, EwArgs& a, const ffq_layout_t* layout, cudaStream_t st) {
  Plan plan;
  int rc = make_plan(layout, &plan);
  if (rc != FFQ_OK) return rc;
  if (plan.numel == 0) return FFQ_OK;
  a.numel = (unsigned long long)plan.numel;
  a.tile_numel = (unsigned long long)plan.tile_numel;
  a.big = plan.numel >= (1ll << 32) ? 1 : 0;
  a.tdiv = make_fast_div(a.big || plan.tile_numel >= (1ll << 32) ? 1u : (unsigned int)plan.tile_numel);
  if (!a.big && plan.tile_numel >= (1ll << 32)) a.big = 1;
  a.gl = make_generic_layout(plan);

  bool done = false;
  const bool fast_ok = plan.row && aligned16(a.in) && aligned16(a.out);
  if (fast_ok) {
    if (op == OP_QUANT) done = dispatch_row<OP_QUANT>(a, st);
    else if (op == OP_DEQUANT) done = dispatch_row<OP_DEQUANT>(a, st);
    else done = dispatch_row<OP_FAKEQUANT>(a, st);
  }
  if (!done) {
    const unsigned long long blocks = (a.numel + 255) / 256;
    if (blocks > 0x7fffffffull) {
      set_error("tensor too large for the generic elementwise kernel");
      return FFQ_ERR_UNSUPPORTED;
    }
    if (op == OP_QUANT) ew_generic_kernel<OP_QUANT><<<(unsigned int)blocks, 256, 0, st>>>(a);
    else if (op == OP_DEQUANT) ew_generic_kernel<OP_DEQUANT><<<(unsigned int)blocks, 256, 0, st>>>(a);
    else ew_generic_kernel<OP_FAKEQUANT><<<(unsigned int)blocks, 256, 0, st>>>(a);
  }
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
  return run_elementwise(OP_QUANT, a, layout, static_cast<cudaStream_t>(stream));
}

int ffq_dequantize(const void* q, int q_dtype, void* y, int y_dtype, const void* scale, int scale_dtype,
                   const void* offset, int offset_dtype, const ffq_layout_t* layout, void* stream) {
  int rc;
  if ((rc = check_dt(q_dtype, "dequantize: codes")) || (rc = check_dt(y_dtype, "dequantize: output")) ||
      (rc = check_dt(scale_dtype, "dequantize: scale")))
    return rc;
  if (!is_float_dt(y_dtype)) { set_error("dequantize: output dtype must be floating point"); return FFQ_ERR_UNSUPPORTED; }
  if (offset == nullptr) offset_dtype = FFQ_NONE;
  else if ((rc = check_dt(offset_dtype, "dequantize: offset"))) return rc;
  EwArgs a{};
  a.in = q; a.out = y; a.codes = nullptr;
  a.in_dt = q_dtype; a.out_dt = y_dtype; a.codes_dt = FFQ_NONE;
  a.scale = scale; a.offset = offset; a.s_dt = scale_dtype; a.o_dt = offset_dtype;
  a.dp = make_dparams(q_dtype, scale_dtype, offset_dtype);
  return run_elementwise(OP_DEQUANT, a, layout, static_cast<cudaStream_t>(stream));
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
  return run_elementwise(OP_FAKEQUANT, a, layout, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

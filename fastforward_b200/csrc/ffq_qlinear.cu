// ffq_qlinear.cu -- W8A8 quantized linear on Blackwell tensor cores (a12).
//
//   y[m,n] = sx*sw[n] * ( sum_k qx[m,k]*qw[n,k] + ox*rowsum_w[n] + ow[n]*rowsum_x[m] + K*ox*ow[n] ) + bias[n]
//
// int8 x int8 -> int32 on tcgen05 (kind::i8), accumulators in TMEM, operands staged by TMA into
// 128B-swizzled shared memory, dequantisation fused into the epilogue.  The reference has no
// such kernel: it dequantises both operands to float tensors in HBM and calls a float GEMM
// (_gen/fallback.py:94-108).  int32 accumulation is exact, so this path is *more* accurate than
// the fallback it replaces.
//
// Structure (one CTA per SM, persistent over output tiles, warp-specialised):
//   warp 0     : TMA producer   -- cp.async.bulk.tensor of the A (128 x 128B) and B (256 x 128B)
//                                  k-blocks into a 4-stage ring, completion on mbarriers
//   warp 1     : MMA issuer     -- one elected lane issues 4 x tcgen05.mma (K=32 each) per k-block
//                                  into one of two 128x256 fp32-column TMEM accumulators;
//                                  tcgen05.commit releases the smem stage / publishes the tile
//   warps 2..5 : epilogue       -- tcgen05.ld 32 columns at a time, y = alpha[n]*float(acc + c[n] +
//                                  ow[n]*rowsum_x[m]) + bias[n], vector stores; overlaps the next tile's
//                                  MMAs through the second accumulator.  All offset corrections are
//                                  added in int32 (exact); one int->float conversion and one FMA
//                                  per output follow
// Roofline: tensor pipe; 2*M*N*K ops.  A 128x256 tile needs 96 B/clk/SM of operand traffic at the
// full MMA rate, so L2->SM bandwidth is the secondary bound.
#include <cuda.h>

#include <cstdlib>

#include <atomic>
#include <cmath>
#include <cstring>

#include "ffq_common.cuh"
#include "ffq_umma.cuh"

namespace ffq {

constexpr int BM = 128, BN = 256, BK = 128;      // BK in bytes == int8 elements: one 128B swizzle atom
constexpr int UMMA_K = 32;                       // int8 elements per tcgen05.mma
constexpr int STAGES = 4;
constexpr int A_STAGE = BM * BK, B_STAGE = BN * BK;
constexpr int STAGE_BYTES = A_STAGE + B_STAGE;   // 48 KB
constexpr int GEMM_THREADS = 192;                // 6 warps
constexpr int TMEM_COLS = 512;                   // 2 accumulators x 256 columns
constexpr int COL_BYTES = 5 * BN * 4;            // per-column epilogue parameters (COL_SLOTS x BN words)
constexpr int OUT_STAGE_BYTES = 4 * 32 * 128;    // epilogue staging: one [32 rows x 128 B] box per epilogue warp
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + OUT_STAGE_BYTES + COL_BYTES + 256 + 1024;   // + staging + column params + barriers + align

struct GemmArgs {
  int M, N, K;
  void* y; int y_dt;                            // may be null when only the requantized codes are wanted
  const float* sx; const float* ox;             // activation scale / offset (one element each; ox may be null)
  const float* sw; const float* ow;             // per output column (ow may be null)
  const int32_t* rowsum_w;                      // per output column: sum_k qw[n,k]
  const void* bias; int bias_dt;                // per output column, optional
  const int32_t* rowsum_x;                      // per output row; used with ow
  int bn;                                       // output columns per tile (256; 224 / 128 when that evens out the waves)
  // fused output quantizer (ffq_requant_t): per-tensor, int8 codes of the output rounded to y_dt
  const float* rq_scale; const float* rq_offset; float rq_lo, rq_hi; int8_t* rq_codes; int32_t* rq_rowsum;
  // test hook (ffq_debug_gemm_profile): per CTA PROF_SLOTS x u64 of SM clocks -- [0] producer waiting for a free stage, [1] producer
  // total, [2] MMA issuer waiting for operands, [3] MMA issuer waiting for a free accumulator, [4] MMA issuer total,
  // [5] epilogue waiting for a finished tile, [6] epilogue total
  unsigned long long* prof;
  int tma_store;                                // the output has a tensor map: staged TMA-store epilogue
  int dbg;                                      // test hook (FFQ_GEMM_DEBUG): 1 skip the epilogue, 2 skip its stores, 4 skip its column parameters
};

constexpr int PROF_SLOTS = 16;                   // u64 counters per CTA of the ffq_debug_gemm_profile hook
constexpr int COL_SLOTS = 5;                     // alpha, bias, int constant, weight offset, float constant

// Per-column epilogue parameters of one tile, derived where they are consumed (no separate launch, no scratch):
//   alpha[n] = sx*sw[n];  bias[n] as float;  cnst[n] = ox*rowsum_w[n] + K*ox*ow[n];  own[n] = ow[n]
// with ox, ow rounded to integers exactly as dequantize_by_tile rounds them.  The int32 form of the constant and of
// own*rowsum_x is exact; when a column's terms could leave int32 (offsets far from zero relative to the range's
// width) the function returns true and the tile's epilogue adds the same terms in float, as the reference's
// dequantize-then-float path does.
__device__ __forceinline__ bool stage_col_params(const GemmArgs& g, int n0, int bn, int tid, int nthreads,
                                                 float* col_params, int32_t* col_ints) {
  // layout: uint4 per column {alpha, bias, int constant, weight offset} (one 16-byte shared load per column in the
  // epilogue), followed by the float form of the constant for the wide path at word 4*256 + c
  const float sx = g.sx[0];
  const float oxf = g.ox ? rintf(g.ox[0]) : 0.f;
  const long long o_x = (long long)oxf;
  bool wide = !(fabsf(oxf) < 1.0e9f);
  for (int c = tid; c < bn; c += nthreads) {
    const int n = n0 + c;
    const bool in = n < g.N;
    const float owf = (in && g.ow) ? rintf(g.ow[n]) : 0.f;
    const long long o_w = (long long)owf;
    const long long c64 = in ? o_x * (long long)g.rowsum_w[n] + (long long)g.K * o_x * o_w : 0;
    // |acc| <= K * 2^14, |own * rowsum_x| <= |o_w| * K * 2^7
    const long long bound = (c64 < 0 ? -c64 : c64) + (o_w < 0 ? -o_w : o_w) * (long long)g.K * 128 + (long long)g.K * 16384;
    wide = wide || !(fabsf(owf) < 1.0e9f) || bound >= 0x7fffffffll;
    float4 v;
    v.x = in ? sx * g.sw[n] : 0.f;
    v.y = (in && g.bias) ? load_as_float(g.bias, g.bias_dt, n) : 0.f;
    v.z = __int_as_float((int32_t)c64);
    v.w = __int_as_float((int32_t)o_w);
    reinterpret_cast<float4*>(col_params)[c] = v;
    col_params[4 * BN + c] = in ? (oxf * (float)g.rowsum_w[n] + (float)g.K * oxf * owf) : 0.f;
  }
  (void)col_ints;
  return wide;
}

__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}

// barrier over the NT epilogue threads that also ORs a predicate (the tile's "wide constants" decision)
template <int NT = 128>
__device__ __forceinline__ bool epilogue_bar_or(bool pred) {
  uint32_t r;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %1, 0;\n\t"
      "barrier.cta.red.or.pred q, 1, %2, p;\n\t"
      "selp.b32 %0, 1, 0, q;\n\t}" : "=r"(r) : "r"((uint32_t)pred), "n"(NT) : "memory");
  return r != 0;
}

// One 32-column chunk of one accumulator row: offset corrections, dequantisation, optional requantisation, stores.
template <typename OutT>
__device__ __forceinline__ void epilogue_chunk(const GemmArgs& g, const uint32_t (&acc)[32], const float* col_params,
                                               const int32_t* col_ints, int bn, int c0, bool wide, int32_t rx, int row,
                                               int n0, float rq_s, float rq_o, int& rq_sum, uint8_t* stage) {
  float v[32];
  const uint32_t cp = smem_u32(col_params) + (uint32_t)c0 * 16u;
  if (g.dbg & 4) {                       // test hook: no per-column parameters (isolates their shared-memory traffic)
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = (float)(int32_t)acc[j] * 0.001f;
  } else if (!wide) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const uint4 p = lds128(cp + (uint32_t)j * 16u);
      const int32_t t = (int32_t)acc[j] + (int32_t)p.z + (int32_t)p.w * rx;
      v[j] = fmaf(__uint_as_float(p.x), (float)t, __uint_as_float(p.y));
    }
  } else {
    const float rxf = (float)rx;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const uint4 p = lds128(cp + (uint32_t)j * 16u);
      const float t = (float)(int32_t)acc[j] + col_params[4 * BN + c0 + j] + (float)(int32_t)p.w * rxf;
      v[j] = fmaf(__uint_as_float(p.x), t, __uint_as_float(p.y));
    }
  }
  if (g.dbg & 2) return;
  if (stage != nullptr) {
    // ---- staged store: this lane's 32 values go into its row of the warp's [32 rows x 128 B] box, 128B-swizzled (16-byte
    // chunk index XOR row%8: conflict-free writes, undone by the TMA store); one TMA store per full box ----
    constexpr int PER = 16 / sizeof(OutT);                    // elements per 16-byte chunk
    constexpr int CHUNKS = 32 / PER;                          // 16-byte chunks this call fills (4 for 16-bit, 8 for fp32)
    const int lane = threadIdx.x & 31;
    const int q0 = (sizeof(OutT) == 2) ? ((c0 >> 5) & 1) * 4 : 0;   // 16-bit: two 32-column calls fill one 128-byte row
    uint8_t* rowp = stage + lane * 128;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
      Vec<OutT, PER> o;
#pragma unroll
      for (int i = 0; i < PER; ++i) o.v[i] = Elem<OutT>::from_f(v[c * PER + i]);
      *reinterpret_cast<Vec<OutT, PER>*>(rowp + (((q0 + c) ^ (lane & 7)) << 4)) = o;
    }
  } else if (row < g.M && n0 < g.N && g.y) {
    const int ncols = (g.N - n0) < 32 ? (g.N - n0) : 32;
    store_chunk<OutT>(static_cast<OutT*>(g.y) + (size_t)row * g.N + n0, v, ncols);
  }
  if (row >= g.M || n0 >= g.N) return;
  const int ncols = (g.N - n0) < 32 ? (g.N - n0) : 32;
  if (g.rq_codes) {
    const SharedRcp rq_k = make_shared_rcp(rq_s);
    // output_quantizer(y): y is first rounded to the output dtype (what the quantizer would read back), then
    // quantize_by_tile's arithmetic in the promoted dtype of (y, fp32 scale) = fp32.  Exact y / s with the reciprocal
    // shared by the whole tensor (ffq_common.cuh: shared_div); round-to-nearest-even and the conversion are ONE
    // cvt.rni.s32.f32 (saturating, NaN -> 0: what clamp(rint(t)) followed by the cast to the integer code dtype gives),
    // the clamp is done on integers, packing saturates two codes at a time and the row sum is one dp4a per word --
    // 9 instead of 17 instructions per element on top of the dequantisation, which keeps the epilogue of a tile
    // shorter than the tile's MMAs (it was longer: the fused GEMM ran at 0.76x the speed of the plain one).
    const int ilo = (int)g.rq_lo, ihi = (int)g.rq_hi;
    int c[32];
    bool ok = rq_k.ok;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float yr = Elem<OutT>::to_f(Elem<OutT>::from_f(v[j]));
      const float quo = shared_div<false>(yr, rq_k, ok);
      c[j] = __float2int_rn(__fsub_rn(quo, rq_o));
    }
    if (!ok) {                           // outside the guard of the shared reciprocal (never for sane scales): IEEE division
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float yr = Elem<OutT>::to_f(Elem<OutT>::from_f(v[j]));
        c[j] = __float2int_rn(__fsub_rn(__fdiv_rn(yr, rq_s), rq_o));
      }
    }
    if (!(ilo == -128 && ihi == 127)) {
#pragma unroll
      for (int j = 0; j < 32; ++j) c[j] = min(max(c[j], ilo), ihi);
    }
    uint32_t packed[8];
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      uint32_t hi16, word;
      asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(hi16) : "r"(c[4 * w + 3]), "r"(c[4 * w + 2]), "r"(0));
      asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(word) : "r"(c[4 * w + 1]), "r"(c[4 * w]), "r"(hi16));
      packed[w] = word;
    }
    if (ncols == 32) {
#pragma unroll
      for (int w = 0; w < 8; ++w) rq_sum = __dp4a((int)packed[w], 0x01010101, rq_sum);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) rq_sum += c[j];
    }
    int8_t* dst = g.rq_codes + (size_t)row * g.N + n0;
    if (ncols == 32 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
      *reinterpret_cast<uint4*>(dst) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
      *reinterpret_cast<uint4*>(dst + 16) = make_uint4(packed[4], packed[5], packed[6], packed[7]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) dst[j] = (int8_t)((packed[j >> 2] >> (8 * (j & 3))) & 0xffu);
    }
  }
}


// TMA store of one [32 rows x 128 B] box from shared memory (bulk async-group completion)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_u32(src)),
               "r"(c0), "r"(c1) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// The epilogue of one tile for one warp (32 accumulator rows): TMEM -> registers -> dequantisation -> either direct
// vector stores (64-byte row segments per lane) or, when the output allows a tensor map (`map_y`), a 128B-swizzled
// staging box in shared memory drained by TMA stores: whole 128-byte lines reach L2 instead of 16-byte pieces of 32
// different rows per instruction, and the stores leave the LSU path the operand loads' completions share.
template <typename OutT>
__device__ __forceinline__ void epilogue_tile(const GemmArgs& g, const CUtensorMap* map_y, uint32_t taddr, int bn, int row,
                                              int row0_warp, int n_tile0, bool wide, int32_t rx, float* col_params,
                                              const int32_t* col_ints, uint8_t* stage, float rq_s, float rq_o, bool& pending,
                                              long long* pt = nullptr, int c_begin = 0, int c_end = -1) {
  // [c_begin, c_end): this warp's share of the tile's columns (the pair kernel splits a lane quadrant between two warps);
  // both bounds are multiples of a staged box
  // pt (test hook, one thread): [0] tcgen05.ld, [1] waiting for the staging box to be free, [2] arithmetic + staging,
  // [3] fence + TMA store issue
  const int lane = threadIdx.x & 31;
  constexpr int COLS_PER_BOX = 128 / (int)sizeof(OutT);       // 64 (16-bit) or 32 (fp32) columns per staged box
  int rq_sum = 0;
  if (c_end < 0) c_end = bn;
#pragma unroll 1
  for (int c0 = c_begin; c0 < ((g.dbg & 1) ? 0 : c_end); c0 += 32) {
    uint32_t acc[32];
    long long q0 = pt ? clock64() : 0;
    tmem_ld32(taddr + (uint32_t)c0, acc);
    if (pt) { const long long q1 = clock64(); pt[0] += q1 - q0; q0 = q1; }
    // a lone trailing 32-column chunk of a 16-bit tile (bn = 224) cannot fill a box: direct stores
    const bool box_first = (c0 % COLS_PER_BOX) == 0;
    const bool staged = stage != nullptr && (sizeof(OutT) == 4 || !box_first || c0 + 32 < c_end);
    if (staged && box_first && pending) {
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous box has been read
      __syncwarp();
      pending = false;
    }
    if (pt) { const long long q1 = clock64(); pt[1] += q1 - q0; q0 = q1; }
    epilogue_chunk<OutT>(g, acc, col_params, col_ints, bn, c0, wide, rx, row, n_tile0 + c0, rq_s, rq_o, rq_sum,
                         staged ? stage : nullptr);
    if (pt) { const long long q1 = clock64(); pt[2] += q1 - q0; q0 = q1; }
    if (staged && ((c0 + 32) % COLS_PER_BOX) == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      const int nb = n_tile0 + c0 + 32 - COLS_PER_BOX;
      if (lane == 0 && nb < g.N && row0_warp < g.M) tma_store_2d(map_y, stage, nb, row0_warp);
      pending = true;
    }
    if (pt) { const long long q1 = clock64(); pt[3] += q1 - q0; q0 = q1; }
  }
  if (g.rq_rowsum && row < g.M) atomicAdd(&g.rq_rowsum[row], rq_sum);     // integer: exact, order independent
}

template <typename OutT>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
w8a8_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_y, const GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint8_t* out_stage = smem + STAGES * STAGE_BYTES;                                // [4 warps][32 rows][128 B], swizzled
  float* col_params = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + OUT_STAGE_BYTES);
  int32_t* col_ints = reinterpret_cast<int32_t*>(col_params);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + OUT_STAGE_BYTES + COL_BYTES);
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* tmem_full = bars + 2 * STAGES;   // [2]
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;   // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bn = g.bn;                       // columns per tile (256 or 128); the smem / TMEM layout keeps its 256-column pitch
  const int tiles_m = (g.M + BM - 1) / BM, tiles_n = (g.N + bn - 1) / bn;
  const uint32_t stage_tx = (uint32_t)(A_STAGE + bn * BK);
  const int num_tiles = tiles_m * tiles_n;
  const int k_blocks = (g.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)),
                 "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  pdl_wait();                   // barriers and TMEM are set up while the previous kernel drains (no-ops under FFQ_PDL=0)
  pdl_trigger();

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        // consecutive CTAs share the same N panel of B (tile index runs fastest over m)
        const int tm = tile % tiles_m, tn = tile / tiles_m;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = stage_base + stage * STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], stage_tx);
          tma_load_2d(sa, &map_a, &full_bar[stage], kb * BK, tm * BM);
          tma_load_2d(sa + A_STAGE, &map_b, &full_bar[stage], kb * BK, tn * bn);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      // instruction descriptor: D=S32, A=B=signed int8, both K-major, N=bn, M=128
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t use = (uint32_t)(it >> 1);            // how many times this buffer was used before
        mbar_wait(&tmem_empty[buf], (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + stage * STAGE_BYTES);
          const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sa + A_STAGE);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advancing K inside the swizzle atom: +32 bytes == +2 in the (>>4) start-address field
            umma_i8(tmem_d, da + (uint64_t)(k * (UMMA_K >> 4)), db + (uint64_t)(k * (UMMA_K >> 4)), idesc,
                    (kb | k) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);                    // frees the smem stage when the MMAs retire
          if (kb == k_blocks - 1) umma_commit(&tmem_full[buf]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===== epilogue (warps 2..5): TMEM lane quadrant = warp % 4 =====
    const int quad = warp & 3;
    const int ep_tid = threadIdx.x - 64;                     // 0..127
    const float rq_s = g.rq_codes ? g.rq_scale[0] : 1.f;
    const float rq_o = (g.rq_codes && g.rq_offset) ? rintf(g.rq_offset[0]) : 0.f;
    bool store_pending = false;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int tm = tile % tiles_m, tn = tile / tiles_m;
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      // stage this tile's column parameters in shared memory (named barrier over the 4 epilogue warps)
      asm volatile("bar.sync 1, 128;" ::: "memory");         // previous tile's readers are done
      const bool wide = epilogue_bar_or(stage_col_params(g, tn * bn, bn, ep_tid, 128, col_params, col_ints));
      const int row = tm * BM + quad * 32 + lane;
      const int32_t rx = (g.ow && row < g.M) ? g.rowsum_x[row] : 0;

      mbar_wait(&tmem_full[buf], use & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * BN);
      epilogue_tile<OutT>(g, &map_y, taddr, bn, row, tm * BM + quad * 32, tn * bn, wide, rx, col_params, col_ints,
                          g.tma_store ? out_stage + quad * 4096 : nullptr, rq_s, rq_o, store_pending);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);          // 4 arrivals (one per epilogue warp) free the accumulator
    }
    if (store_pending && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

// ================================================================================================
// CTA-pair variant, in clusters of P pairs.  A CTA pair (two SMs of one TPC) owns a 256 x bn output tile: each
// CTA stages its own 128 rows of A and its own half of the tile's B rows (32 KB per k-block instead of 48 KB:
// operand delivery from L2 is what bounds the kernel), the pair's leader issues tcgen05.mma.cta_group::2 (M = 256)
// reading both CTAs' shared memory, and each CTA's TMEM holds its 128 accumulator rows.
// With P > 1 the P pairs of a cluster work on P consecutive M tiles of the SAME B panel, and each B half is
// fetched from L2 once per cluster: CTA (pair p, rank r) loads slice p of "B half r" and TMA-multicasts it to the
// rank-r CTA of every pair, so a CTA receives 16 KB of A + 16/P KB of B per k-block from L2 (P = 1: 32 KB,
// P = 2: 24 KB, P = 4: 20 KB) while its shared memory still fills with the whole 32 KB.
// Barriers: TMA completions of a pair (its own A loads and every B slice multicast into it) land on the pair
// leader's "full" barrier; a stage is refilled only after ALL P pairs have consumed it, because every pair writes
// into every other pair's stage (tcgen05.commit multicast to the whole cluster, "empty" barriers count P);
// "tile ready" goes to the pair only, and the pair's epilogue warps (EP2_WARPS per CTA) release the accumulator on the leader.
// ================================================================================================
// Epilogue warps per CTA: 4 (one per TMEM lane quadrant) or 8 (two per quadrant, each taking half of the tile's columns:
// the drain of a tile -- all of it exposed after a launch's last wave -- takes half as long).  Eight warps need eight
// staging boxes (32 KB), which fit beside five ring stages instead of six.  -DFFQ_GEMM2_EPW=4 builds the former shape.
#ifndef FFQ_GEMM2_EPW
#define FFQ_GEMM2_EPW 8
#endif
constexpr int EP2_WARPS = FFQ_GEMM2_EPW;
static_assert(EP2_WARPS == 4 || EP2_WARPS == 8, "FFQ_GEMM2_EPW must be 4 or 8");
constexpr int EP2_THREADS = 32 * EP2_WARPS;
constexpr int GEMM2_THREADS = 64 + EP2_THREADS;    // TMA producer warp + MMA warp + epilogue warps
constexpr int STAGES2 = EP2_WARPS == 8 ? 5 : 6;
constexpr int HALF_STAGE = A_STAGE + BM * BK;      // A 128x128B + B half 128x128B = 32 KB
constexpr int OUT_STAGE2_BYTES = EP2_WARPS * 32 * 128;
constexpr int SMEM2_BYTES = STAGES2 * HALF_STAGE + OUT_STAGE2_BYTES + COL_BYTES + 256 + 1024;
static_assert(SMEM2_BYTES <= 232448, "pair kernel: shared memory over the 227 KB limit");
template <typename OutT, int P>
__global__ void __launch_bounds__(GEMM2_THREADS, 1)
w8a8_gemm2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  const __grid_constant__ CUtensorMap map_y, const GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint8_t* out_stage = smem + STAGES2 * HALF_STAGE;                                // [EP2_WARPS][32 rows][128 B], swizzled
  float* col_params = reinterpret_cast<float*>(smem + STAGES2 * HALF_STAGE + OUT_STAGE2_BYTES);
  int32_t* col_ints = reinterpret_cast<int32_t*>(col_params);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES2 * HALF_STAGE + OUT_STAGE2_BYTES + COL_BYTES);
  uint64_t* full_bar = bars;                  // [STAGES2]  (the pair leader's copy is the one in use)
  uint64_t* empty_bar = bars + STAGES2;       // [STAGES2]  (each CTA waits on its own copy; P arrivals)
  uint64_t* tmem_full = bars + 2 * STAGES2;   // [2]        (each CTA waits on its own copy)
  uint64_t* tmem_empty = bars + 2 * STAGES2 + 2;   // [2]   (pair leader's copy, 8 arrivals)
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES2 + 4);

  constexpr int CSIZE = 2 * P;                // CTAs per cluster
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const uint32_t cta = crank & 1u;            // rank inside the pair
  const int p = (int)(crank >> 1);            // pair inside the cluster
  const int cluster = blockIdx.x / CSIZE, num_clusters = gridDim.x / CSIZE;
  constexpr int TM = 2 * BM;                  // 256 rows per pair tile
  const int bn = g.bn;                        // columns per tile; the smem / TMEM layout keeps its 256-column pitch
  const int tiles_m = (g.M + TM - 1) / TM, tiles_n = (g.N + bn - 1) / bn;
  const int groups_m = (tiles_m + P - 1) / P; // P consecutive M tiles share one B panel
  const uint32_t stage_tx = 2u * (uint32_t)(A_STAGE + (bn / 2) * BK);
  const int num_tiles = groups_m * tiles_n;   // cluster tiles
  const int k_blocks = (g.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES2; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], P); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 2 * EP2_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)),
                 "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  cluster_sync_all();                         // barriers of ALL CTAs are initialised before any remote use
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  pdl_wait();                   // barriers and TMEM are set up while the previous kernel drains (no-ops under FFQ_PDL=0)
  pdl_trigger();

  if (warp == 0) {
    // ===== TMA producer (every CTA; completions count on the pair leaders' full barriers) =====
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
      const int slice = (bn / 2) / P;         // B rows this CTA fetches for all rank-`cta` CTAs of the cluster
      uint16_t mc_mask = 0;
#pragma unroll
      for (int q = 0; q < P; ++q) mc_mask |= (uint16_t)(1u << (2 * q + (int)cta));
      int stage = 0; uint32_t phase = 0;
      const bool prof = g.prof != nullptr;
      long long t_wait = 0;
      const long long t_begin = prof ? clock64() : 0;
      for (int tile = cluster; tile < num_tiles; tile += num_clusters) {
        const int tm = (tile % groups_m) * P + p, tn = tile / groups_m;     // tm >= tiles_m: zero-filled rows
        for (int kb = 0; kb < k_blocks; ++kb) {
          const long long w0 = prof ? clock64() : 0;
          mbar_wait_bounded(&empty_bar[stage], phase ^ 1);
          if (prof) t_wait += clock64() - w0;
          uint8_t* sa = stage_base + stage * HALF_STAGE;
          if (cta == 0) mbar_expect_tx(&full_bar[stage], stage_tx);
          tma_load_2d_pair(sa, &map_a, &full_bar[stage], kb * BK, tm * TM + (int)cta * BM);
          if constexpr (P == 1) {
            tma_load_2d_pair(sa + A_STAGE, &map_b, &full_bar[stage], kb * BK, tn * bn + (int)cta * (bn / 2));
          } else {
            tma_load_2d_pair_mc(sa + A_STAGE + p * slice * BK, &map_b, &full_bar[stage], kb * BK,
                                tn * bn + (int)cta * (bn / 2) + p * slice, mc_mask);
          }
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
      }
      if (prof) { g.prof[blockIdx.x * PROF_SLOTS + 0] = (unsigned long long)t_wait; g.prof[blockIdx.x * PROF_SLOTS + 1] = (unsigned long long)(clock64() - t_begin); }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (pair leaders only) =====
    if (cta == 0 && lane == 0) {
      // D=S32, A=B=signed int8, K-major, N=bn, M=256 (128 rows in each CTA)
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
      const uint16_t all_mask = (uint16_t)((1u << CSIZE) - 1u), pair_mask = (uint16_t)(3u << (2 * p));
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      const bool prof = g.prof != nullptr;
      long long t_full = 0, t_acc = 0;
      const long long t_begin = prof ? clock64() : 0;
      for (int tile = cluster; tile < num_tiles; tile += num_clusters, ++it) {
        const int buf = it & 1;
        const uint32_t use = (uint32_t)(it >> 1);
        long long w0 = prof ? clock64() : 0;
        mbar_wait_bounded(&tmem_empty[buf], (use & 1) ^ 1);
        if (prof) t_acc += clock64() - w0;
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
        for (int kb = 0; kb < k_blocks; ++kb) {
          w0 = prof ? clock64() : 0;
          mbar_wait_bounded(&full_bar[stage], phase);
          if (prof) t_full += clock64() - w0;
          tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + stage * HALF_STAGE);
          const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sa + A_STAGE);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            umma_i8_pair(tmem_d, da + (uint64_t)(k * (UMMA_K >> 4)), db + (uint64_t)(k * (UMMA_K >> 4)), idesc,
                         (kb | k) ? 1u : 0u);
          }
          umma_commit_mc(&empty_bar[stage], all_mask);       // every producer of the cluster may refill the slot
          if (kb == k_blocks - 1) umma_commit_mc(&tmem_full[buf], pair_mask);
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
      }
      if (prof) {
        g.prof[blockIdx.x * PROF_SLOTS + 2] = (unsigned long long)t_full; g.prof[blockIdx.x * PROF_SLOTS + 3] = (unsigned long long)t_acc;
        g.prof[blockIdx.x * PROF_SLOTS + 4] = (unsigned long long)(clock64() - t_begin);
      }
    }
  } else {
    // ===== epilogue (warps 2.. of every CTA): this CTA's 128 rows; with eight warps, warps 2..5 take the first half of
    // a tile's 32-column chunks and warps 6..9 the second (TMEM lane quadrant = warp % 4 either way) =====
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int ep_tid = threadIdx.x - 64;
    const int split = EP2_WARPS == 8 ? (((bn >> 5) + 1) >> 1) << 5 : bn;     // first column of the second half
    const int c_begin = half ? split : 0, c_end = half ? bn : split;
    const float rq_s = g.rq_codes ? g.rq_scale[0] : 1.f;
    const float rq_o = (g.rq_codes && g.rq_offset) ? rintf(g.rq_offset[0]) : 0.f;
    int it = 0;
    bool store_pending = false;
    const bool prof = g.prof != nullptr && ep_tid == 0;
    long long t_tile = 0, t_cols = 0;
    long long pt[4] = {0, 0, 0, 0};
    const long long t_begin = prof ? clock64() : 0;
    for (int tile = cluster; tile < num_tiles; tile += num_clusters, ++it) {
      const int tm = (tile % groups_m) * P + p, tn = tile / groups_m;
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      const long long c0t = prof ? clock64() : 0;
      asm volatile("bar.sync 1, %0;" ::"n"(EP2_THREADS) : "memory");
      const bool wide = epilogue_bar_or<EP2_THREADS>(stage_col_params(g, tn * bn, bn, ep_tid, EP2_THREADS, col_params, col_ints));
      if (prof) t_cols += clock64() - c0t;
      const int row = tm * TM + (int)cta * BM + quad * 32 + lane;
      const int32_t rx = (g.ow && row < g.M) ? g.rowsum_x[row] : 0;

      const long long w0 = prof ? clock64() : 0;
      mbar_wait_bounded(&tmem_full[buf], use & 1);
      if (prof) t_tile += clock64() - w0;
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * BN);
      epilogue_tile<OutT>(g, &map_y, taddr, bn, row, tm * TM + (int)cta * BM + quad * 32, tn * bn, wide, rx, col_params,
                          col_ints, g.tma_store ? out_stage + (half * 4 + quad) * 4096 : nullptr, rq_s, rq_o, store_pending,
                          prof ? pt : nullptr, c_begin, c_end);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tmem_empty[buf]);   // EP2_WARPS x 2 CTAs arrivals free the accumulator
    }
    if (store_pending && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (prof) {
      unsigned long long* o = g.prof + (size_t)blockIdx.x * PROF_SLOTS;
      o[5] = (unsigned long long)t_tile; o[6] = (unsigned long long)(clock64() - t_begin); o[7] = (unsigned long long)t_cols;
      for (int i = 0; i < 4; ++i) o[8 + i] = (unsigned long long)pt[i];
      o[12] = (unsigned long long)it;
    }
  }

  tc_fence_before();
  cluster_sync_all();                         // nobody leaves while a peer may still touch its smem / barriers
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}


// ---- small helper kernels ------------------------------------------------------------------------
// rowsum[r] = sum_k q[r,k]   (one warp per row, 16-byte loads, dp4a against ones)
__global__ void __launch_bounds__(256) rowsum_i8_kernel(const int8_t* __restrict__ q, int32_t* __restrict__ out,
                                                        long long R, long long K) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= R) return;
  const int lane = threadIdx.x & 31;
  const int8_t* p = q + row * K;
  int acc = 0;
  const bool vec = (K % 16 == 0) && ((reinterpret_cast<uintptr_t>(p) & 15u) == 0);
  if (vec) {
    for (long long i = lane * 16; i < K; i += 32 * 16) {
      const int4 v = *reinterpret_cast<const int4*>(p + i);
      acc = __dp4a(v.x, 0x01010101, acc);
      acc = __dp4a(v.y, 0x01010101, acc);
      acc = __dp4a(v.z, 0x01010101, acc);
      acc = __dp4a(v.w, 0x01010101, acc);
    }
  } else {
    for (long long i = lane; i < K; i += 32) acc += p[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[row] = acc;
}

// ---- host side -------------------------------------------------------------------------------------
static int make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t K, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("qlinear: cuTensorMapEncodeTiled is not available from the driver"); return FFQ_ERR_CUDA; }
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)K};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("qlinear: cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return FFQ_ERR_CUDA; }
  return FFQ_OK;
}

// [M, N] output, boxes of 32 rows x 128 bytes, 128B-swizzled in shared memory
static int make_out_map(CUtensorMap* map, void* y, int64_t M, int64_t N, int y_dtype) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("qlinear: cuTensorMapEncodeTiled is not available from the driver"); return FFQ_ERR_CUDA; }
  const int es = y_dtype == FFQ_F32 ? 4 : 2;
  const cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
  const cuuint64_t strides[1] = {(cuuint64_t)N * es};
  const cuuint32_t box[2] = {(cuuint32_t)(128 / es), 32};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = y_dtype == FFQ_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : (y_dtype == FFQ_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
  const CUresult r = enc(map, dt, 2, y, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("qlinear: cuTensorMapEncodeTiled (output) failed with CUresult %d", (int)r); return FFQ_ERR_CUDA; }
  return FFQ_OK;
}

}  // namespace ffq

using namespace ffq;

static unsigned long long* g_gemm_prof = nullptr;     // test hook, see ffq_debug_gemm_profile

extern "C" {

int ffq_rowsum_i8(const int8_t* q, int32_t* rowsum, int64_t R, int64_t K, void* stream) {
  if (R <= 0) return FFQ_OK;
  rowsum_i8_kernel<<<(unsigned int)((R + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(q, rowsum, R, K);
  FFQ_LAUNCH_CHECK();
  return FFQ_OK;
}

void ffq_debug_gemm_profile(unsigned long long* counters_dev) { g_gemm_prof = counters_dev; }

}  // extern "C"

// cluster launch of the pair kernel: P pairs (2P CTAs) per cluster
template <typename OutT, int P>
static int launch_pairs(const CUtensorMap& map_a, const CUtensorMap& map_b, const CUtensorMap& map_y, const GemmArgs& g,
                        long long cluster_tiles, cudaStream_t st) {
  static std::atomic<uint64_t> attr_done{0};
  static int max_clusters[64] = {0};
  auto kern = w8a8_gemm2_kernel<OutT, P>;
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2 * P; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.blockDim = dim3(GEMM2_THREADS); cfg.dynamicSmemBytes = SMEM2_BYTES; cfg.stream = st; cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  const cudaError_t e = once_per_device(attr_done, [&]() -> cudaError_t {
    cudaError_t r = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES);
    if (r != cudaSuccess) return r;
    if (P > 4) { r = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1); if (r != cudaSuccess) return r; }
    cfg.gridDim = dim3(2 * P * (sm_count() / (2 * P)));
    int n = 0;
    r = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);      // GPCs with an SM count that 2P does not divide fit fewer
    if (r != cudaSuccess) return r;
    max_clusters[dev] = n > 0 ? n : 1;
    return cudaSuccess;
  });
  if (e != cudaSuccess) { set_error("qlinear_w8a8: cannot configure the %d-CTA cluster kernel: %s", 2 * P, cudaGetErrorString(e)); return FFQ_ERR_CUDA; }
  const long long clusters = cluster_tiles < max_clusters[dev] ? cluster_tiles : max_clusters[dev];
  cfg.gridDim = dim3((unsigned)(2 * P * clusters));
  const cudaError_t le = cudaLaunchKernelEx(&cfg, kern, map_a, map_b, map_y, g);
  count_launch();
  if (le != cudaSuccess) { cudaGetLastError(); set_error("qlinear_w8a8: cluster launch failed: %s", cudaGetErrorString(le)); return FFQ_ERR_CUDA; }
  return FFQ_OK;
}

template <int P>
static int launch_pairs_dt(int y_dtype, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& y, const GemmArgs& g,
                           long long t, cudaStream_t st) {
  switch (y_dtype) {
    case FFQ_F32: return launch_pairs<float, P>(a, b, y, g, t, st);
    case FFQ_BF16: return launch_pairs<__nv_bfloat16, P>(a, b, y, g, t, st);
    default: return launch_pairs<__half, P>(a, b, y, g, t, st);
  }
}

// pairs per cluster for the pair kernel: FFQ_GEMM_CLUSTER = 2 | 4 | 8 CTAs overrides the shape heuristic
static int env_cluster_pairs() {
  const char* e = getenv("FFQ_GEMM_CLUSTER");       // read per call: tests and A/B runs switch it inside one process
  const int c = e ? atoi(e) : 0;
  return (c == 2 || c == 4 || c == 8) ? c / 2 : 0;
}


extern "C" {

int ffq_qlinear_w8a8(const int8_t* qx, const int8_t* qw, void* y, int y_dtype, int64_t M, int64_t N, int64_t K,
                     const float* sx, const float* ox, const float* sw, const float* ow, const int32_t* rowsum_w,
                     const int32_t* rowsum_x, const void* bias, int bias_dtype, const ffq_requant_t* requant,
                     void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (M <= 0 || N <= 0) return FFQ_OK;
  if (K <= 0 || K % 16 != 0) { set_error("qlinear_w8a8: K must be a positive multiple of 16 (got %lld)", (long long)K); return FFQ_ERR_UNSUPPORTED; }
  if ((reinterpret_cast<uintptr_t>(qx) & 15u) || (reinterpret_cast<uintptr_t>(qw) & 15u)) {
    set_error("qlinear_w8a8: operand pointers must be 16-byte aligned"); return FFQ_ERR_UNSUPPORTED;
  }
  if (!(y_dtype == FFQ_F32 || y_dtype == FFQ_BF16 || y_dtype == FFQ_F16)) {
    set_error("qlinear_w8a8: output dtype must be float32/bfloat16/float16"); return FFQ_ERR_UNSUPPORTED;
  }
  if (M > 0x7fffffffll || N > 0x7fffffffll || K > 0x7fffffffll) { set_error("qlinear_w8a8: dimension too large"); return FFQ_ERR_UNSUPPORTED; }
  if (ow != nullptr && rowsum_x == nullptr) { set_error("qlinear_w8a8: rowsum_x is required when the weight has an offset"); return FFQ_ERR_INVALID; }
  if (y == nullptr && (requant == nullptr || requant->codes == nullptr)) { set_error("qlinear_w8a8: no output requested"); return FFQ_ERR_INVALID; }

  CUtensorMap map_a;
  int rc;
  if ((rc = make_map(&map_a, qx, M, K, BM)) != FFQ_OK) return rc;
  GemmArgs g{};
  g.M = (int)M; g.N = (int)N; g.K = (int)K; g.y = y; g.y_dt = y_dtype;
  g.sx = sx; g.ox = ox; g.sw = sw; g.ow = ow; g.rowsum_w = rowsum_w; g.bias = bias; g.bias_dt = bias_dtype;
  g.rowsum_x = rowsum_x;
  g.prof = g_gemm_prof;
  { const char* e = getenv("FFQ_GEMM_DEBUG"); g.dbg = e ? atoi(e) : 0; }
  // staged TMA-store epilogue whenever the output can be described by a tensor map (16-byte aligned rows);
  // FFQ_GEMM_DIRECT_STORE=1 keeps the per-lane vector stores (A/B switch)
  CUtensorMap map_y;
  memset(&map_y, 0, sizeof(map_y));
  const int es_y = y_dtype == FFQ_F32 ? 4 : 2;
  if (y != nullptr && (N * es_y) % 16 == 0 && (reinterpret_cast<uintptr_t>(y) & 15u) == 0 && getenv("FFQ_GEMM_DIRECT_STORE") == nullptr) {
    if ((rc = make_out_map(&map_y, y, M, N, y_dtype)) != FFQ_OK) return rc;
    g.tma_store = 1;
  }
  if (requant != nullptr && requant->codes != nullptr) {
    if (requant->scale == nullptr) { set_error("qlinear_w8a8: requant needs a scale"); return FFQ_ERR_INVALID; }
    if (!(requant->num_bits >= 1 && requant->num_bits <= 8)) { set_error("qlinear_w8a8: requant codes are int8: num_bits must be in [1, 8]"); return FFQ_ERR_BITWIDTH; }
    g.rq_scale = requant->scale; g.rq_offset = requant->offset; g.rq_codes = requant->codes; g.rq_rowsum = requant->rowsum;
    g.rq_lo = -(float)exp2(requant->num_bits - 1.0); g.rq_hi = (float)exp2(requant->num_bits - 1.0) - 1.f;
  }
  const long long tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const long long pair_tiles = ((M + 2 * BM - 1) / (2 * BM)) * ((N + BN - 1) / BN);
  // the pair kernel needs M > 128 to have work for both CTAs; FFQ_GEMM_1CTA=1 forces the single-CTA kernel
  static const bool force_1cta = getenv("FFQ_GEMM_1CTA") != nullptr;
  // ... and when the pair tiles would leave CTA pairs idle while the 128-row tiles still fit in one wave (e.g.
  // the k/v projections, N = 1024 at M = 2048: 32 pair tiles on 74 pairs vs 64 tiles on 148 SMs), the single-CTA
  // kernel finishes the same work in half-size tiles, all at once
  const bool underfilled = pair_tiles < sm_count() / 2 && tiles <= sm_count();
  const bool use_pair = !force_1cta && M > BM && !underfilled;
  if (use_pair) {
    // columns per pair tile: 256, or 224 when that removes a nearly empty last wave (e.g. N = 14336 at M = 2048:
    // 448 tiles on 74 pairs = 6.05 waves -> 512 tiles = 6.92 waves of 0.94x the per-tile cost; operand delivery,
    // A 128 rows + B bn/2 rows per CTA and k-block, is what a tile costs)
    const long long pairs = sm_count() / 2;
    const long long tiles_m2 = (M + 2 * BM - 1) / (2 * BM);
    // pairs per cluster (B multicast): only when the M tiles fill the cluster
    int P = env_cluster_pairs();
    if (P == 0) P = 1;
    while (P > 1 && tiles_m2 % P != 0) P >>= 1;
    auto cost = [&](int bn) {
      const long long t = tiles_m2 * ((N + bn - 1) / bn);
      return (double)((t + pairs - 1) / pairs) * (128.0 + bn / 2.0 / P);
    };
    static const bool force_256 = getenv("FFQ_GEMM_BN256") != nullptr;
    g.bn = (!force_256 && P <= 2 && N % 32 == 0 && cost(224) < 0.97 * cost(256)) ? 224 : BN;
    CUtensorMap map_b2;     // B box = the slice of the tile's columns this CTA fetches
    if ((rc = make_map(&map_b2, qw, N, K, g.bn / 2 / P)) != FFQ_OK) return rc;
    const long long cluster_tiles = (tiles_m2 / P) * ((N + g.bn - 1) / g.bn);
    switch (P) {
      case 4: return launch_pairs_dt<4>(y_dtype, map_a, map_b2, map_y, g, cluster_tiles, st);
      case 2: return launch_pairs_dt<2>(y_dtype, map_a, map_b2, map_y, g, cluster_tiles, st);
      default: return launch_pairs_dt<1>(y_dtype, map_a, map_b2, map_y, g, cluster_tiles, st);
    }
  }
  static std::atomic<uint64_t> attr_done{0};
  const cudaError_t attr_err = once_per_device(attr_done, []() -> cudaError_t {
    cudaError_t e1 = cudaFuncSetAttribute(w8a8_gemm_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    cudaError_t e2 = cudaFuncSetAttribute(w8a8_gemm_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    cudaError_t e3 = cudaFuncSetAttribute(w8a8_gemm_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    return e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3);
  });
  if (attr_err != cudaSuccess) { set_error("qlinear_w8a8: cannot reserve %d bytes of shared memory: %s", SMEM_BYTES, cudaGetErrorString(attr_err)); return FFQ_ERR_CUDA; }
  // single-CTA kernel: 128 x 256 tiles, or 128 x 128 when the wider ones would leave SMs idle (e.g. the k/v
  // projections, N = 1024 at M = 2048: 64 tiles -> 128 tiles on 148 SMs at two thirds of the per-tile operand traffic)
  {
    const long long sms = sm_count();
    const long long tm1 = (M + BM - 1) / BM;
    auto cost1 = [&](int bn) {
      const long long t = tm1 * ((N + bn - 1) / bn);
      return (double)((t + sms - 1) / sms) * (128.0 + bn);
    };
    static const bool force_256 = getenv("FFQ_GEMM_BN256") != nullptr;
    g.bn = (!force_256 && N % 32 == 0 && cost1(128) < 0.97 * cost1(BN)) ? 128 : BN;
  }
  CUtensorMap map_b1;
  if ((rc = make_map(&map_b1, qw, N, K, g.bn)) != FFQ_OK) return rc;
  const long long tiles1 = ((M + BM - 1) / BM) * ((N + g.bn - 1) / g.bn);
  const int grid1 = (int)(tiles1 < sm_count() ? tiles1 : sm_count());
  switch (y_dtype) {
    case FFQ_F32: launch_pdl(w8a8_gemm_kernel<float>, dim3(grid1), dim3(GEMM_THREADS), SMEM_BYTES, st, map_a, map_b1, map_y, g); break;
    case FFQ_BF16: launch_pdl(w8a8_gemm_kernel<__nv_bfloat16>, dim3(grid1), dim3(GEMM_THREADS), SMEM_BYTES, st, map_a, map_b1, map_y, g); break;
    default: launch_pdl(w8a8_gemm_kernel<__half>, dim3(grid1), dim3(GEMM_THREADS), SMEM_BYTES, st, map_a, map_b1, map_y, g); break;
  }
  FFQ_LAUNCH_CHECK();
  return FFQ_OK;
}

}  // extern "C"

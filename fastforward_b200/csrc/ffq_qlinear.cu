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

#include <map>
#include <mutex>
#include <tuple>

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
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 4 * BN * 4 + 256 + 1024;   // + column params + barriers + align

struct GemmArgs {
  int M, N, K;
  void* y; int y_dt;
  const float* sx; const float* ox;             // activation scale / offset (one element each; ox may be null)
  const float* sw; const float* ow;             // per output column (ow may be null)
  const int32_t* rowsum_w;                      // per output column: sum_k qw[n,k]
  const void* bias; int bias_dt;                // per output column, optional
  const int32_t* rowsum_x;                      // per output row; used with ow
  int bn;                                       // pair kernel: output columns per tile (256, or 224 when that evens out the waves)
};

// Per-column epilogue parameters of one tile, derived where they are consumed (no separate launch, no scratch):
//   alpha[n] = sx*sw[n];  bias[n] as float;  cnst[n] = ox*rowsum_w[n] + K*ox*ow[n];  own[n] = ow[n]
// with ox, ow rounded to integers exactly as dequantize_by_tile rounds them.
__device__ __forceinline__ void stage_col_params(const GemmArgs& g, int n0, int bn, int tid, int nthreads,
                                                 float* col_params, int32_t* col_ints) {
  const float sx = g.sx[0];
  const int o_x = g.ox ? __float2int_rn(rintf(g.ox[0])) : 0;
  for (int c = tid; c < bn; c += nthreads) {
    const int n = n0 + c;
    const bool in = n < g.N;
    const int o_w = (in && g.ow) ? __float2int_rn(rintf(g.ow[n])) : 0;
    col_params[c] = in ? sx * g.sw[n] : 0.f;
    col_params[bn + c] = (in && g.bias) ? load_as_float(g.bias, g.bias_dt, n) : 0.f;
    col_ints[2 * bn + c] = in ? o_x * g.rowsum_w[n] + g.K * o_x * o_w : 0;
    col_ints[3 * bn + c] = o_w;
  }
}

template <typename OutT>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
w8a8_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  float* col_params = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);       // [4][BN]
  int32_t* col_ints = reinterpret_cast<int32_t*>(col_params);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + 4 * BN * 4);
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
    OutT* __restrict__ y = static_cast<OutT*>(g.y);
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int tm = tile % tiles_m, tn = tile / tiles_m;
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      // stage this tile's column parameters in shared memory (named barrier over the 4 epilogue warps)
      asm volatile("bar.sync 1, 128;" ::: "memory");         // previous tile's readers are done
      stage_col_params(g, tn * bn, bn, ep_tid, 128, col_params, col_ints);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int row = tm * BM + quad * 32 + lane;
      const int32_t rx = (g.ow && row < g.M) ? g.rowsum_x[row] : 0;

      mbar_wait(&tmem_full[buf], use & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * BN);
#pragma unroll 1
      for (int c0 = 0; c0 < bn; c0 += 32) {
        uint32_t acc[32];
        tmem_ld32(taddr + (uint32_t)c0, acc);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int32_t t = (int32_t)acc[j] + col_ints[2 * bn + c0 + j] + col_ints[3 * bn + c0 + j] * rx;
          v[j] = fmaf(col_params[c0 + j], (float)t, col_params[bn + c0 + j]);
        }
        const int n0 = tn * bn + c0;
        if (row < g.M && n0 < g.N) {
          const int ncols = (g.N - n0) < 32 ? (g.N - n0) : 32;
          store_chunk<OutT>(y + (size_t)row * g.N + n0, v, ncols);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);          // 4 arrivals (one per epilogue warp) free the accumulator
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

// ================================================================================================
// 2-CTA variant: a CTA pair (cluster 2x1x1, two SMs of one TPC) owns a 256x256 output tile.
// Each CTA stages its own 128 rows of A and its own 128-row half of B (32 KB per k-block instead of
// 48 KB: the operand traffic per MMA drops by a third, which is what bounds the 1-CTA kernel), the
// leader CTA issues tcgen05.mma.cta_group::2 (M=256) reading both CTAs' shared memory, and each
// CTA's TMEM holds its 128 accumulator rows.  TMA completions of both CTAs land on the leader's
// "full" barrier (cta_group::2 loads), tcgen05.commit multicasts "slot free" / "tile ready" to both
// CTAs, and the peer's epilogue warps release the accumulator on the leader's barrier.
// ================================================================================================
constexpr int STAGES2 = 6;
constexpr int HALF_STAGE = A_STAGE + BM * BK;      // A 128x128B + B half 128x128B = 32 KB
constexpr int SMEM2_BYTES = STAGES2 * HALF_STAGE + 4 * BN * 4 + 256 + 1024;
template <typename OutT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
w8a8_gemm2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  float* col_params = reinterpret_cast<float*>(smem + STAGES2 * HALF_STAGE);       // [4][BN]
  int32_t* col_ints = reinterpret_cast<int32_t*>(col_params);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES2 * HALF_STAGE + 4 * BN * 4);
  uint64_t* full_bar = bars;                  // [STAGES2]  (the leader's copy is the one in use)
  uint64_t* empty_bar = bars + STAGES2;       // [STAGES2]  (each CTA waits on its own copy)
  uint64_t* tmem_full = bars + 2 * STAGES2;   // [2]        (each CTA waits on its own copy)
  uint64_t* tmem_empty = bars + 2 * STAGES2 + 2;   // [2]   (leader's copy, 8 arrivals)
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES2 + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta = cluster_ctarank();
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  constexpr int TM = 2 * BM;                  // 256 rows per pair tile
  const int bn = g.bn;                       // columns per tile; the smem / TMEM layout keeps its 256-column pitch
  const int tiles_m = (g.M + TM - 1) / TM, tiles_n = (g.N + bn - 1) / bn;
  const uint32_t stage_tx = 2u * (uint32_t)(A_STAGE + (bn / 2) * BK);
  const int num_tiles = tiles_m * tiles_n;
  const int k_blocks = (g.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES2; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)),
                 "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  cluster_sync_all();                         // barriers of BOTH CTAs are initialised before any remote use
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===== TMA producer (both CTAs; completions count on the leader's full barrier) =====
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
      int stage = 0; uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int tm = tile % tiles_m, tn = tile / tiles_m;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = stage_base + stage * HALF_STAGE;
          if (cta == 0) mbar_expect_tx(&full_bar[stage], stage_tx);
          tma_load_2d_pair(sa, &map_a, &full_bar[stage], kb * BK, tm * TM + (int)cta * BM);
          tma_load_2d_pair(sa + A_STAGE, &map_b, &full_bar[stage], kb * BK, tn * bn + (int)cta * (bn / 2));
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only) =====
    if (cta == 0 && lane == 0) {
      // D=S32, A=B=signed int8, K-major, N=bn, M=256 (128 rows in each CTA)
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
        const int buf = it & 1;
        const uint32_t use = (uint32_t)(it >> 1);
        mbar_wait(&tmem_empty[buf], (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + stage * HALF_STAGE);
          const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sa + A_STAGE);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            umma_i8_pair(tmem_d, da + (uint64_t)(k * (UMMA_K >> 4)), db + (uint64_t)(k * (UMMA_K >> 4)), idesc,
                         (kb | k) ? 1u : 0u);
          }
          umma_commit_pair(&empty_bar[stage]);               // both CTAs' producers may refill the slot
          if (kb == k_blocks - 1) umma_commit_pair(&tmem_full[buf]);
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===== epilogue (warps 2..5 of both CTAs): this CTA's 128 rows =====
    const int quad = warp & 3;
    const int ep_tid = threadIdx.x - 64;
    OutT* __restrict__ y = static_cast<OutT*>(g.y);
    int it = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
      const int tm = tile % tiles_m, tn = tile / tiles_m;
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      stage_col_params(g, tn * bn, bn, ep_tid, 128, col_params, col_ints);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int row = tm * TM + (int)cta * BM + quad * 32 + lane;
      const int32_t rx = (g.ow && row < g.M) ? g.rowsum_x[row] : 0;

      mbar_wait(&tmem_full[buf], use & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * BN);
#pragma unroll 1
      for (int c0 = 0; c0 < bn; c0 += 32) {
        uint32_t acc[32];
        tmem_ld32(taddr + (uint32_t)c0, acc);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int32_t t = (int32_t)acc[j] + col_ints[2 * bn + c0 + j] + col_ints[3 * bn + c0 + j] * rx;
          v[j] = fmaf(col_params[c0 + j], (float)t, col_params[bn + c0 + j]);
        }
        const int n0 = tn * bn + c0;
        if (row < g.M && n0 < g.N) {
          const int ncols = (g.N - n0) < 32 ? (g.N - n0) : 32;
          store_chunk<OutT>(y + (size_t)row * g.N + n0, v, ncols);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tmem_empty[buf]);   // 8 arrivals (4 warps x 2 CTAs) free the accumulator
    }
  }

  tc_fence_before();
  cluster_sync_all();                         // nobody leaves while the peer may still touch its smem / barriers
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

}  // namespace ffq

using namespace ffq;

extern "C" {

size_t ffq_qlinear_workspace_bytes(int64_t N) { (void)N; return 0; }

int ffq_rowsum_i8(const int8_t* q, int32_t* rowsum, int64_t R, int64_t K, void* stream) {
  if (R <= 0) return FFQ_OK;
  rowsum_i8_kernel<<<(unsigned int)((R + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(q, rowsum, R, K);
  FFQ_LAUNCH_CHECK();
  return FFQ_OK;
}

int ffq_qlinear_w8a8(const int8_t* qx, const int8_t* qw, void* y, int y_dtype, int64_t M, int64_t N, int64_t K,
                     const float* sx, const float* ox, const float* sw, const float* ow, const int32_t* rowsum_w,
                     const int32_t* rowsum_x, const void* bias, int bias_dtype, void* workspace,
                     size_t workspace_bytes, void* stream) {
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
  (void)workspace; (void)workspace_bytes;      // kept in the ABI; the column parameters are derived inside the kernel

  CUtensorMap map_a;
  int rc;
  if ((rc = make_map(&map_a, qx, M, K, BM)) != FFQ_OK) return rc;
  GemmArgs g{};
  g.M = (int)M; g.N = (int)N; g.K = (int)K; g.y = y; g.y_dt = y_dtype;
  g.sx = sx; g.ox = ox; g.sw = sw; g.ow = ow; g.rowsum_w = rowsum_w; g.bias = bias; g.bias_dt = bias_dtype;
  g.rowsum_x = rowsum_x;
  const long long tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const long long pair_tiles = ((M + 2 * BM - 1) / (2 * BM)) * ((N + BN - 1) / BN);
  // the pair kernel needs M > 128 to have work for both CTAs; FFQ_GEMM_1CTA=1 forces the single-CTA kernel
  static const bool force_1cta = getenv("FFQ_GEMM_1CTA") != nullptr;
  // ... and when the pair tiles would leave CTA pairs idle while the 128-row tiles still fit in one wave (e.g.
  // the k/v projections, N = 1024 at M = 2048: 32 pair tiles on 74 pairs vs 64 tiles on 148 SMs), the single-CTA
  // kernel finishes the same work in half-size tiles, all at once
  const bool underfilled = pair_tiles < sm_count() / 2 && tiles <= sm_count();
  const bool use_pair = !force_1cta && M > BM && !underfilled;
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    cudaError_t e1 = cudaFuncSetAttribute(w8a8_gemm_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    cudaError_t e2 = cudaFuncSetAttribute(w8a8_gemm_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    cudaError_t e3 = cudaFuncSetAttribute(w8a8_gemm_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    attr_err = e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3);
    cudaError_t f1 = cudaFuncSetAttribute(w8a8_gemm2_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES);
    cudaError_t f2 = cudaFuncSetAttribute(w8a8_gemm2_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES);
    cudaError_t f3 = cudaFuncSetAttribute(w8a8_gemm2_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES);
    if (attr_err == cudaSuccess) attr_err = f1 != cudaSuccess ? f1 : (f2 != cudaSuccess ? f2 : f3);
  });
  if (attr_err != cudaSuccess) { set_error("qlinear_w8a8: cannot reserve %d bytes of shared memory: %s", SMEM_BYTES, cudaGetErrorString(attr_err)); return FFQ_ERR_CUDA; }
  if (use_pair) {
    // columns per pair tile: 256, or 224 when that removes a nearly empty last wave (e.g. N = 14336 at M = 2048:
    // 448 tiles on 74 pairs = 6.05 waves -> 512 tiles = 6.92 waves of 0.94x the per-tile cost; operand delivery,
    // A 128 rows + B bn/2 rows per CTA and k-block, is what a tile costs)
    const long long pairs = sm_count() / 2;
    const long long tiles_m2 = (M + 2 * BM - 1) / (2 * BM);
    auto cost = [&](int bn) {
      const long long t = tiles_m2 * ((N + bn - 1) / bn);
      return (double)((t + pairs - 1) / pairs) * (128.0 + bn / 2.0);
    };
    static const bool force_256 = getenv("FFQ_GEMM_BN256") != nullptr;
    g.bn = (!force_256 && N % 32 == 0 && cost(224) < 0.97 * cost(256)) ? 224 : BN;
    CUtensorMap map_b2;     // B box = this CTA's half of the tile's columns
    if ((rc = make_map(&map_b2, qw, N, K, g.bn / 2)) != FFQ_OK) return rc;
    const int max_pairs = sm_count() / 2;
    const long long pair_tiles_bn = tiles_m2 * ((N + g.bn - 1) / g.bn);
    const int grid2 = 2 * (int)(pair_tiles_bn < max_pairs ? pair_tiles_bn : max_pairs);
    switch (y_dtype) {
      case FFQ_F32: w8a8_gemm2_kernel<float><<<grid2, GEMM_THREADS, SMEM2_BYTES, st>>>(map_a, map_b2, g); break;
      case FFQ_BF16: w8a8_gemm2_kernel<__nv_bfloat16><<<grid2, GEMM_THREADS, SMEM2_BYTES, st>>>(map_a, map_b2, g); break;
      default: w8a8_gemm2_kernel<__half><<<grid2, GEMM_THREADS, SMEM2_BYTES, st>>>(map_a, map_b2, g); break;
    }
    FFQ_LAUNCH_CHECK();
    return FFQ_OK;
  }
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
    case FFQ_F32: w8a8_gemm_kernel<float><<<grid1, GEMM_THREADS, SMEM_BYTES, st>>>(map_a, map_b1, g); break;
    case FFQ_BF16: w8a8_gemm_kernel<__nv_bfloat16><<<grid1, GEMM_THREADS, SMEM_BYTES, st>>>(map_a, map_b1, g); break;
    default: w8a8_gemm_kernel<__half><<<grid1, GEMM_THREADS, SMEM_BYTES, st>>>(map_a, map_b1, g); break;
  }
  FFQ_LAUNCH_CHECK();
  return FFQ_OK;
}

}  // extern "C"

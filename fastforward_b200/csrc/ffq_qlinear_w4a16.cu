// ffq_qlinear_w4a16.cu -- weight-only quantized linear (W4A16; any <= 8-bit integer weight codes)
// on Blackwell tensor cores (a12, section 8b "ffq_qlinear_w4a16").
//
//   y[m,n] = sum_k x[m,k] * w[n,k] + bias[n],   w[n,k] = round_to_dtype_of_x((qw[n,k] + rint(ow[n,g])) * sw[n,g]),  g = k / group
//
// The reference's route for this op is its fallback (_gen/fallback.py:94-108): dequantize the
// weight into a full bf16 tensor in HBM (1 B/elem read + 2 B/elem written), then F.linear re-reads
// it (2 B/elem).  Here the int8 codes are the only weight bytes that ever cross HBM: the codes
// tile is dequantized INSIDE the k-loop into the 128B-swizzled bf16 B operand in shared memory,
// with exactly dequantize_by_tile's arithmetic (fp32 add, fp32 multiply, one rounding to the
// 16-bit dtype), and consumed by tcgen05.mma.kind::f16 with fp32 accumulation in TMEM -- so the
// result equals the fallback's up to the GEMM's accumulation order.
//
// One CTA pair (cluster 2x1x1) per 256x256 output tile, persistent, warp-specialised, with THREE
// decoupled shared-memory rings so that neither HBM latency nor the dequantisation sits on the MMA's
// critical path:
//   warp 0      A producer: TMA of the A tile (128 rows x 64 elem, 128B swizzle) into a 6-deep ring;
//               completes on the LEADER's barrier, slots freed by tcgen05.commit
//   warp 2      raw-code producer: TMA of this CTA's 128 weight rows x 64 B into an 8-deep ring,
//               slots freed by the dequantizers as soon as the codes are in registers
//   warps 8..15 dequantizers: raw codes -> (q + o) * s -> bf16/f16 -> 3-deep swizzled B ring of this CTA
//               (slot freed by tcgen05.commit), fence.proxy.async, arrive on the leader's barrier
//   warp 1      MMA issuer (leader CTA): 4 x tcgen05.mma.cta_group::2 (K = 16) per k-block
//   warps 4..7  epilogue: tcgen05.ld, + bias, convert, vector stores; overlaps the next tile's k-loop
// Roofline: tensor pipe (bf16 dense), 2*M*N*K flops; the dequantizers need ~4 issue slots per
// weight element per M-tile, about half of the issue capacity left beside the MMAs.
#include <cuda.h>

#include <atomic>
#include <cstdlib>

#include "ffq_common.cuh"
#include "ffq_umma.cuh"

namespace ffq {
namespace w4 {

constexpr int BM = 128, BN = 256, BK = 64;        // BK in 16-bit elements: one 128B swizzle atom
constexpr int TM = 2 * BM;
constexpr int UMMA_K = 16;
constexpr int SB = 3, SR = 8;                     // ring depths: dequantized B / raw codes
// MT = M tiles (of 256 rows) per pair tile.  MT = 2: ONE dequantized B stage feeds the MMAs of two 256-row tiles (two
// 128 x 256 fp32 accumulators per CTA = all 512 TMEM columns), so the dequantisation -- what bounds the kernel, see the
// header -- is done once per 512 activation rows; the price is an epilogue that no longer overlaps the next tile's MMAs.
template <int MT> struct RingA { static constexpr int depth = MT == 1 ? 6 : 3; };   // A ring: stages of MT x 16 KB
constexpr int A_BYTES = BM * BK * 2;              // 16 KB
constexpr int B_BYTES = (BN / 2) * BK * 2;        // 16 KB: this CTA's half of B, as 16-bit floats
constexpr int RAW_BYTES = (BN / 2) * BK;          // 8 KB: the same half as int8 codes
constexpr int DQ_WARPS = 8;
// warp roles (TMEM lane quadrant of an epilogue warp = warp % 4; dequantizers spread 2 per scheduler)
constexpr int WARP_TMA_A = 0, WARP_MMA = 1, WARP_TMA_RAW = 2, WARP_EPI0 = 4, WARP_DQ0 = 8;
constexpr int THREADS = (WARP_DQ0 + DQ_WARPS) * 32;   // 512 (warp 3 idles)
constexpr int TMEM_COLS = 512;
constexpr int SMEM_BYTES = 6 * A_BYTES + SB * B_BYTES + SR * RAW_BYTES + BN * 4 + 512 + 1024;     // same for MT = 1 and 2

__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (done) break;
    __nanosleep(500);
  }
}

struct Args {
  int M, N, K;
  void* y;
  const float* sw; const float* ow;     // [N][groups]
  int group, groups, kb_per_group;       // k elements per parameter; K / group; group / BK
  const void* bias; int bias_dt;
};

__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc),
      "r"(acc) : "memory");
}

template <typename T> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));   // upper half <- first source operand
  return r;
}
template <> __device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

// 4 signed bytes -> 4 floats holding (q + o) * s.  The byte is planted into the mantissa of 2^23
// (after flipping its sign bit: u = q + 128), so  as_float(0x4B0000uu) + (o - 2^23 - 128) == q + o
// exactly (integers below 2^24) with one PRMT and one FADD instead of an int->float conversion.
// Valid for |o| < 2^22; larger offsets take the plain conversion path.
__device__ __forceinline__ void dequant4_fast(uint32_t w, float c_fast, float s, float (&f)[4]) {
  const uint32_t u = w ^ 0x80808080u;
  f[0] = __fmul_rn(__fadd_rn(__uint_as_float(__byte_perm(u, 0x4B000000u, 0x7440)), c_fast), s);
  f[1] = __fmul_rn(__fadd_rn(__uint_as_float(__byte_perm(u, 0x4B000000u, 0x7441)), c_fast), s);
  f[2] = __fmul_rn(__fadd_rn(__uint_as_float(__byte_perm(u, 0x4B000000u, 0x7442)), c_fast), s);
  f[3] = __fmul_rn(__fadd_rn(__uint_as_float(__byte_perm(u, 0x4B000000u, 0x7443)), c_fast), s);
}

template <typename T, int MT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
w4a16_gemm2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_raw, const Args g) {
  constexpr int SA = RingA<MT>::depth;
  constexpr int A_STAGE = MT * A_BYTES;
  constexpr int TMQ = MT * TM;                                // rows of a pair tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_base = smem;                                   // [SA][128 rows x 128 B]   A operand, TMA, 128B swizzle
  uint8_t* b_base = a_base + SA * A_STAGE;                  // [SB][128 rows x 128 B]   this CTA's half of B, written by the dequantizers
  uint8_t* r_base = b_base + SB * B_BYTES;                  // [SR][128 rows x 64 B]    raw int8 codes, TMA, unswizzled
  float* col_bias = reinterpret_cast<float*>(r_base + SR * RAW_BYTES);      // [BN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(col_bias + BN);
  uint64_t* a_full = bars;                       // [SA] leader's copy: A bytes of both CTAs
  uint64_t* a_empty = a_full + SA;               // [SA] local copy, signalled by the leader's tcgen05.commit multicast
  uint64_t* b_full = a_empty + SA;               // [SB] leader's copy: 2 x DQ_WARPS dequantizer warps
  uint64_t* b_empty = b_full + SB;               // [SB] local copy, commit multicast
  uint64_t* raw_full = b_empty + SB;             // [SR] local: this CTA's raw code tile has landed
  uint64_t* raw_empty = raw_full + SR;           // [SR] local: DQ_WARPS dequantizer warps have read it
  uint64_t* tmem_full = raw_empty + SR;          // [2]
  uint64_t* tmem_empty = tmem_full + 2;          // [2]  leader's copy, 8 arrivals
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta = cluster_ctarank();
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int tiles_m = (g.M + TMQ - 1) / TMQ, tiles_n = (g.N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int k_blocks = g.K / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < SB; ++s) { mbar_init(&b_full[s], 2 * DQ_WARPS); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < SR; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], DQ_WARPS); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == WARP_MMA) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)),
                 "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == WARP_TMA_A) {
    // ===== A producer (both CTAs; completions count on the leader's barrier) =====
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
      int stage = 0; uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int tm = tile % tiles_m;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&a_empty[stage], phase ^ 1);
          if (cta == 0) mbar_expect_tx(&a_full[stage], 2 * A_STAGE);
#pragma unroll
          for (int h = 0; h < MT; ++h)
            tma_load_2d_pair(a_base + stage * A_STAGE + h * A_BYTES, &map_a, &a_full[stage], kb * BK,
                             tm * TMQ + h * TM + (int)cta * BM);
          if (++stage == SA) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == WARP_TMA_RAW) {
    // ===== raw-code producer: its own, deeper ring, released by the dequantizers (not by the MMAs) =====
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_raw) : "memory");
      int stage = 0; uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int tn = tile / tiles_m;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&raw_empty[stage], phase ^ 1);
          mbar_expect_tx(&raw_full[stage], RAW_BYTES);
          tma_load_2d(r_base + stage * RAW_BYTES, &map_raw, &raw_full[stage], kb * BK, tn * BN + (int)cta * (BN / 2));
          if (++stage == SR) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == WARP_MMA) {
    // ===== MMA issuer (leader CTA only) =====
    if (cta == 0 && lane == 0) {
      constexpr uint32_t fmt = Elem<T>::dt == FFQ_BF16 ? 1u : 0u;
      // D = F32, A = B = bf16/f16, both K-major, N = 256, M = 256 (128 rows in each CTA)
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
      int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
      int it = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
        // MT = 1: two accumulators alternate between tiles; MT = 2: both belong to this tile (one per 256-row half)
        const int buf = MT == 1 ? (it & 1) : 0;
        const uint32_t use = MT == 1 ? (uint32_t)(it >> 1) : (uint32_t)it;
        mbar_wait(&tmem_empty[buf], (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&a_full[sa], pa);
          mbar_wait(&b_full[sb], pb);
          tc_fence_after();
          const uint64_t db = make_smem_desc(smem_u32(b_base + sb * B_BYTES));
#pragma unroll
          for (int h = 0; h < MT; ++h) {
            const uint64_t da = make_smem_desc(smem_u32(a_base + sa * A_STAGE + h * A_BYTES));
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              // +16 elements = +32 bytes inside the swizzle atom == +2 in the (>>4) start-address field
              umma_f16_pair(tmem_d + (uint32_t)(h * BN), da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
            }
          }
          umma_commit_pair(&a_empty[sa]);
          umma_commit_pair(&b_empty[sb]);
          if (kb == k_blocks - 1) umma_commit_pair(&tmem_full[buf]);
          if (++sa == SA) { sa = 0; pa ^= 1; }
          if (++sb == SB) { sb = 0; pb ^= 1; }
        }
      }
    }
  } else if (warp >= WARP_EPI0 && warp < WARP_EPI0 + 4) {
    // ===== epilogue (4 warps of both CTAs): this CTA's 128 rows; TMEM lane quadrant = warp % 4 =====
    const int quad = warp & 3;
    const int ep_tid = threadIdx.x - WARP_EPI0 * 32;
    T* __restrict__ y = static_cast<T*>(g.y);
    int it = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
      const int tm = tile % tiles_m, tn = tile / tiles_m;
      const int buf = MT == 1 ? (it & 1) : 0;
      const uint32_t use = MT == 1 ? (uint32_t)(it >> 1) : (uint32_t)it;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int c = ep_tid; c < BN; c += 128) {
        const int n = tn * BN + c;
        col_bias[c] = (g.bias && n < g.N) ? load_as_float(g.bias, g.bias_dt, n) : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");

      mbar_wait_backoff(&tmem_full[buf], use & 1);           // a whole k-loop away: do not burn issue slots
      tc_fence_after();
#pragma unroll 1
      for (int h = 0; h < MT; ++h) {
        const int row = tm * TMQ + h * TM + (int)cta * BM + quad * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((MT == 1 ? buf : h) * BN);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t acc[32];
          tmem_ld32(taddr + (uint32_t)c0, acc);
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __fadd_rn(__uint_as_float(acc[j]), col_bias[c0 + j]);
          const int n0 = tn * BN + c0;
          if (row < g.M && n0 < g.N) {
            const int ncols = (g.N - n0) < 32 ? (g.N - n0) : 32;
            store_chunk<T>(y + (size_t)row * g.N + n0, v, ncols);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tmem_empty[buf]);
    }
  } else if (warp >= WARP_DQ0) {
    // ===== dequantizers (8 warps): raw int8 codes -> swizzled 16-bit B stage =====
    // Two warps per scheduler so that one warp's shared-memory / fence latency hides behind the
    // other's arithmetic.  Thread t owns 16 codes (16 raw bytes -> two 16-byte B chunks) of rows
    // t/4 and t/4 + 64 of this CTA's 128-row half.
    constexpr int ITEMS = (BN / 2) * 4 / (DQ_WARPS * 32);     // 2
    constexpr int ROW_STEP = DQ_WARPS * 8;                      // 64
    const int t = threadIdx.x - WARP_DQ0 * 32;
    const int chunk = t & 3;
    const int row0 = t >> 2;
    const uint32_t b0 = smem_u32(b_base), r0 = smem_u32(r_base);
    int sr = 0, sb = 0; uint32_t pr = 0, pb = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int tn = tile / tiles_m;
      const int nbase = tn * BN + (int)cta * (BN / 2) + row0;
      float s_cur[ITEMS], o_cur[ITEMS], s_nxt[ITEMS], o_nxt[ITEMS];
      bool fast = true;
      // raw parameter loads only: nothing here may consume the values, or the prefetch turns into a stall
      auto fetch = [&](int gi, float (&s)[ITEMS], float (&o)[ITEMS]) {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
          const int n = nbase + ROW_STEP * i;
          const bool in = n < g.N && gi < g.groups;
          s[i] = in ? __ldg(g.sw + (size_t)n * g.groups + gi) : 0.f;
          o[i] = (in && g.ow) ? __ldg(g.ow + (size_t)n * g.groups + gi) : 0.f;
        }
      };
      fetch(0, s_nxt, o_nxt);
      int gi = 0, left = 0;                          // k-blocks left in the current group
      for (int kb = 0; kb < k_blocks; ++kb) {
        if (left == 0) {                             // parameters of the NEXT group are requested one group ahead
          bool ok = true;
#pragma unroll
          for (int i = 0; i < ITEMS; ++i) { s_cur[i] = s_nxt[i]; o_cur[i] = rintf(o_nxt[i]); ok = ok && fabsf(o_cur[i]) < 4194304.f; }
          fast = __all_sync(0xffffffffu, ok);       // warp-uniform: the magic-number path needs |offset| < 2^22
          fetch(++gi, s_nxt, o_nxt);
          left = g.kb_per_group;
        }
        --left;
        mbar_wait(&raw_full[sr], pr);
        uint4 v[ITEMS];
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
          const int r = row0 + ROW_STEP * i;
          asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v[i].x), "=r"(v[i].y), "=r"(v[i].z), "=r"(v[i].w)
                       : "r"(r0 + (uint32_t)(sr * RAW_BYTES + r * BK + chunk * 16)));
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&raw_empty[sr]);            // the codes are in registers: the raw slot may be refilled
        if (++sr == SR) { sr = 0; pr ^= 1; }
        float f[ITEMS][16];
        if (fast) {                                  // ONE warp-uniform branch per k-block
#pragma unroll
          for (int i = 0; i < ITEMS; ++i) {
            const float cf = __fsub_rn(o_cur[i], 8388736.f);   // o - 2^23 - 128, exact for |o| < 2^22
            const uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) { float q[4]; dequant4_fast(w[j], cf, s_cur[i], q); f[i][4 * j] = q[0]; f[i][4 * j + 1] = q[1]; f[i][4 * j + 2] = q[2]; f[i][4 * j + 3] = q[3]; }
          }
        } else {
#pragma unroll
          for (int i = 0; i < ITEMS; ++i) {
            const uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
            for (int j = 0; j < 16; ++j) f[i][j] = __fmul_rn(__fadd_rn((float)(int8_t)(w[j >> 2] >> (8 * (j & 3))), o_cur[i]), s_cur[i]);
          }
        }
        mbar_wait(&b_empty[sb], pb ^ 1);                        // the MMAs that read this B slot have retired
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
          const int r = row0 + ROW_STEP * i;
          // 128B swizzle: 16-byte chunk j of row r lives at chunk j ^ (r % 8); rows are 128 B apart
          const uint32_t rowp = b0 + (uint32_t)(sb * B_BYTES + r * 128);
          const int sw7 = r & 7;
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(rowp + (uint32_t)(((2 * chunk) ^ sw7) << 4)),
                       "r"(pack2<T>(f[i][0], f[i][1])), "r"(pack2<T>(f[i][2], f[i][3])), "r"(pack2<T>(f[i][4], f[i][5])),
                       "r"(pack2<T>(f[i][6], f[i][7])) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(rowp + (uint32_t)(((2 * chunk + 1) ^ sw7) << 4)),
                       "r"(pack2<T>(f[i][8], f[i][9])), "r"(pack2<T>(f[i][10], f[i][11])), "r"(pack2<T>(f[i][12], f[i][13])),
                       "r"(pack2<T>(f[i][14], f[i][15])) : "memory");
        }
        // generic-proxy stores -> visible to the async proxy (tcgen05.mma), then a plain remote arrive: a
        // .release.cluster arrive would lower to MEMBAR.ALL.GPU + ERRBAR per k-block (measured: 4x slower kernel)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&b_full[sb]);
        if (++sb == SB) { sb = 0; pb ^= 1; }
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == WARP_MMA) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

static int make_map2(CUtensorMap* map, CUtensorMapDataType dt, int elem_bytes, const void* base, int64_t rows, int64_t K,
                     int box_k, int box_rows, CUtensorMapSwizzle swz) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("qlinear_w4a16: cuTensorMapEncodeTiled is not available from the driver"); return FFQ_ERR_CUDA; }
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)K * elem_bytes};
  const cuuint32_t box[2] = {(cuuint32_t)box_k, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("qlinear_w4a16: cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return FFQ_ERR_CUDA; }
  return FFQ_OK;
}

}  // namespace w4
}  // namespace ffq

using namespace ffq;

extern "C" int ffq_qlinear_w4a16(const void* x, int x_dtype, const int8_t* qw, void* y, int64_t M, int64_t N, int64_t K,
                                 const float* sw, const float* ow, int64_t group, const void* bias, int bias_dtype,
                                 void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (M <= 0 || N <= 0) return FFQ_OK;
  if (!(x_dtype == FFQ_BF16 || x_dtype == FFQ_F16)) {
    set_error("qlinear_w4a16: activations must be bfloat16 or float16"); return FFQ_ERR_UNSUPPORTED;
  }
  if (K <= 0 || K % w4::BK != 0) { set_error("qlinear_w4a16: K must be a positive multiple of %d (got %lld)", w4::BK, (long long)K); return FFQ_ERR_UNSUPPORTED; }
  if (group <= 0 || K % group != 0 || group % w4::BK != 0) {
    set_error("qlinear_w4a16: the group size must divide K and be a multiple of %d (got %lld)", w4::BK, (long long)group); return FFQ_ERR_UNSUPPORTED;
  }
  if ((reinterpret_cast<uintptr_t>(x) & 15u) || (reinterpret_cast<uintptr_t>(qw) & 15u) || (reinterpret_cast<uintptr_t>(y) & 15u)) {
    set_error("qlinear_w4a16: pointers must be 16-byte aligned"); return FFQ_ERR_UNSUPPORTED;
  }
  if (M > 0x7fffffffll || N > 0x7fffffffll || K > 0x7fffffffll) { set_error("qlinear_w4a16: dimension too large"); return FFQ_ERR_UNSUPPORTED; }
  CUtensorMap map_a, map_raw;
  int rc;
  const CUtensorMapDataType adt = x_dtype == FFQ_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  if ((rc = w4::make_map2(&map_a, adt, 2, x, M, K, w4::BK, w4::BM, CU_TENSOR_MAP_SWIZZLE_128B)) != FFQ_OK) return rc;
  if ((rc = w4::make_map2(&map_raw, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, qw, N, K, w4::BK, w4::BN / 2, CU_TENSOR_MAP_SWIZZLE_NONE)) != FFQ_OK) return rc;
  w4::Args g{};
  g.M = (int)M; g.N = (int)N; g.K = (int)K; g.y = y; g.sw = sw; g.ow = ow; g.group = (int)group; g.groups = (int)(K / group); g.kb_per_group = (int)(group / w4::BK);
  g.bias = bias; g.bias_dt = bias_dtype;
  static std::atomic<uint64_t> attr_done{0};
  const cudaError_t attr_err = once_per_device(attr_done, []() -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(w4::w4a16_gemm2_kernel<__nv_bfloat16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, w4::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(w4::w4a16_gemm2_kernel<__half, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, w4::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(w4::w4a16_gemm2_kernel<__nv_bfloat16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, w4::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(w4::w4a16_gemm2_kernel<__half, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, w4::SMEM_BYTES);
    return e;
  });
  if (attr_err != cudaSuccess) { set_error("qlinear_w4a16: cannot reserve %d bytes of shared memory: %s", w4::SMEM_BYTES, cudaGetErrorString(attr_err)); return FFQ_ERR_CUDA; }
  // two 256-row tiles per dequantized weight stage from 257 activation rows on (FFQ_W4A16_MT=1|2 overrides: A/B switch)
  int mt = M > w4::TM ? 2 : 1;
  { const char* e = getenv("FFQ_W4A16_MT"); if (e && (e[0] == '1' || e[0] == '2')) mt = e[0] - '0'; }
  const long long tmq = (long long)mt * w4::TM;
  const long long pair_tiles = ((M + tmq - 1) / tmq) * ((N + w4::BN - 1) / w4::BN);
  const int max_pairs = sm_count() / 2;
  const int grid = 2 * (int)(pair_tiles < max_pairs ? pair_tiles : max_pairs);
  if (x_dtype == FFQ_BF16) {
    if (mt == 2) w4::w4a16_gemm2_kernel<__nv_bfloat16, 2><<<grid, w4::THREADS, w4::SMEM_BYTES, st>>>(map_a, map_raw, g);
    else w4::w4a16_gemm2_kernel<__nv_bfloat16, 1><<<grid, w4::THREADS, w4::SMEM_BYTES, st>>>(map_a, map_raw, g);
  } else {
    if (mt == 2) w4::w4a16_gemm2_kernel<__half, 2><<<grid, w4::THREADS, w4::SMEM_BYTES, st>>>(map_a, map_raw, g);
    else w4::w4a16_gemm2_kernel<__half, 1><<<grid, w4::THREADS, w4::SMEM_BYTES, st>>>(map_a, map_raw, g);
  }
  FFQ_LAUNCH_CHECK();
  return FFQ_OK;
}

// ubench_store.cu -- how fast can all SMs push a GEMM epilogue's output tile to global memory?
// One CTA per SM, 4 storing warps (like the epilogue warps of the W8A8 kernels).  Every CTA writes `tiles` output tiles of
// [rows x 256 B] (128 bf16 columns) into a row-major [M, N] bf16 matrix.  Patterns:
//   0  transposed 2-byte stores: a warp instruction writes 32 consecutive bf16 (64 B) of ONE row, next instruction next row
//   1  transposed 4-byte stores: 128 B of one row per warp instruction
//   2  row-per-lane 16-byte stores: a warp instruction writes 16 B of 32 different rows
//   3  TMA bulk stores from shared memory, boxes of 32 rows x 128 B
//   4  like 0 with st.global.cs (evict-first)
// Output: bytes/clk/SM (SM clocks from clock64) and GB/s (events), for a small (L2-resident) and a large output.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench_store tools/ubench_store.cu -lcuda && ./ubench_store
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int PATTERN>
__global__ void __launch_bounds__(128, 1) store_kernel(uint16_t* y, int M, int N, int tiles_per_cta, int rows_per_tile,
                                                       const __grid_constant__ CUtensorMap map, unsigned long long* clocks) {
  __shared__ __align__(1024) uint8_t stage[4][4096];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = N / 128, tiles_m = M / rows_per_tile;
  const long long t0 = clock64();
  for (int t = 0; t < tiles_per_cta; ++t) {
    const int tile = (blockIdx.x + t * gridDim.x) % (tiles_n * tiles_m);
    const int tn = tile % tiles_n, tm = tile / tiles_n;
    const int n0 = tn * 128, m0 = tm * rows_per_tile;
    if (PATTERN == 0 || PATTERN == 4) {
      // warp w covers columns n0 + 32w .. +31; loops over rows
      uint16_t* p = y + (size_t)m0 * N + n0 + warp * 32 + lane;
      for (int r = 0; r < rows_per_tile; ++r) {
        if (PATTERN == 0) p[(size_t)r * N] = (uint16_t)(r + lane);
        else asm volatile("st.global.cs.u16 [%0], %1;" ::"l"(p + (size_t)r * N), "h"((uint16_t)(r + lane)) : "memory");
      }
    } else if (PATTERN == 1) {
      // warp w covers rows r = w, w+4, ...; 32 lanes x 4 B = 128 B... only half the tile width per instruction: two halves
      for (int r = warp; r < rows_per_tile; r += 4) {
        uint32_t* p = reinterpret_cast<uint32_t*>(y + (size_t)(m0 + r) * N + n0);
        p[lane] = (uint32_t)(r + lane);
        p[32 + lane] = (uint32_t)(r - lane);
      }
    } else if (PATTERN == 2) {
      // lane = row (32 rows per warp pass), 16-byte stores walking along the row: 64 columns = 128 B ... covers 256 B per row in 16 stores
      for (int rb = warp * 32; rb < rows_per_tile; rb += 128) {
        uint4* p = reinterpret_cast<uint4*>(y + (size_t)(m0 + rb + lane) * N + n0);
#pragma unroll
        for (int c = 0; c < 16; ++c) p[c] = make_uint4(c, lane, rb, 7);
      }
    } else if (PATTERN == 3) {
      // each warp: boxes of 32 rows x 128 B (64 bf16 columns); tile = rows_per_tile/32 x 2 boxes, split over 4 warps
      const int boxes = (rows_per_tile / 32) * 2;
      for (int b = warp; b < boxes; b += 4) {
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        reinterpret_cast<uint4*>(stage[warp])[lane] = make_uint4(b, lane, t, 1);      // token fill (not the point here)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          const int c0 = n0 + (b & 1) * 64, c1 = m0 + (b >> 1) * 32;
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&map),
                       "r"(smem_u32(stage[warp])), "r"(c0), "r"(c1) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
  }
  if (PATTERN == 3 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) clocks[blockIdx.x] = (unsigned long long)(clock64() - t0);
}

// Patterns 5/6: the transposed epilogue staged through shared memory.  Every lane owns one output column and has 32
// consecutive rows of it in registers (what tcgen05.ld.32x32b gives the swapped GEMM); it writes them with 32 2-byte
// shared stores (a warp instruction = 64 consecutive bytes of one staged row), then TMA stores drain the boxes.
//   5: one box [32 rows][64 B] per warp, NBUF-deep ring per warp            6: one box [32 rows][256 B] per CTA, NBUF-deep
template <int PATTERN, int NBUF>
__global__ void __launch_bounds__(128, 1) staged_kernel(int M, int N, int tiles_per_cta, int rows_per_tile,
                                                        const __grid_constant__ CUtensorMap map64, const __grid_constant__ CUtensorMap map256,
                                                        unsigned long long* clocks) {
  extern __shared__ __align__(1024) uint8_t stage[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = N / 128, tiles_m = M / rows_per_tile;
  const long long t0 = clock64();
  int buf = 0;
  for (int t = 0; t < tiles_per_cta; ++t) {
    const int tile = (blockIdx.x + t * gridDim.x) % (tiles_n * tiles_m);
    const int tn = tile % tiles_n, tm = tile / tiles_n;
    const int n0 = tn * 128, m0 = tm * rows_per_tile;
    for (int c0 = 0; c0 < rows_per_tile; c0 += 32) {
      if (PATTERN == 5) {
        uint8_t* box = stage + (warp * NBUF + buf) * 2048;
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NBUF - 1) : "memory");
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; ++j) *reinterpret_cast<uint16_t*>(box + j * 64 + lane * 2) = (uint16_t)(j + lane + c0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&map64),
                       "r"(smem_u32(box)), "r"(n0 + warp * 32), "r"(m0 + c0) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      } else {
        uint8_t* box = stage + buf * 8192;
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NBUF - 1) : "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) *reinterpret_cast<uint16_t*>(box + j * 256 + warp * 64 + lane * 2) = (uint16_t)(j + lane + c0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 0) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&map256),
                       "r"(smem_u32(box)), "r"(n0), "r"(m0 + c0) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      if (++buf == NBUF) buf = 0;
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) clocks[blockIdx.x] = (unsigned long long)(clock64() - t0);
}

// Pattern 7/8: direct transposed stores from 8 warps (2 per 32-column stripe, alternating 32-row chunks): is the request
// rate a per-warp or a per-SM limit?   7: 2-byte stores, 8: 2-byte st.global.cs
template <int PATTERN>
__global__ void __launch_bounds__(256, 1) store8_kernel(uint16_t* y, int M, int N, int tiles_per_cta, int rows_per_tile,
                                                        unsigned long long* clocks) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = N / 128, tiles_m = M / rows_per_tile;
  const long long t0 = clock64();
  for (int t = 0; t < tiles_per_cta; ++t) {
    const int tile = (blockIdx.x + t * gridDim.x) % (tiles_n * tiles_m);
    const int tn = tile % tiles_n, tm = tile / tiles_n;
    const int n0 = tn * 128, m0 = tm * rows_per_tile;
    uint16_t* p = y + (size_t)m0 * N + n0 + (warp & 3) * 32 + lane;
    for (int c0 = (warp >> 2) * 32; c0 < rows_per_tile; c0 += 64) {
#pragma unroll 8
      for (int j = 0; j < 32; ++j) {
        if (PATTERN == 7) p[(size_t)(c0 + j) * N] = (uint16_t)(j + lane);
        else asm volatile("st.global.cs.u16 [%0], %1;" ::"l"(p + (size_t)(c0 + j) * N), "h"((uint16_t)(j + lane)) : "memory");
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) clocks[blockIdx.x] = (unsigned long long)(clock64() - t0);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int P>
static void run(const char* name, uint16_t* y, int M, int N, int rows_per_tile, const CUtensorMap& map, unsigned long long* clk_d, int sms) {
  const int tiles = (M / rows_per_tile) * (N / 128);
  const int per_cta = (tiles + sms - 1) / sms;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int w = 0; w < 2; ++w) store_kernel<P><<<sms, 128>>>(y, M, N, per_cta, rows_per_tile, map, clk_d);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  store_kernel<P><<<sms, 128>>>(y, M, N, per_cta, rows_per_tile, map, clk_d);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  unsigned long long* clk = (unsigned long long*)malloc(sizeof(unsigned long long) * sms);
  CK(cudaMemcpy(clk, clk_d, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost));
  double avg = 0;
  for (int i = 0; i < sms; ++i) avg += (double)clk[i];
  avg /= sms;
  const double bytes = (double)per_cta * sms * rows_per_tile * 256.0;
  printf("  %-44s %8.1f us  %7.1f GB/s  %6.2f B/clk/SM  (%.0f clks per [%d x 256 B] tile)\n", name, ms * 1e3, bytes / ms / 1e6,
         bytes / sms / avg, avg / per_cta, rows_per_tile);
  free(clk);
}

static void report(const char* name, float ms, unsigned long long* clk_d, int sms, int per_cta, int rows_per_tile) {
  unsigned long long* clk = (unsigned long long*)malloc(sizeof(unsigned long long) * sms);
  CK(cudaMemcpy(clk, clk_d, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost));
  double avg = 0;
  for (int i = 0; i < sms; ++i) avg += (double)clk[i];
  avg /= sms;
  const double bytes = (double)per_cta * sms * rows_per_tile * 256.0;
  printf("  %-44s %8.1f us  %7.1f GB/s  %6.2f B/clk/SM  (%.0f clks per [%d x 256 B] tile)\n", name, ms * 1e3, bytes / ms / 1e6,
         bytes / sms / avg, avg / per_cta, rows_per_tile);
  free(clk);
}

template <typename F>
static void timed(const char* name, F launch, unsigned long long* clk_d, int sms, int per_cta, int rows_per_tile) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int w = 0; w < 2; ++w) launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  launch();
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  report(name, ms, clk_d, sms, per_cta, rows_per_tile);
}

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  EncodeTiledFn enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
  unsigned long long* clk_d;
  CK(cudaMalloc(&clk_d, sizeof(unsigned long long) * sms));
  const int shapes[3][2] = {{2048, 4096}, {2048, 14336}, {8192, 14336}};
  for (int s = 0; s < 3; ++s) {
    const int M = shapes[s][0], N = shapes[s][1];
    uint16_t* y;
    CK(cudaMalloc(&y, (size_t)M * N * 2));
    CUtensorMap map;
    const cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
    const cuuint64_t strides[1] = {(cuuint64_t)N * 2};
    const cuuint32_t box[2] = {64, 32};
    const cuuint32_t estr[2] = {1, 1};
    if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, y, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      printf("tensor map failed\n"); return 1;
    }
    printf("output [%d, %d] bf16 = %.1f MB, %d SMs x 4 warps\n", M, N, (double)M * N * 2 / 1e6, sms);
    run<0>("transposed 2-byte stores (64 B of a row/instr)", y, M, N, 256, map, clk_d, sms);
    run<4>("same, st.global.cs", y, M, N, 256, map, clk_d, sms);
    run<1>("4-byte stores (128 B of a row/instr)", y, M, N, 256, map, clk_d, sms);
    run<2>("row-per-lane 16-byte stores", y, M, N, 256, map, clk_d, sms);
    run<3>("TMA stores, boxes of 32 rows x 128 B", y, M, N, 256, map, clk_d, sms);
    {
      CUtensorMap map64, map256;
      const cuuint32_t box64[2] = {32, 32}, box256[2] = {128, 32};
      if (enc(&map64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, y, dims, strides, box64, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
          enc(&map256, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, y, dims, strides, box256, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
        printf("tensor map (staged) failed\n"); return 1;
      }
      const int tiles = (M / 256) * (N / 128), per_cta = (tiles + sms - 1) / sms;
      CK(cudaFuncSetAttribute(staged_kernel<5, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
      CK(cudaFuncSetAttribute(staged_kernel<5, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
      CK(cudaFuncSetAttribute(staged_kernel<6, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
      CK(cudaFuncSetAttribute(staged_kernel<6, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
      timed("staged, per-warp boxes 32 x 64 B, 2 deep", [&] { staged_kernel<5, 2><<<sms, 128, 4 * 2 * 2048 + 1024>>>(M, N, per_cta, 256, map64, map256, clk_d); }, clk_d, sms, per_cta, 256);
      timed("staged, per-warp boxes 32 x 64 B, 4 deep", [&] { staged_kernel<5, 4><<<sms, 128, 4 * 4 * 2048 + 1024>>>(M, N, per_cta, 256, map64, map256, clk_d); }, clk_d, sms, per_cta, 256);
      timed("staged, CTA boxes 32 x 256 B, 2 deep", [&] { staged_kernel<6, 2><<<sms, 128, 2 * 8192 + 1024>>>(M, N, per_cta, 256, map64, map256, clk_d); }, clk_d, sms, per_cta, 256);
      timed("staged, CTA boxes 32 x 256 B, 4 deep", [&] { staged_kernel<6, 4><<<sms, 128, 4 * 8192 + 1024>>>(M, N, per_cta, 256, map64, map256, clk_d); }, clk_d, sms, per_cta, 256);
      timed("8 warps, transposed 2-byte stores", [&] { store8_kernel<7><<<sms, 256>>>(y, M, N, per_cta, 256, clk_d); }, clk_d, sms, per_cta, 256);
      timed("8 warps, transposed 2-byte st.global.cs", [&] { store8_kernel<8><<<sms, 256>>>(y, M, N, per_cta, 256, clk_d); }, clk_d, sms, per_cta, 256);
    }
    CK(cudaFree(y));
  }
  return 0;
}

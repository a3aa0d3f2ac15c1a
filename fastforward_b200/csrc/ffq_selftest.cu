// ffq_selftest.cu -- device self-test hook: sweeps the shared-reciprocal division of
// ffq_common.cuh against __fdiv_rn.  Test infrastructure, exported so that tests/ can call it
// through the C ABI.
#include "ffq_common.cuh"

namespace ffq {

__device__ __forceinline__ unsigned int mix(unsigned int h) {
  h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
  return h;
}

// counts[0]: quotients accepted by the guard that differ from __fdiv_rn (must be 0)
// counts[1]: accepted (non-strict guard)   counts[2]: accepted (strict guard)
// counts[3]: strict-accepted quotients that differ (must be 0)
__global__ void selftest_div_kernel(unsigned long long n, unsigned int seed, unsigned long long* counts) {
  unsigned long long bad = 0, acc = 0, acc_strict = 0, bad_strict = 0;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned int a = mix((unsigned int)i * 2u + seed), b = mix((unsigned int)(i >> 7) * 2u + 1u + seed * 3u);
    // scale: sign, exponent in a band around the guard's edges, adversarial mantissas now and then
    unsigned int se = 127u - 44u + (b % 89u);                 // 2^-44 .. 2^44
    unsigned int sm = (b >> 9) & 0x7fffffu;
    const unsigned int pick = (b >> 5) & 7u;
    if (pick == 0) sm = 0x7fffffu; else if (pick == 1) sm = 0u; else if (pick == 2) sm = 0x7ffffeu; else if (pick == 3) sm = 1u;
    const float s = __uint_as_float(((b & 16u) << 27) | (se << 23) | sm);
    // dividend: any bit pattern half of the time, otherwise a "data-like" magnitude
    float x;
    if (a & 1u) x = __uint_as_float(mix(a + 77u));
    else x = __uint_as_float((a & 0x80000000u) | ((100u + ((a >> 1) % 56u)) << 23) | ((a >> 8) & 0x7fffffu));
    const float want = __fdiv_rn(x, s);
    const SharedRcp k = make_shared_rcp(s);
    bool ok = k.ok;
    const float got = shared_div<false>(x, k, ok);
    if (ok) { ++acc; if (__float_as_uint(got) != __float_as_uint(want) && !(fabsf(want) < 0x1p-50f)) ++bad; }
    bool ok2 = k.ok;
    const float got2 = shared_div<true>(x, k, ok2);
    if (ok2) { ++acc_strict; if (__float_as_uint(got2) != __float_as_uint(want)) ++bad_strict; }
  }
  atomicAdd(&counts[0], bad);
  atomicAdd(&counts[1], acc);
  atomicAdd(&counts[2], acc_strict);
  atomicAdd(&counts[3], bad_strict);
}

}  // namespace ffq

extern "C" int ffq_selftest_shared_div(unsigned long long n, unsigned int seed, unsigned long long* counts_dev,
                                       void* stream) {
  ffq::selftest_div_kernel<<<148 * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(n, seed, counts_dev);
  FFQ_LAUNCH_CHECK();
  return FFQ_OK;
}

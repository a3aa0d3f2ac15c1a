/* ffq_oracle.c -- plain-C scalar restatement of the fp32 hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * An independent cross-check of the IEEE claims the CUDA kernels rely on: with round-to-nearest
 * binary32 arithmetic, `x / s`, `- rint(o)`, rintf (half-to-even), NaN-propagating clamp and
 * `(q + rint(o)) * s` evaluated one operation at a time give exactly the bits aten produces.
 * Compiled with -O2 -ffp-contract=off -fno-fast-math (oracle/Makefile); tests/test_oracle_golden.py
 * runs it over every fp32 golden vector recorded from the reference.
 *
 * Layout: dense row-major tensor `dims[rank]`, tile `tile[rank]`, parameters indexed by tile,
 * tile index row-major over the block grid (quantization/tiled_tensor.py:71-98).
 * Citations are into /root/reference/src/fastforward/.                                          */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

#define MAXR 8

static int64_t tile_of(int64_t e, int rank, const int64_t* dims, const int64_t* tile) {
  int64_t idx[MAXR];
  for (int d = rank - 1; d >= 0; --d) { idx[d] = e % dims[d]; e /= dims[d]; }
  int64_t p = 0;
  for (int d = 0; d < rank; ++d) p = p * (dims[d] / tile[d]) + idx[d] / tile[d];
  return p;
}

static float clamp_nan(float v, float lo, float hi) {      /* torch.clamp propagates NaN */
  if (v != v) return v;
  return fminf(fmaxf(v, lo), hi);
}

/* quantization/_quantizer_impl.py:144-169 (fp32 data, fp32 params, fp32 codes) */
void ffqo_quantize(const float* x, float* q, int64_t n, int rank, const int64_t* dims, const int64_t* tile,
                   const float* scale, const float* offset, double num_bits) {
  const float lo = (float)(-pow(2.0, num_bits - 1.0)), hi = (float)(pow(2.0, num_bits - 1.0) - 1.0);
  for (int64_t e = 0; e < n; ++e) {
    const int64_t p = tile_of(e, rank, dims, tile);
    const float o = offset ? rintf(offset[p]) : 0.0f;
    volatile float t = x[e] / scale[p];        /* volatile: one rounding per operation, no contraction */
    t = t - o;
    q[e] = clamp_nan(rintf(t), lo, hi);
  }
}

/* quantization/_quantizer_impl.py:172-190 */
void ffqo_dequantize(const float* q, float* y, int64_t n, int rank, const int64_t* dims, const int64_t* tile,
                     const float* scale, const float* offset) {
  for (int64_t e = 0; e < n; ++e) {
    const int64_t p = tile_of(e, rank, dims, tile);
    const float o = offset ? rintf(offset[p]) : 0.0f;
    volatile float u = q[e] + o;
    y[e] = u * scale[p];
  }
}

/* quantization/_quantizer_impl.py:193-237: dx and the per-tile sums (accumulated in double: the
 * order-independent target) */
void ffqo_backward(const float* x, const float* g, float* dx, double* dscale, double* doffset, int64_t n,
                   int64_t ntiles, int rank, const int64_t* dims, const int64_t* tile, const float* scale,
                   const float* offset, double num_bits) {
  const float lo = (float)(-pow(2.0, num_bits - 1.0)), hi = (float)(pow(2.0, num_bits - 1.0) - 1.0);
  for (int64_t p = 0; p < ntiles; ++p) { dscale[p] = 0.0; if (doffset) doffset[p] = 0.0; }
  for (int64_t e = 0; e < n; ++e) {
    const int64_t p = tile_of(e, rank, dims, tile);
    const float s = scale[p], o = offset ? rintf(offset[p]) : 0.0f;
    volatile float pre = x[e] / s;
    pre = pre - o;
    const float q = rintf(pre);
    const int below = q < lo, above = q > hi, clip = below || above;
    dx[e] = clip ? 0.0f : g[e];
    volatile float bound = (below ? lo : hi) + o;
    volatile float resid = q - pre;
    volatile float term = (clip ? bound : resid) * g[e];
    dscale[p] += (double)term;
    if (doffset) { volatile float so = s * g[e]; doffset[p] += clip ? (double)so : 0.0; }
  }
}

/* range_setting/minmax.py:229-230 */
void ffqo_minmax(const float* x, float* mn, float* mx, int64_t n, int64_t ntiles, int rank, const int64_t* dims,
                 const int64_t* tile) {
  for (int64_t p = 0; p < ntiles; ++p) { mn[p] = INFINITY; mx[p] = -INFINITY; }
  for (int64_t e = 0; e < n; ++e) {
    const int64_t p = tile_of(e, rank, dims, tile);
    const float v = x[e];
    if (v != v) { mn[p] = v; mx[p] = v; continue; }
    if (!(mn[p] != mn[p]) && v < mn[p]) mn[p] = v;
    if (!(mx[p] != mx[p]) && v > mx[p]) mx[p] = v;
  }
}

/* quantization/affine/range.py:54-122; returns 1 when an offset was produced */
int ffqo_params_for_range(const float* mn, const float* mx, int64_t n, double num_bits, int symmetric,
                          int allow_one_sided, float* scale, float* offset) {
  float gmin = INFINITY;
  int has_nan = 0;
  for (int64_t i = 0; i < n; ++i) { if (mn[i] != mn[i]) has_nan = 1; else if (mn[i] < gmin) gmin = mn[i]; }
  const int one_sided = (!has_nan && gmin >= 0.0f) && allow_one_sided;
  const float int_min = (float)(-pow(2.0, num_bits - 1.0));
  if (symmetric && !one_sided) {
    const float a = fabsf(int_min), b = fabsf(-int_min - 1.0f);
    for (int64_t i = 0; i < n; ++i) {
      const float neg = fabsf(mn[i]) / a, pos = fabsf(mx[i]) / b;
      scale[i] = (neg != neg) ? neg : ((pos != pos) ? pos : fmaxf(neg, pos));
    }
    return 0;
  }
  const float steps = (float)(pow(2.0, num_bits) - 1.0);
  for (int64_t i = 0; i < n; ++i) {
    const float lo = (symmetric && one_sided) ? 0.0f : mn[i];
    volatile float len = mx[i] - lo;
    float s = len / steps;
    if (!(s != s) && s < 1.1920928955078125e-07f) s = 1.1920928955078125e-07f;
    volatile float r = lo / s;
    scale[i] = s;
    offset[i] = r - int_min;
  }
  return 1;
}

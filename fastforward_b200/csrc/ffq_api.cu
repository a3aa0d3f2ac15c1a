// ffq_api.cu -- library-level pieces of the C ABI: errors, launch accounting, dtype algebra,
// tile-layout planning (the host-side "layout classification" of SURVEY.md section 7 step 2).
#include <cstdarg>
#include <cstring>
#include <mutex>

#include "ffq_common.cuh"

namespace ffq {

static thread_local char t_error[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_error, sizeof(t_error), fmt, ap);
  va_end(ap);
}

const char* dt_name(int dt) {
  switch (dt) {
    case FFQ_F32: return "float32";
    case FFQ_F16: return "float16";
    case FFQ_BF16: return "bfloat16";
    case FFQ_F64: return "float64";
    case FFQ_I8: return "int8";
    case FFQ_I16: return "int16";
    case FFQ_I32: return "int32";
    case FFQ_U8: return "uint8";
    case FFQ_I64: return "int64";
    case FFQ_NONE: return "none";
  }
  return "?";
}

int promote(int a, int b) {
  if (a == FFQ_NONE) return b;
  if (b == FFQ_NONE) return a;
  if (a == b) return a;
  const bool fa = is_float_dt(a), fb = is_float_dt(b);
  if (fa && !fb) return a;
  if (fb && !fa) return b;
  if (fa && fb) {
    if (a == FFQ_F64 || b == FFQ_F64) return FFQ_F64;
    if (a == FFQ_F32 || b == FFQ_F32) return FFQ_F32;
    return FFQ_F32;  // float16 x bfloat16
  }
  // integers
  auto rank = [](int d) { return d == FFQ_I64 ? 4 : d == FFQ_I32 ? 3 : d == FFQ_I16 ? 2 : 1; };
  if (a == FFQ_U8 || b == FFQ_U8) {
    const int other = (a == FFQ_U8) ? b : a;
    return other == FFQ_I8 ? FFQ_I16 : other;  // uint8 x int8 -> int16, else the wider signed type
  }
  return rank(a) >= rank(b) ? a : b;
}

int make_plan(const ffq_layout_t* L, Plan* p) {
  if (L == nullptr || L->rank < 0 || L->rank > FFQ_MAX_RANK) {
    set_error("layout rank must be in [0, %d]", FFQ_MAX_RANK);
    return FFQ_ERR_INVALID;
  }
  int64_t d[FFQ_MAX_RANK], t[FFQ_MAX_RANK];
  int r = 0;
  int64_t numel = 1, tnumel = 1;
  for (int i = 0; i < L->rank; ++i) {
    const int64_t di = L->dims[i], ti = L->tile[i];
    if (di < 0 || ti <= 0 || (di % ti) != 0) {
      if (di == 0) { numel = 0; continue; }
      set_error("Each dimension of tile_size must divide the corresponding input dimension. Got %lld and %lld for dimension %d.",
                (long long)di, (long long)ti, i);
      return FFQ_ERR_INVALID;
    }
    numel *= di;
    tnumel *= ti;
    if (di == 1) continue;  // contributes nothing to indexing
    d[r] = di; t[r] = ti; ++r;
  }
  // merge (i, i+1) when tile[i]==1 or tile[i+1]==dims[i+1]
  bool merged = true;
  while (merged && r > 1) {
    merged = false;
    for (int i = 0; i + 1 < r; ++i) {
      if (t[i] == 1 || t[i + 1] == d[i + 1]) {
        d[i] = d[i] * d[i + 1];
        t[i] = t[i] * t[i + 1];
        for (int j = i + 1; j + 1 < r; ++j) { d[j] = d[j + 1]; t[j] = t[j + 1]; }
        --r;
        merged = true;
        break;
      }
    }
  }
  p->rank = r;
  for (int i = 0; i < r; ++i) { p->dims[i] = d[i]; p->tile[i] = t[i]; }
  p->numel = numel;
  p->tile_numel = numel == 0 ? 0 : tnumel;
  p->num_tiles = numel == 0 ? 0 : numel / tnumel;
  p->row = r <= 1;
  return FFQ_OK;
}

GenericLayout make_generic_layout(const Plan& p) {
  GenericLayout g;
  memset(&g, 0, sizeof(g));
  g.rank = p.rank;
  unsigned long long stride = 1;
  for (int i = p.rank - 1; i >= 0; --i) {
    g.dims[i] = (unsigned long long)p.dims[i];
    g.tile[i] = (unsigned long long)p.tile[i];
    g.grid[i] = (unsigned long long)(p.dims[i] / p.tile[i]);
    g.stride[i] = stride;
    stride *= (unsigned long long)p.dims[i];
  }
  return g;
}

FastDiv make_fast_div(unsigned int d) {
  FastDiv f{};
  f.d = d;
  if ((d & (d - 1)) == 0) {
    f.pow2 = 1;
    unsigned int s = 0;
    while ((1u << s) < d) ++s;
    f.shift = s;
    f.mul = 0;
    return f;
  }
  unsigned int L = 0;
  while ((1ull << L) < d) ++L;  // ceil(log2 d)
  const unsigned long long m = ((1ull << 32) * ((1ull << L) - d)) / d + 1;
  f.mul = (unsigned int)m;
  f.shift = L - 1;
  f.pow2 = 0;
  return f;
}

int sm_count() {
  static int cached[64];
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  std::lock_guard<std::mutex> lock(mu);
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace ffq

extern "C" {

int ffq_abi_version(void) { return FFQ_ABI_VERSION; }
const char* ffq_last_error(void) { return ffq::t_error; }
uint64_t ffq_launch_count(void) { return ffq::g_launches.load(std::memory_order_relaxed); }

int64_t ffq_num_tiles(const ffq_layout_t* layout) {
  ffq::Plan p;
  if (ffq::make_plan(layout, &p) != FFQ_OK) return -1;
  return p.num_tiles;
}

}  // extern "C"

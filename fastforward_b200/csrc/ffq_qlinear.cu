// placeholder until the tcgen05 kernel lands
#include "ffq_common.cuh"
extern "C" int ffq_qlinear_w8a8(const int8_t*, const int8_t*, void*, int, int64_t, int64_t, int64_t, const float*,
                                const float*, const float*, const float*, const int32_t*, const int32_t*,
                                const void*, int, void*) {
  ffq::set_error("qlinear_w8a8: not built yet");
  return FFQ_ERR_UNSUPPORTED;
}
extern "C" int ffq_rowsum_i8(const int8_t*, int32_t*, int64_t, int64_t, void*) {
  ffq::set_error("rowsum_i8: not built yet");
  return FFQ_ERR_UNSUPPORTED;
}

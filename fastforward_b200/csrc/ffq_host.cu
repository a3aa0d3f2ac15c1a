// ffq_host.cu -- host-buffer entry point: the end-to-end path a caller with HOST tensors takes
// (H2D copies, the two fused kernels, D2H copies).  bench.py's "e2e" number goes through here.
#include <cstdlib>
#include <mutex>

#include "ffq_common.cuh"

namespace {

// Grow-only per-device staging buffers (device memory + pinned host mirror are owned here so the
// caller can hand in ordinary pageable memory).
constexpr int CHUNKS = 8;      // pipeline depth of the host entry point
struct Staging {
  void* dev = nullptr;
  size_t dev_bytes = 0;
  cudaStream_t stream = nullptr;            // compute
  cudaStream_t in = nullptr, out = nullptr; // H2D / D2H copy streams (the two PCIe directions run concurrently)
  cudaEvent_t ev_in[CHUNKS] = {}, ev_out[CHUNKS] = {}, ev_par = nullptr, ev_start = nullptr;
};
Staging g_staging[64];
std::mutex g_mu;

int ensure(Staging& s, size_t bytes) {
  if (s.stream == nullptr) {
    FFQ_CUDA_CHECK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    FFQ_CUDA_CHECK(cudaStreamCreateWithFlags(&s.in, cudaStreamNonBlocking));
    FFQ_CUDA_CHECK(cudaStreamCreateWithFlags(&s.out, cudaStreamNonBlocking));
    for (int i = 0; i < CHUNKS; ++i) {
      FFQ_CUDA_CHECK(cudaEventCreateWithFlags(&s.ev_in[i], cudaEventDisableTiming));
      FFQ_CUDA_CHECK(cudaEventCreateWithFlags(&s.ev_out[i], cudaEventDisableTiming));
    }
    FFQ_CUDA_CHECK(cudaEventCreateWithFlags(&s.ev_par, cudaEventDisableTiming));
    FFQ_CUDA_CHECK(cudaEventCreateWithFlags(&s.ev_start, cudaEventDisableTiming));
  }
  if (s.dev_bytes < bytes) {
    if (s.dev) FFQ_CUDA_CHECK(cudaFree(s.dev));
    s.dev = nullptr; s.dev_bytes = 0;
    FFQ_CUDA_CHECK(cudaMalloc(&s.dev, bytes));
    s.dev_bytes = bytes;
  }
  return FFQ_OK;
}

size_t align_up(size_t v) { return (v + 255) / 256 * 256; }

}  // namespace

extern "C" int ffq_fakequant_fwd_bwd_host(const void* x_host, const void* g_host, int dtype, void* y_host,
                                          void* dx_host, float* dscale_host, float* doffset_host,
                                          const float* scale_host, const float* offset_host,
                                          const ffq_layout_t* layout, double num_bits, int device) {
  using namespace ffq;
  if (!(dtype == FFQ_F32 || dtype == FFQ_F16 || dtype == FFQ_BF16)) {
    set_error("fakequant_fwd_bwd_host: dtype must be float32/float16/bfloat16");
    return FFQ_ERR_UNSUPPORTED;
  }
  if (device < 0 || device >= 64) { set_error("bad device index %d", device); return FFQ_ERR_INVALID; }
  Plan plan;
  int rc = make_plan(layout, &plan);
  if (rc != FFQ_OK) return rc;
  if (plan.numel == 0) return FFQ_OK;
  FFQ_CUDA_CHECK(cudaSetDevice(device));
  std::lock_guard<std::mutex> lock(g_mu);
  Staging& s = g_staging[device];
  const size_t nbytes = align_up((size_t)plan.numel * dt_size(dtype));
  const size_t pbytes = align_up((size_t)plan.num_tiles * sizeof(float));
  size_t wneed = ffq_workspace_bytes(FFQ_WS_QUANTIZE_BWD, layout, dtype);
  if (plan.row && plan.num_tiles >= 2 * CHUNKS) {      // the pipelined path runs the kernels on groups of whole tiles
    const long long per = (plan.num_tiles + CHUNKS - 1) / CHUNKS;
    for (long long nt : {per, plan.num_tiles - (CHUNKS - 1) * per}) {
      if (nt <= 0) continue;
      ffq_layout_t sub{};
      sub.rank = 1; sub.dims[0] = nt * plan.tile_numel; sub.tile[0] = plan.tile_numel;
      const size_t w = ffq_workspace_bytes(FFQ_WS_QUANTIZE_BWD, &sub, dtype);
      if (w > wneed) wneed = w;
    }
  }
  const size_t wbytes = align_up(wneed);
  // [x | g | y | dx | scale | offset | dscale | doffset | workspace]
  rc = ensure(s, 4 * nbytes + 4 * pbytes + wbytes);
  if (rc != FFQ_OK) return rc;
  char* base = static_cast<char*>(s.dev);
  char *dx_x = base, *dx_g = base + nbytes, *dx_y = base + 2 * nbytes, *dx_dx = base + 3 * nbytes;
  char* p = base + 4 * nbytes;
  float *d_scale = (float*)p, *d_off = (float*)(p + pbytes), *d_dsc = (float*)(p + 2 * pbytes),
        *d_doff = (float*)(p + 3 * pbytes);
  void* d_ws = p + 4 * pbytes;
  const size_t raw = (size_t)plan.numel * dt_size(dtype), praw = (size_t)plan.num_tiles * sizeof(float);
  const size_t es = (size_t)dt_size(dtype);
  // ---- pipelined: tiles that are contiguous runs are cut into CHUNKS groups of whole tiles; chunk c's kernels run while
  // chunk c+1 arrives and chunk c-1 leaves (H2D and D2H use separate copy engines), so the call costs about one
  // direction's transfer time instead of H2D + kernels + D2H in sequence.  Same kernels, same per-tile results. ----
  if (plan.row && plan.num_tiles >= 2 * CHUNKS && raw >= (size_t)(8u << 20) && getenv("FFQ_HOST_SERIAL") == nullptr) {
    const long long per = (plan.num_tiles + CHUNKS - 1) / CHUNKS;
    FFQ_CUDA_CHECK(cudaMemcpyAsync(d_scale, scale_host, praw, cudaMemcpyHostToDevice, s.stream));
    if (offset_host) FFQ_CUDA_CHECK(cudaMemcpyAsync(d_off, offset_host, praw, cudaMemcpyHostToDevice, s.stream));
    FFQ_CUDA_CHECK(cudaEventRecord(s.ev_par, s.stream));
    int used = 0;
    for (int c = 0; c < CHUNKS; ++c) {
      const long long t0 = (long long)c * per, t1 = (t0 + per < plan.num_tiles) ? t0 + per : plan.num_tiles;
      if (t0 >= t1) break;
      used = c + 1;
      const size_t off = (size_t)t0 * plan.tile_numel * es, len = (size_t)(t1 - t0) * plan.tile_numel * es;
      FFQ_CUDA_CHECK(cudaMemcpyAsync(dx_x + off, static_cast<const char*>(x_host) + off, len, cudaMemcpyHostToDevice, s.in));
      FFQ_CUDA_CHECK(cudaMemcpyAsync(dx_g + off, static_cast<const char*>(g_host) + off, len, cudaMemcpyHostToDevice, s.in));
      FFQ_CUDA_CHECK(cudaEventRecord(s.ev_in[c], s.in));
      FFQ_CUDA_CHECK(cudaStreamWaitEvent(s.stream, s.ev_in[c], 0));
      ffq_layout_t sub{};
      sub.rank = 1; sub.dims[0] = (t1 - t0) * plan.tile_numel; sub.tile[0] = plan.tile_numel;
      rc = ffq_fakequant_fwd(dx_x + off, dtype, dx_y + off, dtype, nullptr, dtype, d_scale + t0, FFQ_F32,
                             offset_host ? d_off + t0 : nullptr, FFQ_F32, &sub, num_bits, s.stream);
      if (rc != FFQ_OK) return rc;
      rc = ffq_quantize_bwd(dx_x + off, dtype, dx_g + off, dtype, dx_dx + off, d_dsc + t0, offset_host ? d_doff + t0 : nullptr,
                            d_scale + t0, FFQ_F32, offset_host ? d_off + t0 : nullptr, FFQ_F32, &sub, num_bits, d_ws, wbytes, s.stream);
      if (rc != FFQ_OK) return rc;
      FFQ_CUDA_CHECK(cudaEventRecord(s.ev_out[c], s.stream));
      FFQ_CUDA_CHECK(cudaStreamWaitEvent(s.out, s.ev_out[c], 0));
      FFQ_CUDA_CHECK(cudaMemcpyAsync(static_cast<char*>(y_host) + off, dx_y + off, len, cudaMemcpyDeviceToHost, s.out));
      FFQ_CUDA_CHECK(cudaMemcpyAsync(static_cast<char*>(dx_host) + off, dx_dx + off, len, cudaMemcpyDeviceToHost, s.out));
    }
    (void)used;
    FFQ_CUDA_CHECK(cudaMemcpyAsync(dscale_host, d_dsc, praw, cudaMemcpyDeviceToHost, s.stream));
    if (offset_host && doffset_host)
      FFQ_CUDA_CHECK(cudaMemcpyAsync(doffset_host, d_doff, praw, cudaMemcpyDeviceToHost, s.stream));
    FFQ_CUDA_CHECK(cudaStreamSynchronize(s.stream));
    FFQ_CUDA_CHECK(cudaStreamSynchronize(s.out));
    return FFQ_OK;
  }
  FFQ_CUDA_CHECK(cudaMemcpyAsync(dx_x, x_host, raw, cudaMemcpyHostToDevice, s.stream));
  FFQ_CUDA_CHECK(cudaMemcpyAsync(dx_g, g_host, raw, cudaMemcpyHostToDevice, s.stream));
  FFQ_CUDA_CHECK(cudaMemcpyAsync(d_scale, scale_host, praw, cudaMemcpyHostToDevice, s.stream));
  if (offset_host) FFQ_CUDA_CHECK(cudaMemcpyAsync(d_off, offset_host, praw, cudaMemcpyHostToDevice, s.stream));
  rc = ffq_fakequant_fwd(dx_x, dtype, dx_y, dtype, nullptr, dtype, d_scale, FFQ_F32, offset_host ? d_off : nullptr,
                         FFQ_F32, layout, num_bits, s.stream);
  if (rc != FFQ_OK) return rc;
  rc = ffq_quantize_bwd(dx_x, dtype, dx_g, dtype, dx_dx, d_dsc, offset_host ? d_doff : nullptr, d_scale, FFQ_F32,
                        offset_host ? d_off : nullptr, FFQ_F32, layout, num_bits, d_ws, wbytes, s.stream);
  if (rc != FFQ_OK) return rc;
  FFQ_CUDA_CHECK(cudaMemcpyAsync(y_host, dx_y, raw, cudaMemcpyDeviceToHost, s.stream));
  FFQ_CUDA_CHECK(cudaMemcpyAsync(dx_host, dx_dx, raw, cudaMemcpyDeviceToHost, s.stream));
  FFQ_CUDA_CHECK(cudaMemcpyAsync(dscale_host, d_dsc, praw, cudaMemcpyDeviceToHost, s.stream));
  if (offset_host && doffset_host)
    FFQ_CUDA_CHECK(cudaMemcpyAsync(doffset_host, d_doff, praw, cudaMemcpyDeviceToHost, s.stream));
  FFQ_CUDA_CHECK(cudaStreamSynchronize(s.stream));
  return FFQ_OK;
}

#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 quantization hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): Llama-3-8B-shape random-init decoder stack, W8 per-channel /
A8 per-tensor RunningMinMax calibration, seq 2048, one batch per step per GPU (data-parallel
calibration: weak scaling; the ranges of all ranks are all-reduced once when the estimate_ranges
block ends, inside the timed region).  A step = one calibration forward = the hot path of all 224
quantized linears (min/max + range->params + quantize for 448 quantizers, 224 quantized linears)
plus the library ops around them.  Prints ONE JSON line (see the task contract)."""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

SEQ = 2048


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference+plugin"],
                    help="ours: this repository.  reference: the unmodified reference (oracle/_ref) on the host cores.  "
                         "reference+plugin: the unmodified reference's host code on the GPU with plugin.install() underneath")
    ap.add_argument("--layers", type=int, default=None, help="decoder layers to instantiate (default: all 32)")
    ap.add_argument("--seq", type=int, default=SEQ)
    ap.add_argument("--shape", default="8b", choices=["8b", "70b", "tiny"])
    ap.add_argument("--no-graph", action="store_true", help="run the steps eagerly instead of replaying a CUDA graph")
    ap.add_argument("--cpu-sample-layers", type=int, default=1)
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-extras", action="store_true")
    ap.add_argument("--skip-compiled-baseline", action="store_true",
                    help="skip the reference's compiled_quant_funcs timing of cfg1 (saves ~1 min of inductor compilation)")
    ap.add_argument("--skip-drop-in", action="store_true", help="skip the reference+plugin measurements of the default line")
    ap.add_argument("--memoize-parameters", action="store_true",
                    help="headline with the estimator's default memoisation of unchanged weights ON (default: OFF, so that every "
                         "timed step re-quantizes every weight exactly as the reference's step does)")
    ap.add_argument("--overlap-parameters", type=int, default=0,
                    help="estimator option overlap_parameters: launch each weight's fused calibration step this many weight "
                         "quantizers ahead of its use on a side stream (0: in line on the step's stream)")
    ap.add_argument("--overlap-sweep", default="", help="comma-separated overlap_parameters values timed as extra ablations")
    ap.add_argument("--workload", default="calib", choices=["calib", "w4a16-calib", "wq4", "cfg5"],
                    help="calib: configs[1] (default, W8A8 8B-shape).  w4a16-calib: configs[4] recipe (W4 g=128 / A16) on "
                         "--shape.  wq4: configs[2], W4 g=128 weight fake-quant of all linears sharded by layer")
    return ap.parse_args()


def shape_of(name):
    import bench_workloads as bw
    return {"8b": bw.LLAMA3_8B, "70b": bw.LLAMA3_70B, "tiny": bw.TINY}[name]


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int) -> None:
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.device_index = device_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.device_index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); smax.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(smax), reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------
# per-kernel timing for the roofline object
# ---------------------------------------------------------------------------------------------
class KernelCensus:
    """Records every hot-path launch of ONE eager step (the callable and its live arguments), then
    replays each kind of launch back to back from a CUDA graph bracketed by CUDA events on the
    launching stream.  Eager per-launch events would time the host (the eager step is
    dispatch-bound); the graph replay keeps the device busy, so total/launches is the kernel's
    average duration at exactly the shapes and operands of the timed step."""

    def __init__(self, ff):
        from fastforward_b200 import _cabi
        from fastforward_b200.nn import qlinear
        self.ff, self.C, self.qlinear = ff, _cabi, qlinear
        self.calls = {}     # kind -> list of (closure, algorithmic units)
        self._undo = []

    def _nbytes(self, *ts):
        return sum(t.numel() * t.element_size() for t in ts if isinstance(t, torch.Tensor))

    def install(self):
        ops, C = self.ff.ops, self.C
        self.qlinear.keepalive = []          # recorded C-ABI pointers must outlive the replay

        def wrap(obj, name, kind, units_fn, c_abi=False):
            orig = getattr(obj, name)

            def rec(*a, **k):
                out = orig(*a, **k)
                if c_abi:     # last argument is the stream: re-resolve it at replay time (graph capture stream)
                    call = lambda: orig(*a[:-1], torch.cuda.current_stream().cuda_stream)
                else:
                    call = lambda: orig(*a, **k)
                self.calls.setdefault(kind, []).append((call, units_fn(a, k, out)))
                return out
            setattr(obj, name, rec)
            self._undo.append((obj, name, orig))

        wrap(ops, "quantize_by_tile", "quantize (ew_row_kernel<QUANT>)", lambda a, k, o: self._nbytes(a[0], o))
        wrap(ops, "running_minmax_update_", "running min/max (mm_row_*_kernel)", lambda a, k, o: self._nbytes(a[2]))
        wrap(ops, "calibrate_quantize_", "fused calibration step (calq_rows/calq_tensor_kernel)",
             lambda a, k, o: self._nbytes(a[2], o[0]))
        wrap(ops, "dequantize_by_tile", "dequantize (ew_row_kernel<DEQUANT>)", lambda a, k, o: self._nbytes(a[0], o))
        # the GEMM at C-ABI level: args 4..6 are M, N, K
        wrap(C.lib, "ffq_qlinear_w8a8", "w8a8 linear (w8a8_gemm_kernel)", lambda a, k, o: 2.0 * a[4] * a[5] * a[6], c_abi=True)
        wrap(C.lib, "ffq_rowsum_i8", "rowsum (rowsum_i8_kernel)", lambda a, k, o: float(a[2] * a[3]), c_abi=True)

    def remove(self):
        for obj, name, orig in self._undo:
            setattr(obj, name, orig)
        self._undo = []
        self._alive, self.qlinear.keepalive = self.qlinear.keepalive, None

    def replay(self, reps=3):
        out = {}
        for kind, calls in self.calls.items():
            for fn, _ in calls[:3]:
                fn()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for fn, _ in calls:
                    fn()
            g.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                g.replay()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            units = sum(u for _, u in calls)
            out[kind] = dict(launches=len(calls), total_ms=round(ms, 3), avg_us=round(1e3 * ms / len(calls), 2),
                             units=units, rate=units / (ms * 1e-3))
            del g
        return out


# ---------------------------------------------------------------------------------------------
# the other two pieces of BASELINE.json's metric, measured in isolation on this rank's GPU
# ---------------------------------------------------------------------------------------------
def _time_cuda(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    return ts[len(ts) // 2] * 1e-3


def _time_graph(fn, calls_per_replay, reps=5):
    """Device time of `fn` per call with the launches replayed from a CUDA graph (no host gaps)."""
    for _ in range(calls_per_replay):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(calls_per_replay):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / (reps * calls_per_replay)


def measure_int8_library_peak(dev):
    """Sustained int8 tensor throughput of the library GEMM (cuBLASLt through torch._int_mm) at 8192^3: the measured
    stand-in for an int8 peak, which MEASURED_PEAKS.json does not carry.  Context for `frac`, never on the product path."""
    n = 8192
    a = torch.randint(-128, 128, (n, n), dtype=torch.int8, device=dev)
    b = torch.randint(-128, 128, (n, n), dtype=torch.int8, device=dev)
    t = _time_cuda(lambda: torch._int_mm(a, b.t()), iters=30, warm=5)
    return 2.0 * n ** 3 / t / 1e12


def measure_extras(ff, dev, hbm_peak, int8_peak, int8_lib_peak):
    """fake-quant fwd+bwd GB/s (configs[0]/cfg1 shape and a weight-sized bf16 tensor), W8A8 linear TOPS
    (configs[3]) and W4 g=128 weight QDQ GB/s (configs[2], one Llama-3-8B gate_proj).  Inputs exceed L2
    or are cycled so that no iteration re-reads L2-resident data.  Runs on EVERY rank (replicas)."""
    import ctypes
    from fastforward_b200 import _cabi as C, ops
    out = {}
    for name, shape, dt in (("cfg1_4096x4096_fp32_perchannel8", (4096, 4096), torch.float32),
                            ("16384x4096_fp32_perchannel8", (16384, 4096), torch.float32),
                            ("14336x4096_bf16_perchannel8", (14336, 4096), torch.bfloat16)):
        torch.manual_seed(0)
        nbuf = max(1, int(400e6 // (shape[0] * shape[1] * torch.empty(0, dtype=dt).element_size() * 3)) + 1)
        xs = [torch.randn(shape, device=dev, dtype=dt) for _ in range(nbuf)]
        gs = [torch.randn(shape, device=dev, dtype=dt) for _ in range(nbuf)]
        tile = (1, shape[1])
        mn, mx = ops.tile_minmax(xs[0], tile)
        scale = torch.empty(shape[0], device=dev); offset = torch.empty(shape[0], device=dev)
        ops.parameters_for_range_(mn, mx, 8, True, True, scale, offset)
        it = [0]

        def step():
            i = it[0] % nbuf; it[0] += 1
            ops.fake_quantize_by_tile(xs[i], scale, tile, 8.0, None, offset)
            ops.quantize_by_tile_backward(xs[i], gs[i], scale, tile, 8.0, offset)
        t = _time_graph(step, calls_per_replay=max(4, 2 * nbuf))
        t_eager = _time_cuda(step)
        by = 5 * xs[0].numel() * xs[0].element_size()
        out[name] = {"fwd_bwd_us": round(t * 1e6, 1), "GBps": round(by / t / 1e9, 1), "frac_of_measured_hbm": round(by / t / 1e9 / hbm_peak, 3),
                     "algorithmic_MB": round(by / 1e6, 1), "buffers_cycled": nbuf, "timing": "CUDA-graph replay of the two launches",
                     "eager_python_us": round(t_eager * 1e6, 1)}
        if name.startswith("cfg1"):
            # end to end with HOST buffers through the C ABI (H2D + kernels + D2H inside the call)
            xh, gh = xs[0].cpu().pin_memory(), gs[0].cpu().pin_memory()
            yh, dxh = torch.empty_like(xh).pin_memory(), torch.empty_like(xh).pin_memory()
            sh_, oh_ = scale.cpu(), offset.cpu()
            dsc, dof = torch.empty(shape[0]), torch.empty(shape[0])
            lay = C.make_layout(shape, tile)

            def host_step():
                C.check(C.lib.ffq_fakequant_fwd_bwd_host(xh.data_ptr(), gh.data_ptr(), C.dtype_tag(dt), yh.data_ptr(), dxh.data_ptr(),
                                                          dsc.data_ptr(), dof.data_ptr(), sh_.data_ptr(), oh_.data_ptr(),
                                                          ctypes.byref(lay), 8.0, dev.index or 0))
            host_step()
            t0 = time.perf_counter()
            for _ in range(5):
                host_step()
            th = (time.perf_counter() - t0) / 5
            out[name]["e2e_host_buffers"] = {"ms": round(th * 1e3, 2), "GBps_algorithmic": round(by / th / 1e9, 1),
                                             "h2d_bytes": 2 * xh.numel() * 4, "d2h_bytes": 2 * xh.numel() * 4}
        if name.startswith("cfg1"):
            # asymmetric (doffset live) and clipping (range = 0.5 x true range: clip branches taken) variant
            ops.parameters_for_range_(mn * 0.5, mx * 0.5, 8, False, True, scale, offset)
            t2 = _time_graph(step, calls_per_replay=max(4, 2 * nbuf))
            out[name]["asymmetric_clipping_variant"] = {"fwd_bwd_us": round(t2 * 1e6, 1), "GBps": round(by / t2 / 1e9, 1),
                                                        "frac_of_measured_hbm": round(by / t2 / 1e9 / hbm_peak, 3)}
        del xs, gs
    # W8A8 linear, configs[3]
    M, N, K = 8192, 14336, 4096
    qx = torch.randint(-128, 128, (M, K), dtype=torch.int8, device=dev)
    qw = torch.randint(-128, 128, (N, K), dtype=torch.int8, device=dev)
    y = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    sx = torch.tensor([0.01], device=dev); ox = torch.tensor([3.0], device=dev); sw = torch.rand(N, device=dev) * 0.01
    rs = torch.empty(N, dtype=torch.int32, device=dev)
    st = C.current_stream(dev)
    C.check(C.lib.ffq_rowsum_i8(qw.data_ptr(), rs.data_ptr(), N, K, st))

    def gemm():
        C.check(C.lib.ffq_qlinear_w8a8(qx.data_ptr(), qw.data_ptr(), y.data_ptr(), 2, M, N, K, sx.data_ptr(), ox.data_ptr(), sw.data_ptr(),
                                       None, rs.data_ptr(), None, None, 255, None, st))
    t = _time_cuda(gemm)
    t_lib = _time_cuda(lambda: torch._int_mm(qx, qw.t()))
    # ... with the output quantizer fused into the epilogue (int8 codes + their row sums instead of the bf16 tensor)
    codes = torch.empty(M, N, dtype=torch.int8, device=dev)
    rs_out = torch.zeros(M, dtype=torch.int32, device=dev)
    oq_s = torch.tensor([0.05], device=dev); oq_o = torch.tensor([-2.0], device=dev)
    rq = C.Requant(oq_s.data_ptr(), oq_o.data_ptr(), 8.0, codes.data_ptr(), rs_out.data_ptr())

    def gemm_requant():
        C.check(C.lib.ffq_qlinear_w8a8(qx.data_ptr(), qw.data_ptr(), None, 2, M, N, K, sx.data_ptr(), ox.data_ptr(), sw.data_ptr(),
                                       None, rs.data_ptr(), None, None, 255, ctypes.byref(rq), st))
    t_rq = _time_cuda(gemm_requant)
    out["w8a8_linear_8192x14336x4096"] = {"us": round(t * 1e6, 1), "TOPS": round(2 * M * N * K / t / 1e12, 1),
                                          "frac_of_int8_peak": round(2 * M * N * K / t / 1e12 / int8_peak, 3),
                                          "frac_of_library_int8_sustained": round(2 * M * N * K / t / 1e12 / int8_lib_peak, 3),
                                          "context_cublaslt_int_mm_TOPS": round(2 * M * N * K / t_lib / 1e12, 1),
                                          "fused_output_requant_us": round(t_rq * 1e6, 1),
                                          "fused_output_requant_TOPS": round(2 * M * N * K / t_rq / 1e12, 1)}
    del qx, qw, y, codes
    # W4A16 linear (configs[4] recipe at the configs[3] shape): bf16 activations x 4-bit g=128 codes
    for M in (256, 2048):
        x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
        qw = torch.randint(-8, 8, (N, K), dtype=torch.int8, device=dev)
        sw4 = torch.rand(N * (K // 128), device=dev) * 0.01 + 1e-3
        ow4 = torch.randint(-3, 4, (N * (K // 128),), device=dev).float()
        y = torch.empty(M, N, dtype=torch.bfloat16, device=dev)

        def w4():
            C.check(C.lib.ffq_qlinear_w4a16(x.data_ptr(), 2, qw.data_ptr(), y.data_ptr(), M, N, K, sw4.data_ptr(), ow4.data_ptr(), 128,
                                            None, 255, st))

        def w4_fallback():      # the reference's route with our dequantize kernel underneath + the library bf16 GEMM
            wd = ops.dequantize_by_tile(qw, sw4, (1, 128), ow4, torch.bfloat16)
            torch.nn.functional.linear(x, wd)
        t = _time_cuda(w4)
        t_fb = _time_cuda(w4_fallback)
        wd = ops.dequantize_by_tile(qw, sw4, (1, 128), ow4, torch.bfloat16)
        t_mm = _time_cuda(lambda: torch.nn.functional.linear(x, wd))
        out[f"w4a16_linear_{M}x{N}x{K}_g128"] = {
            "us": round(t * 1e6, 1), "TFLOPS": round(2 * M * N * K / t / 1e12, 1),
            "frac_of_bf16_peak": round(2 * M * N * K / t / 1e12 / (int8_peak / 2), 3),
            "context_dequant_plus_cublas_us": round(t_fb * 1e6, 1), "context_cublas_bf16_gemm_only_us": round(t_mm * 1e6, 1)}
        del x, qw, y, wd
    # W4 g=128 weight QDQ of one 14336x4096 bf16 weight (configs[2] unit of work): min/max + params + fused QDQ in place
    w = [torch.randn(14336, 4096, device=dev, dtype=torch.bfloat16) * 0.02 for _ in range(3)]
    tile = (1, 128)
    nt = w[0].numel() // 128
    scale = torch.empty(nt, device=dev); offset = torch.empty(nt, device=dev)
    it = [0]

    def qdq_fused():        # calibrate + snap in ONE launch (ffq_calibrate_fakequant), 2s bytes/element
        i = it[0] % 3; it[0] += 1
        ops.calibrate_fake_quantize_(w[i], tile, 4, True, True, scale, offset, None, out=w[i])
    t = _time_graph(qdq_fused, calls_per_replay=6)
    by = 2 * w[0].numel() * 2
    out["w4_g128_weight_qdq_fused_14336x4096_bf16"] = {"us": round(t * 1e6, 1), "GBps": round(by / t / 1e9, 1),
                                                        "frac_of_measured_hbm": round(by / t / 1e9 / hbm_peak, 3),
                                                        "what": "one launch (+ the early-exit fix-up): read once, write once"}
    del w
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------------
# configs[2] and configs[4] as measurements that run on every rank of `bench.py --gpus N`
# ---------------------------------------------------------------------------------------------
def _dist_reduce(vals, op, dev, world):
    import torch.distributed as dist
    t = torch.tensor(vals, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=op)
    return t.tolist()


def measure_cfg3(ff, sh, dev, rank, world, hbm_peak, steps=5, warmup=2):
    """configs[2]: W4 per-group (g=128) weight fake-quant of every decoder linear of the 8B-shape model, sharded by
    layer (layer i -> rank i mod N), no data-path collective: strong scaling.  A step = calibrate each owned weight
    quantizer on its weight and snap the weight in place (one fused launch per weight, 2s bytes per element)."""
    import torch.distributed as dist

    import bench_workloads as bw
    from fastforward_b200 import _cabi
    from fastforward_b200.quantization.fuse import calibrate_and_fuse_qdq_weights

    mine = list(range(rank, sh.layers, world))
    model = torch.nn.ModuleList(bw.DecoderLayer(sh, torch.bfloat16, dev) for _ in mine)   # only this rank's layers
    bw.init_weights_(model, seed=rank)
    ff.quantize_model(model, extra_conversion=ff.surrogate_quantized_modules(model))
    ff.find_quantizers(model, "**/[quantizer:parameter/weight]").initialize(
        ff.nn.LinearQuantizer, num_bits=4, granularity=ff.PerBlock(block_dims=1, block_sizes=128, per_channel_dims=0))
    model.to(dev)                    # the quantizers were created on the CPU
    n_weights = sum(m.weight.numel() for m in model.modules() if isinstance(m, torch.nn.Linear))

    def step():
        calibrate_and_fuse_qdq_weights(model)
    for _ in range(warmup):
        step()
    l0 = _cabi.launch_count()
    step()
    launches = _cabi.launch_count() - l0
    torch.cuda.synchronize()
    cg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(cg):
        step()
    cg.replay()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        cg.replay()
    t1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = t0.elapsed_time(t1) * 1e-3 / steps
    dt_max = _dist_reduce([dt], dist.ReduceOp.MAX, dev, world)[0]
    n_all = _dist_reduce([float(n_weights)], dist.ReduceOp.SUM, dev, world)[0]
    del cg, model
    torch.cuda.empty_cache()
    by = 2 * 2 * n_all
    return {"what": f"{sh.name}: {sh.layers} layers x 7 linears = {n_all / 1e9:.2f} G weights, LinearQuantizer(4, PerBlock g=128), "
                    "calibrate_and_fuse_qdq_weights (one fused launch per weight, in place, 2s bytes/element); layer i -> rank i mod N; "
                    "CUDA-graph replay; time = max over ranks",
            "n_gpus": world, "scaling": "strong", "ms_whole_model": round(dt_max * 1e3, 3),
            "aggregate_GBps": round(by / dt_max / 1e9, 1), "per_gpu_GBps": round(by / dt_max / 1e9 / world, 1),
            "frac_of_measured_hbm_per_gpu": round(by / dt_max / 1e9 / world / hbm_peak, 3), "launches_per_step_per_rank": int(launches)}


def measure_cfg5(ff, dev, rank, world, steps=2, layers=None, seq=SEQ):
    """configs[4]: Llama-3-70B-shape W4 (g=128, int8 container) / A16 (per-tensor asymmetric, fp32 codes) data-parallel
    calibration.  The 80 layers are STREAMED (random-init one layer, calibrate it on this rank's batches inside its own
    estimate_ranges block with the NCCL MIN/MAX range exchange at block exit, free it): 137 GB of bf16 weights never have
    to be resident.  Timed: every layer's estimate_ranges block (CUDA events, summed; weight initialisation excluded);
    tokens/s = N x steps x seq / max-over-ranks time."""
    import torch.distributed as dist

    import bench_workloads as bw
    from fastforward_b200 import _cabi
    from fastforward_b200.nn import qlinear

    sh = bw.LLAMA3_70B
    n_layers = layers or sh.layers
    qlinear.install()
    g = torch.Generator().manual_seed(4321 + rank)
    hidden = [(torch.randn(1, seq, sh.hidden, generator=g) * 0.5).to(dev, torch.bfloat16) for _ in range(steps)]
    total_ms, exit_ms = 0.0, 0.0
    calls0 = qlinear.stats().get("calls_w4a16", 0)
    l0 = _cabi.launch_count()
    for li in range(n_layers):
        layer = bw.DecoderLayer(sh, torch.bfloat16, dev)
        bw.init_weights_(layer, seed=1000 + li)
        ff.quantize_model(layer, extra_conversion=ff.surrogate_quantized_modules(layer))
        ff.set_strict_quantization(False)
        ff.find_quantizers(layer, "**/[quantizer:parameter/weight]").initialize(
            ff.nn.LinearQuantizer, num_bits=4, quantized_dtype=torch.int8,
            granularity=ff.PerBlock(block_dims=1, block_sizes=128, per_channel_dims=0))
        ff.find_quantizers(layer, "**/[quantizer:activation/input]").initialize(
            ff.nn.LinearQuantizer, num_bits=16, symmetric=False, granularity=ff.PerTensor(), quantized_dtype=torch.float32)
        layer.to(dev)
        est = ff.range_setting.running_minmax(sync_ranges=world > 1, memoize_parameters=False)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        with torch.no_grad(), ff.estimate_ranges(layer, est):
            for b in range(steps):
                hidden[b] = layer(hidden[b])
            e1.record()
        e2.record()
        torch.cuda.synchronize()
        total_ms += e0.elapsed_time(e2)
        exit_ms += e1.elapsed_time(e2)
        del layer, est
    w4_calls = qlinear.stats().get("calls_w4a16", 0) - calls0
    launches = _cabi.launch_count() - l0
    torch.cuda.empty_cache()
    t_max, exit_max = _dist_reduce([total_ms, exit_ms], dist.ReduceOp.MAX, dev, world)
    finite = bool(torch.isfinite(hidden[-1].float()).all())
    return {"what": f"{sh.name}, {n_layers} of {sh.layers} layers streamed, W4 g=128 (int8 container) + A16 per-tensor asymmetric "
                    f"LinearQuantizers, estimate_ranges(running_minmax) per layer, {steps} batch(es) [1,{seq}] per GPU, eager (no CUDA graph), "
                    "weights re-quantized every step, NCCL MIN/MAX of the activation ranges at every block exit",
            "n_gpus": world, "scaling": "weak", "tokens_per_s": round(world * steps * seq / (t_max * 1e-3 * sh.layers / n_layers), 1),
            "ms_per_step_all_layers": round(t_max / steps, 2), "block_exit_ms_total": round(exit_max, 2),
            "w4a16_kernel_calls": int(w4_calls), "linears_x_steps": 7 * n_layers * steps, "ffq_launches": int(launches),
            "outputs_finite": finite}


# ---------------------------------------------------------------------------------------------
# ours
# ---------------------------------------------------------------------------------------------
METRIC_8B = "calib tokens/s (Llama-3-8B-shape W8 per-channel / A8 per-tensor RunningMinMax calibration, seq 2048)"


def workload_config(sh, layers, seq, world):
    """The `config` object both arms print (the reference arm describes its sampling in `cpu_baseline.sample`)."""
    return {"workload": f"{sh.name} decoder stack ({layers} layers, 7 quantized linears each), W8 PerChannel(0) symmetric + "
                        f"A8 PerTensor asymmetric LinearQuantizers (int8 codes), estimate_ranges(running_minmax), "
                        f"batch [1,{seq}] per GPU per step, random-init normal(0,0.02), no lm_head",
            "parallelism": f"dp{world} calibration, one MIN/MAX all-reduce of ranges at block exit",
            "l2": "per-step working set (>= 14 GB of weights re-quantized every step) exceeds the 126 MB L2"}


def run_ours(args):
    import torch.distributed as dist

    import bench_workloads as bw
    import fastforward_b200 as ff
    from fastforward_b200 import _cabi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    sh = shape_of(args.shape)
    layers = args.layers or sh.layers
    seq = args.seq
    torch.manual_seed(0)
    model = bw.DecoderStack(sh, layers=layers, dtype=torch.bfloat16, device=dev)
    bw.init_weights_(model, seed=0)
    if args.workload == "w4a16-calib":
        # W4 per-group (g=128) symmetric weights in an int8 container, A16 per-tensor asymmetric (fp32 codes:
        # bf16 cannot hold 16 bits, _quantizer_impl.py:44-75); the linear is the W4A16 tcgen05 kernel
        extra = ff.surrogate_quantized_modules(model)
        ff.quantize_model(model, extra_conversion=extra)
        ff.set_strict_quantization(False)
        ff.find_quantizers(model, "**/layers/**/[quantizer:parameter/weight]").initialize(
            ff.nn.LinearQuantizer, num_bits=4, quantized_dtype=torch.int8,
            granularity=ff.PerBlock(block_dims=1, block_sizes=128, per_channel_dims=0))
        ff.find_quantizers(model, "**/layers/**/[quantizer:activation/input]").initialize(
            ff.nn.LinearQuantizer, num_bits=16, symmetric=False, granularity=ff.PerTensor(), quantized_dtype=torch.float32)
    else:
        bw.quantize_for_w8a8(ff, model)
    model.to(dev)
    from fastforward_b200.nn import qlinear
    qlinear.install()            # W8A8 tcgen05 kernel behind dispatcher "linear"

    # per-rank synthetic batches in pinned host memory (seed per rank)
    g = torch.Generator().manual_seed(1234 + rank)
    n_batches = args.warmup + 2 * args.steps + 2
    host_tokens = [torch.randint(0, sh.vocab, (1, seq), generator=g).pin_memory() for _ in range(n_batches)]
    dev_tokens = [t.to(dev) for t in host_tokens]
    static_tokens = torch.empty_like(dev_tokens[0])

    def make_estimator(memoize, overlap=None):
        return ff.range_setting.running_minmax(sync_ranges=world > 1, memoize_parameters=memoize,
                                               overlap_parameters=args.overlap_parameters if overlap is None else overlap)

    def barrier():
        if world > 1:
            dist.barrier()

    def region(steps, tokens_src, e2e, graph, memoize=False, dedupe=True, align_ranks_before_exit=False, overlap=None):
        """Enter estimate_ranges, warm up, time `steps` steps + block exit.  Returns seconds (device)."""
        out_host = torch.empty(max(steps, 1), dtype=torch.float32).pin_memory()    # one pinned slot per step
        estimator = make_estimator(memoize, overlap)
        estimator.dedupe = dedupe
        estimator._state.dedupe = dedupe
        with torch.no_grad(), ff.estimate_ranges(model, estimator):
            for i in range(args.warmup):
                static_tokens.copy_(dev_tokens[i])
                y = model(static_tokens)
            cg = None
            if graph:
                torch.cuda.synchronize()
                cg = torch.cuda.CUDAGraph()
                with torch.cuda.graph(cg):
                    y = model(static_tokens)
                    # the step's result (one scalar) is produced inside the captured step; only its read-back is per step
                    metric = y.float().abs().mean().reshape(1) if e2e else None
            if world > 1:   # NCCL sets up its MIN/MAX channels lazily: do that outside the timed region
                for op in (dist.ReduceOp.MIN, dist.ReduceOp.MAX):
                    dist.all_reduce(torch.zeros(1 << 14, dtype=torch.bfloat16, device=dev), op=op)
                dist.all_reduce(torch.zeros(1, dtype=torch.int32, device=dev), op=dist.ReduceOp.MAX)
                dist.all_reduce(torch.zeros(2, 8, dtype=torch.int64, device=dev), op=dist.ReduceOp.MAX)
            barrier(); torch.cuda.synchronize()
            t0, t1, tmid = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            wall0 = time.perf_counter()
            lc_timed0 = _cabi.launch_count()
            t0.record()
            for i in range(steps):
                static_tokens.copy_(tokens_src[args.warmup + i], non_blocking=True)
                if cg is not None:
                    cg.replay()
                else:
                    y = model(static_tokens)
                    metric = y.float().abs().mean().reshape(1) if e2e else None
                if e2e:
                    # D2H read of the step's result into that step's pinned slot; asynchronous like the H2D copy of the
                    # inputs, all of them complete before the region's closing synchronize
                    out_host[i:i + 1].copy_(metric, non_blocking=True)
            region.launches_per_timed_step = (_cabi.launch_count() - lc_timed0) / max(steps, 1)   # eager regions only
            if align_ranks_before_exit:      # diagnostic regions only: take the ranks' skew out of the exit time
                torch.cuda.synchronize(); barrier()
            tmid.record()
            wall_exit0 = time.perf_counter()
        # leaving the block: +-inf check (one sync) and, for N>1, the MIN/MAX all-reduce of the activation ranges
        t1.record()
        torch.cuda.synchronize()
        region.exit_wall_ms = (time.perf_counter() - wall_exit0) * 1e3
        barrier()
        if e2e and not bool(torch.isfinite(out_host[:steps]).all()):
            raise RuntimeError("bench: a step's result read back from the device is not finite")
        wall = time.perf_counter() - wall0
        dt = t0.elapsed_time(t1) * 1e-3
        region.exit_ms = tmid.elapsed_time(t1)
        region.stats = estimator.last_stats
        region.exit_detail = {k: (round(v, 3) if isinstance(v, float) else v) for k, v in estimator.last_exit.items()}
        return max(dt, 0.0), wall

    def reset_quantizers():
        for _, q in ff.nn.named_quantizers(model):
            q.reset_parameters()

    use_graph = not args.no_graph
    memo = bool(args.memoize_parameters)
    # ---- timed regions: (1) inputs resident in HBM, (2) end to end (pinned host tokens in, scalar out every step).
    # Each samples nvidia-smi clocks during its own steps.
    def timed(e2e, **kw):
        sampler = ClockSampler(local_rank)
        sampler.start()
        d, w = region(args.steps, host_tokens if e2e else dev_tokens, e2e=e2e, graph=use_graph, memoize=memo, **kw)
        return d, w, region.exit_ms, sampler.stop()

    launches0 = _cabi.launch_count()
    # one complete untimed estimate_ranges block first (enter, steps, exit): lazy one-time set-up that belongs to no
    # step -- NCCL's channels for the exit's collectives at their real sizes, the allocator's pools, cuDNN plans
    region(1, dev_tokens, e2e=False, graph=use_graph, memoize=memo)
    reset_quantizers()
    dt, wall, exit_ms, clk = timed(False)
    exit_detail = region.exit_detail
    exit_wall_ms = region.exit_wall_ms
    est_stats = region.stats
    reset_quantizers()
    dt_e2e, _, _, clk_e2e = timed(True)
    # ---- the same steps under other schedules (context, each its own estimate_ranges block) -----------------
    ablation = {}
    short = max(3, min(args.steps, 5))
    sweep = [int(v) for v in args.overlap_sweep.split(",") if v.strip()]
    if args.overlap_parameters and 0 not in sweep:
        sweep.insert(0, 0)
    for name, kw in [("eager_no_cuda_graph", dict(graph=False, memoize=memo)),
                     ("memoize_parameters_on" if not memo else "memoize_parameters_off", dict(graph=use_graph, memoize=not memo)),
                     ("no_dedupe_no_memoize", dict(graph=use_graph, memoize=False, dedupe=False))] + \
            [(f"overlap_parameters_{v}", dict(graph=use_graph, memoize=False, overlap=v)) for v in sweep]:
        reset_quantizers()
        d, _ = region(short, dev_tokens, e2e=False, **kw)
        d = allmax_value(d, dev, world)
        ablation[name] = {"tokens_per_s": round(world * short * seq / d, 1), "ms_per_step": round(1e3 * d / short, 3), "steps": short}
    # the block exit alone (ranks aligned first: inside the timed regions the exit also absorbs the ranks' skew)
    reset_quantizers()
    region(2, dev_tokens, e2e=False, graph=use_graph, memoize=memo, align_ranks_before_exit=True)
    exit_aligned_ms = allmax_value(region.exit_ms, dev, world)
    exit_aligned_wall_ms = allmax_value(region.exit_wall_ms, dev, world)
    # ---- instrumented eager pass for the roofline of the dominant kernel ------------------------
    reset_quantizers()
    # launches of this library per STEADY-STATE step: the timed steps of an eager region (its warm-up steps still issue
    # the fix-up launches of weights whose `settled` flag the host has not seen yet; a replayed graph does not pass
    # through the library's counter at all)
    region(2, dev_tokens, e2e=False, graph=False, memoize=memo)
    launches_per_step = region.launches_per_timed_step
    reset_quantizers()
    census = KernelCensus(ff)
    with torch.no_grad(), ff.estimate_ranges(model, make_estimator(memo)):
        static_tokens.copy_(dev_tokens[0])
        model(static_tokens)                      # first forward of a block materialises the lazy parameters
        census.install()
        model(static_tokens)                      # ONE recorded steady-state eager step
        census.remove()
    per_op = census.replay()
    qlin = qlinear.stats()

    dt, dt_e2e = allmax_value(dt, dev, world), allmax_value(dt_e2e, dev, world)
    exit_ms = allmax_value(exit_ms, dev, world)
    tokens = world * args.steps * seq

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    bf16_peak = float(peaks.get("bf16_tflops", 1590.0))
    int8_peak = 2.0 * bf16_peak      # kind::i8 issues at twice the kind::f16 rate on sm_100

    # ---- pieces measured on EVERY rank (replicas / shards) and aggregated ------------------------------------
    extras = cfg3 = cfg5 = None
    int8_lib_peak = None
    if not args.skip_extras:
        del census
        torch.cuda.empty_cache()
        int8_lib_peak = measure_int8_library_peak(dev)
        extras = measure_extras(ff, dev, hbm_peak, int8_peak, int8_lib_peak)
        if world > 1:      # configs[3] sharded over M: every rank runs its own 8192-token shard; aggregate = sum
            ent = extras["w8a8_linear_8192x14336x4096"]
            tot = _dist_reduce([ent["TOPS"]], dist.ReduceOp.SUM, dev, world)[0]
            mn = _dist_reduce([-ent["TOPS"]], dist.ReduceOp.MAX, dev, world)[0]
            ent.update(aggregate_TOPS_all_ranks=round(tot, 1), min_rank_TOPS=round(-mn, 1), ranks=world,
                       sharding="rows of X (8192 tokens per rank), W replicated, no collective")
        if args.shape == "8b" and args.workload == "calib":
            cfg3 = measure_cfg3(ff, sh, dev, rank, world, hbm_peak)
            cfg5 = measure_cfg5(ff, dev, rank, world)

    if rank == 0:
        dominant = max(per_op.items(), key=lambda kv: kv[1]["total_ms"]) if per_op else (None, None)
        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of one
        # decoder layer -- a profiler figure, never measured inside this run
        ncu_traffic = {}
        for fn in ("r02_traffic.json", "r01_traffic.json"):
            try:
                ncu_traffic = json.load(open(os.path.join(ROOT, "profiles", fn)))
                break
            except OSError:
                continue

        def traffic_of(kind):
            # the layer's 7 linears run on two kernels (CTA pairs; single CTAs for the k / v projections): launch-weighted mean
            ents = [v for k, v in ncu_traffic.items() if kind.startswith("w8a8") and k.startswith("w8a8_gemm") and isinstance(v, dict)]
            n = sum(e["launches_in_capture"] for e in ents)
            if not n or not (layers == sh.layers and seq == SEQ and args.shape == "8b"):
                return None
            return int(sum(e["dram_bytes_per_launch"] * e["launches_in_capture"] for e in ents) / n)
        roofline = None
        if dominant[0]:
            d = dominant[1]
            if dominant[0].startswith("w8a8"):
                roofline = dict(bound="tensor", kernel=dominant[0], achieved=round(d["rate"] / 1e12, 1), peak=int8_peak,
                                unit="TOP/s", frac=round(d["rate"] / 1e12 / int8_peak, 4), traffic=traffic_of(dominant[0]),
                                traffic_unit="DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over the "
                                             "7 linears of a layer); int8 operands + bf16 output are 67.7 MB per launch algorithmic, "
                                             "the rest is served by L2",
                                peak_source="2 x MEASURED_PEAKS.json bf16_tflops (burst): int8 MMA issues at twice the bf16 rate",
                                frac_of_library_int8_sustained=(round(d["rate"] / 1e12 / int8_lib_peak, 4) if int8_lib_peak else None),
                                # the 224 launches are timed back to back inside a 12 ms replay (the power-capped regime of a
                                # long step): against the SUSTAINED bf16 figure of the same file the fraction is higher; `frac`
                                # stays on the burst figure, the stricter of the two
                                frac_of_2x_bf16_sustained=(round(d["rate"] / 1e12 / (2.0 * float(peaks["bf16_tflops_sustained"])), 4)
                                                           if "bf16_tflops_sustained" in peaks else None),
                                algorithmic="2*M*N*K ops per launch, M=2048 (the step's 224 linears)")
            else:
                roofline = dict(bound="hbm", kernel=dominant[0], achieved=round(d["rate"] / 1e9, 1), peak=hbm_peak, unit="GB/s",
                                frac=round(d["rate"] / 1e9 / hbm_peak, 4), traffic=None, peak_source=peak_src)
            roofline.update(avg_launch_us=d["avg_us"], launches_per_step=d["launches"], kernel_ms_per_step=d["total_ms"],
                            method="all launches of this kind in one step replayed back to back from a CUDA graph, CUDA events on the launching stream")
        cpu_baseline = drop_in = cfg1_ref = None
        if not args.skip_cpu_baseline:
            import bench_reference as br
            cpu_baseline = br.cpu_calibration(sh, seq, layers, args.cpu_sample_layers, steps=3, warmup=1)
            if br.reference_available() and args.shape == "8b" and not args.skip_drop_in:
                del model
                torch.cuda.empty_cache()
                cfg1_ref = {}
                drop_in = br.gpu_calibration(sh, seq, layers, dev, cfg1=cfg1_ref, compiled=not args.skip_compiled_baseline)
        config = workload_config(sh, layers, seq, world)
        schedule = {"cuda_graph": use_graph, "dedupe_shared_inputs": True, "memoize_parameters": memo,
                    "overlap_parameters": args.overlap_parameters,
                    "note": "how THIS arm runs the config's steps; every timed step re-quantizes every weight unless "
                            "memoize_parameters is true"}
        line = {
            "metric": METRIC_8B if args.shape == "8b" else f"calib tokens/s ({sh.name})",
            "value": round(tokens / dt, 1), "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(1e3 * dt / args.steps, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16 data / fp32 quantizer arithmetic / int8 codes", "data": "synthetic",
            "config": config, "schedule": schedule,
            "e2e": {"value": round(tokens / dt_e2e, 1), "unit": "tokens/s", "h2d_bytes_per_step": seq * 8,
                    "d2h_bytes_per_step": 4, "clocks": clk_e2e},
            "gpu_launches": int(round(launches_per_step * args.steps)),
            "gpu_launches_per_step": round(launches_per_step, 1),
            "clocks": clk, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "kernels": {k: dict(launches=v["launches"], ms_per_step=v["total_ms"], avg_us=v["avg_us"],
                                rate=(f"{v['rate'] / 1e12:.0f} TOP/s" if k.startswith("w8a8") else f"{v['rate'] / 1e9:.0f} G(B|elem)/s"))
                        for k, v in per_op.items()},
            "estimator": est_stats, "ablation": ablation,
            "qlinear": qlin, "extras": extras, "cfg3_wq4": cfg3, "cfg5_70b_w4a16": cfg5,
            "cfg1_reference": cfg1_ref, "drop_in": drop_in,
            "peaks": {"hbm_GBps": hbm_peak, "int8_TOPS": int8_peak, "int8_library_sustained_TOPS": (round(int8_lib_peak, 1) if int8_lib_peak else None),
                      "note": "hbm and bf16 from MEASURED_PEAKS.json; int8 = 2 x measured bf16 burst; int8_library_sustained = "
                              "torch._int_mm (cuBLASLt) 8192^3 measured in this run"},
            "wall_s_timed_region": round(wall, 3), "block_exit_ms": round(exit_ms, 2),
            "block_exit_host_breakdown_ms": exit_detail,
            "block_exit_alone": {"device_ms": round(exit_aligned_ms, 2), "host_wall_ms": round(exit_aligned_wall_ms, 2),
                                 "note": "ranks aligned by a barrier first; block_exit_ms (inside the timed region) also contains "
                                         "the skew between ranks that the exit's collectives absorb"},
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def allmax_value(v, dev, world):
    if world == 1:
        return v
    import torch.distributed as dist
    t = torch.tensor([v], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ---------------------------------------------------------------------------------------------
# reference arms: the unmodified reference (oracle/_ref) through its own public API
# ---------------------------------------------------------------------------------------------
def run_reference(args):
    """`--impl reference`: the reference's own CPU eager path on the host cores.  K timed steps are really run; a step
    is one calibration forward of a bounded sample of the workload (`--cpu-sample-layers` of the decoder layers, same
    per-layer shapes), `ms_per_step` is that step's measured time and tokens/s is scaled to the full depth."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import bench_reference as br

    sh = shape_of(args.shape)
    layers = args.layers or sh.layers
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cb = br.cpu_calibration(sh, args.seq, layers, args.cpu_sample_layers, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference",
        "metric": METRIC_8B if args.shape == "8b" else f"calib tokens/s ({sh.name})",
        "value": cb["value"], "unit": "tokens/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_sample_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16 (CPU eager)", "data": "synthetic",
        "config": workload_config(sh, layers, args.seq, world),
        "sampling": {"layers_per_step": args.cpu_sample_layers, "of_layers": layers,
                     "note": "ms_per_step is the measured time of one sampled step; value = seq / (ms_per_step x of_layers / layers_per_step)"},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def run_reference_plugin(args):
    """`--impl reference+plugin`: the drop-in measured -- the unmodified reference's host code on cuda:0, alone and with
    this repository's kernels registered through plugin.install().  Eager, no CUDA graph (the reference's estimator
    synchronises the host twice per quantizer per forward)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import bench_reference as br

    sh = shape_of(args.shape)
    layers = args.layers or sh.layers
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    res = br.gpu_calibration(sh, args.seq, layers, dev, steps=args.steps, warmup=args.warmup)
    best = res["reference_plus_plugin_estimators"]
    line = {
        "impl": "reference+plugin",
        "metric": METRIC_8B if args.shape == "8b" else f"calib tokens/s ({sh.name})",
        "value": best["tokens_per_s"], "unit": "tokens/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": best["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16 data / fp32 quantizer arithmetic / int8 codes", "data": "synthetic",
        "config": workload_config(sh, layers, args.seq, 1), "drop_in": res,
    }
    print(json.dumps(line))


def run_wq4(args):
    """configs[2]: W4 per-group (g=128) weight fake-quant of every decoder linear of the 8B-shape model,
    sharded by layer across the ranks (layer i -> rank i mod N), no data-path collective.  A step = calibrate
    each owned weight quantizer on its weight (min/max + range->params) and snap the weight in place with the
    fused QDQ kernel (fuse_qdq_weights).  Reports GB/s of algorithmic traffic (3 x 2 bytes per weight)."""
    import torch.distributed as dist

    import bench_workloads as bw
    import fastforward_b200 as ff
    from fastforward_b200 import _cabi
    from fastforward_b200.quantization.fuse import calibrate_and_fuse_qdq_weights, calibrate_weight_quantizers, fuse_qdq_weights

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sh = shape_of(args.shape)
    layers = args.layers or sh.layers
    mine = list(range(rank, layers, world))
    model = torch.nn.ModuleList(bw.DecoderLayer(sh, torch.bfloat16, dev) for _ in mine)   # only this rank's layers
    bw.init_weights_(model, seed=rank)
    ff.quantize_model(model, extra_conversion=ff.surrogate_quantized_modules(model))
    ff.find_quantizers(model, "**/[quantizer:parameter/weight]").initialize(
        ff.nn.LinearQuantizer, num_bits=4, granularity=ff.PerBlock(block_dims=1, block_sizes=128, per_channel_dims=0))
    model.to(dev)
    n_weights = sum(m.weight.numel() for m in model.modules() if isinstance(m, torch.nn.Linear))

    two_step = os.environ.get("FFQ_WQ4_TWO_STEP") == "1"     # the separate calibrate + fuse launches, for comparison

    def step():
        if two_step:
            calibrate_weight_quantizers(model)
            fuse_qdq_weights(model)
        else:
            calibrate_and_fuse_qdq_weights(model)

    for _ in range(args.warmup):
        step()
    l0 = _cabi.launch_count()
    step()
    launches = (_cabi.launch_count() - l0) * args.steps
    # the step has no host sync and only in-place updates of persistent buffers: replay it from a CUDA graph
    # (--no-graph times the eager step, which is bound by Python/launch overhead: ~900 launches of 5-100 us)
    cg = None
    if not args.no_graph:
        torch.cuda.synchronize()
        cg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(cg):
            step()
        cg.replay()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        if cg is not None:
            cg.replay()
        else:
            step()
    t1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = t0.elapsed_time(t1) * 1e-3
    tot = torch.tensor([dt, float(n_weights)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = tot.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tot.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dt, n_all = float(mx[0]), float(sm[1])
    else:
        n_all = float(n_weights)
    if rank == 0:
        by = (3 if two_step else 2) * 2 * n_all * args.steps     # algorithmic: one read + one write (+ the min/max read)
        print(json.dumps({
            "metric": "W4 g=128 weight fake-quant GB/s (Llama-3-8B-shape, all decoder linears, sharded by layer)",
            "value": round(by / dt / 1e9, 1), "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(1e3 * dt / args.steps, 3), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{sh.name}: {layers} layers x 7 linears = {n_all / 1e9:.2f} G weights, LinearQuantizer(4, PerBlock g=128), "
                                   + ("calibrate_weight_quantizers + fuse_qdq_weights (3s bytes/element)" if two_step else
                                    "calibrate_and_fuse_qdq_weights: one fused launch per weight, in place (2s bytes/element)")
                                   + "; layer i -> rank i mod N",
                       "cuda_graph": cg is not None},
            "gpu_launches": int(launches), "per_gpu_GBps": round(by / dt / 1e9 / world, 1)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    elif a.impl == "reference+plugin":
        run_reference_plugin(a)
    elif a.workload == "wq4":
        run_wq4(a)
    else:
        run_ours(a)

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
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--layers", type=int, default=None, help="decoder layers to instantiate (default: all 32)")
    ap.add_argument("--seq", type=int, default=SEQ)
    ap.add_argument("--shape", default="8b", choices=["8b", "70b", "tiny"])
    ap.add_argument("--no-graph", action="store_true", help="run the steps eagerly instead of replaying a CUDA graph")
    ap.add_argument("--cpu-sample-layers", type=int, default=1)
    ap.add_argument("--skip-cpu-baseline", action="store_true")
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
# per-kernel timing (CUDA events on the launching stream) for the roofline object
# ---------------------------------------------------------------------------------------------
class OpTimer:
    """Wraps the op entry points with CUDA events; algorithmic bytes are computed from the
    tensors each call touches (DESIGN.md section 'algorithmic bytes')."""

    def __init__(self, ops):
        self.ops = ops
        self.records = {}   # name -> list of (start, end, bytes)
        self._orig = {}

    @staticmethod
    def _nbytes(*tensors):
        return sum(t.numel() * t.element_size() for t in tensors if isinstance(t, torch.Tensor))

    def install(self):
        ops = self.ops

        def wrap(name, fn, bytes_fn):
            def timed(*a, **k):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                out = fn(*a, **k)
                e.record()
                self.records.setdefault(name, []).append((s, e, bytes_fn(a, k, out)))
                return out
            self._orig[name] = fn
            setattr(ops, name, timed)

        wrap("quantize_by_tile", ops.quantize_by_tile, lambda a, k, o: self._nbytes(a[0], o))
        wrap("dequantize_by_tile", ops.dequantize_by_tile, lambda a, k, o: self._nbytes(a[0], o))
        wrap("running_minmax_update_", ops.running_minmax_update_, lambda a, k, o: self._nbytes(a[2]))
        wrap("fake_quantize_by_tile", ops.fake_quantize_by_tile, lambda a, k, o: self._nbytes(a[0], o))

    def remove(self):
        for name, fn in self._orig.items():
            setattr(self.ops, name, fn)

    def summary(self):
        out = {}
        for name, recs in self.records.items():
            ms = [s.elapsed_time(e) for s, e, _ in recs]
            by = [b for _, _, b in recs]
            out[name] = dict(launches=len(recs), total_ms=sum(ms), bytes=sum(by),
                             gbps=(sum(by) / (sum(ms) * 1e-3) / 1e9) if sum(ms) > 0 else 0.0,
                             avg_us=1e3 * sum(ms) / max(1, len(ms)))
        return out


# ---------------------------------------------------------------------------------------------
# ours
# ---------------------------------------------------------------------------------------------
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

    estimator = ff.range_setting.running_minmax(sync_ranges=world > 1)

    def barrier():
        if world > 1:
            dist.barrier()

    def region(steps, tokens_src, e2e, graph):
        """Enter estimate_ranges, warm up, time `steps` steps + block exit.  Returns seconds (device)."""
        out_host = torch.empty(1, dtype=torch.float32).pin_memory()
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
            barrier(); torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            wall0 = time.perf_counter()
            t0.record()
            for i in range(steps):
                static_tokens.copy_(tokens_src[args.warmup + i], non_blocking=True)
                if cg is not None:
                    cg.replay()
                else:
                    y = model(static_tokens)
                if e2e:
                    out_host.copy_(y.float().abs().mean().reshape(1), non_blocking=False)   # D2H read of the step's result
        # leaving the block: +-inf check (one sync) and, for N>1, the MIN/MAX all-reduce of all ranges
        t1.record()
        torch.cuda.synchronize(); barrier()
        wall = time.perf_counter() - wall0
        dt = t0.elapsed_time(t1) * 1e-3
        return max(dt, 0.0), wall

    def reset_quantizers():
        for _, q in ff.nn.named_quantizers(model):
            q.reset_parameters()

    use_graph = not args.no_graph
    # ---- timed region 1: inputs resident in HBM ------------------------------------------------
    clocks = ClockSampler(local_rank)
    launches0 = _cabi.launch_count()
    clocks.start()
    dt, wall = region(args.steps, dev_tokens, e2e=False, graph=use_graph)
    clk = clocks.stop()
    launches_eager_part = _cabi.launch_count() - launches0
    # ---- timed region 2: end to end (pinned host tokens in, scalar out every step) -------------
    reset_quantizers()
    dt_e2e, _ = region(args.steps, host_tokens, e2e=True, graph=use_graph)
    # ---- instrumented eager pass for the roofline of the dominant kernel ------------------------
    reset_quantizers()
    timer = OpTimer(ff.ops)
    timer.install()
    lc0 = _cabi.launch_count()
    region(2, dev_tokens, e2e=False, graph=False)
    launches_per_step = (_cabi.launch_count() - lc0) / (2 + args.warmup)
    timer.remove()
    per_op = timer.summary()
    qlin = qlinear.stats()

    def allmax(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    dt, dt_e2e = allmax(dt), allmax(dt_e2e)
    tokens = world * args.steps * seq
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        dominant = max(per_op.items(), key=lambda kv: kv[1]["total_ms"]) if per_op else (None, None)
        roofline = None
        if dominant[0]:
            d = dominant[1]
            roofline = dict(bound="hbm", kernel=dominant[0], achieved=round(d["gbps"], 1), peak=hbm_peak, unit="GB/s",
                            frac=round(d["gbps"] / hbm_peak, 4), traffic=None, peak_source=peak_src,
                            avg_launch_us=round(d["avg_us"], 2), launches_timed=d["launches"])
        cpu_baseline = None
        if not args.skip_cpu_baseline:
            cpu_baseline = run_reference_sample(args, sh, sample_layers=args.cpu_sample_layers, steps=1, warmup=0)
        line = {
            "metric": "calib tokens/s (Llama-3-8B-shape W8 per-channel / A8 per-tensor RunningMinMax calibration, seq 2048)"
            if args.shape == "8b" else f"calib tokens/s ({sh.name})",
            "value": round(tokens / dt, 1), "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(1e3 * dt / args.steps, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16 data / fp32 quantizer arithmetic / int8 codes", "data": "synthetic",
            "config": {"workload": f"{sh.name} decoder stack ({layers} layers, 7 quantized linears each), W8 PerChannel(0) symmetric + "
                                   f"A8 PerTensor asymmetric LinearQuantizers (int8 codes), estimate_ranges(running_minmax), "
                                   f"batch [1,{seq}] per GPU per step, random-init normal(0,0.02), no lm_head",
                       "parallelism": f"dp{world} calibration, one MIN/MAX all-reduce of ranges at block exit",
                       "cuda_graph": use_graph,
                       "l2": "per-step working set (>= 14 GB of weights re-quantized every step) exceeds the 126 MB L2"},
            "e2e": {"value": round(tokens / dt_e2e, 1), "unit": "tokens/s", "h2d_bytes_per_step": seq * 8,
                    "d2h_bytes_per_step": 4},
            "gpu_launches": int(round(launches_per_step * args.steps)),
            "gpu_launches_per_step": round(launches_per_step, 1),
            "clocks": clk, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "kernels": {k: {kk: (round(vv, 2) if isinstance(vv, float) else vv) for kk, vv in v.items()} for k, v in per_op.items()},
            "qlinear": qlin,
            "wall_s_timed_region": round(wall, 3),
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the same workload on the host cores
# ---------------------------------------------------------------------------------------------
def run_reference_sample(args, sh, sample_layers, steps, warmup):
    import bench_workloads as bw
    from oracle import workload as ow

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    seq = args.seq
    model = bw.DecoderStack(sh, layers=sample_layers, dtype=torch.bfloat16, device="cpu")
    bw.init_weights_(model, seed=0)
    ow.oracle_calibration_model(model)
    g = torch.Generator().manual_seed(1234)
    toks = [torch.randint(0, sh.vocab, (1, seq), generator=g) for _ in range(warmup + steps)]
    with torch.no_grad():
        for i in range(warmup):
            model(toks[i])
        t0 = time.perf_counter()
        for i in range(steps):
            model(toks[warmup + i])
        dt = time.perf_counter() - t0
    full_layers = args.layers or sh.layers
    tok_s = steps * seq / (dt * full_layers / sample_layers)
    return dict(value=round(tok_s, 2), unit="tokens/s", cores=cores, kind="port",
                sample=f"{sample_layers} of {full_layers} decoder layers at seq {seq}, {steps} step(s), bf16, "
                       f"torch CPU eager ops in the reference's order (oracle/workload.py); tokens/s scaled by "
                       f"{sample_layers}/{full_layers}", seconds=round(dt, 2))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sh = shape_of(args.shape)
    cb = run_reference_sample(args, sh, sample_layers=args.cpu_sample_layers, steps=max(1, min(args.steps, 3)),
                              warmup=min(args.warmup, 1))
    line = {
        "impl": "reference",
        "metric": "calib tokens/s (Llama-3-8B-shape W8 per-channel / A8 per-tensor RunningMinMax calibration, seq 2048)"
        if args.shape == "8b" else f"calib tokens/s ({sh.name})",
        "value": cb["value"], "unit": "tokens/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * args.seq / cb["value"], 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16 (CPU eager)", "data": "synthetic",
        "config": {"workload": f"{sh.name} decoder stack, same quantizers and batches as the GPU arm; CPU sample: {cb['sample']}"},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)

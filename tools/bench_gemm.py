#!/usr/bin/env python
"""W8A8 GEMM shapes of the calibration step and of BASELINE configs[3], per cluster size (B multicast across 1, 2 or 4
CTA pairs), next to the library int8 GEMM (context).  CUDA events, inputs cycled so operands do not stay in L2.

    python tools/bench_gemm.py [--json gpurun_out/gemm.json]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from fastforward_b200 import _cabi as C  # noqa: E402

SHAPES = [(8192, 14336, 4096), (2048, 4096, 4096), (2048, 14336, 4096), (2048, 4096, 14336), (2048, 1024, 4096),
          (2048, 8192, 8192), (2048, 28672, 8192)]


def time_fn(fn, iters=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--clusters", default="")
    ap.add_argument("--shapes", default="", help="comma list of MxNxK (default: the built-in list)")
    a = ap.parse_args()
    dev = torch.device("cuda")
    res = {}
    shapes = [tuple(int(v) for v in t.split("x")) for t in a.shapes.split(",") if t] or SHAPES
    for (M, N, K) in shapes:
        nbuf = 3
        qx = [torch.randint(-128, 128, (M, K), dtype=torch.int8, device=dev) for _ in range(nbuf)]
        qw = [torch.randint(-128, 128, (N, K), dtype=torch.int8, device=dev) for _ in range(nbuf)]
        y = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        sx = torch.tensor([0.01], device=dev); ox = torch.tensor([3.0], device=dev); sw = torch.rand(N, device=dev) * 0.01
        rs = torch.empty(N, dtype=torch.int32, device=dev)
        st = C.current_stream(dev)
        C.check(C.lib.ffq_rowsum_i8(qw[0].data_ptr(), rs.data_ptr(), N, K, st))
        it = [0]

        def gemm():
            i = it[0] % nbuf; it[0] += 1
            C.check(C.lib.ffq_qlinear_w8a8(qx[i].data_ptr(), qw[i].data_ptr(), y.data_ptr(), 2, M, N, K, sx.data_ptr(), ox.data_ptr(),
                                           sw.data_ptr(), None, rs.data_ptr(), None, None, 255, None, st))

        def lib():
            i = it[0] % nbuf; it[0] += 1
            torch._int_mm(qx[i], qw[i].t())
        ent = {}
        for c in [c for c in a.clusters.split(",") if c]:
            os.environ["FFQ_GEMM_CLUSTER"] = c
            try:
                t = time_fn(gemm)
                ent[f"cluster{c}"] = {"us": round(t * 1e6, 1), "TOPS": round(2 * M * N * K / t / 1e12, 1)}
            except Exception as e:  # noqa: BLE001
                ent[f"cluster{c}"] = f"error: {e}"
                torch.cuda.synchronize()
        os.environ.pop("FFQ_GEMM_CLUSTER", None)
        t = time_fn(gemm)
        ent["default"] = {"us": round(t * 1e6, 1), "TOPS": round(2 * M * N * K / t / 1e12, 1)}
        t = time_fn(lib)
        ent["cublaslt_int_mm"] = {"us": round(t * 1e6, 1), "TOPS": round(2 * M * N * K / t / 1e12, 1)}
        res[f"{M}x{N}x{K}"] = ent
        print(f"{M}x{N}x{K}", json.dumps(ent), flush=True)
        del qx, qw, y
    if a.json:
        os.makedirs(os.path.dirname(os.path.abspath(a.json)), exist_ok=True)
        json.dump(res, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()

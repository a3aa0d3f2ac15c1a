#!/usr/bin/env python
"""Where the warp-specialised W8A8 pair kernel waits: per role, the fraction of its lifetime spent blocked on a
barrier (ffq_debug_gemm_profile).  producer blocked = the MMA side is the bottleneck (stages full); MMA issuer blocked on
operands = L2->SM delivery is the bottleneck; on the accumulator = the epilogue is.

    python tools/prof_gemm_roles.py [M N K] [--cluster 2|4|8]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from fastforward_b200 import _cabi as C  # noqa: E402


def run(M, N, K, cluster=None):
    dev = torch.device("cuda")
    if cluster:
        os.environ["FFQ_GEMM_CLUSTER"] = str(cluster)
    else:
        os.environ.pop("FFQ_GEMM_CLUSTER", None)
    qx = torch.randint(-128, 128, (M, K), dtype=torch.int8, device=dev)
    qw = torch.randint(-128, 128, (N, K), dtype=torch.int8, device=dev)
    y = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    sx = torch.tensor([0.01], device=dev); ox = torch.tensor([3.0], device=dev); sw = torch.rand(N, device=dev) * 0.01
    rs = torch.empty(N, dtype=torch.int32, device=dev)
    st = C.current_stream(dev)
    C.check(C.lib.ffq_rowsum_i8(qw.data_ptr(), rs.data_ptr(), N, K, st))

    def gemm():
        C.check(C.lib.ffq_qlinear_w8a8(qx.data_ptr(), qw.data_ptr(), y.data_ptr(), 2, M, N, K, sx.data_ptr(), ox.data_ptr(),
                                       sw.data_ptr(), None, rs.data_ptr(), None, None, 255, None, st))
    for _ in range(3):
        gemm()
    prof = torch.zeros(16 * 8 * 148, dtype=torch.int64, device=dev)
    C.lib.ffq_debug_gemm_profile(prof.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gemm(); e1.record()
    torch.cuda.synchronize()
    C.lib.ffq_debug_gemm_profile(None)
    us = e0.elapsed_time(e1) * 1e3
    p = prof.view(-1, 16).double().cpu()
    live = p[p[:, 1] > 0]
    leaders = live[live[:, 4] > 0]
    out = {
        "shape": f"{M}x{N}x{K}", "cluster_ctas": cluster or 2, "us": round(us, 1), "TOPS": round(2 * M * N * K / us / 1e6, 1),
        "ctas": int(live.shape[0]),
        "kernel_clocks_median": float(live[:, 1].median()),
        "clock_MHz_implied": round(float(live[:, 1].median()) / us, 1),
        "producer_blocked_on_free_stage": round(float((live[:, 0] / live[:, 1]).mean()), 3),
        "mma_blocked_on_operands": round(float((leaders[:, 2] / leaders[:, 4]).mean()), 3),
        "mma_blocked_on_accumulator": round(float((leaders[:, 3] / leaders[:, 4]).mean()), 3),
        "epilogue_blocked_on_tile": round(float((live[:, 5] / live[:, 6].clamp_min(1)).mean()), 3),
    }
    ep = live[live[:, 12] > 0]
    if ep.shape[0]:
        tiles = ep[:, 12]
        out["epilogue_clocks_per_tile"] = {
            "total_busy": round(float(((ep[:, 6] - ep[:, 5]) / tiles).mean())),
            "column_parameters_and_barriers": round(float((ep[:, 7] / tiles).mean())),
            "tcgen05_ld": round(float((ep[:, 8] / tiles).mean())),
            "wait_staging_box_free": round(float((ep[:, 9] / tiles).mean())),
            "arithmetic_and_staging": round(float((ep[:, 10] / tiles).mean())),
            "fence_and_store_issue": round(float((ep[:, 11] / tiles).mean())),
            "tiles_per_cta": round(float(tiles.mean()), 1),
        }
    return out


if __name__ == "__main__":
    import json
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    shapes = [tuple(int(v) for v in args[:3])] if len(args) >= 3 else [(8192, 14336, 4096), (2048, 4096, 4096), (2048, 14336, 4096)]
    for shp in shapes:
        print(json.dumps(run(*shp)), flush=True)

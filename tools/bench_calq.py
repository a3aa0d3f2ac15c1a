"""Per-shape timing of the fused calibration step (ffq_calibrate_quantize) against the separate kernels.
Weights cycle through enough buffers to defeat L2 (they stream from HBM in the real step); activations
reuse one buffer (in the real step they were just written by the previous kernel)."""
import sys

import torch

sys.path.insert(0, ".."); sys.path.insert(0, ".")
from fastforward_b200 import ops  # noqa: E402

dev = torch.device("cuda")


def time_graph(fn, n, reps=5):
    for i in range(n):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(n):
            fn(i)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / (reps * n)


def run(shape, tile, symmetric, nbuf, dtype=torch.bfloat16, reset=False):
    """reset=True: the running range is reset to (+inf, -inf) before every step (two fill kernels, timed apart and
    subtracted), so every step is a 'the batch widens the range' step -- the slow case of the optimistic per-tensor pair."""
    xs = [(torch.randn(shape, device=dev) * 0.02).to(dtype) for _ in range(nbuf)]
    nt = (shape[0] // tile[0]) * (shape[1] // tile[1])
    mn = torch.full((nt,), float("inf"), dtype=dtype, device=dev); mx = -mn
    scale, offset = torch.empty(nt, device=dev), torch.empty(nt, device=dev)
    settled = torch.zeros(1, dtype=torch.int32, device=dev)
    keep = []

    def fills(i):
        mn.fill_(float("inf")); mx.fill_(float("-inf"))

    def fused(i):
        if reset:
            fills(i)
        keep.append(ops.calibrate_quantize_(mn, mx, xs[i % nbuf], tile, 8, symmetric, True, scale, offset, None, settled, rowsum=True))

    def unfused(i):
        x = xs[i % nbuf]
        ops.running_minmax_update_(mn, mx, x, tile)
        ops.parameters_for_range_(mn, mx, 8, symmetric, True, scale, offset)
        q = ops.quantize_by_tile(x, scale, tile, 8.0, torch.int8, offset)
        rs = torch.empty(shape[0], dtype=torch.int32, device=dev)
        from fastforward_b200 import _cabi as C
        C.check(C.lib.ffq_rowsum_i8(q.data_ptr(), rs.data_ptr(), shape[0], shape[1], C.current_stream(dev)))
        keep.append((q, rs))
    n = max(8, nbuf)
    tf = time_graph(fused, n)
    if reset:
        tf -= time_graph(fills, n)
    keep.clear()
    tu = time_graph(unfused, n)
    keep.clear()
    by = xs[0].numel() * (xs[0].element_size() + 1)
    print(("range reset every step: " if reset else "") + f"{str(shape):>16} tile={str(tile):>14} sym={symmetric!s:5} fused {tf * 1e6:7.1f} us ({by / tf / 1e9:6.0f} GB/s)   "
          f"unfused {tu * 1e6:7.1f} us ({by / tu / 1e9:6.0f} GB/s)", flush=True)


if __name__ == "__main__":
    run((4096, 4096), (1, 4096), True, 8)
    run((1024, 4096), (1, 4096), True, 16)
    run((14336, 4096), (1, 4096), True, 4)
    run((4096, 14336), (1, 14336), True, 4)
    run((2048, 4096), (2048, 4096), False, 1)
    run((2048, 14336), (2048, 14336), False, 1)
    run((8192, 4096), (8192, 4096), False, 1)
    run((2048, 4096), (2048, 4096), False, 1, reset=True)
    run((2048, 14336), (2048, 14336), False, 1, reset=True)
    run((4096, 4096), (1, 4096), True, 4, torch.float32)

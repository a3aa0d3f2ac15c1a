"""Time the fused MSE grid-search kernel (ops.grid_mse) against the candidate-by-candidate loop."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastforward_b200 import ops


def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for dt in (torch.float32, torch.bfloat16):
    for name, shape, tile in [("pc", (4096, 4096), (1, 4096)), ("g128", (4096, 4096), (1, 128)), ("pt", (2048, 4096), (2048, 4096))]:
        x = torch.randn(shape, device="cuda", dtype=dt)
        nt = x.numel() // (tile[0] * tile[1])
        C = 100
        mn, mx = ops.tile_minmax(x, tile)
        steps = torch.linspace(0.01, 1, C, device="cuda")
        cs = torch.empty(C, nt, device="cuda"); co = torch.empty(C, nt, device="cuda")
        for i in range(C):
            ops.parameters_for_range_(mn.float() * steps[i], mx.float() * steps[i], 4, False, True, cs[i], co[i])
        t_f = timeit(lambda: ops.grid_mse(x, cs, co, tile, 4.0))

        def loop():
            for i in range(C):
                y = ops.fake_quantize_by_tile(x, cs[i], tile, 4.0, None, co[i])
                torch.mean((y.view(nt, -1) - x.view(nt, -1)) ** 2, dim=1)
        t_l = timeit(loop, 2)
        el = x.numel() * C
        print(f"{str(dt):15s} {name:5s} fused {t_f:8.3f} ms ({el / t_f / 1e6:7.1f} G cand-elem/s)  loop(fused fake-quant + torch mse) {t_l:8.3f} ms  x{t_l / t_f:.1f}")

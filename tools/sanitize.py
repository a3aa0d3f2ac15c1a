"""One pass over every kernel family at small sizes (for compute-sanitizer)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fastforward_b200 as ff
from fastforward_b200 import ops
from fastforward_b200.nn import qlinear

dev = "cuda"
torch.manual_seed(0)
for shape, tile, dt in [((64, 1024), (1, 1024), torch.bfloat16), ((64, 1024), (1, 128), torch.float32), ((33, 77), (33, 77), torch.float32),
                        ((256, 2048), (256, 2048), torch.bfloat16), ((16, 32, 24), (4, 8, 6), torch.float16), ((40, 96), (40, 1), torch.float32)]:
    x = torch.randn(shape, device=dev, dtype=dt); g = torch.randn(shape, device=dev, dtype=dt)
    mn, mx = ops.tile_minmax(x, tile)
    s = torch.empty(mn.numel(), device=dev); o = torch.empty(mn.numel(), device=dev)
    ops.parameters_for_range_(mn, mx, 8, False, True, s, o)
    q = ops.quantize_by_tile(x, s, tile, 8.0, torch.int8, o)
    ops.dequantize_by_tile(q, s, tile, o, dt)
    ops.fake_quantize_by_tile(x, s, tile, 8.0, None, o)
    ops.quantize_by_tile_backward(x, g, s, tile, 8.0, o)
    ops.quantize_dynamic_by_tile(x, tile, 8.0, True, True, dt)
for m, k, n in [(300, 1040, 700), (128, 128, 256), (17, 96, 40)]:
    lin = torch.nn.Linear(k, n, bias=True, dtype=torch.bfloat16)
    ff.quantize_model(lin)
    lin.input_quantizer = ff.nn.LinearQuantizer(8, symmetric=False, quantized_dtype=torch.int8)
    lin.weight_quantizer = ff.nn.LinearQuantizer(8, symmetric=False, granularity=ff.PerChannel(0), quantized_dtype=torch.int8)
    lin.to(dev)
    x = torch.randn(m, k, device=dev, dtype=torch.bfloat16)
    lin.input_quantizer.quantization_range = (x.min(), x.max())
    lin.weight_quantizer.quantization_range = (lin.weight.min(1).values, lin.weight.max(1).values)
    qlinear.install()
    with torch.no_grad():
        lin(x)
# fused calibration step: per-channel rows (all vector-per-thread variants, deferred one-sided rows), per-tensor
# (single chunk, many chunks, ragged tail, rows straddling warp spans)
for shape, tile, dt, sym in [((7, 512), (1, 512), torch.bfloat16, True), ((5, 1024), (1, 1024), torch.float32, False),
                             ((3, 2056), (1, 2056), torch.float16, True), ((2, 4096), (1, 4096), torch.bfloat16, True),
                             ((2100, 4096), (1, 4096), torch.bfloat16, True), ((3, 20480), (1, 20480), torch.bfloat16, True),
                             ((8, 512), (8, 512), torch.bfloat16, False), ((3, 77, 264), (3, 77, 264), torch.float16, True),
                             ((1500, 4096), (1500, 4096), torch.bfloat16, False), ((700, 14336), (700, 14336), torch.bfloat16, False)]:
    nt = 1
    for d, t in zip(shape, tile):
        nt *= d // t
    for variant in ("mixed", "positive"):
        x = torch.randn(shape, device=dev).to(dt)
        x = x.abs() if variant == "positive" else x
        mn = torch.full((nt,), float("inf"), dtype=dt, device=dev); mx = -mn
        s = torch.empty(nt, device=dev); o = torch.empty(nt, device=dev)
        flags = torch.zeros(1, dtype=torch.int32, device=dev); settled = torch.zeros(1, dtype=torch.int32, device=dev)
        for _ in range(2):
            ops.calibrate_quantize_(mn, mx, x, tile, 8, sym, True, s, o, flags, settled, rowsum=True)
        assert int(flags.item()) == 0
# per-group tiles (int8 codes and fake-quant in place) and long rows with fake-quant output, incl. deferred tiles
for shape, tile, dt in [((24, 512), (1, 128), torch.bfloat16), ((33, 256), (1, 32), torch.float16), ((5, 64), (1, 4), torch.float32),
                        ((9, 1024), (1, 1024), torch.bfloat16), ((3, 20480), (1, 20480), torch.bfloat16), ((1001, 128), (1, 8), torch.bfloat16)]:
    nt = (shape[0] // tile[0]) * (shape[1] // tile[1])
    for variant in ("mixed", "positive", "some"):
        x = torch.randn(shape, device=dev)
        x = x.abs() if variant == "positive" else x
        if variant == "some":
            x[::2] = x[::2].abs()
        x = x.to(dt)
        s = torch.empty(nt, device=dev); o = torch.empty(nt, device=dev)
        mn = torch.full((nt,), float("inf"), dtype=dt, device=dev); mx = -mn
        for sym in (True, False):
            if ops.calibrate_quantize_mode(shape, tile, dt) == 3:
                ops.calibrate_quantize_(mn, mx, x, tile, 4, sym, True, s, o)
            ops.calibrate_fake_quantize_(x.clone(), tile, 4, sym, True, s, o, None, run_min=mn, run_max=mx)
            y = x.clone()
            ops.calibrate_fake_quantize_(y, tile, 8, sym, True, s, o, torch.int8, out=y)
# GPTQ block kernel: full and ragged blocks, few and many rows
from fastforward_b200.quantization import gptq as G
for rows, ncols in [(5, 128), (1000, 37), (33, 64)]:
    w = torch.randn(rows, 256, device=dev) * 0.1
    blk = w[:, 64:64 + ncols].contiguous(); q = torch.zeros_like(w); e = torch.zeros_like(w)
    hinv = torch.eye(256, device=dev) + torch.triu(torch.randn(256, 256, device=dev) * 0.01, 1)
    sc = torch.rand(rows * 2, device=dev) * 0.02 + 1e-3; of = torch.randn(rows * 2, device=dev)
    G.gptq_block_(blk, q[:, 64:64 + ncols], e[:, 64:64 + ncols], hinv[64:64 + ncols, 64:64 + ncols], sc, of,
                  torch.arange(64, 64 + ncols, dtype=torch.int32, device=dev), 1, 128, 2, 4)
# optimistic per-tensor pair without row sums (ragged last task) and with rows at the window's edges; LPBQ scale compression
for shape, rowsum in [((1000, 1004), True), ((333, 1000), False), ((64, 65536), True), ((9000, 256), True)]:
    x = torch.randn(shape, device=dev).bfloat16() if shape[1] % 8 == 0 else torch.randn(shape, device=dev)
    mn = torch.full((1,), float("inf"), dtype=x.dtype, device=dev); mx = -mn
    s = torch.empty(1, device=dev); o = torch.empty(1, device=dev)
    for scale in (1.0, 0.5, 2.0):
        ops.calibrate_quantize_(mn, mx, x * scale, shape, 8, False, True, s, o, None, None, rowsum=rowsum)
for shape, axis in [((300, 7), 0), ((5, 1000), 1), ((1, 1), 0), ((4096, 32), 0)]:
    ops.lpbq_encode(torch.rand(shape, device=dev) + 1e-3, axis, 4)
torch.cuda.synchronize()
print("sanitize pass done")

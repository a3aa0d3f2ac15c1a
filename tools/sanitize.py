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
torch.cuda.synchronize()
print("sanitize pass done")

import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastforward_b200 import ops
import fastforward_b200 as ff
dev = "cuda"
x = torch.randn(64, 256, device=dev, dtype=torch.bfloat16); g = torch.randn_like(x)
tile = (1, 256)
s = torch.rand(64, device=dev) + 0.1; o = torch.zeros(64, device=dev)
rmin = torch.full((64,), float("inf"), device=dev, dtype=torch.bfloat16); rmax = -rmin
flags = torch.zeros(1, dtype=torch.int32, device=dev)
def t(name, fn, n=2000):
    for _ in range(50): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
    print(f"{name:32s} {dt*1e6:7.1f} us/call")
t("quantize_by_tile", lambda: ops.quantize_by_tile(x, s, tile, 8.0, torch.int8, o))
q = ops.quantize_by_tile(x, s, tile, 8.0, torch.int8, o)
t("dequantize_by_tile", lambda: ops.dequantize_by_tile(q, s, tile, o, torch.bfloat16))
t("fake_quantize_by_tile", lambda: ops.fake_quantize_by_tile(x, s, tile, 8.0, None, o))
t("quantize_by_tile_backward", lambda: ops.quantize_by_tile_backward(x, g, s, tile, 8.0, o))
t("running_minmax_update_", lambda: ops.running_minmax_update_(rmin, rmax, x, tile, flags))
t("parameters_for_range_", lambda: ops.parameters_for_range_(rmin, rmax, 8, True, True, s, o))
t("torch.empty_like", lambda: torch.empty_like(x))
t("torch add (reference launch cost)", lambda: x + x)
lq = ff.nn.LinearQuantizer(8, granularity=ff.PerChannel(0), quantized_dtype=torch.int8, device=dev)
lq.quantization_range = (x.min(1).values.float(), x.max(1).values.float())
with torch.no_grad():
    t("LinearQuantizer(x)", lambda: lq(x))
    qt = lq(x)
    t("QuantizedTensor.dequantize()", lambda: qt.dequantize())
    with ff.estimate_ranges(lq, ff.range_setting.running_minmax):
        t("calibration step (estimator+quantize)", lambda: lq(x))

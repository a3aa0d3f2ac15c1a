"""Run one hot-path op a few times (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastforward_b200 import ops

dt = {"f32": torch.float32, "bf16": torch.bfloat16}[sys.argv[1]]
tile = {"pc": (1, 4096), "g128": (1, 128), "pt": (14336, 4096)}[sys.argv[2]]
which = sys.argv[3].split(",")
x = torch.randn(14336, 4096, device="cuda", dtype=dt)
g = torch.randn(14336, 4096, device="cuda", dtype=dt)
mn, mx = ops.tile_minmax(x, tile)
scale = torch.empty(mn.numel(), device="cuda"); offset = torch.empty(mn.numel(), device="cuda")
ops.parameters_for_range_(mn, mx, 8, False, True, scale, offset)
q = ops.quantize_by_tile(x, scale, tile, 8.0, torch.int8, offset)
for _ in range(3):
    if "fq" in which: ops.fake_quantize_by_tile(x, scale, tile, 8.0, None, offset)
    if "bwd" in which: ops.quantize_by_tile_backward(x, g, scale, tile, 8.0, offset)
    if "q8" in which: ops.quantize_by_tile(x, scale, tile, 8.0, torch.int8, offset)
    if "dq8" in which: ops.dequantize_by_tile(q, scale, tile, offset, dt)
    if "mm" in which: ops.tile_minmax(x, tile)
torch.cuda.synchronize()

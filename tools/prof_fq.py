"""Fake-quant forward + STE backward of one 14336x4096 bf16 per-channel tensor for ncu (source page)."""
import sys

import torch

sys.path.insert(0, ".")
from fastforward_b200 import ops  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(0)
x = torch.randn(14336, 4096, device=dev, dtype=torch.bfloat16)
g = torch.randn(14336, 4096, device=dev, dtype=torch.bfloat16)
tile = (1, 4096)
mn, mx = ops.tile_minmax(x, tile)
scale = torch.empty(14336, device=dev); offset = torch.empty(14336, device=dev)
ops.parameters_for_range_(mn, mx, 8, True, True, scale, offset)
for _ in range(3):
    ops.fake_quantize_by_tile(x, scale, tile, 8.0, None, offset)
    ops.quantize_by_tile_backward(x, g, scale, tile, 8.0, offset)
torch.cuda.synchronize()

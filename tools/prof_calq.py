"""One launch of each fused-calibration kernel shape for ncu."""
import sys

import torch

sys.path.insert(0, ".")
from fastforward_b200 import ops  # noqa: E402

dev = torch.device("cuda")
for shape, tile, sym in [((4096, 4096), (1, 4096), True), ((4096, 14336), (1, 14336), True), ((2048, 4096), (2048, 4096), False),
                         ((2048, 14336), (2048, 14336), False)]:
    x = (torch.randn(shape, device=dev) * 0.02).bfloat16()
    nt = (shape[0] // tile[0]) * (shape[1] // tile[1])
    mn = torch.full((nt,), float("inf"), dtype=torch.bfloat16, device=dev); mx = -mn
    scale, offset = torch.empty(nt, device=dev), torch.empty(nt, device=dev)
    settled = torch.zeros(1, dtype=torch.int32, device=dev)
    for _ in range(3):
        ops.calibrate_quantize_(mn, mx, x, tile, 8, sym, True, scale, offset, None, settled, rowsum=True)
    torch.cuda.synchronize()

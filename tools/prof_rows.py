"""One launch of the fused calibration rows kernel on a Llama-3-8B gate_proj weight for ncu."""
import sys

import torch

sys.path.insert(0, ".")
from fastforward_b200 import ops  # noqa: E402

dev = torch.device("cuda")
w = (torch.randn(14336, 4096, device=dev) * 0.02).bfloat16()
nt = 14336
mn = torch.full((nt,), float("inf"), dtype=torch.bfloat16, device=dev); mx = -mn
scale, offset = torch.empty(nt, device=dev), torch.empty(nt, device=dev)
settled = torch.zeros(1, dtype=torch.int32, device=dev)
for _ in range(3):
    ops.calibrate_quantize_(mn, mx, w, (1, 4096), 8, True, True, scale, offset, None, settled, rowsum=True)
torch.cuda.synchronize()

"""One launch of the fused calibrate + fake-quantize kernel on a Llama-3-8B gate_proj weight (g=128) for ncu."""
import sys

import torch

sys.path.insert(0, ".")
from fastforward_b200 import ops  # noqa: E402

dev = torch.device("cuda")
w = (torch.randn(14336, 4096, device=dev) * 0.02).bfloat16()
nt = w.numel() // 128
scale, offset = torch.empty(nt, device=dev), torch.empty(nt, device=dev)
for _ in range(3):
    ops.calibrate_fake_quantize_(w, (1, 128), 4, True, True, scale, offset, None, out=w)
torch.cuda.synchronize()

"""W4 g=128 weight fake-quant (ffq_calibrate_fakequant, per-group tiles) and int8-code calibration of per-group weights:
the one-thread-per-tile kernel against the sub-warp kernel (FFQ_CALQ_GROUP_SUBWARP=1 in the environment selects the
latter for the whole process).  Buffers are cycled so the weights stream from HBM."""
import os
import sys

import torch

sys.path.insert(0, ".."); sys.path.insert(0, ".")
from fastforward_b200 import ops  # noqa: E402

dev = torch.device("cuda")
from bench_calq import time_graph  # noqa: E402


def run(shape, g, dtype, nbuf=4, bits=4, symmetric=True):
    ws = [(torch.randn(shape, device=dev) * 0.02).to(dtype) for _ in range(nbuf)]
    tile = (1, g)
    nt = ws[0].numel() // g
    scale, offset = torch.empty(nt, device=dev), torch.empty(nt, device=dev)

    def fq(i):
        ops.calibrate_fake_quantize_(ws[i % nbuf], tile, bits, symmetric, True, scale, offset, None, out=ws[i % nbuf])
    t = time_graph(fq, max(8, nbuf))
    by = 2 * ws[0].numel() * ws[0].element_size()
    mn = torch.full((nt,), float("inf"), dtype=dtype, device=dev); mx = -mn
    keep = []

    def codes(i):
        keep.append(ops.calibrate_quantize_(mn, mx, ws[i % nbuf], tile, bits, symmetric, True, scale, offset))
    t2 = time_graph(codes, max(8, nbuf))
    keep.clear()
    by2 = ws[0].numel() * (ws[0].element_size() + 1)
    kind = "sub-warp" if os.environ.get("FFQ_CALQ_GROUP_SUBWARP") else "thread-per-tile"
    print(f"{kind:>16} {str(shape):>16} g={g} {str(dtype)[6:]:>8} sym={symmetric!s:5} fake-quant {t * 1e6:7.1f} us ({by / t / 1e9:6.0f} GB/s)   "
          f"int8 codes {t2 * 1e6:7.1f} us ({by2 / t2 / 1e9:6.0f} GB/s)", flush=True)


if __name__ == "__main__":
    run((14336, 4096), 128, torch.bfloat16)
    run((4096, 4096), 128, torch.bfloat16)
    run((14336, 4096), 64, torch.bfloat16)
    run((14336, 4096), 128, torch.bfloat16, symmetric=False)
    run((28672, 8192), 128, torch.bfloat16, nbuf=2)
    run((14336, 4096), 128, torch.float32)

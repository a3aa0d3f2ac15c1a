"""W4A16 linear: fused in-loop dequantisation kernel vs (our dequantize kernel + library bf16 GEMM) across M."""
import sys

import torch

sys.path.insert(0, ".."); sys.path.insert(0, ".")
from fastforward_b200 import _cabi as C, ops  # noqa: E402

dev = torch.device("cuda")
N, K = 14336, 4096
st = C.current_stream(dev)


def t_cuda(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


qw = torch.randint(-8, 8, (N, K), dtype=torch.int8, device=dev)
sw = torch.rand(N * (K // 128), device=dev) * 0.01 + 1e-3
ow = torch.randint(-3, 4, (N * (K // 128),), device=dev).float()
for M in (16, 64, 128, 256, 512, 1024, 2048, 4096):
    x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    y = torch.empty(M, N, dtype=torch.bfloat16, device=dev)

    def fused():
        C.check(C.lib.ffq_qlinear_w4a16(x.data_ptr(), 2, qw.data_ptr(), y.data_ptr(), M, N, K, sw.data_ptr(), ow.data_ptr(), 128, None, 255, st))

    def fallback():
        wd = ops.dequantize_by_tile(qw, sw, (1, 128), ow, torch.bfloat16)
        torch.nn.functional.linear(x, wd)
    print(f"M={M:5d}  fused {t_cuda(fused):8.1f} us   dequant+cuBLAS {t_cuda(fallback):8.1f} us", flush=True)

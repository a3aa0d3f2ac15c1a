"""Launch the W4A16 linear a few times (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastforward_b200 import _cabi as C

M, N, K = int(sys.argv[1]) if len(sys.argv) > 1 else 2048, 14336, 4096
dev = torch.device("cuda", 0)
x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
qw = torch.randint(-8, 8, (N, K), dtype=torch.int8, device=dev)
sw = torch.rand(N * (K // 128), device=dev) * 0.01 + 1e-3
ow = torch.randint(-3, 4, (N * (K // 128),), device=dev).float()
y = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
st = C.current_stream(dev)
for _ in range(3):
    C.check(C.lib.ffq_qlinear_w4a16(x.data_ptr(), 2, qw.data_ptr(), y.data_ptr(), M, N, K, sw.data_ptr(), ow.data_ptr(), 128, None, 255, st))
torch.cuda.synchronize()

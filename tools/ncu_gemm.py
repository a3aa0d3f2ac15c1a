#!/usr/bin/env python
"""One launch of our W8A8 GEMM and one of the library int8 GEMM at BASELINE configs[3], for `ncu` (run under the profiler:
tools/gpu_round2_b.sh).  Nothing here is a timing."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from fastforward_b200 import _cabi as C  # noqa: E402

M, N, K = (int(v) for v in (sys.argv[1:4] if len(sys.argv) >= 4 else (8192, 14336, 4096)))
dev = torch.device("cuda")
qx = torch.randint(-128, 128, (M, K), dtype=torch.int8, device=dev)
qw = torch.randint(-128, 128, (N, K), dtype=torch.int8, device=dev)
y = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
sx = torch.tensor([0.01], device=dev); ox = torch.tensor([3.0], device=dev); sw = torch.rand(N, device=dev) * 0.01
rs = torch.empty(N, dtype=torch.int32, device=dev)
st = C.current_stream(dev)
C.check(C.lib.ffq_rowsum_i8(qw.data_ptr(), rs.data_ptr(), N, K, st))
for _ in range(2):
    C.check(C.lib.ffq_qlinear_w8a8(qx.data_ptr(), qw.data_ptr(), y.data_ptr(), 2, M, N, K, sx.data_ptr(), ox.data_ptr(),
                                   sw.data_ptr(), None, rs.data_ptr(), None, None, 255, None, st))
    torch._int_mm(qx, qw.t())
torch.cuda.synchronize()

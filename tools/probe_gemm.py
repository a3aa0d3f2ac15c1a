import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastforward_b200 import _cabi as C

def run(M, N, K, iters=20):
    dev = "cuda"
    qx = torch.randint(-128, 128, (M, K), dtype=torch.int8, device=dev)
    qw = torch.randint(-128, 128, (N, K), dtype=torch.int8, device=dev)
    y = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    sx = torch.tensor([0.01], device=dev); ox = torch.tensor([3.0], device=dev)
    sw = torch.rand(N, device=dev) * 0.01
    rs = torch.empty(N, dtype=torch.int32, device=dev)
    ws = torch.empty(4 * N, device=dev)
    st = C.current_stream(qx.device)
    C.check(C.lib.ffq_rowsum_i8(qw.data_ptr(), rs.data_ptr(), N, K, st))
    def f():
        C.check(C.lib.ffq_qlinear_w8a8(qx.data_ptr(), qw.data_ptr(), y.data_ptr(), 2, M, N, K, sx.data_ptr(), ox.data_ptr(),
                                       sw.data_ptr(), None, rs.data_ptr(), None, None, 255, ws.data_ptr(), ws.numel() * 4, st))
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): f()
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / iters * 1e-3
    ref = torch._int_mm(qx, qw.t())
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters): torch._int_mm(qx, qw.t())
    e1.record(); torch.cuda.synchronize()
    t2 = e0.elapsed_time(e1) / iters * 1e-3
    exp = (ref.float() + 3.0 * rs.float()[None, :]) * (0.01 * sw)[None, :]
    err = (y.float() - exp).abs().max().item() / exp.abs().max().item()
    print(f"M{M} N{N} K{K}: ours {t*1e6:.1f} us {2*M*N*K/t/1e12:.0f} TOPS | torch._int_mm {t2*1e6:.1f} us {2*M*N*K/t2/1e12:.0f} TOPS | relerr {err:.2e}", flush=True)

for shp in [(8192, 14336, 4096), (2048, 4096, 4096), (2048, 14336, 4096), (2048, 4096, 14336), (2048, 1024, 4096), (8192, 8192, 8192)]:
    run(*shp)

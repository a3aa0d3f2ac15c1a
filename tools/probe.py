"""Quick device-time probe of every hot-path kernel (CUDA events, inputs larger than L2)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastforward_b200 import ops


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    return ts[len(ts) // 2] * 1e-3


def main():
    dev = "cuda"
    res = {}
    for name, shape, tile, dtype, bits in [
        ("cfg1_f32_perchannel", (4096, 4096), (1, 4096), torch.float32, 8),
        ("big_f32_perchannel", (16384, 4096), (1, 4096), torch.float32, 8),
        ("bf16_perchannel", (14336, 4096), (1, 4096), torch.bfloat16, 8),
        ("bf16_g128", (14336, 4096), (1, 128), torch.bfloat16, 4),
        ("f32_g128", (14336, 4096), (1, 128), torch.float32, 4),
        ("bf16_pertensor", (8192, 4096), (8192, 4096), torch.bfloat16, 8),
        ("f32_pertensor", (8192, 4096), (8192, 4096), torch.float32, 8),
    ]:
        torch.manual_seed(0)
        x = torch.randn(shape, device=dev, dtype=dtype)
        g = torch.randn(shape, device=dev, dtype=dtype)
        n = x.numel()
        s = x.element_size()
        mn, mx = ops.tile_minmax(x, tile)
        nt = mn.numel()
        scale = torch.empty(nt, device=dev)
        offset = torch.empty(nt, device=dev)
        ops.parameters_for_range_(mn, mx, bits, False, True, scale, offset)
        r = {}
        t = timeit(lambda: ops.fake_quantize_by_tile(x, scale, tile, float(bits), None, offset)); r["fakequant_fwd"] = (2 * s * n / t / 1e9, t * 1e6)
        t = timeit(lambda: ops.quantize_by_tile_backward(x, g, scale, tile, float(bits), offset)); r["bwd"] = (3 * s * n / t / 1e9, t * 1e6)
        t = timeit(lambda: ops.quantize_by_tile(x, scale, tile, float(bits), torch.int8, offset)); r["quant_i8"] = ((s + 1) * n / t / 1e9, t * 1e6)
        q = ops.quantize_by_tile(x, scale, tile, float(bits), torch.int8, offset)
        t = timeit(lambda: ops.dequantize_by_tile(q, scale, tile, offset, dtype)); r["dequant_i8"] = ((s + 1) * n / t / 1e9, t * 1e6)
        t = timeit(lambda: ops.tile_minmax(x, tile)); r["minmax"] = (s * n / t / 1e9, t * 1e6)
        y = torch.empty_like(x)
        t = timeit(lambda: y.copy_(x)); r["torch_copy"] = (2 * s * n / t / 1e9, t * 1e6)
        res[name] = {k: f"{v[0]:.0f} GB/s ({v[1]:.1f} us)" for k, v in r.items()}
        print(name, json.dumps(res[name]), flush=True)


if __name__ == "__main__":
    main()

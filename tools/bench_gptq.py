"""GPTQ of one Llama-3-8B-shape linear on the GPU: whole-layer time and the share of the block kernel."""
import sys
import time

import torch

sys.path.insert(0, ".")
import fastforward_b200 as ff  # noqa: E402
from fastforward_b200 import _cabi  # noqa: E402
from fastforward_b200.quantization import gptq as G  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(0)
for rows, cols, gran, name in [(4096, 4096, ff.PerChannel(0), "q_proj W4 per-channel"),
                               (4096, 4096, ff.PerBlock(block_dims=1, block_sizes=128, per_channel_dims=0), "q_proj W4 g=128"),
                               (14336, 4096, ff.PerBlock(block_dims=1, block_sizes=128, per_channel_dims=0), "gate_proj W4 g=128")]:
    layer = ff.nn.QuantizedLinear(cols, rows, bias=False)
    with torch.no_grad():
        layer.weight.copy_(torch.randn(rows, cols) * 0.02)
    layer.weight_quantizer = ff.nn.LinearQuantizer(4, granularity=gran)
    layer.to(dev)
    data = [((torch.randn(1, 2048, cols, device=dev),), {}) for _ in range(2)]
    w0 = layer.weight.detach().clone()
    for it in range(2):
        with torch.no_grad():
            layer.weight.copy_(w0)
        layer.weight_quantizer.reset_parameters()
        torch.cuda.synchronize(); l0 = _cabi.launch_count(); t0 = time.perf_counter()
        with torch.no_grad():
            G.gptq(layer, data)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    # block kernel alone
    hinv = torch.eye(128, device=dev) + torch.triu(torch.randn(128, 128, device=dev) * 0.01, 1)
    blk = w0[:, :128].float().contiguous(); q = torch.empty_like(blk); e = torch.empty_like(blk)
    sc = layer.weight_quantizer.scale.data; off = layer.weight_quantizer.offset.data
    tile = gran.tile_size(w0.shape)
    oc = torch.arange(128, dtype=torch.int32, device=dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    G.gptq_block_(blk, q, e, hinv, sc, off, oc, tile[0], tile[1], cols // tile[1], 4)
    ev0.record()
    for _ in range(20):
        G.gptq_block_(blk, q, e, hinv, sc, off, oc, tile[0], tile[1], cols // tile[1], 4)
    ev1.record(); torch.cuda.synchronize()
    print(f"{name:28s} [{rows}x{cols}] gptq() {dt * 1e3:8.1f} ms, {_cabi.launch_count() - l0 - 21} launches of this library; "
          f"block kernel {ev0.elapsed_time(ev1) / 20 * 1e3:6.1f} us per 128-column block ({cols // 128} blocks)", flush=True)

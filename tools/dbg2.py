import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, 'tests')
import torch
import fastforward_b200 as ff
from fastforward_b200.nn import qlinear
from oracle import ref_ops as R
m,k,n,dt = 2048,4096,1024,torch.bfloat16
g = torch.Generator().manual_seed(0)
x = torch.randn(m, k, generator=g).to(dt); w = (torch.randn(n, k, generator=g) * 0.05).to(dt)
lin = torch.nn.Linear(k, n, bias=False, dtype=dt)
with torch.no_grad(): lin.weight.copy_(w)
ff.quantize_model(lin)
lin.input_quantizer = ff.nn.LinearQuantizer(8, symmetric=False, quantized_dtype=torch.int8)
lin.weight_quantizer = ff.nn.LinearQuantizer(8, symmetric=True, granularity=ff.PerChannel(0), quantized_dtype=torch.int8)
lin.to("cuda"); xc = x.cuda()
lin.input_quantizer.quantization_range = (xc.min(), xc.max())
lin.weight_quantizer.quantization_range = (lin.weight.min(1).values, lin.weight.max(1).values)
with torch.no_grad():
    qlinear.install(); y = lin(xc); qlinear.uninstall(); yf = lin(xc)
    xq, wq = lin.input_quantizer(xc), lin.weight_quantizer(lin.weight)
px, pw = xq.quant_args(), wq.quant_args()
y64 = R.exact_linear_f64(xq.raw_data.cpu(), px.scale.detach().cpu(), px.offset.detach().cpu(), (m, k), wq.raw_data.cpu(), pw.scale.detach().cpu(), None if pw.offset is None else pw.offset.detach().cpu(), (1, k), None)
ek = (y.double().cpu()-y64).abs(); ef = (yf.double().cpu()-y64).abs()
print("max|y|", y64.abs().max().item(), "err_k max", ek.max().item(), "err_f max", ef.max().item())
i = ek.argmax(); r, c = divmod(i.item(), n)
print("worst at", r, c, "y64", y64[r,c].item(), "yk", y[r,c].item(), "yf", yf[r,c].item())
bad = (ek > 0.02*y64.abs().clamp_min(0.1)).nonzero()
print("n bad", bad.shape[0], bad[:10].tolist())
if bad.shape[0]:
    rows = bad[:,0].unique(); cols = bad[:,1].unique()
    print("rows", rows[:20].tolist(), len(rows), "cols", cols[:20].tolist(), len(cols))

import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, 'tests')
import torch
from conftest import load_golden
from fastforward_b200 import ops
c = load_golden("static")[38]
x, s, o = c["x"].cuda(), c["scale"].cuda(), c["offset"].cuda()
yf, codes = ops.fake_quantize_by_tile(x, s, c["tile"], float(c["num_bits"]), c["qdtype"], o, c["ddtype"], return_codes=True)
y = c["y"]; q = c["q"]
bad = (yf.cpu() != y) | (codes.cpu() != q)
idx = bad.nonzero()
print("nbad", idx.shape[0])
for i in idx[:10]:
    i = tuple(i.tolist())
    print(i, "x", c["x"][i].item(), "s", c["scale"][i[0]].item(), "o", c["offset"][i[0]].item(), "ours", yf[i].item(), codes[i].item(), "ref", y[i].item(), q[i].item())
yf2 = ops.fake_quantize_by_tile(x, s, c["tile"], float(c["num_bits"]), c["qdtype"], o, c["ddtype"])
print("without codes equal:", torch.equal(yf2.cpu(), y), "bits", torch.equal(yf2.cpu().view(torch.int32), y.view(torch.int32)))
print("with codes bits:", torch.equal(yf.cpu().view(torch.int32), y.view(torch.int32)), torch.equal(codes.cpu(), q))

import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench_workloads as bw, fastforward_b200 as ff
from fastforward_b200.nn import qlinear
dev = "cuda"
m = bw.DecoderStack(bw.TINY, dtype=torch.bfloat16, device=dev)
bw.init_weights_(m)
bw.quantize_for_w8a8(ff, m); m.to(dev); qlinear.install()
def hook(name):
    def f(mod, inp, out):
        print(name, type(mod).__name__, [getattr(i,'dtype',None) for i in inp], getattr(out,'dtype',None), type(out).__name__)
    return f
for n, mod in m.named_modules():
    if n and n.count('.') <= 3 and n.startswith(("embed", "layers.0")): mod.register_forward_hook(hook(n))
with torch.no_grad(), ff.estimate_ranges(m, ff.range_setting.running_minmax):
    y = m(torch.randint(0, 1024, (1, 16), device=dev))
print(y.dtype)

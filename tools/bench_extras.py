"""Run only bench.py's `extras` micro-benchmarks (per-kernel numbers) on one GPU."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import fastforward_b200 as ff

peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
hbm = float(peaks.get("hbm_gbs", 6453.7))
out = bench.measure_extras(ff, dev, hbm, 3321.8, bench.measure_int8_library_peak(dev))
for k, v in out.items():
    print(k, json.dumps(v))

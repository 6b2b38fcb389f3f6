"""Fused cell in its bandwidth-bound regime (16 384 and 4 096 rows x 4096, bf16), exact and faithful modes."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
peak = bench.measured_peak_gbs()[0]
for rows, ln_scale in ((16384, 1.0), (4096, 1.0), (16384, 4.0)):   # ln_scale 4: ~5 % of the GELU inputs have |y| >= 8
    r = bench.bench_cell_hbm_regime(dev, peak, rows=rows, ln_scale=ln_scale)
    print(json.dumps({"rows": rows, "ln_scale": ln_scale, "exact": r["exact"], "faithful": r["faithful"]}))

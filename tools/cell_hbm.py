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
for rows in (16384, 4096):
    r = bench.bench_cell_hbm_regime(dev, peak, rows=rows)
    print(json.dumps({"rows": rows,
                      "exact": r["exact"], "faithful": r["faithful"]}))

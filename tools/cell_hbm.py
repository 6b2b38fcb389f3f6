"""Fused cell in its bandwidth-bound regime (16 384 rows x 4096, bf16), exact and faithful modes.
AB200_CELL_GELU_TABLE=0 selects the polynomial + guard GELU for the exact mode (A/B against the table)."""
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
    print(json.dumps({"table": os.environ.get("AB200_CELL_GELU_TABLE", "1"), "variant": os.environ.get("AB200_CELL_VARIANT", "0"), "rows": rows,
                      "exact": r["exact"], "faithful": r["faithful"]}))

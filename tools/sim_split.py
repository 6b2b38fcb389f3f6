import sys, json
sys.path.insert(0, '/root/repo')
import torch, bench
dev = torch.device("cuda:0")
for world in (4, 8):
    r3 = bench.bench_cfg3_70b(dev, world, 0, 10, 3, 6557.1)
    r4 = bench.bench_cfg4_gptoss(dev, world, 0, 10, 3, 6557.1)
    print(world, {k: (v["ms_per_proposal"], v["head"], v["bound"], round(v["frac_of_roofline"], 3)) for k, v in r3.items() if k.startswith("D")},
          r4["per_gpu_batch"], r4["speculator"]["ms_per_proposal"], r4["speculator"]["head"], r4["kv_write"]["ms"], flush=True)

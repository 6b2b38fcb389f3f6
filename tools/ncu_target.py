"""Small ncu target: a few launches of each hot kernel at the BASELINE sizes (cfg2 step + cfg3 KV)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from arcticinference_b200 import py_custom_ops
from arcticinference_b200.speculator import ArcticLSTMSpeculator
import bench

dev = torch.device("cuda:0")
V, H, D, S, B = (bench.CFG[k] for k in "VHDSB")
model = ArcticLSTMSpeculator(vocab_size=V, input_hidden_dim=H, inner_dim=D, n_predict=S, max_num_seqs=B,
                             device=dev, enforce_eager=True)
bench.synthetic_weights(model, 0)
g = torch.Generator(device="cpu").manual_seed(1)
ids = torch.randint(0, V, (B,), generator=g).to(dev)
hid = torch.randn(B, H, generator=g).to(torch.bfloat16).to(dev)
for _ in range(3):
    out = model.generate_proposals(ids, hid, S)       # eager: rms, 3x(gate, cell, head), unpack
torch.cuda.synchronize()
print("tokens", out[0].tolist())
kv = bench.bench_kv_write(dev, iters=3, peak_gbs=6557.1)
print("kv GB/s", kv["value"])
# round-1 additions: the fused cell in its HBM-bound regime, the FP8 KV write, the MLP speculator step
cell = bench.bench_cell_hbm_regime(dev, 6557.1, iters=1)
print("cell 16384 rows: exact %.0f GB/s, faithful %.0f GB/s" % (cell["exact"]["gbs"], cell["faithful"]["gbs"]))
from arcticinference_b200.mlp_speculator import ArcticMLPSpeculator
mlp = ArcticMLPSpeculator(vocab_size=V, emb_dim=H, inner_dim=D, n_predict=S, max_num_seqs=B, device=dev, enforce_eager=True)
bench.synthetic_weights(mlp, 1)
print("mlp tokens", mlp.generate_proposals(ids, hid, S)[0].tolist())
torch.cuda.synchronize()

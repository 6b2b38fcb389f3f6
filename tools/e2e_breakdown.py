"""Where the end-to-end gap of ArcticProposer.propose() goes (B = 64, 8B shape): wall-clock of sub-sequences of the call,
each ending in a stream synchronize, median of 200."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch, bench
from arcticinference_b200.speculator import ArcticLSTMSpeculator, ArcticProposer
dev = torch.device("cuda:0")
V, H, D, S = (bench.CFG[k] for k in "VHDS")
B = 64
model = ArcticLSTMSpeculator(vocab_size=V, input_hidden_dim=H, inner_dim=D, n_predict=S, max_num_seqs=B, device=dev)
bench.synthetic_weights(model, 0)
prop = ArcticProposer(model)
g = torch.Generator(device="cpu").manual_seed(1)
ids = torch.randint(0, V, (B,), generator=g).numpy()
hid = torch.randn(B, H, generator=g).to(torch.bfloat16).pin_memory()
for _ in range(20):
    prop.propose(ids, hid, S)
hid_dev, ids_dev, ids_pin, ids_np, graph, out_dev, out, out_np = prop._plans[(B, S)][:8]
stream = torch.cuda.current_stream(dev)

def wall(fn, n=200):
    ts = []
    for _ in range(n):
        stream.synchronize()
        t0 = time.perf_counter()
        fn()
        stream.synchronize()
        ts.append((time.perf_counter() - t0) * 1e6)
    return float(np.median(ts))

def f_sync(): pass
def f_hid(): hid_dev.copy_(hid, non_blocking=True)
def f_ids(): ids_np[:] = ids; ids_dev.copy_(ids_pin, non_blocking=True)
def f_graph(): graph.replay()
def f_d2h(): out.copy_(out_dev, non_blocking=True)
def f_h2d_both(): f_hid(); f_ids()
def f_all(): f_hid(); f_ids(); f_graph(); f_d2h()
for name, fn in (("empty sync", f_sync), ("H2D hidden 512 KB", f_hid), ("H2D ids 512 B", f_ids), ("both H2D", f_h2d_both),
                 ("graph replay only", f_graph), ("D2H tokens 1.5 KB", f_d2h), ("all four", f_all),
                 ("propose()", lambda: prop.propose(ids, hid, S))):
    print(f"{name:22s} {wall(fn):8.1f} us wall (enqueue -> synchronize returns)")

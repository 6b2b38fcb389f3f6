"""Dev/profiling: device-side timeline (globaltimer) of one graph-replayed proposal at cfg2."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench
from arcticinference_b200 import _capi
from arcticinference_b200.speculator import ArcticLSTMSpeculator
dev = torch.device("cuda:0"); lib = _capi.lib()
V, H, D, S = (bench.CFG[k] for k in "VHDS")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
model = ArcticLSTMSpeculator(vocab_size=V, input_hidden_dim=H, inner_dim=D, n_predict=S, max_num_seqs=B, device=dev)
bench.synthetic_weights(model, 0)
g = torch.Generator(device="cpu").manual_seed(1)
ids = torch.randint(0, V, (B,), generator=g).to(dev)
hid = torch.randn(B, H, generator=g).to(torch.bfloat16).to(dev)
N = 16
trace = torch.empty(N, 4, dtype=torch.int64, device=dev)
def reset():
    trace[:, :2] = torch.iinfo(torch.int64).max; trace[:, 2:] = 0   # signed view of the u64 slots (ns << 2^63)
reset()
# warm-up must not consume trace slots: enable tracing only around the capture
model._get_workspace(B, S)
out = torch.empty(B, S, dtype=torch.long, device=dev)
model._propose_into(ids, hid, out, B, S); torch.cuda.synchronize()
lib.ab200_set_trace_buffer(trace.data_ptr(), N)
gph = torch.cuda.CUDAGraph()
with torch.cuda.graph(gph):
    model._propose_into(ids, hid, out, B, S)
lib.ab200_set_trace_buffer(None, 0)
names = ["rms"] + [f"{k}{s}" for s in range(S) for k in ("gate", "cell", "head")]
import numpy as np
runs = []
for rep in range(7):
    for _ in range(20): gph.replay()
    torch.cuda.synchronize(); reset(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gph.replay(); e1.record(); torch.cuda.synchronize()
    t = trace.cpu().numpy().astype("float64")
    runs.append((t[:len(names), 2] - t[:len(names), 1]) / 1e3)
med = np.median(np.stack(runs), 0)
print("median run(dep->exit) over 7 replays:", " ".join(f"{n}={v:.1f}" for n, v in zip(names, med)))
t0 = t[0, 0]
print(f"graph replay {e0.elapsed_time(e1)*1e3:.1f} us (events); timeline in us relative to first CTA entry")
print(f"{'kernel':8s} {'entry':>8s} {'dep-ok':>8s} {'exit':>8s} | {'run(dep->exit)':>14s} {'early-start':>11s}")
prev_exit = t0
for i, n in enumerate(names):
    en, dp, ex = (t[i, 0] - t0) / 1e3, (t[i, 1] - t0) / 1e3, (t[i, 2] - t0) / 1e3
    print(f"{n:8s} {en:8.1f} {dp:8.1f} {ex:8.1f} | {ex - dp:14.1f} {dp - en:11.1f}")
print(f"total {(t[len(names)-1, 2] - t0)/1e3:.1f} us")

"""Dev/profiling: device-side timeline (globaltimer) of one graph-replayed proposal at cfg2 (bench.device_timeline)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench
from arcticinference_b200.speculator import ArcticLSTMSpeculator
dev = torch.device("cuda:0")
V, H, D, S = (bench.CFG[k] for k in "VHDS")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
model = ArcticLSTMSpeculator(vocab_size=V, input_hidden_dim=H, inner_dim=D, n_predict=S, max_num_seqs=B, device=dev)
bench.synthetic_weights(model, 0)
g = torch.Generator(device="cpu").manual_seed(1)
ids = torch.randint(0, V, (B,), generator=g).to(dev)
hid = torch.randn(B, H, generator=g).to(torch.bfloat16).to(dev)
tl = bench.device_timeline(model, ids, hid, B, S)
print("median run(dep->exit) over 7 replays:", " ".join(f"{n}={v:.1f}" for n, v in zip(tl["names"], tl["run_us"])))
print(f"graph replay {tl['replay_us']:.1f} us (events); timeline in us relative to first CTA entry")
print(f"{'kernel':8s} {'entry':>8s} {'dep-ok':>8s} {'exit':>8s} | {'run(dep->exit)':>14s} {'early-start':>11s} | "
      f"{'1st data':>8s} {'last data':>9s} {'1st exit':>8s}  (GEMMs: operand bytes landed in smem)")
f = lambda v: f"{v:8.1f}" if v is not None else "       -"
for n, (en, dp, ex, fd, ld, fx) in zip(tl["names"], tl["last"]):
    print(f"{n:8s} {en:8.1f} {dp:8.1f} {ex:8.1f} | {ex - dp:14.1f} {dp - en:11.1f} | {f(fd)} {f(ld)}  {f(fx)}")
print(f"total {tl['last'][-1][2]:.1f} us")

"""Dev (GPU box): graph-replay time of the cfg2 draft loop for settings of the debug library's launch tunables
(ab200_debug_set_tunable).  The speculator class is pointed at the DEBUG flavour for this run only.
    python tools/graph_tune.py l2_prefetch_kb=0,2,4,8 [B=64] [cfg=cfg2|cfg3|cfg3w|cfg4]"""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from arcticinference_b200 import _capi
_capi._lib = _capi.debug_lib()                       # tools only: the product never does this
import bench
from arcticinference_b200.speculator import ArcticLSTMSpeculator

dev = torch.device("cuda:0")
lib = _capi.debug_lib()
sweeps, B, cfg = [], 64, "cfg2"
for a in sys.argv[1:]:
    k, v = a.split("=")
    if k == "B":
        B = int(v)
    elif k == "cfg":
        cfg = v
    else:
        sweeps.append((k, [int(x) for x in v.split(",")]))
V, H, D, S = {"cfg2": tuple(bench.CFG[k] for k in "VHDS"), "cfg3": (128256, 8192, 4096, 4),
              "cfg3w": (128256, 8192, 8192, 4), "cfg4": (201088, 2880, 2880, 3)}[cfg]
model = ArcticLSTMSpeculator(vocab_size=V, input_hidden_dim=H, inner_dim=D, n_predict=S, max_num_seqs=B, device=dev)
bench.synthetic_weights(model, 0)
g = torch.Generator(device="cpu").manual_seed(1)
ids = torch.randint(0, V, (B,), generator=g).to(dev)
hid = torch.randn(B, H, generator=g).to(torch.bfloat16).to(dev)


def replay_us():
    model._invalidate_graphs()
    model.generate_proposals(ids, hid, S)                      # capture
    gph = model.cuda_graphs[(model.cuda_graph_max_batch_size if False else __import__("arcticinference_b200.speculator", fromlist=["padding_size"]).padding_size(B), S)]
    for _ in range(50):
        gph.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(100):
            gph.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 10)
    return statistics.median(ts), min(ts)


def per_kernel():
    """median dep->exit per kernel kind from the device-side trace (resolution ~0.1 us, unlike the replay time)"""
    tl = bench.device_timeline(model, ids, hid, B, S, replays=9)
    out = {}
    for n, v in zip(tl["names"], tl["run_us"]):
        out.setdefault(n.rstrip("0123456789"), []).append(v)
    return " ".join(f"{k}={statistics.median(v):.2f}" for k, v in out.items()) + f" | traced replay {tl['replay_us']:.1f}"


base = replay_us()
print(f"B={B} defaults: {base[0]:.1f} us (min {base[1]:.1f}) | {per_kernel()}", flush=True)
for name, values in sweeps:
    for v in values:
        assert lib.ab200_debug_set_tunable(name.encode(), v) == 0, name
        med, mn = replay_us()
        print(f"  {name}={v}: {med:.1f} us (min {mn:.1f}) | {per_kernel()}", flush=True)
    # back to the compiled-in default of this tunable (first value listed is expected to be the default)
    lib.ab200_debug_set_tunable(name.encode(), values[0])
again = replay_us()
print(f"defaults again: {again[0]:.1f} us (min {again[1]:.1f})")

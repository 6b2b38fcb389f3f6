"""Dev (GPU box): the same draft loop on the RELEASE and the DEBUG flavour of the library on ONE box (the debug flavour
also with the CTA-pair kernel's k-blocks per request set back to 1) - separates box-to-box spread from real changes.
    python tools/flavour_probe.py [cfg2|cfg3|cfg3n2|cfg4|cfg4n4|cfg2b256]"""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from arcticinference_b200 import _capi
import bench

which = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
V, H, D, S, B = {"cfg3": (128256, 8192, 4096, 4, 256), "cfg4": (201088, 2880, 2880, 3, 512),
                 "cfg2b256": (128256, 4096, 4096, 3, 256), "cfg2": (128256, 4096, 4096, 3, 64),
                 "cfg3n2": (128256, 8192, 4096, 4, 128), "cfg4n4": (201088, 2880, 2880, 3, 128)}[which]
dev = torch.device("cuda:0")
rel, dbg = _capi.lib(), _capi.debug_lib()
OLD = os.path.join(ROOT, "arcticinference_b200", "libarctic_b200_old.so")     # optional: a previous build, side by side
old = _capi._load(OLD, _capi.SIGNATURES) if os.path.exists(OLD) else None


def run(lib, label, **tun):
    _capi._lib = lib
    for k, v in tun.items():
        assert dbg.ab200_debug_set_tunable(k.encode(), v) == 0
    from arcticinference_b200.speculator import ArcticLSTMSpeculator
    m = ArcticLSTMSpeculator(vocab_size=V, input_hidden_dim=H, inner_dim=D, n_predict=S, max_num_seqs=B, device=dev)
    bench.synthetic_weights(m, 0)
    g = torch.Generator(device="cpu").manual_seed(1)
    ids = torch.randint(0, V, (B,), generator=g).to(dev)
    hid = torch.randn(B, H, generator=g).to(torch.bfloat16).to(dev)
    out = m.generate_proposals(ids, hid, S)
    for _ in range(10):
        m.generate_proposals(ids, hid, S)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            m.generate_proposals(ids, hid, S)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 20 * 1e3)
    print(f"{which} {label}: {statistics.median(ts):.1f} us per proposal (min {min(ts):.1f})", flush=True)
    del m
    torch.cuda.empty_cache()
    return out


if old is not None:
    run(old, "PREVIOUS release build (libarctic_b200_old.so)")
NEW1 = os.path.join(ROOT, "arcticinference_b200", "libarctic_b200_new1.so")   # optional: current code, pair kbs = 1 compiled in
new1 = _capi._load(NEW1, _capi.SIGNATURES) if os.path.exists(NEW1) else None
a = run(rel, "release (compiled-in defaults)")
if new1 is not None:
    run(new1, "release, pair kbs = 1 compiled in")
b = run(dbg, "debug, defaults")
c = run(dbg, "debug, pair kbs = 1", pair_head_kbs=1, pair_gate_kbs=1)
d = run(dbg, "debug, pair kbs = 2", pair_head_kbs=2, pair_gate_kbs=2)
if new1 is not None:
    run(new1, "release, pair kbs = 1 compiled in, again")
e = run(rel, "release again")
if old is not None:
    run(old, "PREVIOUS release build again")
print("tokens equal across all:", all(torch.equal(a, x) for x in (b, c, d, e)))
_capi._lib = rel

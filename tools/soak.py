"""Soak test (GPU box): replay the captured draft loop for `seconds` per configuration with ALTERNATING inputs and check
every result against the eager result of the same input - a race in the PDL-chained graph (weights prefetched before the
dependency wait, the cell's side inputs fetched early, packed arg-max slots reset by the cell) would show up as a
token mismatch sooner or later.      python tools/soak.py [seconds=20]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from arcticinference_b200.speculator import ArcticLSTMSpeculator, ArcticProposer

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0
dev = torch.device("cuda:0")
V, H, D, S = (bench.CFG[k] for k in "VHDS")
total_bad = 0
for B, kw in ((64, {}), (33, {}), (8, {}), (200, {}), (300, {}), (512, {}), (64, {"fp8_head_max_batch": 1 << 30})):
    model = ArcticLSTMSpeculator(vocab_size=V, input_hidden_dim=H, inner_dim=D, n_predict=S, max_num_seqs=B, device=dev, **kw)
    bench.synthetic_weights(model, 0)
    eager = ArcticLSTMSpeculator(vocab_size=V, input_hidden_dim=H, inner_dim=D, n_predict=S, max_num_seqs=B, device=dev,
                                 enforce_eager=True, allocate_parameters=False, **kw)
    eager.share_parameters(model.named_parameters())
    g = torch.Generator(device="cpu").manual_seed(B)
    inputs, expect = [], []
    for i in range(4):
        ids = torch.randint(0, V, (B,), generator=g).to(dev)
        hid = torch.randn(B, H, generator=g).to(torch.bfloat16).to(dev)
        inputs.append((ids, hid))
        expect.append(eager.generate_proposals(ids, hid, S).clone())
    assert not torch.equal(expect[0], expect[1])
    prop = ArcticProposer(model)
    n = bad = 0
    t_end = time.time() + seconds
    while time.time() < t_end:
        outs = []
        for k in range(64):                                # 64 replays enqueued back to back, results checked afterwards
            ids, hid = inputs[(n + k) % 4]
            outs.append(model.generate_proposals(ids, hid, S))
        torch.cuda.synchronize()
        for k, o in enumerate(outs):
            bad += int(not torch.equal(o, expect[(n + k) % 4]))
        n += 64
        # and the host round trip (H2D into the static buffers, replay, D2H)
        ids, hid = inputs[n % 4]
        got = prop.propose(ids.cpu().numpy(), hid, S)
        bad += int(not (torch.from_numpy(got).to(dev) == expect[n % 4]).all())
    print(f"B={B} {kw}: {n} graph replays + {n // 64} propose() calls in {seconds:.0f} s, mismatches: {bad}", flush=True)
    total_bad += bad
    del model, eager, prop
    torch.cuda.empty_cache()
print("SOAK", "OK" if total_bad == 0 else f"FAILED ({total_bad})")
sys.exit(0 if total_bad == 0 else 1)

import sys, statistics
sys.path.insert(0, '/root/repo')
import torch, bench
from arcticinference_b200.speculator import ArcticLSTMSpeculator, padding_size
dev = torch.device("cuda:0")
V, H, D, S, B = (bench.CFG[k] for k in "VHDSB")
m = ArcticLSTMSpeculator(vocab_size=V, input_hidden_dim=H, inner_dim=D, n_predict=S, max_num_seqs=B, device=dev)
bench.synthetic_weights(m, 0)
g0 = torch.Generator(device="cpu").manual_seed(1)
ids = torch.randint(0, V, (B,), generator=g0).to(dev); hid = torch.randn(B, H, generator=g0).to(torch.bfloat16).to(dev)
sb = m.static_cuda_buffers; sb["last_tokens"][:B].copy_(ids); sb["previous_hidden_states"][:B].copy_(hid)
ids_r, hid_r = sb["last_tokens"][:B], sb["previous_hidden_states"][:B]
m.generate_proposals(ids_r, hid_r, S)
gph = m.cuda_graphs[(padding_size(B), S)]
out_view = sb["next_tokens"].view(-1)[:B * S].view(B, S)
ring = [torch.empty(B, S, dtype=torch.long, device=dev) for _ in range(8)]
def timed(fn, n=200):
    for _ in range(20): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / n * 1e3)
    return statistics.median(ts)
i = [0]
def ring_copy():
    gph.replay(); ring[i[0] % 8].copy_(out_view); i[0] += 1
print("replay only            %.1f us" % timed(lambda: gph.replay()))
print("replay + clone         %.1f us" % timed(lambda: (gph.replay(), out_view.clone())))
print("replay + copy_ to ring %.1f us" % timed(ring_copy))
print("generate_proposals     %.1f us" % timed(lambda: m.generate_proposals(ids_r, hid_r, S)))

"""Small run of every kernel for compute-sanitizer (memcheck / racecheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as g
g.smoke()
from oracle import speculator_ref as O, mlp_speculator_ref as M
from arcticinference_b200.speculator import ArcticLSTMSpeculator
from arcticinference_b200.mlp_speculator import ArcticMLPSpeculator
from arcticinference_b200 import py_custom_ops, swiftkv
dev = torch.device("cuda:0")
V, H, D, S = 2048, 256, 256, 3
w = O.make_weights(V, H, D, torch.bfloat16, seed=0, device=dev)
for B, q in ((7, True), (96, False), (200, False), (300, False)):   # FP8 head; 16 KiB slabs with 2 k-blocks per request; CTA pairs; two m-blocks per tile
    m = ArcticLSTMSpeculator(vocab_size=V, input_hidden_dim=H, inner_dim=D, n_predict=S, max_num_seqs=B, device=dev,
                             enforce_eager=True, quantize_lm_head=q)
    m.load_weights(w.items())
    ids, hid = O.make_inputs(B, V, H, torch.bfloat16, seed=1, device=dev)
    print(B, m.generate_proposals(ids, hid, S)[0].tolist())
wm = M.make_mlp_weights(V, H, D, S, False, torch.bfloat16, seed=0, device=dev)
mm = ArcticMLPSpeculator(vocab_size=V, emb_dim=H, inner_dim=D, n_predict=S, tie_weights=False, max_num_seqs=16, device=dev,
                         enforce_eager=True)
mm.load_weights(wm.items())
ids, hid = O.make_inputs(9, V, H, torch.bfloat16, seed=1, device=dev)
print("mlp", mm.generate_proposals(ids, hid, S)[0].tolist())
# fp8 kv + gather
L, T, Hkv, Dh, bs = 3, 50, 2, 32, 16
n = Hkv * Dh; nb = 8
k = torch.randn(T, L * n, device=dev).to(torch.bfloat16); v = torch.randn(T, L * n, device=dev).to(torch.bfloat16)
caches = [torch.zeros(2, nb, bs, Hkv, Dh, dtype=torch.uint8, device=dev) for _ in range(L)]
sc = [torch.full((), 0.1, device=dev) for _ in range(L)]
assert py_custom_ops.try_load_torch_library()
py_custom_ops.reshape_and_cache_flash_bulk(k, v, [c[0] for c in caches], [c[1] for c in caches],
                                           torch.randperm(nb * bs, device=dev)[:T], "fp8", sc, sc, Hkv, Dh)
print("gather", swiftkv.gather_rows([k, v], torch.tensor([3, 1, 49], device=dev))[0].shape)
# fused K-rotary + cache write (NeoX / GPT-J, both rounding contracts, fp8 cache, selected rows)
from oracle import kv_rope_ref as R
for Dh2, rot, neox, rounding, kvd in ((64, 64, True, "fp32", "auto"), (128, 64, False, "storage", "auto"), (32, 32, True, "fp32", "fp8")):
    n2 = Hkv * Dh2
    kvl = [torch.randn(T, 2 * n2, device=dev).to(torch.bfloat16) for _ in range(L)]
    cdt = torch.uint8 if kvd == "fp8" else torch.bfloat16
    cch = [torch.zeros(2, nb, bs, Hkv, Dh2, dtype=cdt, device=dev) for _ in range(L)]
    pos = torch.randint(0, 128, (T,), device=dev)
    cs = R.make_cos_sin_cache(128, rot, torch.bfloat16, device=dev)
    slots = torch.randperm(nb * bs, device=dev)[:T]; slots[::7] = -1
    ks, vs = swiftkv.kv_rope_write_bulk(kvl, pos, cs, [c[0] for c in cch], [c[1] for c in cch], slots, kvd, sc, sc, Hkv, Dh2,
                                        rotary_dim=rot, is_neox_style=neox, rounding=rounding,
                                        logits_indices=torch.tensor([0, 5, -1, 49], device=dev))
    print("rope", Dh2, rot, neox, rounding, kvd, ks.shape)
torch.cuda.synchronize()
print("sanitize target done")

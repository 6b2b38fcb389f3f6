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
for B, q in ((7, True), (200, False)):          # FP8 head path; CTA-pair GEMM path
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
torch.cuda.synchronize()
print("sanitize target done")

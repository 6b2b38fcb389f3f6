"""Dev probe (GPU box): which arithmetic does the installed vLLM wheel's rotary_embedding CUDA op use?
Candidates: per-op rounding in the storage dtype (forward_native) vs fp32 internally with one final rounding."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vllm, vllm._custom_ops as vops
from oracle import kv_rope_ref as R
dev = "cuda"
print("vllm", vllm.__version__)
for dtype in (torch.bfloat16, torch.float16):
    for neox in (True, False):
        T, Hkv, Dh = 2048, 8, 128
        g = torch.Generator(device=dev).manual_seed(1)
        k = torch.randn(T, Hkv * Dh, device=dev, generator=g).to(dtype)
        q = torch.randn(T, Hkv * Dh, device=dev, generator=g).to(dtype)
        pos = torch.randint(0, 8192, (T,), device=dev, generator=g)
        cs = R.make_cos_sin_cache(8192, Dh, dtype, device=dev)
        live = k.clone(); vops.rotary_embedding(pos, q.clone(), live, Dh, cs, neox)
        per_op = R.rope_torch(pos, k, cs, Dh, Dh, neox)
        f32 = R.rope_torch(pos, k.float(), cs.float(), Dh, Dh, neox).to(dtype)            # fp32 ops, one final rounding
        f64 = R.rope_torch(pos, k.double(), cs.double(), Dh, Dh, neox)
        via32 = f64.float().to(dtype)                                                      # exact -> f32 -> storage
        direct = f64.to(dtype)                                                             # exact -> storage
        eq = lambda a: (a.view(torch.int16) == live.view(torch.int16)).float().mean().item()
        print(dtype, "neox" if neox else "gptj", "per_op", eq(per_op), "fp32", eq(f32), "exact->f32->st", eq(via32),
              "exact->st", eq(direct))

"""TEST INFRASTRUCTURE ONLY - golden vectors for the FP8 KV-cache write from vLLM's own per-layer op
(torch.ops._C_cache_ops.reshape_and_cache_flash, the reference's fallback path
llama_swiftkv.py:629-656) run on a B200.

    gpurun -- python oracle/gen_golden_kv_fp8_gpu.py     # writes gpurun_out/golden/kvfp8_*.npz
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vllm import _custom_ops as ops  # noqa: E402

# name: (L, T, Hkv, Dh, bs, src dtype, kv_cache_dtype, value std)
CASES = {
    "e4m3_bf16": (3, 45, 2, 32, 16, torch.bfloat16, "fp8", 1.0),
    "e4m3_fp16": (2, 33, 4, 16, 8, torch.float16, "fp8_e4m3", 3.0),
    "e4m3_fp32": (2, 21, 2, 16, 16, torch.float32, "fp8", 0.05),
    "e5m2_bf16": (2, 40, 2, 32, 16, torch.bfloat16, "fp8_e5m2", 20.0),
}


def main():
    out = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out, exist_ok=True)
    dev = "cuda"
    for name, (L, T, Hkv, Dh, bs, dt, kvd, std) in CASES.items():
        g = torch.Generator(device="cpu").manual_seed(7)
        n = Hkv * Dh
        nb = 2 * ((T + bs - 1) // bs)
        keys = (torch.randn(T, L * n, generator=g) * std).to(dt)
        values = (torch.randn(T, L * n, generator=g) * std).to(dt)
        # specials: zeros, tiny, values that saturate, exact ties of the fp8 grid
        keys[0, :8] = torch.tensor([0.0, -0.0, 1e-9, -1e-9, 1e6, -1e6, 0.4375, 0.46875]).to(dt)
        slots = torch.randperm(nb * bs, generator=g)[:T].to(torch.int64)
        slots[torch.rand(T, generator=g) < 0.1] = -1
        k_scales = [torch.tensor(0.5 + 0.37 * l, dtype=torch.float32) for l in range(L)]
        v_scales = [torch.tensor(0.011 * (l + 1), dtype=torch.float32) for l in range(L)]
        init = [torch.randint(0, 256, (2, nb, bs, Hkv, Dh), generator=g, dtype=torch.uint8) for _ in range(L)]
        caches = [c.clone().to(dev) for c in init]
        dk, dv, ds = keys.to(dev), values.to(dev), slots.to(dev)
        for l in range(L):
            ops.reshape_and_cache_flash(dk[:, l * n:(l + 1) * n].reshape(T, Hkv, Dh).contiguous(),
                                        dv[:, l * n:(l + 1) * n].reshape(T, Hkv, Dh).contiguous(),
                                        caches[l][0], caches[l][1], ds, kvd, k_scales[l].to(dev),
                                        v_scales[l].to(dev))
        torch.cuda.synchronize()
        view = torch.int16 if dt != torch.float32 else torch.int32
        rec = {"meta": np.array([L, T, Hkv, Dh, bs, nb], dtype=np.int64), "kv_cache_dtype": np.array(kvd),
               "src_dtype": np.array(str(dt)), "slots": slots.numpy(),
               "keys": keys.view(view).numpy(), "values": values.view(view).numpy(),
               "k_scales": np.array([float(s) for s in k_scales], dtype=np.float32),
               "v_scales": np.array([float(s) for s in v_scales], dtype=np.float32)}
        for l in range(L):
            rec[f"init{l}"] = init[l].numpy()
            rec[f"final{l}"] = caches[l].cpu().numpy()
        path = os.path.join(out, f"kvfp8_{name}.npz")
        np.savez_compressed(path, **rec)
        print(name, "->", path, os.path.getsize(path))


if __name__ == "__main__":
    main()

"""TEST INFRASTRUCTURE ONLY - generates tests/golden/rope_vllm_*.npz on a GPU box: vLLM's own rotary_embedding CUDA
op (the op the reference reaches through ``rotary_emb(positions, q, k)``, llama_swiftkv.py:270; installed wheel: vllm
0.22 - the reference pins 0.9.2, absent) applied to small seeded inputs.  The goldens pin oracle/kv_rope_ref.py
(``compute="fp32"``) on the CPU; run with
    gpurun -- python oracle/gen_golden_rope_gpu.py && cp gpurun_out/rope_vllm_*.npz tests/golden/
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import kv_rope_ref as R  # noqa: E402

CASES = {  # name: (T, Hkv, Dh, rot_dim, neox, dtype, max_pos)
    "neox_d128_bf16": (24, 4, 128, 128, True, torch.bfloat16, 4096),
    "neox_d64_bf16": (17, 8, 64, 64, True, torch.bfloat16, 32768),
    "neox_d128_f16": (9, 2, 128, 128, True, torch.float16, 4096),
    "gptj_d128_rot64_bf16": (12, 2, 128, 64, False, torch.bfloat16, 2048),
}


def to_np(t):
    return t.cpu().contiguous().view(torch.int16).numpy() if t.element_size() == 2 else t.cpu().numpy()


def main():
    import vllm
    import vllm._custom_ops as vops
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    dev = "cuda"
    for name, (T, Hkv, Dh, rot, neox, dtype, max_pos) in CASES.items():
        g = torch.Generator(device="cpu").manual_seed(len(name) * 7 + T)
        k = (torch.randn(T, Hkv * Dh, generator=g) * 1.5).to(dtype)
        q = torch.randn(T, Hkv * Dh, generator=g).to(dtype)
        pos = torch.randint(0, max_pos, (T,), generator=g, dtype=torch.int64)
        cs = R.make_cos_sin_cache(max_pos, rot, dtype)
        kd, qd = k.to(dev).clone(), q.to(dev).clone()
        vops.rotary_embedding(pos.to(dev), qd, kd, Dh, cs.to(dev), neox)
        torch.cuda.synchronize()
        path = os.path.join(out_dir, f"rope_vllm_{name}.npz")
        np.savez_compressed(path, k=to_np(k), positions=pos.numpy(), cos_sin_rows=to_np(cs.index_select(0, pos)),
                            k_rot=to_np(kd), meta=np.array([T, Hkv, Dh, rot, int(neox), max_pos], dtype=np.int64),
                            vllm_version=np.array([int(x) for x in vllm.__version__.split(".")[:2]]))
        print(name, "->", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

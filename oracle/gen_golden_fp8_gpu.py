"""TEST INFRASTRUCTURE ONLY - golden vectors for the FP8 head oracle from vLLM's own ops
(the backend of the reference's qhead) run on a B200.

    gpurun -- python oracle/gen_golden_fp8_gpu.py     # writes gpurun_out/golden/fp8_vllm_*.npz
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vllm import _custom_ops as ops  # noqa: E402

CASES = {"small": (512, 256, 24, 0.02, 1.0), "wide": (1024, 512, 32, 0.05, 3.0), "tinyrows": (256, 128, 5, 0.02, 1e-4)}


def main():
    out = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out, exist_ok=True)
    dev = "cuda"
    for name, (V, D, M, wstd, xstd) in CASES.items():
        g = torch.Generator(device="cpu").manual_seed(11)
        w = (torch.randn(V, D, generator=g) * wstd).to(torch.bfloat16)
        x = (torch.randn(M, D, generator=g) * xstd).to(torch.bfloat16)
        if name == "tinyrows":
            x[0] = 0          # all-zero row -> minimum scale
        qw, sw = ops.scaled_fp8_quant(w.to(dev), scale=None)
        qx, sx = ops.scaled_fp8_quant(x.to(dev), scale=None, use_per_token_if_dynamic=True)
        mm = ops.cutlass_scaled_mm(qx, qw.t(), scale_a=sx, scale_b=sw, out_dtype=torch.bfloat16)
        rec = {"w": w.view(torch.int16).numpy(), "x": x.view(torch.int16).numpy(),
               "qw": qw.view(torch.uint8).cpu().numpy(), "sw": sw.float().cpu().numpy(),
               "qx": qx.view(torch.uint8).cpu().numpy(), "sx": sx.float().cpu().numpy(),
               "out": mm.view(torch.int16).cpu().numpy()}
        path = os.path.join(out, f"fp8_vllm_{name}.npz")
        np.savez_compressed(path, **rec)
        print(name, "->", path, os.path.getsize(path))


if __name__ == "__main__":
    main()

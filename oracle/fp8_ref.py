"""TEST INFRASTRUCTURE ONLY - oracle for the FP8 (e4m3) quantised head.

The reference's ``qhead`` (arctic_speculator.py:455-465, used iff the padded batch <= 32, :726-727)
is vLLM's Fp8LinearMethod applied to an unquantised checkpoint
(arctic_inference/vllm/spec_dec/fp8.py:207-221, 276-308).  The arithmetic lives in vLLM
(``ops.scaled_fp8_quant`` + ``cutlass_scaled_mm``; pinned dependency vllm==0.9.2, absent here - the
image has vllm 0.22 whose ops implement the same published algorithm).  This module restates it:

    weights      s_w = amax|W| / 448                       q = sat_e4m3(w * (1 / s_w))
    activations  s_x = max(amax|x_row| / 448, 1/(448*512))  q = sat_e4m3(x / s_x)
    output       bf16( s_x * (s_w * (q_x @ q_w^T in fp32)) )

Pinned against vLLM 0.22's own ops run on a B200: tests/golden/fp8_vllm_*.npz
(oracle/gen_golden_fp8_gpu.py).  vLLM computes its scales with fast-math division, so scales may
differ from the IEEE value by 1 ulp and a ~1e-5 fraction of quantised elements by one code; the
tests state these tolerances.  Parity with the reference's exact vllm 0.9.2 binary: unpinned.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from oracle.speculator_ref import RefLSTMSpeculator

FP8_MAX = 448.0
MIN_SCALE = 1.0 / (448.0 * 512.0)
F8 = torch.float8_e4m3fn


def _sat_cast(x32: torch.Tensor) -> torch.Tensor:
    return x32.clamp(-FP8_MAX, FP8_MAX).to(F8)


def _fp8_max(ref: torch.Tensor) -> torch.Tensor:
    # a tensor divisor: ATen's CUDA kernel turns division by a *Python scalar* into multiplication
    # by its reciprocal, which is not what the kernels (IEEE division) do
    return torch.tensor(FP8_MAX, dtype=torch.float32, device=ref.device)


def quantize_weight(w: torch.Tensor):
    """fp8.py:207-209 -> ops.scaled_fp8_quant(weight, scale=None): dynamic per-tensor."""
    w32 = w.float()
    scale = (w32.abs().max() / _fp8_max(w32)).reshape(1)
    inv = torch.ones_like(scale) / scale
    return _sat_cast(w32 * inv), scale


def quantize_rows(x: torch.Tensor):
    """Fp8LinearOp.apply -> ops.scaled_fp8_quant(x, use_per_token_if_dynamic=True)."""
    x32 = x.float()
    scale = torch.clamp(x32.abs().amax(-1, keepdim=True) / _fp8_max(x32), min=MIN_SCALE)
    return _sat_cast(x32 / scale), scale


def scaled_mm(qx: torch.Tensor, sx: torch.Tensor, qw: torch.Tensor, sw: torch.Tensor) -> torch.Tensor:
    """cutlass_scaled_mm(qx, qw.t(), scale_a=sx, scale_b=sw, out_dtype=bf16)."""
    acc = qx.float() @ qw.float().t()
    return (sx * (sw * acc)).to(torch.bfloat16)


def fp8_linear(x: torch.Tensor, qw: torch.Tensor, sw: torch.Tensor) -> torch.Tensor:
    qx, sx = quantize_rows(x)
    return scaled_mm(qx, sx, qw, sw)


class RefLSTMSpeculatorFp8(RefLSTMSpeculator):
    """Op-for-op speculator restatement with the reference's head selection rule: FP8 ``qhead``
    for batch <= 32, bf16 ``head`` above (arctic_speculator.py:726-727)."""

    def __init__(self, weights: Dict[str, torch.Tensor], vocab_size: int, n_predict: int,
                 max_batch_fp8: int = 32, **kw):
        super().__init__(weights, vocab_size, n_predict, **kw)
        self.qw, self.sw = quantize_weight(weights["head.0.weight"])
        self.max_batch_fp8 = max_batch_fp8

    def logits(self, states_2d):
        if states_2d.shape[0] <= self.max_batch_fp8:
            return fp8_linear(states_2d, self.qw, self.sw)[..., :self.vocab_size]
        return super().logits(states_2d)

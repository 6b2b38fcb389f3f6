"""TEST INFRASTRUCTURE ONLY - oracle for the FP8 KV-cache variant of the bulk reshape-and-cache
(SURVEY.md 8f rank 2).

The reference's bulk kernel contains the FP8 branch (csrc/custom_ops/kernels.cu:57-66 ->
``fp8::scaled_convert`` -> csrc/custom_ops/quant_utils.cuh:456-489) but compiles it out
(``ENABLE_FP8`` is never defined); its *caller* then falls back to vLLM's per-layer op
``torch.ops._C_cache_ops.reshape_and_cache_flash`` (llama_swiftkv.py:629-656), which performs the
same published conversion:

    cache_byte = cvt_rn_satfinite_fp8( float(x) / scale_layer )      e4m3 ("fp8", "fp8_e4m3") or e5m2

This module restates that with torch CPU ops (IEEE fp32 division by a tensor divisor, clamp to the
format's finite range, round-to-nearest-even cast).  Pinned against golden vectors written by
vLLM 0.22's own op on a B200 (tests/golden/kvfp8_*.npz, oracle/gen_golden_kv_fp8_gpu.py;
the reference pins vllm 0.9.2, absent here - same algorithm).  Nothing under
arcticinference_b200/ may import this module.
"""
from __future__ import annotations

from typing import List

import torch

FMT = {"fp8": (torch.float8_e4m3fn, 448.0), "fp8_e4m3": (torch.float8_e4m3fn, 448.0),
       "fp8_e5m2": (torch.float8_e5m2, 57344.0)}


def quantize(x: torch.Tensor, scale: torch.Tensor, kv_cache_dtype: str) -> torch.Tensor:
    """bytes of cvt_satfinite(float(x) / scale); NaN stays NaN (0x7f for e4m3, as the hardware cvt)."""
    f8, fmax = FMT[kv_cache_dtype]
    q = x.float() / scale.float().reshape(())        # tensor divisor: true division, not * (1/s)
    q = torch.where(q.isnan(), q, q.clamp(-fmax, fmax))
    return q.to(f8).view(torch.uint8)


def kv_write_fp8(keys: torch.Tensor, values: torch.Tensor, key_caches: List[torch.Tensor],
                 value_caches: List[torch.Tensor], slot_mapping: torch.Tensor, kv_cache_dtype: str,
                 k_scales: List[torch.Tensor], v_scales: List[torch.Tensor], num_heads: int,
                 head_size: int) -> None:
    """key_caches[l] / value_caches[l]: uint8 views [nb, bs, Hkv, Dh] (any block stride)."""
    L = len(key_caches)
    if L == 0:
        return
    T = slot_mapping.shape[0]
    bs = key_caches[0].shape[1]
    n = num_heads * head_size
    sm = slot_mapping[:T]
    rows = torch.nonzero(sm >= 0).flatten()
    s = sm[rows]
    blk, off = s // bs, s % bs
    for l in range(L):
        kq = quantize(keys[rows, l * n:(l + 1) * n], k_scales[l], kv_cache_dtype)
        vq = quantize(values[rows, l * n:(l + 1) * n], v_scales[l], kv_cache_dtype)
        key_caches[l][blk, off] = kq.reshape(-1, num_heads, head_size)
        value_caches[l][blk, off] = vq.reshape(-1, num_heads, head_size)

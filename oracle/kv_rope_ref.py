"""TEST INFRASTRUCTURE ONLY - oracle for the fused K-rotary + paged-cache write.

Restates, with plain torch ops (CPU or CUDA), the reference sequence the fused kernel replaces:

* ``rope_torch``         vLLM's rotary embedding as the reference calls it
                         (arctic_inference/vllm/swiftkv/llama_swiftkv.py:263,270 -> ``rotary_emb(positions,
                         q, k)`` -> vllm RotaryEmbedding.forward).  The formula is vLLM's
                         ``ApplyRotaryEmb.forward_static`` (o1 = x1*cos - x2*sin, o2 = x2*cos + x1*sin) in two
                         rounding contracts: ``compute="fp32"`` - fp32 arithmetic, one final rounding: what the
                         rotary_embedding CUDA op does (measured on a B200 against the installed vllm 0.22 wheel:
                         100 % of the elements bit-equal for bf16 / f16, NeoX / GPT-J, tools/rope_probe.py;
                         tests/test_kv_rope_gpu.py re-checks it live); ``compute="storage"`` - the same formula
                         in the storage dtype (forward_native: a rounding after every op; equals the CUDA op on
                         only ~72 % of the elements, the rest is off by the rounding of the two products).  vLLM is a third-party dependency
                         of the reference (pinned vllm==0.9.2, pyproject.toml:42, absent here): parity with that
                         exact binary is unpinned; the installed wheel's op is the anchor.
* ``reference_sequence`` llama_swiftkv.py:263-274: per layer ``k, v = kv.chunk(2, -1)``, rotate k,
                         ``cat`` -> k_states, v_states ``[T, L*n]``.
* ``write_caches``       llama_swiftkv.py:599-628: the bulk cache write (torch index_put restatement of
                         csrc/custom_ops/kernels.cu:29-68; the numpy/C oracle of kv_write_ref.py is the
                         pinned one - this is the same scatter on torch tensors so it can run on the GPU).
* ``make_cos_sin_cache`` vLLM RotaryEmbedding._compute_cos_sin_cache: inv_freq = base^(-2i/rot), cache =
                         cat(cos(t * inv_freq), sin(t * inv_freq)) cast to the model dtype.

Nothing under arcticinference_b200/ may import this.
"""
from __future__ import annotations

from typing import List, Tuple

import torch


def make_cos_sin_cache(max_position: int, rot_dim: int, dtype: torch.dtype, base: float = 500000.0,
                       device="cpu") -> torch.Tensor:
    inv_freq = 1.0 / (base ** (torch.arange(0, rot_dim, 2, dtype=torch.float) / rot_dim))
    t = torch.arange(max_position, dtype=torch.float)
    freqs = torch.einsum("i,j -> ij", t, inv_freq)
    return torch.cat((freqs.cos(), freqs.sin()), dim=-1).to(dtype).to(device)


def rope_torch(positions: torch.Tensor, k: torch.Tensor, cos_sin_cache: torch.Tensor, head_size: int,
               rot_dim: int, is_neox: bool, compute: str = "fp32") -> torch.Tensor:
    """k [T, Hkv*Dh] -> rotated copy (see the module docstring for ``compute``)."""
    if compute == "fp32" and k.dtype != torch.float32:
        return rope_torch(positions, k.float(), cos_sin_cache.float(), head_size, rot_dim, is_neox, "storage").to(k.dtype)
    T = k.shape[0]
    cos_sin = cos_sin_cache.index_select(0, positions)
    cos, sin = cos_sin[:, :rot_dim].chunk(2, dim=-1)
    kk = k.reshape(T, -1, head_size)
    k_rot, k_pass = kk[..., :rot_dim], kk[..., rot_dim:]
    cos = cos.unsqueeze(-2).to(k.dtype)
    sin = sin.unsqueeze(-2).to(k.dtype)
    if is_neox:
        x1, x2 = torch.chunk(k_rot, 2, dim=-1)
    else:
        x1, x2 = k_rot[..., ::2], k_rot[..., 1::2]
    o1 = x1 * cos - x2 * sin
    o2 = x2 * cos + x1 * sin
    out = torch.cat((o1, o2), dim=-1) if is_neox else torch.stack((o1, o2), dim=-1).flatten(-2)
    return torch.cat((out, k_pass), dim=-1).reshape(T, -1)


def reference_sequence(kv_list: List[torch.Tensor], positions: torch.Tensor, cos_sin_cache: torch.Tensor,
                       num_heads: int, head_size: int, rot_dim: int, is_neox: bool, compute: str = "fp32"
                       ) -> Tuple[torch.Tensor, torch.Tensor]:
    """llama_swiftkv.py:263-274."""
    k_states, v_states = [], []
    for kv in kv_list:
        k, v = kv.chunk(2, dim=-1)
        k_states.append(rope_torch(positions, k, cos_sin_cache, head_size, rot_dim, is_neox, compute))
        v_states.append(v)
    return torch.cat(k_states, dim=-1), torch.cat(v_states, dim=-1)


def write_caches(k_states: torch.Tensor, v_states: torch.Tensor, key_caches: List[torch.Tensor],
                 value_caches: List[torch.Tensor], slot_mapping: torch.Tensor, num_heads: int, head_size: int):
    """Bulk write (kernels.cu:29-68, 'auto'): slot < 0 skipped; caches [nb, bs, Hkv, Dh] views."""
    n = num_heads * head_size
    bs = key_caches[0].shape[1]
    valid = slot_mapping >= 0
    s = slot_mapping[valid]
    blk, off = s // bs, s % bs
    for l in range(len(key_caches)):
        key_caches[l][blk, off] = k_states[valid][:, l * n:(l + 1) * n].reshape(-1, num_heads, head_size)
        value_caches[l][blk, off] = v_states[valid][:, l * n:(l + 1) * n].reshape(-1, num_heads, head_size)

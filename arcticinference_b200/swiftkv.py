"""SwiftKV select step - host mirror of the tail of
``LlamaSwiftKVModel.swiftkv_select`` (arctic_inference/vllm/swiftkv/llama_swiftkv.py:599-685):

1. bulk-write the K/V of every SwiftKV layer into the paged caches (``reshape_and_cache_flash_bulk``,
   :599-628 - the torch op of arcticinference_b200.py_custom_ops), then
2. compact ``hidden_states, residual, positions, k_states, v_states`` to the surviving tokens
   (``index_fn`` :669-685: five ``torch.index_select`` calls, into the decode runner's CUDA-graph
   input buffers when the batch fits, padded to the graph batch size).

Step 2 is ONE launch here (``ab200_gather_rows``); there is no PyTorch fallback.

``kv_rope_write_bulk`` goes one step further up the call chain (llama_swiftkv.py:263-274): it takes the
per-layer ``kv_proj_swiftkv`` outputs themselves, applies the K rotary with vLLM's rounding and scatters
K / V into the paged caches in ONE launch - the ``[T, L*n]`` ``k_states`` / ``v_states`` tensors are never
built; only the rows of the surviving tokens are emitted (for the decode runner).
"""
from __future__ import annotations

import ctypes
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import _capi


def gather_rows(tensors: Sequence[torch.Tensor], indices: torch.Tensor,
                outs: Optional[Sequence[Optional[torch.Tensor]]] = None) -> List[torch.Tensor]:
    """``[t.index_select(0, indices) for t in tensors]`` in one launch.  ``outs[i]`` (optional) is a
    preallocated buffer with at least ``indices.numel()`` rows (``torch.index_select(..., out=)``)."""
    lib = _capi.lib()
    n = len(tensors)
    if n == 0:
        return []
    if n > 8:
        return gather_rows(tensors[:8], indices, outs[:8] if outs else None) + \
            gather_rows(tensors[8:], indices, outs[8:] if outs else None)
    dev = tensors[0].device
    if dev.type != "cuda":
        raise NotImplementedError("gather_rows: CUDA tensors only (no CPU fallback)")
    if indices.dtype != torch.int64 or indices.dim() != 1:
        raise RuntimeError("gather_rows: indices must be a 1-D int64 tensor")
    if indices.device != dev or not indices.is_contiguous():
        indices = indices.to(dev).contiguous()
    m = indices.numel()
    rows0 = tensors[0].shape[0]
    results: List[torch.Tensor] = []
    srcs, dsts, ss, ds, rb = [], [], [], [], []
    for i, t in enumerate(tensors):
        if t.shape[0] != rows0:
            raise RuntimeError("gather_rows: all tensors must have the same number of rows")
        t2 = t if t.dim() > 1 else t.unsqueeze(1)
        if t2.shape[0] and t2[0].numel() and not t2[0].is_contiguous():
            t2 = t2.contiguous()
        out = outs[i] if outs is not None and outs[i] is not None else None
        if out is None:
            out = torch.empty((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        else:
            if out.dtype != t.dtype or out.shape[1:] != t.shape[1:] or out.shape[0] < m:
                raise RuntimeError("gather_rows: out buffer has the wrong dtype / shape")
            out = out[:m]
        o2 = out if out.dim() > 1 else out.unsqueeze(1)
        if o2.shape[0] and o2[0].numel() and not o2[0].is_contiguous():
            raise RuntimeError("gather_rows: out rows must be dense")
        results.append(out)
        srcs.append(t2.data_ptr()); dsts.append(o2.data_ptr())
        ss.append(t2.stride(0) * t2.element_size()); ds.append(o2.stride(0) * o2.element_size())
        rb.append((t2.numel() // t2.shape[0] if t2.shape[0] else 0) * t2.element_size())
    if m == 0 or rows0 == 0:
        return results
    keep = [i for i in range(n) if rb[i] > 0]
    if not keep:
        return results
    k = len(keep)
    # one host buffer for the five per-tensor arrays of the C ABI (pointers are 8 bytes, like the int64s)
    flat = (ctypes.c_int64 * (5 * k))(*[srcs[i] for i in keep], *[dsts[i] for i in keep], *[ss[i] for i in keep],
                                      *[ds[i] for i in keep], *[rb[i] for i in keep])
    base = ctypes.addressof(flat)
    if torch.cuda.current_device() == dev.index:
        rc = lib.ab200_gather_rows(base, base + 8 * k, base + 16 * k, base + 24 * k, base + 32 * k, k,
                                   indices.data_ptr(), m, rows0, _capi.current_stream_ptr(dev))
    else:
        with torch.cuda.device(dev):
            rc = lib.ab200_gather_rows(base, base + 8 * k, base + 16 * k, base + 24 * k, base + 32 * k, k,
                                       indices.data_ptr(), m, rows0, _capi.current_stream_ptr(dev))
    _capi.check(rc, "ab200_gather_rows")
    return results


def swiftkv_select(hidden_states: torch.Tensor, residual: torch.Tensor, positions: torch.Tensor,
                   k_states: torch.Tensor, v_states: torch.Tensor, *,
                   key_caches: List[torch.Tensor], value_caches: List[torch.Tensor],
                   slot_mapping: torch.Tensor, kv_cache_dtype: str, k_scales: List[torch.Tensor],
                   v_scales: List[torch.Tensor], num_kv_heads: int, head_size: int,
                   logits_indices: torch.Tensor,
                   graph_buffers: Optional[Dict[str, torch.Tensor]] = None,
                   cuda_graph_max_batch_size: int = 0,
                   pad_for_cudagraph: Optional[Callable[[int], int]] = None
                   ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """llama_swiftkv.py:599-685 without the vLLM objects: the per-layer caches, the slot mapping and
    the surviving-token indices are passed in directly.  ``graph_buffers`` = the decode runner's
    ``inputs`` dict (:676-681): used when 0 < batch <= cuda_graph_max_batch_size, results are views
    padded to ``pad_for_cudagraph(batch)`` rows as in the reference."""
    from . import py_custom_ops
    if len(key_caches) > 0:                                            # :622-628
        if not py_custom_ops.try_load_torch_library():
            raise RuntimeError("libarctic_b200 custom ops are not available (no fallback on this path)")
        py_custom_ops.reshape_and_cache_flash_bulk(k_states, v_states, key_caches, value_caches, slot_mapping,
                                                   kv_cache_dtype, k_scales, v_scales, num_kv_heads, head_size)
    names = ("hidden_states", "residual", "positions", "k_states", "v_states")
    tensors = (hidden_states, residual, positions, k_states, v_states)
    b = logits_indices.numel()
    use_buf = graph_buffers is not None and 0 < b <= cuda_graph_max_batch_size          # :676
    outs = [graph_buffers[nm] for nm in names] if use_buf else None
    got = gather_rows(tensors, logits_indices, outs)
    if use_buf:
        padded = pad_for_cudagraph(b) if pad_for_cudagraph else b                        # :679-680
        got = [graph_buffers[nm][:padded] for nm in names]
    return tuple(got)  # type: ignore[return-value]


_KV_DTYPES = {"auto": _capi.AB200_KV_AUTO, "fp8": _capi.AB200_KV_FP8_E4M3,
              "fp8_e4m3": _capi.AB200_KV_FP8_E4M3, "fp8_e5m2": _capi.AB200_KV_FP8_E5M2}


def kv_rope_write_bulk(kv_list: Sequence[torch.Tensor], positions: torch.Tensor, cos_sin_cache: torch.Tensor,
                       key_caches: List[torch.Tensor], value_caches: List[torch.Tensor],
                       slot_mapping: torch.Tensor, kv_cache_dtype: str, k_scales: List[torch.Tensor],
                       v_scales: List[torch.Tensor], num_kv_heads: int, head_size: int, *,
                       rotary_dim: Optional[int] = None, is_neox_style: bool = True, rounding: str = "fp32",
                       logits_indices: Optional[torch.Tensor] = None,
                       k_states_out: Optional[torch.Tensor] = None, v_states_out: Optional[torch.Tensor] = None
                       ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """Fused replacement of llama_swiftkv.py:263-274 (+ the cache write :599-628, + the k/v index_select
    :669-685).  ``kv_list[l]`` = ``layer.self_attn.kv_proj_swiftkv(h)[0]`` of SwiftKV layer l, ``[T, 2*n]`` (K
    half, then V half: ``kv.chunk(2, dim=-1)``); ``cos_sin_cache`` = ``rotary_emb.cos_sin_cache`` in the kv
    dtype.  ``rounding``: ``"fp32"`` (default) = vLLM's rotary_embedding CUDA op, fp32 internally and one
    rounding - what ``rotary_emb(positions, q, k)`` runs on a GPU; ``"storage"`` = ``forward_native`` in the
    storage dtype (a rounding after every operation).  Caches / slot_mapping / scales exactly as
    ``reshape_and_cache_flash_bulk``.  With
    ``logits_indices`` the rotated K rows and the V rows of those tokens are returned as ``[m, L*n]`` tensors
    (written into ``k_states_out`` / ``v_states_out`` when given)."""
    L = len(kv_list)
    if L == 0:
        return None, None
    kv0 = kv_list[0]
    dev = kv0.device
    if dev.type != "cuda":
        raise NotImplementedError("kv_rope_write_bulk: CUDA tensors only (no CPU fallback)")
    if not (L == len(key_caches) == len(value_caches)):
        raise RuntimeError("kv_rope_write_bulk: one key cache and one value cache per kv tensor expected")
    if kv_cache_dtype not in _KV_DTYPES:
        raise RuntimeError(f"Unsupported data type of kv cache: {kv_cache_dtype}")
    if rounding not in ("fp32", "storage"):
        raise ValueError("rounding must be 'fp32' or 'storage'")
    n = num_kv_heads * head_size
    T = slot_mapping.shape[0]
    for kv in kv_list:
        if kv.dim() != 2 or kv.shape[1] != 2 * n or kv.stride(1) != 1 or kv.stride(0) != kv0.stride(0) \
                or kv.dtype != kv0.dtype or kv.shape[0] < T:
            raise RuntimeError("kv_rope_write_bulk: every kv tensor must be [T, 2*num_kv_heads*head_size] with the "
                               "same dtype and row stride")
    if cos_sin_cache.dtype != kv0.dtype or cos_sin_cache.dim() != 2 or cos_sin_cache.stride(1) != 1:
        raise RuntimeError("kv_rope_write_bulk: cos_sin_cache must be [max_position, rot_dim] in the kv dtype")
    rot = rotary_dim if rotary_dim is not None else cos_sin_cache.shape[1]
    if positions.dtype != torch.int64 or slot_mapping.dtype != torch.int64:
        raise RuntimeError("kv_rope_write_bulk: positions and slot_mapping must be int64")
    positions = positions.contiguous()
    slot_mapping = slot_mapping.contiguous()
    if T == 0:
        return None, None
    block_size, block_stride = key_caches[0].size(1), key_caches[0].stride(0)
    if block_stride != value_caches[0].stride(0):
        raise RuntimeError("block_stride == value_caches[0].stride(0)")
    m = 0
    if logits_indices is not None:
        logits_indices = logits_indices.to(device=dev, dtype=torch.int64).contiguous()
        m = logits_indices.numel()
        if k_states_out is None:
            k_states_out = torch.empty(m, L * n, dtype=kv0.dtype, device=dev)
        if v_states_out is None:
            v_states_out = torch.empty(m, L * n, dtype=kv0.dtype, device=dev)
        if k_states_out.shape[0] < m or k_states_out.shape[1] != L * n or k_states_out.stride(1) != 1 or \
                v_states_out.shape[0] < m or v_states_out.stride() != k_states_out.stride():
            raise RuntimeError("kv_rope_write_bulk: k_states_out / v_states_out must be [>= m, L*n] with equal strides")
    arr = ctypes.c_void_p * L
    with torch.cuda.device(dev):
        rc = _capi.lib().ab200_kv_rope_write_bulk(
            arr(*[t.data_ptr() for t in kv_list]), kv0.stride(0), n, positions.data_ptr(), cos_sin_cache.data_ptr(),
            cos_sin_cache.stride(0), cos_sin_cache.shape[0], rot, int(is_neox_style),
            _capi.AB200_ROPE_FP32 if rounding == "fp32" else _capi.AB200_ROPE_STORAGE,
            arr(*[t.data_ptr() for t in key_caches]), arr(*[t.data_ptr() for t in value_caches]),
            arr(*[t.data_ptr() for t in k_scales]), arr(*[t.data_ptr() for t in v_scales]), L,
            slot_mapping.data_ptr(), T, block_stride, block_size, num_kv_heads, head_size,
            _capi.torch_dtype_code(kv0.dtype), _KV_DTYPES[kv_cache_dtype],
            logits_indices.data_ptr() if m else None, m,
            k_states_out.data_ptr() if m else None, v_states_out.data_ptr() if m else None,
            k_states_out.stride(0) if m else 0, _capi.current_stream_ptr(dev))
    _capi.check(rc, "ab200_kv_rope_write_bulk")
    if logits_indices is None:
        return None, None
    return k_states_out[:m], v_states_out[:m]

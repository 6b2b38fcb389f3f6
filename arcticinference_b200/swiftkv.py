"""SwiftKV select step - host mirror of the tail of
``LlamaSwiftKVModel.swiftkv_select`` (arctic_inference/vllm/swiftkv/llama_swiftkv.py:599-685):

1. bulk-write the K/V of every SwiftKV layer into the paged caches (``reshape_and_cache_flash_bulk``,
   :599-628 - the torch op of arcticinference_b200.py_custom_ops), then
2. compact ``hidden_states, residual, positions, k_states, v_states`` to the surviving tokens
   (``index_fn`` :669-685: five ``torch.index_select`` calls, into the decode runner's CUDA-graph
   input buffers when the batch fits, padded to the graph batch size).

Step 2 is ONE launch here (``ab200_gather_rows``); there is no PyTorch fallback.
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

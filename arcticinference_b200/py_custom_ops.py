"""Drop-in for the reference's ``arctic_inference/py_custom_ops.py`` (hot path B boundary).

Same two public names, same argument order, same behaviour contract:

* ``try_load_torch_library() -> bool``  never raises; True once
  ``torch.ops.arctic_inference.reshape_and_cache_flash_bulk`` is usable
  (reference: py_custom_ops.py:9-37).
* ``reshape_and_cache_flash_bulk(keys, values, key_caches, value_caches, slot_mapping,
  kv_cache_dtype, k_scales, v_scales, num_heads, head_size) -> None``
  (reference: py_custom_ops.py:40-54), forwarding to the torch op.

The op is defined with the reference's schema string, byte for byte
(csrc/custom_ops/torch_bindings.cpp:6-16), implemented for the CUDA dispatch key only
(torch_bindings.cpp:17-18), and backed by ``ab200_kv_write_bulk`` in libarctic_b200.so through
ctypes - hand-written sm_100a CUDA, no CPU fallback.  Argument derivation and error behaviour
follow the reference's host launcher (csrc/custom_ops/kernels.cu:86-155): RuntimeError where it
TORCH_CHECKs, silent return for zero layers.
"""
from __future__ import annotations

import ctypes
import logging
from typing import List

import torch

from . import _capi

logger = logging.getLogger(__name__)

# csrc/custom_ops/torch_bindings.cpp:6-16 (whitespace between tokens is not significant to the
# schema parser; names, alias annotations and order are)
OP_SCHEMA = ("reshape_and_cache_flash_bulk(Tensor keys,"
             " Tensor values,"
             " Tensor(c!)[] key_caches,"
             " Tensor(d!)[] value_caches,"
             " Tensor slot_mapping,"
             " str kv_cache_dtype,"
             " Tensor(e)[] k_scales,"
             " Tensor(f)[] v_scales,"
             " int num_heads,"
             " int head_size) -> ()")

_KV_DTYPES = {"auto": _capi.AB200_KV_AUTO, "fp8": _capi.AB200_KV_FP8_E4M3,
              "fp8_e4m3": _capi.AB200_KV_FP8_E4M3, "fp8_e5m2": _capi.AB200_KV_FP8_E5M2}

_library = None  # keeps the torch.library.Library alive
_registered = False


def _check(cond: bool, msg: str) -> None:
    if not cond:
        raise RuntimeError(msg)  # what TORCH_CHECK raises on the Python side


def _reshape_and_cache_flash_bulk_cuda(keys: torch.Tensor, values: torch.Tensor,
                                       key_caches: List[torch.Tensor],
                                       value_caches: List[torch.Tensor],
                                       slot_mapping: torch.Tensor, kv_cache_dtype: str,
                                       k_scales: List[torch.Tensor], v_scales: List[torch.Tensor],
                                       num_heads: int, head_size: int) -> None:
    num_layers = len(key_caches)
    if num_layers == 0:                                               # kernels.cu:99-101
        return
    _check(num_layers == len(value_caches), "num_layers == value_caches.size()")   # kernels.cu:103-106
    _check(num_layers == len(k_scales), "num_layers == k_scales.size()")
    _check(num_layers == len(v_scales), "num_layers == v_scales.size()")
    _check(kv_cache_dtype in _KV_DTYPES, f"Unsupported data type of kv cache: {kv_cache_dtype}")
    _check(keys.is_cuda and values.is_cuda and slot_mapping.is_cuda,
           "reshape_and_cache_flash_bulk: CUDA tensors expected")
    if kv_cache_dtype == "auto":
        _check(keys.dtype == values.dtype == key_caches[0].dtype == value_caches[0].dtype,
               "keys/values/cache dtypes must match for kv_cache_dtype='auto'")
    else:
        _check(key_caches[0].element_size() == 1 and value_caches[0].element_size() == 1,
               "fp8 kv cache needs a one-byte cache dtype (uint8 / float8)")
    _check(slot_mapping.dtype == torch.int64, "expected scalar type Long for slot_mapping")
    num_tokens = slot_mapping.size(0)                                 # kernels.cu:108
    if num_tokens == 0:
        return
    _check(keys.dim() == 2 and values.dim() == 2 and keys.stride(1) == 1 and values.stride(1) == 1,
           "keys/values must be [num_tokens, num_layers*num_heads*head_size] with a dense last dim")
    _check(keys.size(0) >= num_tokens and values.size(0) >= num_tokens,
           "keys/values have fewer rows than slot_mapping")
    block_size = key_caches[0].size(1)                                # kernels.cu:109
    block_stride = key_caches[0].stride(0)                            # kernels.cu:113
    _check(block_stride == value_caches[0].stride(0), "block_stride == value_caches[0].stride(0)")
    if not slot_mapping.is_contiguous():
        slot_mapping = slot_mapping.contiguous()
    ptr_arr = ctypes.c_void_p * num_layers
    kptrs = ptr_arr(*[t.data_ptr() for t in key_caches])
    vptrs = ptr_arr(*[t.data_ptr() for t in value_caches])
    ksc = ptr_arr(*[t.data_ptr() for t in k_scales])
    vsc = ptr_arr(*[t.data_ptr() for t in v_scales])
    with torch.cuda.device(keys.device):                              # OptionalCUDAGuard, kernels.cu:132-133
        rc = _capi.lib().ab200_kv_write_bulk(
            keys.data_ptr(), values.data_ptr(), keys.stride(0), values.stride(0),
            kptrs, vptrs, ksc, vsc, num_layers, slot_mapping.data_ptr(), num_tokens,
            block_stride, block_size, num_heads, head_size,
            _capi.torch_dtype_code(keys.dtype), _KV_DTYPES[kv_cache_dtype],
            _capi.current_stream_ptr(keys.device))
    _capi.check(rc, "ab200_kv_write_bulk")


def _register() -> None:
    global _library, _registered
    if _registered:
        return
    _capi.lib()  # raises if the native library is not built
    lib = torch.library.Library("arctic_inference", "FRAGMENT")
    lib.define(OP_SCHEMA)
    lib.impl("reshape_and_cache_flash_bulk", _reshape_and_cache_flash_bulk_cuda, "CUDA")
    _library = lib
    _registered = True


def try_load_torch_library() -> bool:
    """Reference contract (py_custom_ops.py:9-37): return False instead of raising when the
    native library cannot be loaded, so the caller falls back to vLLM's per-layer op
    (llama_swiftkv.py:370-371, 629-656)."""
    try:
        _register()
        return True
    except RuntimeError as e:
        logger.info("Unable to load libarctic_b200 custom ops. RuntimeError: %s. "
                    "Falling back to original implementation.", e)
        return False
    except Exception as e:  # noqa: BLE001 - mirror of the reference's catch-all
        logger.info("Unable to load libarctic_b200 custom ops. Exception: %s. "
                    "Falling back to original implementation.", e)
        return False


def reshape_and_cache_flash_bulk(
    keys: torch.Tensor,
    values: torch.Tensor,
    key_caches: List[torch.Tensor],
    value_caches: List[torch.Tensor],
    slot_mapping: torch.Tensor,
    kv_cache_dtype: str,
    k_scales: List[torch.Tensor],
    v_scales: List[torch.Tensor],
    num_heads: int,
    head_size: int,
) -> None:
    torch.ops.arctic_inference.reshape_and_cache_flash_bulk(
        keys, values, key_caches, value_caches, slot_mapping, kv_cache_dtype,
        k_scales, v_scales, num_heads, head_size)

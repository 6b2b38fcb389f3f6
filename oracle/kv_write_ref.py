"""TEST INFRASTRUCTURE ONLY - oracle for hot path B (SwiftKV bulk reshape-and-cache).

The reference has no CPU implementation of this path (its fallback is vLLM's CUDA op,
arctic_inference/vllm/swiftkv/llama_swiftkv.py:629-656), so the oracle restates the CUDA kernel
/root/reference/csrc/custom_ops/kernels.cu:12-69 and its launcher's argument derivation
(kernels.cu:86-155):

* ``kv_write_numpy``  - vectorised numpy restatement (fancy-index scatter per layer).
* ``kv_write_c``      - ctypes call into oracle/kv_write_ref.c (element loop, as the kernel).
* ``launcher_args``   - the quantities the reference launcher derives from its tensors
                        (T from slot_mapping.size(0), block_size = key_caches[0].size(1), strides
                        from layer 0 only).

Caches are modelled as numpy views into a per-layer backing buffer, in vLLM's two layouts
(llama_swiftkv.py:610-619): FlashAttention ``[2, nb, bs, Hkv, Dh]`` and FlashInfer
``[nb, 2, bs, Hkv, Dh]``.  Element type is an opaque 2- or 4-byte word (uint16 / uint32): the
"auto" path is a bit copy.

Parity pinned: against tests/golden/kv_write_*.npz produced on a B200 by the reference kernel
itself (oracle/gen_golden_kv_gpu.py).  Nothing under arcticinference_b200/ may import this.
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "_build", "liboracle_kv.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.check_call(["make", "-C", HERE, "-s"])
        _LIB = ctypes.CDLL(path)
        _LIB.oracle_kv_write_bulk.restype = ctypes.c_int
        _LIB.oracle_kv_write_bulk.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
            ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p), ctypes.c_int32,
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
            ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32]
    return _LIB


def make_caches(L: int, nb: int, bs: int, Hkv: int, Dh: int, dtype=np.uint16, layout: str = "flash",
                fill: str = "zeros", seed: int = 0):
    """Per-layer backing buffers and the key/value views the caller passes
    (llama_swiftkv.py:610-619)."""
    rng = np.random.default_rng(seed + 17)
    backing, kc, vc = [], [], []
    for _ in range(L):
        shape = (2, nb, bs, Hkv, Dh) if layout == "flash" else (nb, 2, bs, Hkv, Dh)
        if fill == "zeros":
            buf = np.zeros(shape, dtype=dtype)
        else:
            buf = rng.integers(0, np.iinfo(dtype).max, size=shape, dtype=dtype)
        backing.append(buf)
        if layout == "flash":
            kc.append(buf[0]); vc.append(buf[1])
        else:
            kc.append(buf[:, 0]); vc.append(buf[:, 1])
    return backing, kc, vc


def make_kv_inputs(L: int, T: int, Hkv: int, Dh: int, bs: int = 16, dtype=np.uint16,
                   slots: str = "perm", frac_pad: float = 0.01, seed: int = 0,
                   extra_rows: int = 0, row_pad: int = 0):
    """Synthetic inputs: keys/values [T(+extra), L*n (+row_pad)] random words,
    unique slots (duplicates make the result order dependent), ~1 % of them -1."""
    rng = np.random.default_rng(seed)
    n = Hkv * Dh
    nb = 2 * ((T + bs - 1) // bs) if T else 2
    width = L * n + row_pad
    hi = np.iinfo(dtype).max
    keys = rng.integers(0, hi, size=(T + extra_rows, width), dtype=dtype)[:, :L * n]
    values = rng.integers(0, hi, size=(T + extra_rows, width), dtype=dtype)[:, :L * n]
    if slots == "perm":
        sm = rng.permutation(nb * bs)[:T].astype(np.int64)
    elif slots == "sequential":
        sm = (np.arange(T, dtype=np.int64) + bs * (nb // 4))
    else:
        raise ValueError(slots)
    if frac_pad > 0 and T:
        pad = rng.random(T) < frac_pad
        sm[pad] = -1
    return keys, values, sm, nb


def launcher_args(keys: np.ndarray, values: np.ndarray, key_caches: List[np.ndarray],
                  value_caches: List[np.ndarray], slot_mapping: np.ndarray):
    """kernels.cu:108-114: everything is taken from layer 0; strides in elements."""
    e = keys.itemsize
    num_tokens = slot_mapping.shape[0]
    block_size = key_caches[0].shape[1]
    key_stride = keys.strides[0] // e
    value_stride = values.strides[0] // e
    block_stride = key_caches[0].strides[0] // e
    if block_stride != value_caches[0].strides[0] // e:
        raise RuntimeError("block_stride == value_caches[0].stride(0)")  # TORCH_CHECK kernels.cu:114
    return num_tokens, block_size, key_stride, value_stride, block_stride


def kv_write_numpy(keys, values, key_caches, value_caches, slot_mapping, num_heads, head_size):
    """Vectorised restatement of kernels.cu:29-68 ("auto": bit copy)."""
    L = len(key_caches)
    if L == 0:
        return
    T, bs, *_ = launcher_args(keys, values, key_caches, value_caches, slot_mapping)
    n = num_heads * head_size
    sm = slot_mapping[:T]
    valid = sm >= 0
    s = sm[valid]
    blk, off = s // bs, s % bs
    rows = np.nonzero(valid)[0]
    for l in range(L):
        key_caches[l][blk, off] = keys[rows, l * n:(l + 1) * n].reshape(-1, num_heads, head_size)
        value_caches[l][blk, off] = values[rows, l * n:(l + 1) * n].reshape(-1, num_heads, head_size)


def kv_write_c(keys, values, key_caches, value_caches, slot_mapping, num_heads, head_size):
    """Element-loop C restatement (oracle/kv_write_ref.c)."""
    L = len(key_caches)
    if L == 0:
        return
    T, bs, ks, vs, bstride = launcher_args(keys, values, key_caches, value_caches, slot_mapping)
    kp = (ctypes.c_void_p * L)(*[c.ctypes.data for c in key_caches])
    vp = (ctypes.c_void_p * L)(*[c.ctypes.data for c in value_caches])
    sm = np.ascontiguousarray(slot_mapping, dtype=np.int64)
    rc = _lib().oracle_kv_write_bulk(keys.ctypes.data, values.ctypes.data, ks, vs, kp, vp, L,
                                     sm.ctypes.data, T, bstride, bs, num_heads, head_size,
                                     keys.itemsize)
    if rc != 0:
        raise RuntimeError(f"oracle_kv_write_bulk rc={rc}")


def algorithmic_bytes(num_valid_tokens: int, num_tokens: int, L: int, n: int, elem_bytes: int) -> int:
    """read K,V + write K,V for every valid (token, layer) + the slot map."""
    return num_valid_tokens * L * n * 2 * (2 * elem_bytes) + 8 * num_tokens

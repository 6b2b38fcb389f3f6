"""ctypes binding of libarctic_b200.so (C ABI declared in include/arctic_b200.h).

There is no fallback: if the library is missing or a call fails, a RuntimeError is raised."""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int32, c_int64, c_size_t, c_void_p, POINTER

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libarctic_b200.so")
DEBUG_LIB_PATH = os.path.join(HERE, "libarctic_b200_dbg.so")

AB200_F32, AB200_F16, AB200_BF16 = 0, 1, 2
AB200_KV_AUTO, AB200_KV_FP8_E4M3, AB200_KV_FP8_E5M2 = 0, 1, 2
AB200_ROPE_FP32, AB200_ROPE_STORAGE = 0, 1
AB200_E_INVALID, AB200_E_UNSUPPORTED, AB200_E_CUDA, AB200_E_WORKSPACE = -1, -2, -3, -4


class LstmSpeculatorDesc(ctypes.Structure):
    """struct ab200_lstm_speculator (include/arctic_b200.h)."""
    _fields_ = [
        ("dtype", c_int32), ("emulate_eager_rounding", c_int32),
        ("vocab", c_int64), ("head_rows", c_int64),
        ("input_hidden_dim", c_int64), ("inner_dim", c_int64),
        ("scale_input", c_int32), ("alpha", c_float), ("eps", c_float),
        ("forget_emb", c_void_p), ("proj0", c_void_p), ("proj1", c_void_p),
        ("cell_ln_w", c_void_p), ("cell_ln_b", c_void_p),
        ("state_ln_w", c_void_p), ("state_ln_b", c_void_p),
        ("head", c_void_p),
        ("head_q", c_void_p), ("head_scale", c_void_p), ("fp8_head_max_batch", c_int32),
    ]


AB200_MLP_MAX_STAGES = 16
_PtrArr = c_void_p * AB200_MLP_MAX_STAGES


class MlpSpeculatorDesc(ctypes.Structure):
    """struct ab200_mlp_speculator (include/arctic_b200.h)."""
    _fields_ = [
        ("dtype", c_int32), ("emulate_eager_rounding", c_int32),
        ("vocab", c_int64), ("head_rows", c_int64),
        ("input_hidden_dim", c_int64), ("inner_dim", c_int64),
        ("scale_input", c_int32), ("alpha", c_float), ("eps", c_float),
        ("emb", _PtrArr), ("proj", _PtrArr), ("ln_w", _PtrArr), ("ln_b", _PtrArr), ("head", _PtrArr),
        ("head_q", _PtrArr), ("head_scale", _PtrArr), ("fp8_head_max_batch", c_int32),
    ]


# name -> (restype, argtypes); mirrors include/arctic_b200.h one to one
SIGNATURES = {
    "ab200_last_error": (c_char_p, []),
    "ab200_abi_version": (c_int32, []),
    "ab200_launch_count": (c_int64, []),
    "ab200_init": (c_int32, [c_void_p]),
    "ab200_set_trace_buffer": (c_int32, [c_void_p, c_int32]),
    "ab200_kv_write_bulk": (c_int32, [c_void_p, c_void_p, c_int64, c_int64,
                                      POINTER(c_void_p), POINTER(c_void_p),
                                      POINTER(c_void_p), POINTER(c_void_p), c_int32,
                                      c_void_p, c_int64, c_int64, c_int32, c_int32, c_int32,
                                      c_int32, c_int32, c_void_p]),
    "ab200_kv_rope_write_bulk": (c_int32, [POINTER(c_void_p), c_int64, c_int64, c_void_p, c_void_p, c_int64, c_int64,
                                           c_int32, c_int32, c_int32, POINTER(c_void_p), POINTER(c_void_p),
                                           POINTER(c_void_p), POINTER(c_void_p), c_int32, c_void_p, c_int64,
                                           c_int64, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p,
                                           c_int64, c_void_p, c_void_p, c_int64, c_void_p]),
    "ab200_rms_scale": (c_int32, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64,
                                  c_float, c_float, c_int32, c_int32, c_void_p]),
    "ab200_sum_lstm_cell": (c_int32, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_int64, c_int64, c_int64, c_float, c_float, c_int32, c_int32,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ab200_linear": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                               c_int64, c_int64, c_int64, c_int32, c_void_p]),
    "ab200_linear_simt_bf16": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                                         c_int64, c_int64, c_int64, c_void_p]),
    "ab200_head_argmax": (c_int32, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64,
                                    c_int64, c_int64, c_void_p, c_void_p, c_int64, c_void_p,
                                    c_int32, c_void_p, c_size_t, c_void_p]),
    "ab200_fp8_quantize_weight": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                                            c_int64, c_void_p]),
    "ab200_fp8_quantize_rows": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                                          c_int64, c_void_p]),
    "ab200_linear_fp8": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                                   c_int64, c_int64, c_int64, c_int64, c_void_p]),
    "ab200_head_argmax_fp8": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                                        c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p,
                                        c_int64, c_void_p, c_void_p]),
    "ab200_argmax_shard_key": (c_int32, [c_void_p, c_int64, c_int64, c_void_p]),
    "ab200_argmax_unpack": (c_int32, [c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    "ab200_argmax_rows": (c_int32, [c_void_p, c_int64, c_int64, c_int64, c_int32, c_void_p,
                                    c_void_p]),
    "ab200_gemm_block_n": (c_int32, [c_int64, c_int32]),
    "ab200_lstm_workspace_bytes": (c_size_t, [POINTER(LstmSpeculatorDesc), c_int64, c_int32]),
    "ab200_lstm_propose": (c_int32, [POINTER(LstmSpeculatorDesc), c_void_p, c_void_p, c_int64,
                                     c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_size_t, c_void_p]),
    "ab200_select_hidden_rows": (c_int32, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int32, c_void_p,
                                           c_int64, c_void_p, c_int64, c_int64, c_void_p]),
    "ab200_gather_rows": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p,
                                    c_int64, c_int64, c_void_p]),
    "ab200_sum_mlp_cell": (c_int32, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_int64, c_int64, c_int64, c_float, c_float, c_int32,
                                     c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ab200_mlp_workspace_bytes": (c_size_t, [POINTER(MlpSpeculatorDesc), c_int64, c_int32]),
    "ab200_mlp_propose": (c_int32, [POINTER(MlpSpeculatorDesc), c_void_p, c_void_p, c_int64, c_int64,
                                    c_int32, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
}

# exported by the debug flavour only (libarctic_b200_dbg.so, -DAB200_DEBUG): tests/ and tools/
DEBUG_SIGNATURES = {
    "ab200_debug_cell_generic": (c_int32, [c_int32]),
    "ab200_debug_sigmoid_search": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_float, c_void_p]),
    "ab200_debug_unary_bf16": (c_int32, [c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_int64, c_void_p]),
    "ab200_debug_mlp_generic": (c_int32, [c_int32]),
    "ab200_debug_set_tunable": (c_int32, [c_char_p, c_int32]),
    "ab200_debug_max_clusters_of_4": (c_int32, []),
}

_lib = None
_debug_lib = None


def _load(path: str, signatures) -> ctypes.CDLL:
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing - build it with `python -m arcticinference_b200._build` "
            "(there is no CPU or PyTorch fallback for this path)")
    handle = ctypes.CDLL(path)
    for name, (res, args) in signatures.items():
        fn = getattr(handle, name)
        fn.restype = res
        fn.argtypes = args
    return handle


def lib() -> ctypes.CDLL:
    """Load (once) and return the native library. Raises RuntimeError if it is not built."""
    global _lib
    if _lib is None:
        _lib = _load(LIB_PATH, SIGNATURES)
    return _lib


def debug_lib() -> ctypes.CDLL:
    """The debug flavour (same kernels + ab200_debug_* hooks and settable launch tunables).  Test and
    tool infrastructure only: nothing in the product package calls this."""
    global _debug_lib
    if _debug_lib is None:
        _debug_lib = _load(DEBUG_LIB_PATH, {**SIGNATURES, **DEBUG_SIGNATURES})
    return _debug_lib


def init_device(device=None) -> None:
    """ab200_init on `device` (function attributes, constant tables): call before capturing graphs."""
    import torch
    with torch.cuda.device(device):
        check(lib().ab200_init(current_stream_ptr(device)), "ab200_init")


def check(rc: int, what: str = "", handle=None) -> None:
    if rc != 0:
        msg = (handle or lib()).ab200_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libarctic_b200 {what} failed (code {rc}): {msg}")


def torch_dtype_code(dtype) -> int:
    import torch
    try:
        return {torch.float32: AB200_F32, torch.float16: AB200_F16, torch.bfloat16: AB200_BF16}[dtype]
    except KeyError:
        raise RuntimeError(f"unsupported dtype {dtype} (expected float32, float16 or bfloat16)")


def current_stream_ptr(device=None) -> int:
    import torch
    return torch.cuda.current_stream(device).cuda_stream

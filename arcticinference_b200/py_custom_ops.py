"""Drop-in for the reference's ``arctic_inference/py_custom_ops.py`` (hot path B boundary) plus the
ADDITIVE torch ops of hot path A (the reference has no op for the draft head, so
these names are new in the ``arctic_inference`` namespace and collide with nothing).

Same two public names, same argument order, same behaviour contract:

* ``try_load_torch_library() -> bool``  never raises; True once
  ``torch.ops.arctic_inference.reshape_and_cache_flash_bulk`` is usable
  (reference: py_custom_ops.py:9-37).
* ``reshape_and_cache_flash_bulk(keys, values, key_caches, value_caches, slot_mapping,
  kv_cache_dtype, k_scales, v_scales, num_heads, head_size) -> None``
  (reference: py_custom_ops.py:40-54), forwarding to the torch op.

Path A ops (``torch.ops.arctic_inference.*``, CUDA dispatch key only, fake/meta kernels registered so
that ``torch.compile`` / ``torch.export`` consumers can trace them):
``speculator_rms_scale``, ``speculator_linear`` (alias ``linear_bf16``), ``sum_lstm_cell``,
``lm_head_argmax``, ``lm_head_argmax_fp8`` and the whole draft loop ``lstm_speculator_propose`` (out
variant: writes a caller-owned token buffer, so it can be captured into a CUDA graph over static
buffers).  ``ArcticLSTMSpeculator`` (speculator.py) is routed through them.

The bulk-KV op is defined with the reference's schema string, byte for byte
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


# ---------------------------------------------------------------------------------------------
# Hot path A: additive ops (no reference op exists - SURVEY F1; pattern of torch_bindings.cpp:5-19:
# one schema per op, CUDA key only)
# ---------------------------------------------------------------------------------------------
SPECULATOR_OP_SCHEMAS = {
    # y = x * rsqrt(mean(x^2, -1) + eps) * post_scale          (arctic_speculator.py:88-91, 656-657)
    "speculator_rms_scale": "speculator_rms_scale(Tensor x, float eps, float post_scale, "
                            "int emulate_eager_rounding=1) -> Tensor",
    # out = x @ w^T                                             (arctic_speculator.py:555-561, 668)
    "speculator_linear": "speculator_linear(Tensor x, Tensor w) -> Tensor",
    "linear_bf16": "linear_bf16(Tensor x, Tensor w) -> Tensor",
    # fused embedding + alpha-sum + gates + 2 x RMS-LN + 2 x GELU (arctic_speculator.py:667-689)
    "sum_lstm_cell": "sum_lstm_cell(Tensor gates, Tensor emb_table, Tensor tokens, Tensor? cell_prev, "
                     "Tensor w_cell, Tensor b_cell, Tensor w_state, Tensor b_state, float alpha, float eps, "
                     "int emulate_eager_rounding=1) -> (Tensor, Tensor)",
    # head projection + [:vocab] slice + argmax, logits never materialised
    # (logits_processor_opt.py:83-107, arctic_speculator.py:726-732)
    "lm_head_argmax": "lm_head_argmax(Tensor x, Tensor w, int vocab) -> Tensor",
    # FP8 qhead variant (arctic_speculator.py:455-465, 726-727; fp8.py:207-221, 276-308)
    "lm_head_argmax_fp8": "lm_head_argmax_fp8(Tensor x, Tensor wq, Tensor w_scale, int vocab) -> Tensor",
    # the whole S-step loop (arctic_speculator.py:706-751), out variant for CUDA-graph capture
    "lstm_speculator_propose": "lstm_speculator_propose(Tensor input_ids, Tensor hidden, Tensor forget_emb, "
                               "Tensor proj0, Tensor proj1, Tensor cell_ln_w, Tensor cell_ln_b, "
                               "Tensor state_ln_w, Tensor state_ln_b, Tensor head, Tensor? head_q, "
                               "Tensor? head_scale, int vocab, int num_steps, float alpha, float eps, "
                               "bool scale_input, int emulate_eager_rounding, int fp8_head_max_batch, "
                               "Tensor(a!) workspace, Tensor(b!) tokens_out, Tensor(c!)? states_out=None, "
                               "Tensor(d!)? cells_out=None) -> ()",
}


def _stream(t: torch.Tensor) -> int:
    return _capi.current_stream_ptr(t.device)


def _ptr(t):
    return t.data_ptr() if t is not None else None


def _rows2d(x: torch.Tensor) -> torch.Tensor:
    x2 = x.reshape(-1, x.shape[-1])
    return x2 if x2.stride(-1) == 1 else x2.contiguous()


def _speculator_rms_scale_cuda(x, eps, post_scale, emulate_eager_rounding=1):
    x2 = _rows2d(x)
    y = torch.empty(x2.shape, dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        _capi.check(_capi.lib().ab200_rms_scale(
            x2.data_ptr(), y.data_ptr(), x2.shape[0], x2.shape[1], x2.stride(0), y.stride(0), eps, post_scale,
            _capi.torch_dtype_code(x.dtype), int(emulate_eager_rounding), _stream(x)), "ab200_rms_scale")
    return y.view(x.shape)


def _speculator_linear_cuda(x, w):
    _check(x.dtype == w.dtype and x.shape[-1] == w.shape[1] and w.dim() == 2,
           f"speculator_linear: x {tuple(x.shape)} {x.dtype} does not match w {tuple(w.shape)} {w.dtype}")
    x2 = _rows2d(x)
    if w.stride(1) != 1:
        w = w.contiguous()
    out = torch.empty(x2.shape[0], w.shape[0], dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        _capi.check(_capi.lib().ab200_linear(
            x2.data_ptr(), w.data_ptr(), out.data_ptr(), x2.shape[0], w.shape[0], x2.shape[1], x2.stride(0),
            w.stride(0), out.stride(0), _capi.torch_dtype_code(x.dtype), _stream(x)), "ab200_linear")
    return out.view(*x.shape[:-1], w.shape[0])


def _linear_bf16_cuda(x, w):
    _check(x.dtype == torch.bfloat16, "linear_bf16: bf16 tensors expected")
    return _speculator_linear_cuda(x, w)


def _sum_lstm_cell_cuda(gates, emb_table, tokens, cell_prev, w_cell, b_cell, w_state, b_state, alpha, eps,
                        emulate_eager_rounding=1):
    g2 = _rows2d(gates)
    B, D = g2.shape[0], g2.shape[1] // 4
    _check(g2.shape[1] == 4 * D and emb_table.shape[1] == D, "sum_lstm_cell: gates must be [B, 4*D], emb_table [V, D]")
    toks = tokens.reshape(B).to(torch.long).contiguous()
    cprev = None if cell_prev is None else cell_prev.reshape(B, D).contiguous()
    state = torch.empty(B, D, dtype=gates.dtype, device=gates.device)
    cell = torch.empty(B, D, dtype=gates.dtype, device=gates.device)
    with torch.cuda.device(gates.device):
        _capi.check(_capi.lib().ab200_sum_lstm_cell(
            g2.data_ptr(), g2.stride(0), emb_table.data_ptr(), toks.data_ptr(), None, _ptr(cprev),
            w_cell.data_ptr(), b_cell.data_ptr(), w_state.data_ptr(), b_state.data_ptr(), state.data_ptr(),
            cell.data_ptr(), B, D, emb_table.shape[0], alpha, eps, _capi.torch_dtype_code(gates.dtype),
            int(emulate_eager_rounding), None, None, None, None, _stream(gates)), "ab200_sum_lstm_cell")
    return state, cell


def _lm_head_argmax_cuda(x, w, vocab):
    x2 = _rows2d(x)
    B, D = x2.shape
    # [B] arg-max slots + one completion counter, zero on entry (include/arctic_b200.h, ab200_head_argmax)
    packed = torch.zeros(B + 1, dtype=torch.int64, device=x.device)
    out = torch.empty(B, dtype=torch.long, device=x.device)
    ws = torch.empty(B * vocab, dtype=torch.float32, device=x.device) if x.dtype == torch.float32 else None
    with torch.cuda.device(x.device):
        _capi.check(_capi.lib().ab200_head_argmax(
            x2.data_ptr(), w.data_ptr(), B, w.shape[0], D, vocab, x2.stride(0), w.stride(0), packed.data_ptr(),
            out.data_ptr(), 1, packed[B:].data_ptr(), _capi.torch_dtype_code(x.dtype), _ptr(ws),
            ws.numel() * 4 if ws is not None else 0, _stream(x)), "ab200_head_argmax")
    return out


def _lm_head_argmax_fp8_cuda(x, wq, w_scale, vocab):
    _check(x.dtype == torch.bfloat16 and wq.element_size() == 1, "lm_head_argmax_fp8: bf16 x and e4m3 weights expected")
    x2 = _rows2d(x)
    B, D = x2.shape
    lib = _capi.lib()
    xq = torch.empty(B, D, dtype=torch.uint8, device=x.device)
    xs = torch.empty(B, dtype=torch.float32, device=x.device)
    packed = torch.zeros(B + 1, dtype=torch.int64, device=x.device)
    out = torch.empty(B, dtype=torch.long, device=x.device)
    with torch.cuda.device(x.device):
        st = _stream(x)
        _capi.check(lib.ab200_fp8_quantize_rows(x2.data_ptr(), xq.data_ptr(), xs.data_ptr(), B, D, x2.stride(0),
                                                xq.stride(0), st), "ab200_fp8_quantize_rows")
        _capi.check(lib.ab200_head_argmax_fp8(xq.data_ptr(), xs.data_ptr(), wq.data_ptr(), w_scale.data_ptr(), B,
                                              wq.shape[0], D, vocab, xq.stride(0), wq.stride(0), packed.data_ptr(),
                                              out.data_ptr(), 1, packed[B:].data_ptr(), st), "ab200_head_argmax_fp8")
    return out


def _lstm_speculator_propose_cuda(input_ids, hidden, forget_emb, proj0, proj1, cell_ln_w, cell_ln_b, state_ln_w,
                                  state_ln_b, head, head_q, head_scale, vocab, num_steps, alpha, eps, scale_input,
                                  emulate_eager_rounding, fp8_head_max_batch, workspace, tokens_out,
                                  states_out=None, cells_out=None):
    d = _capi.LstmSpeculatorDesc()
    d.dtype = _capi.torch_dtype_code(hidden.dtype)
    d.emulate_eager_rounding = int(emulate_eager_rounding)
    d.vocab, d.head_rows = vocab, head.shape[0]
    d.input_hidden_dim, d.inner_dim = proj0.shape[1], proj1.shape[1]
    d.scale_input, d.alpha, d.eps = int(scale_input), alpha, eps
    d.forget_emb, d.proj0, d.proj1 = forget_emb.data_ptr(), proj0.data_ptr(), proj1.data_ptr()
    d.cell_ln_w, d.cell_ln_b = cell_ln_w.data_ptr(), cell_ln_b.data_ptr()
    d.state_ln_w, d.state_ln_b = state_ln_w.data_ptr(), state_ln_b.data_ptr()
    d.head = head.data_ptr()
    d.head_q, d.head_scale = _ptr(head_q), _ptr(head_scale)
    d.fp8_head_max_batch = int(fp8_head_max_batch) if head_q is not None else 0
    B = input_ids.shape[0]
    _check(input_ids.dtype == torch.int64 and input_ids.is_contiguous(), "lstm_speculator_propose: input_ids must be contiguous int64")
    _check(hidden.dim() == 2 and hidden.shape[0] == B and hidden.stride(1) == 1, "lstm_speculator_propose: hidden must be [B, H], dense rows")
    _check(tokens_out.dtype == torch.int64 and tokens_out.is_contiguous() and tokens_out.numel() >= B * num_steps,
           "lstm_speculator_propose: tokens_out must be a contiguous int64 buffer of B * num_steps elements")
    with torch.cuda.device(hidden.device):
        rc = _capi.lib().ab200_lstm_propose(
            ctypes.byref(d), input_ids.data_ptr(), hidden.data_ptr(), hidden.stride(0), B, num_steps,
            tokens_out.data_ptr(), _ptr(states_out), _ptr(cells_out), workspace.data_ptr(), workspace.numel(),
            _stream(hidden))
    _capi.check(rc, "ab200_lstm_propose")


def lstm_speculator_workspace_bytes(dtype: torch.dtype, vocab: int, head_rows: int, input_hidden_dim: int,
                                    inner_dim: int, batch: int, num_steps: int) -> int:
    """Size of the ``workspace`` (uint8 CUDA tensor, 256-byte aligned) lstm_speculator_propose needs."""
    d = _capi.LstmSpeculatorDesc()
    d.dtype = _capi.torch_dtype_code(dtype)
    d.vocab, d.head_rows, d.input_hidden_dim, d.inner_dim = vocab, head_rows, input_hidden_dim, inner_dim
    return int(_capi.lib().ab200_lstm_workspace_bytes(ctypes.byref(d), batch, num_steps))


_SPECULATOR_IMPLS = {
    "speculator_rms_scale": _speculator_rms_scale_cuda, "speculator_linear": _speculator_linear_cuda,
    "linear_bf16": _linear_bf16_cuda, "sum_lstm_cell": _sum_lstm_cell_cuda,
    "lm_head_argmax": _lm_head_argmax_cuda, "lm_head_argmax_fp8": _lm_head_argmax_fp8_cuda,
    "lstm_speculator_propose": _lstm_speculator_propose_cuda,
}
# shape/dtype propagation for tracing (no data): what torch.compile / export need to see
_SPECULATOR_FAKES = {
    "speculator_rms_scale": lambda x, eps, post_scale, emulate_eager_rounding=1: torch.empty_like(x),
    "speculator_linear": lambda x, w: x.new_empty(*x.shape[:-1], w.shape[0]),
    "linear_bf16": lambda x, w: x.new_empty(*x.shape[:-1], w.shape[0]),
    "sum_lstm_cell": lambda gates, emb_table, tokens, cell_prev, w_cell, b_cell, w_state, b_state, alpha, eps,
    emulate_eager_rounding=1: (gates.new_empty(gates.numel() // gates.shape[-1], gates.shape[-1] // 4),
                               gates.new_empty(gates.numel() // gates.shape[-1], gates.shape[-1] // 4)),
    "lm_head_argmax": lambda x, w, vocab: x.new_empty(x.numel() // x.shape[-1], dtype=torch.long),
    "lm_head_argmax_fp8": lambda x, wq, w_scale, vocab: x.new_empty(x.numel() // x.shape[-1], dtype=torch.long),
    "lstm_speculator_propose": lambda *a, **k: None,
}


def _register() -> None:
    global _library, _registered
    if _registered:
        return
    _capi.lib()  # raises if the native library is not built
    lib = torch.library.Library("arctic_inference", "FRAGMENT")
    lib.define(OP_SCHEMA)
    lib.impl("reshape_and_cache_flash_bulk", _reshape_and_cache_flash_bulk_cuda, "CUDA")
    for name, schema in SPECULATOR_OP_SCHEMAS.items():
        lib.define(schema)
        lib.impl(name, _SPECULATOR_IMPLS[name], "CUDA")
        torch.library.register_fake(f"arctic_inference::{name}", _SPECULATOR_FAKES[name], lib=lib)
    _library = lib
    _registered = True


def load_speculator_ops():
    """Register the ops (raises without the native library - the speculator has no fallback) and
    return the ``torch.ops.arctic_inference`` namespace."""
    _register()
    return torch.ops.arctic_inference


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

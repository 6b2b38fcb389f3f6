"""One-call drop-in for an installed ArcticInference (the patches of INTEGRATION.md 1a and 2a, applied at run
time the way the reference itself patches vLLM).

* ``patch_py_custom_ops()``  - hot path B: ``arctic_inference.py_custom_ops.{try_load_torch_library,
  reshape_and_cache_flash_bulk}`` (arctic_inference/py_custom_ops.py:9-54) become ours: same names, same
  signatures, same torch op schema, backed by ``ab200_kv_write_bulk``.
* ``patch_lstm_speculator()`` - hot path A: ``ArcticLSTMSpeculator.generate_proposals``
  (arctic_inference/vllm/spec_dec/arctic_speculator.py:753-866) is routed through the native draft loop for the
  configurations the native classes cover (``method == "sum_lstm"`` or ``"sum_rnn"`` - the latter with single or
  stacked dims -, ``tp_size == 1``, bf16 / fp32, CUDA); every other
  configuration keeps running the reference's own code - that is the reference's path, not a fallback of ours.
  The native twin SHARES the module's parameter storage (``share_parameters``: no second copy of the weights; only the
  FP8 copy of the head is its own) and is rebuilt when the parameters move or ``load_weights`` ran again;
  ``module.qhead is None`` (the oracle recipe) disables the FP8 head on the twin as well.

Nothing here is needed by the native classes themselves; nothing here imports ``oracle/``.
"""
from __future__ import annotations

import functools
from typing import Optional

import torch

from . import py_custom_ops as _ours


def patch_py_custom_ops(module=None) -> bool:
    """Replace the two public names of ``arctic_inference.py_custom_ops``.  Returns False (and changes nothing)
    when the native library cannot be loaded - the reference then keeps its own loader's outcome."""
    if module is None:
        import arctic_inference.py_custom_ops as module  # type: ignore
    if not _ours.try_load_torch_library():
        return False
    module.try_load_torch_library = _ours.try_load_torch_library
    module.reshape_and_cache_flash_bulk = _ours.reshape_and_cache_flash_bulk
    return True


def _native_twin(module) -> "object":
    from .speculator import ArcticLSTMSpeculator as Native
    params = dict(module.named_parameters())
    stamp = tuple((k, v.data_ptr()) for k, v in params.items())
    twin = getattr(module, "_b200_impl", None)
    if twin is not None and getattr(module, "_b200_stamp", None) == stamp:
        return twin
    if getattr(module, "method", None) == "sum_rnn":
        # :476-536, :693-704: dims are lists on the module; same parameter names, storage shared
        from .sum_rnn_speculator import ArcticSumRNNSpeculator
        p0 = params["proj.0.0.weight"]
        twin = ArcticSumRNNSpeculator(
            vocab_size=module.vocab_size, input_hidden_dim=module.input_hidden_dim, emb_dim=list(module.emb_dim),
            proj_dim=list(module.proj_dim), inner_dim=list(module.inner_dim), n_predict=module.n_predict,
            num_lookahead_tokens=module.max_speculative_tokens, tie_weights=module.tie_weights,
            max_num_seqs=getattr(module, "cuda_graph_max_batch_size", 256), dtype=p0.dtype, device=p0.device,
            scale_input=module.scale_input, enforce_eager=not getattr(module, "cuda_graph_mode", True),
            quantize_lm_head=None if getattr(module, "qhead", None) is not None else False)
        twin.share_parameters((k, v.data) for k, v in params.items() if not k.startswith("qhead"))
        module._b200_impl, module._b200_stamp = twin, stamp
        return twin
    p0 = params["projs.0.weight"]
    twin = Native(vocab_size=module.vocab_size, input_hidden_dim=module.input_hidden_dim,
                  inner_dim=module.inner_dim[-1], n_predict=module.n_predict,
                  num_lookahead_tokens=module.max_speculative_tokens,
                  max_num_seqs=getattr(module, "cuda_graph_max_batch_size", 256), dtype=p0.dtype, device=p0.device,
                  scale_input=module.scale_input, enforce_eager=not getattr(module, "cuda_graph_mode", True),
                  quantize_lm_head=None if getattr(module, "qhead", None) is not None else False,
                  allocate_parameters=False)
    # same names / shapes as the reference registers them (head rows = vocab padded to 64): shared, not copied;
    # qhead.* is rebuilt from head.*
    twin.share_parameters((k, v.data) for k, v in params.items() if not k.startswith("qhead"))
    module._b200_impl, module._b200_stamp = twin, stamp
    return twin


def patch_lstm_speculator(cls=None):
    """Wrap ``cls.generate_proposals`` (default: the installed reference's ``ArcticLSTMSpeculator``).  Idempotent;
    returns the class."""
    if cls is None:
        from arctic_inference.vllm.spec_dec.arctic_speculator import ArcticLSTMSpeculator as cls  # type: ignore
    if getattr(cls.generate_proposals, "_b200_patched", False):
        return cls
    original = cls.generate_proposals

    @functools.wraps(original)
    def generate_proposals(self, input_ids: torch.Tensor, previous_hidden_states: torch.Tensor,
                           num_predict_tokens: int):
        if (getattr(self, "method", None) in ("sum_lstm", "sum_rnn") and getattr(self, "tp_size", 1) == 1
                and previous_hidden_states.is_cuda and previous_hidden_states.dtype in (torch.bfloat16, torch.float32)):
            return _native_twin(self).generate_proposals(input_ids, previous_hidden_states, num_predict_tokens)
        return original(self, input_ids, previous_hidden_states, num_predict_tokens)

    generate_proposals._b200_patched = True
    generate_proposals._b200_original = original
    cls.generate_proposals = generate_proposals
    if hasattr(cls, "load_weights"):
        original_load = cls.load_weights

        @functools.wraps(original_load)
        def load_weights(self, *a, **k):
            out = original_load(self, *a, **k)
            self._b200_impl = None            # new values in the same storage: rebuild the native twin
            self._b200_stamp = None
            return out

        load_weights._b200_original = original_load
        cls.load_weights = load_weights
    return cls


def unpatch_lstm_speculator(cls) -> None:
    gp = cls.generate_proposals
    if getattr(gp, "_b200_patched", False):
        cls.generate_proposals = gp._b200_original
    lw = getattr(cls, "load_weights", None)
    if lw is not None and hasattr(lw, "_b200_original"):
        cls.load_weights = lw._b200_original


def apply(py_custom_ops: bool = True, lstm_speculator: bool = True) -> dict:
    """Patch whatever of the reference is importable; returns what was done."""
    done = {}
    if py_custom_ops:
        try:
            done["py_custom_ops"] = patch_py_custom_ops()
        except ImportError:
            done["py_custom_ops"] = False
    if lstm_speculator:
        try:
            patch_lstm_speculator()
            done["lstm_speculator"] = True
        except ImportError:
            done["lstm_speculator"] = False
    return done

"""Host-side mirror of the reference's ``ArcticLSTMSpeculator`` (hot path A boundary).

The reference has no custom op for the draft head; its drop-in boundary is the
Python method boundary of
``arctic_inference/vllm/spec_dec/arctic_speculator.py::ArcticLSTMSpeculator``:

* ``generate_proposals(input_ids[B], previous_hidden_states[B,H], num_predict_tokens) -> [B,S] int64``
  (:753-866; ``ValueError`` when S > max_speculative_tokens, :759-762)
* ``generate_states(last_tokens[B,1], previous_hidden_states[B,1,K], head_index, cell_states[B,1,D])
  -> (state, cell)`` (:648-691)
* ``generate_token_ids(batch_size, num_predict_tokens, last_tokens, previous_hidden_states,
  next_tokens_tensors, cell_states)`` (:706-751)
* parameter names/shapes as the reference loads them (:874-902): ``forget_emb.0.weight [V,D]``,
  ``projs.0.weight [4D,H]``, ``projs.1.weight [4D,D]``, ``cell_ln.0.{weight,bias} [D]``,
  ``state_ln.0.{weight,bias} [D]``, ``head.0.weight [V_pad64, D]``; ``load_weights`` fuses
  ``forget/input/output/cell_proj`` in the gate order f, i, o, c.

Everything numeric runs in libarctic_b200.so (hand-written sm_100a kernels) through the C ABI;
PyTorch only owns device memory, the stream and the CUDA-graph capture.  There is no PyTorch or
CPU fallback: without the native library construction raises.

Built: ``method="sum_lstm"`` with ``tie_weights`` / ``tie_lstm_embs`` (the configuration the
reference's hot path asserts, :545, :663); the FP8 ``qhead`` for padded batches <= 32 (:726-727)
as an option (``quantize_lm_head=True``; pinned against vLLM 0.22's fp8 ops); a
vocab-sharded head for ``tp_size > 1`` (:733-744) whose per-step exchange is one MAX all-reduce.
``method="sum_rnn"`` / ``ArcticMLPSpeculator`` live in ``mlp_speculator.py``.
"""
from __future__ import annotations

import collections
import ctypes
import math
from typing import Dict, Iterable, List, Optional, Tuple

import torch

from . import _capi, py_custom_ops

SQRT2 = 2 ** 0.5  # arctic_speculator.py:36


def padding_size(size: int) -> int:
    """arctic_speculator.py:39-44 - pad a batch to a multiple of pow2ceil(size)//4 so that few
    CUDA graphs cover all batch sizes (1,2,3,4,6,8,12,16,24,32,48,64,96,128,...)."""
    mult = (1 << (size - 1).bit_length()) // 4
    if mult < 1:
        return size
    return (size + mult - 1) // mult * mult


def pad_vocab_size(vocab_size: int, pad_to: int = 64) -> int:
    """vocab_parallel_embedding.py pad_vocab_size (DEFAULT_VOCAB_PADDING_SIZE = 64)."""
    return (vocab_size + pad_to - 1) // pad_to * pad_to


class ArcticLSTMSpeculator:
    """B200-native Arctic LSTM speculator draft head (see module docstring)."""

    def __init__(self, *, vocab_size: int, input_hidden_dim: int, inner_dim: int, n_predict: int,
                 num_lookahead_tokens: Optional[int] = None, max_num_seqs: int = 256,
                 dtype: torch.dtype = torch.bfloat16, device: str | torch.device = "cuda",
                 scale_input: bool = True, enforce_eager: bool = False,
                 emulate_eager_rounding: bool | int = True, alpha_flavor: str = "cuda",
                 quantize_lm_head: Optional[bool] = None, fp8_head_max_batch: int = 32,
                 tp_size: int = 1, tp_rank: int = 0, tp_group=None, allocate_parameters: bool = True) -> None:
        self._lib = _capi.lib()  # fail loudly if the native library is missing
        # path A goes through the additive torch ops (arctic_inference::sum_lstm_cell, ::lm_head_argmax,
        # ::lstm_speculator_propose, ... - py_custom_ops.py), which call the C ABI
        self._ops = py_custom_ops.load_speculator_ops()
        # Tensor parallelism of the vocab head (the reference shards ParallelLMHead over
        # max(TP, SP) ranks, vocab_parallel_embedding.py:20-35, and merges local top-1s with an
        # all_gather, arctic_speculator.py:733-744).  Everything else is replicated.
        if tp_size < 1 or not 0 <= tp_rank < tp_size:
            raise ValueError("bad tp_size / tp_rank")
        self.tp_size, self.tp_rank, self.tp_group = tp_size, tp_rank, tp_group
        if dtype not in (torch.bfloat16, torch.float32):
            raise ValueError("dtype must be torch.bfloat16 or torch.float32")
        if inner_dim % 8 or input_hidden_dim % 8:
            raise ValueError("inner_dim and input_hidden_dim must be multiples of 8")
        self.device = torch.device(device)
        if self.device.type == "cuda":
            # function attributes + constant tables of this device, before any graph capture
            _capi.init_device(self.device)
        self.dtype = dtype
        self.n_predict = n_predict
        self.vocab_size = vocab_size
        self.input_hidden_dim = input_hidden_dim
        self.inner_dim = [inner_dim]
        self.emb_dim = [inner_dim]
        self.proj_dim = [inner_dim]
        self.max_speculative_tokens = num_lookahead_tokens or n_predict
        self.tie_weights = True
        self.tie_lstm_embs = True
        self.scale_input = scale_input
        self.method = "sum_lstm"
        # 0 / False: fp32 internally; 1 / True: bit-exact eager emulation; 2: "faithful" (same rounding
        # points, unguarded fast sigmoid / GELU) - include/arctic_b200.h, ab200_sum_lstm_cell
        self.emulate_eager_rounding = 0 if dtype == torch.float32 else int(emulate_eager_rounding)
        if self.emulate_eager_rounding not in (0, 1, 2):
            raise ValueError("emulate_eager_rounding must be 0, 1 or 2")
        # arctic_speculator.py:575-577
        self.state_weight = 0.5 ** (0.5 / n_predict)
        self.emb_weight = math.sqrt((1 - self.state_weight ** 2) * (inner_dim / 2))
        alpha = torch.tensor(self.emb_weight / self.state_weight, dtype=torch.float32)
        if alpha_flavor == "cpu" and dtype != torch.float32:
            # ATen's CPU add kernel casts alpha to the tensor dtype; the CUDA kernel keeps fp32
            alpha = alpha.to(dtype).to(torch.float32)
        elif alpha_flavor not in ("cpu", "cuda"):
            raise ValueError("alpha_flavor must be 'cuda' or 'cpu'")
        self.alpha = float(alpha)
        self.eps = 1e-6
        # FP8 duplicate of the head, used iff the (padded) batch <= 32 - the reference's `qhead`.
        # The reference hard-codes `self.quantize_lm_head = True` (arctic_speculator.py:435) and picks
        # `qhead` for batch_size <= 32 (:726-727), so that is the default here too (None = "as the
        # reference": on wherever it is built - bf16, D % 16 == 0, tp_size == 1).  Its arithmetic is
        # pinned to vLLM 0.22's fp8 ops, not to the reference's vLLM 0.9.2 binary (DESIGN.md section
        # 2); the bf16-oracle recipe of SURVEY App. C (`model.qhead = None`) works here unchanged.
        fp8_ok = dtype == torch.bfloat16 and inner_dim % 16 == 0 and tp_size == 1 and self.device.type == "cuda"
        if quantize_lm_head and tp_size > 1:
            raise ValueError("the FP8 head is not built for tp_size > 1")
        self.quantize_lm_head = fp8_ok if quantize_lm_head is None else (bool(quantize_lm_head) and fp8_ok)
        self.fp8_head_max_batch = fp8_head_max_batch
        self._head_q: Optional[torch.Tensor] = None
        self._head_scale: Optional[torch.Tensor] = None
        # bumped whenever something baked into captured graphs changes (cached plans check it)
        self._graph_epoch = 0

        D, H, V = inner_dim, input_hidden_dim, vocab_size
        Vp = pad_vocab_size(V)
        if Vp % tp_size:
            raise ValueError("padded vocab size must be divisible by tp_size")
        self.vocab_padded = Vp
        self.vocab_per_rank = Vp // tp_size          # num_embeddings_per_partition
        # allocate_parameters=False: zero-size placeholders of the right rank; share_parameters() adopts the
        # caller's tensors (a drop-in next to the reference module must not hold a second copy of the weights)
        mk = (lambda *shape: torch.zeros(*shape, dtype=dtype, device=self.device)) if allocate_parameters else \
            (lambda *shape: torch.zeros(*([0] * len(shape)), dtype=dtype, device=self.device))
        self._param_shapes = {"head.0.weight": (self.vocab_per_rank, D), "forget_emb.0.weight": (V, D),
                              "projs.0.weight": (4 * D, H), "projs.1.weight": (4 * D, D), "cell_ln.0.weight": (D,),
                              "cell_ln.0.bias": (D,), "state_ln.0.weight": (D,), "state_ln.0.bias": (D,)}
        self._params: Dict[str, torch.Tensor] = collections.OrderedDict([
            ("head.0.weight", mk(self.vocab_per_rank, D)),
            ("forget_emb.0.weight", mk(V, D)),
            ("projs.0.weight", mk(4 * D, H)),
            ("projs.1.weight", mk(4 * D, D)),
            ("cell_ln.0.weight", mk(D)), ("cell_ln.0.bias", mk(D)),
            ("state_ln.0.weight", mk(D)), ("state_ln.0.bias", mk(D)),
        ])

        # static buffers (arctic_speculator.py:590-610)
        self.cuda_graph_max_batch_size = padding_size(max_num_seqs)
        self.cuda_graph_mode = not enforce_eager
        self.cuda_graphs: Dict[Tuple[int, int], torch.cuda.CUDAGraph] = {}
        n = self.cuda_graph_max_batch_size
        self.static_cuda_buffers = {
            "last_tokens": torch.zeros(n, dtype=torch.long, device=self.device),
            "previous_hidden_states": torch.zeros(n, H, dtype=dtype, device=self.device),
            "next_tokens": torch.zeros(n, self.max_speculative_tokens, dtype=torch.long,
                                       device=self.device),
        }
        self._workspace: Optional[torch.Tensor] = None
        self._desc: Optional[_capi.LstmSpeculatorDesc] = None
        if self.quantize_lm_head:
            # allocated ONCE: captured graphs bake these pointers in, so re-loading weights must
            # re-quantise into the same buffers
            self._head_q = torch.zeros(self.vocab_per_rank, D, dtype=torch.uint8, device=self.device)
            self._head_scale = torch.ones(1, dtype=torch.float32, device=self.device)

    # ---- parameters ---------------------------------------------------------------------------
    def named_parameters(self):
        return list(self._params.items())

    @property
    def qhead(self):
        """The FP8 head ([V_pad, D] e4m3 as uint8, scale [1] f32) or None.  Assigning None disables it
        (the reference's oracle recipe `model.qhead = None`: bf16 head at every batch size)."""
        return None if self._head_q is None else (self._head_q, self._head_scale)

    @qhead.setter
    def qhead(self, value) -> None:
        if value is not None:
            raise ValueError("qhead can only be disabled (set to None); it is rebuilt from head.0.weight")
        self.quantize_lm_head = False
        self._head_q = self._head_scale = None
        self._invalidate_graphs()

    def _invalidate_graphs(self) -> None:
        """Drop everything that baked device pointers / kernel choices in: descriptor, graphs, plans."""
        self._desc = None
        self.cuda_graphs.clear()
        self._graph_epoch += 1

    def uses_fp8_head(self, batch_size: int) -> bool:
        """The reference's selection rule (arctic_speculator.py:726-727): `qhead` iff it exists and the
        batch the loop runs on (the PADDED batch under CUDA graphs, :631-640) is <= 32."""
        b = padding_size(batch_size) if (self.cuda_graph_mode and batch_size <= self.cuda_graph_max_batch_size) \
            else batch_size
        return self._head_q is not None and b <= self.fp8_head_max_batch

    def share_parameters(self, named_tensors: Iterable[Tuple[str, torch.Tensor]]) -> List[str]:
        """Adopt the caller's tensors as this model's parameters WITHOUT copying (same names as ``load_weights``
        after gate fusing: ``head.0.weight [V_pad64, D]``, ``forget_emb.0.weight``, ``projs.{0,1}.weight``, the four
        LayerNorm vectors).  A tensor is shared when it already has the exact shape, dtype, device, a dense row-major
        layout and 16-byte alignment; anything else is copied like ``load_weights`` would.  Later in-place updates
        of a shared tensor are seen by the bf16 path at once; call ``requantize()`` (or ``load_weights``) afterwards
        to refresh the FP8 copy of the head.  Returns the names that were shared."""
        shared = []
        for name, t in named_tensors:
            name = name.replace("speculator.", "")
            if name not in self._params:
                continue
            ok = (tuple(t.shape) == tuple(self._param_shapes[name]) and t.dtype == self.dtype and t.device == self.device
                  and t.is_contiguous() and t.data_ptr() % 16 == 0)
            if ok:
                self._params[name] = t.detach()
                shared.append(name)
            else:
                if self._params[name].numel() == 0:
                    self._params[name] = torch.zeros(*self._param_shapes[name], dtype=self.dtype, device=self.device)
                self._set_param(name, t.to(device=self.device, dtype=self.dtype))
        missing = [n for n, p in self._params.items() if p.numel() == 0]
        if missing:
            raise ValueError(f"share_parameters: no tensor given for {missing}")
        self._invalidate_graphs()
        self.requantize()
        return shared

    def requantize(self) -> None:
        """Refresh the FP8 copy of the head from ``head.0.weight`` (in place: captured graphs stay valid)."""
        if self.quantize_lm_head and self._head_q is not None:
            self._quantize_head()

    def _set_param(self, name: str, value: torch.Tensor) -> None:
        p = self._params[name]
        if p.numel() == 0 and name in getattr(self, "_param_shapes", {}):      # allocate_parameters=False: allocate on demand
            p = self._params[name] = torch.zeros(*self._param_shapes[name], dtype=self.dtype, device=self.device)
        if name.startswith("head") and self.tp_size > 1 and value.shape[0] != p.shape[0]:
            # full (unsharded) head given: keep this rank's vocab partition
            lo = self.tp_rank * self.vocab_per_rank
            part = value[lo:lo + self.vocab_per_rank]
            p.zero_()
            p[:part.shape[0]].copy_(part)
            return
        if tuple(p.shape) != tuple(value.shape):
            if name.startswith("head") and value.shape[0] <= p.shape[0] and value.shape[1:] == p.shape[1:]:
                p.zero_()                       # vocab-padding rows stay zero
                p[:value.shape[0]].copy_(value)
                return
            raise ValueError(f"shape mismatch for {name}: {tuple(value.shape)} vs {tuple(p.shape)}")
        p.copy_(value)

    def load_weights(self, weights: Iterable[Tuple[str, torch.Tensor]]) -> None:
        """arctic_speculator.py:874-902: drop the tied embeddings, fuse forget/input/output/cell
        projections into ``projs.N.weight`` (order f, i, o, c), strip the ``speculator.`` prefix."""
        weights = collections.OrderedDict((k.replace("speculator.", ""), v) for k, v in weights)
        for tied in ("input_emb.0.weight", "cell_emb.0.weight", "output_emb.0.weight"):
            weights.pop(tied, None)
        for name in ("projs.0.weight", "projs.1.weight"):
            parts = [name.replace("projs", p) for p in ("forget_proj", "input_proj", "output_proj", "cell_proj")]
            if all(p in weights for p in parts):
                weights[name] = torch.cat([weights.pop(p) for p in parts])
        for name, w in weights.items():
            if name.startswith("qhead"):
                continue  # rebuilt from head.0.weight below (the reference loads head into qhead, :899-902)
            if name in self._params:
                self._set_param(name, w.to(device=self.device, dtype=self.dtype))
        self._desc = None
        if self.quantize_lm_head:
            self._quantize_head()

    def _quantize_head(self) -> None:
        """process_weights_after_loading of the reference's qhead (fp8.py:207-221): dynamic
        per-tensor e4m3 quantisation of the (padded) head weight - IN PLACE into the buffers the
        captured graphs point at, so a second load_weights() keeps every graph valid."""
        head = self._params["head.0.weight"]
        with torch.cuda.device(self.device):
            _capi.check(self._lib.ab200_fp8_quantize_weight(
                head.data_ptr(), self._head_q.data_ptr(), self._head_scale.data_ptr(), head.shape[0],
                head.shape[1], head.stride(0), self._head_q.stride(0),
                _capi.current_stream_ptr(self.device)), "ab200_fp8_quantize_weight")

    # ---- native descriptor / workspace -------------------------------------------------------
    def _descriptor(self) -> _capi.LstmSpeculatorDesc:
        if self._desc is None:
            p = self._params
            d = _capi.LstmSpeculatorDesc()
            d.dtype = _capi.torch_dtype_code(self.dtype)
            d.emulate_eager_rounding = int(self.emulate_eager_rounding)
            d.vocab = self.vocab_size
            d.head_rows = p["head.0.weight"].shape[0]
            d.input_hidden_dim = self.input_hidden_dim
            d.inner_dim = self.inner_dim[0]
            d.scale_input = int(self.scale_input)
            d.alpha = self.alpha
            d.eps = self.eps
            d.forget_emb = p["forget_emb.0.weight"].data_ptr()
            d.proj0 = p["projs.0.weight"].data_ptr()
            d.proj1 = p["projs.1.weight"].data_ptr()
            d.cell_ln_w = p["cell_ln.0.weight"].data_ptr()
            d.cell_ln_b = p["cell_ln.0.bias"].data_ptr()
            d.state_ln_w = p["state_ln.0.weight"].data_ptr()
            d.state_ln_b = p["state_ln.0.bias"].data_ptr()
            d.head = p["head.0.weight"].data_ptr()
            d.head_q = self._head_q.data_ptr() if self._head_q is not None else None
            d.head_scale = self._head_scale.data_ptr() if self._head_scale is not None else None
            d.fp8_head_max_batch = self.fp8_head_max_batch if self._head_q is not None else 0
            self._desc = d
        return self._desc

    def _get_workspace(self, batch: int, steps: int) -> torch.Tensor:
        need = self._lib.ab200_lstm_workspace_bytes(ctypes.byref(self._descriptor()), batch, steps)
        if self._workspace is None or self._workspace.numel() < need:
            if self.cuda_graphs:
                # a batch above the graph limit (eager path): the captured graphs keep pointing at
                # the original workspace, this call gets its own
                return torch.empty(need, dtype=torch.uint8, device=self.device)
            biggest = self._lib.ab200_lstm_workspace_bytes(
                ctypes.byref(self._descriptor()),
                max(batch, self.cuda_graph_max_batch_size), max(steps, self.max_speculative_tokens))
            self._workspace = torch.empty(max(need, biggest), dtype=torch.uint8, device=self.device)
        return self._workspace

    def _propose_into(self, ids: torch.Tensor, hidden: torch.Tensor, out: torch.Tensor, batch: int,
                      steps: int, states: Optional[torch.Tensor] = None,
                      cells: Optional[torch.Tensor] = None) -> None:
        ws = self._get_workspace(batch, steps)
        p = self._params
        self._ops.lstm_speculator_propose(
            ids[:batch], hidden[:batch], p["forget_emb.0.weight"], p["projs.0.weight"], p["projs.1.weight"],
            p["cell_ln.0.weight"], p["cell_ln.0.bias"], p["state_ln.0.weight"], p["state_ln.0.bias"],
            p["head.0.weight"], self._head_q, self._head_scale, self.vocab_size, steps, self.alpha, self.eps,
            self.scale_input, int(self.emulate_eager_rounding), self.fp8_head_max_batch, ws, out, states, cells)

    # ---- reference method boundary ------------------------------------------------------------
    def generate_states(self, last_tokens: torch.Tensor, previous_hidden_states: torch.Tensor,
                        head_index: int, cell_states: Optional[torch.Tensor] = None
                        ) -> Tuple[torch.Tensor, torch.Tensor]:
        """arctic_speculator.py:648-691 - one step of ln0 (step 0) -> gate GEMM -> fused cell."""
        ops, p = self._ops, self._params
        B = last_tokens.shape[0]
        x = previous_hidden_states.reshape(B, -1)
        emu = int(self.emulate_eager_rounding)
        if head_index == 0 and self.scale_input:                         # :656-657
            inv_sqrt2 = float(torch.tensor(1.0, dtype=torch.float32) / torch.tensor(SQRT2, dtype=torch.float32))
            x = ops.speculator_rms_scale(x, self.eps, inv_sqrt2, emu)
        proj = p["projs.0.weight"] if head_index == 0 else p["projs.1.weight"]   # :659-660
        if x.shape[1] != proj.shape[1]:
            raise RuntimeError(f"hidden width {x.shape[1]} does not match projs weight {tuple(proj.shape)}")
        gates = ops.speculator_linear(x, proj)                           # :668
        state, cell = ops.sum_lstm_cell(gates, p["forget_emb.0.weight"], last_tokens, cell_states,
                                        p["cell_ln.0.weight"], p["cell_ln.0.bias"], p["state_ln.0.weight"],
                                        p["state_ln.0.bias"], self.alpha, self.eps, emu)   # :667-689
        return state.unsqueeze(1), cell.unsqueeze(1)

    def greedy_tokens(self, states: torch.Tensor) -> torch.Tensor:
        """Head projection + arg-max for states [B, D] (logits_processor_opt.py:83-107,
        arctic_speculator.py:726-732) -> [B] int64."""
        if self.tp_size > 1:
            keys = self._all_reduce_max(self.local_argmax_keys(states), states)
            toks = torch.empty(states.shape[0], dtype=torch.long, device=self.device)
            with torch.cuda.device(self.device):
                _capi.check(self._lib.ab200_argmax_unpack(keys.data_ptr(), toks.data_ptr(), states.shape[0], 1,
                                                          _capi.current_stream_ptr(self.device)), "ab200_argmax_unpack")
            return toks
        return self._ops.lm_head_argmax(states, self._params["head.0.weight"], self.vocab_size)

    # ---- tensor-parallel head -----------------------------------------------------------------
    def local_argmax_keys(self, states: torch.Tensor) -> torch.Tensor:
        """Arg-max of this rank's vocab partition as all-reduce keys [B] int64 (global token index,
        signed order = (logit, lowest index) order).  Like the reference (:733-736, logits are only
        cut at org_vocab_size) the partition's zero-weight padding rows take part."""
        lib, head = self._lib, self._params["head.0.weight"]
        B, D = states.shape
        states = states.contiguous()
        packed = torch.zeros(B, dtype=torch.int64, device=self.device)
        code = _capi.torch_dtype_code(self.dtype)
        per = self.vocab_per_rank
        n_valid = per                                      # the whole partition
        ws = torch.empty(B * n_valid, dtype=torch.float32, device=self.device) if self.dtype == torch.float32 else None
        stream = _capi.current_stream_ptr(self.device)
        with torch.cuda.device(self.device):
            _capi.check(lib.ab200_head_argmax(states.data_ptr(), head.data_ptr(), B, head.shape[0], D, n_valid,
                                              states.stride(0), head.stride(0), packed.data_ptr(), None, 0, None,
                                              code, ws.data_ptr() if ws is not None else None,
                                              ws.numel() * 4 if ws is not None else 0, stream), "ab200_head_argmax")
            _capi.check(lib.ab200_argmax_shard_key(packed.data_ptr(), B, self.tp_rank * per, stream),
                        "ab200_argmax_shard_key")
        return packed

    def _all_reduce_max(self, keys: torch.Tensor, states: torch.Tensor) -> torch.Tensor:
        g = self.tp_group
        if hasattr(g, "all_reduce_max"):           # test double: see tests/test_tp_gpu.py
            return g.all_reduce_max(keys, states)
        import torch.distributed as dist
        dist.all_reduce(keys, op=dist.ReduceOp.MAX, group=g)
        return keys

    def _propose_tp(self, input_ids: torch.Tensor, previous_hidden_states: torch.Tensor, B: int, S: int) -> torch.Tensor:
        """generate_token_ids with tp_size > 1 (arctic_speculator.py:706-751, branch :733-744): one MAX
        all-reduce of [B] int64 keys per step instead of topk + all_gather of [2B,1] + argmax + gather."""
        tokens = input_ids.to(device=self.device, dtype=torch.long).reshape(B, 1)
        prev = previous_hidden_states.to(device=self.device, dtype=self.dtype).reshape(B, 1, -1)
        cell = None
        out = torch.empty(B, S, dtype=torch.long, device=self.device)
        for s in range(S):
            state, cell = self.generate_states(tokens, prev, s, cell)
            keys = self._all_reduce_max(self.local_argmax_keys(state.reshape(B, -1)), state.reshape(B, -1))
            with torch.cuda.device(self.device):
                _capi.check(self._lib.ab200_argmax_unpack(keys.data_ptr(), out[:, s].data_ptr(), B, S,
                                                          _capi.current_stream_ptr(self.device)), "ab200_argmax_unpack")
            tokens, prev = out[:, s:s + 1], state
        return out

    def generate_token_ids(self, batch_size: int, num_predict_tokens: int, last_tokens: torch.Tensor,
                           previous_hidden_states: torch.Tensor,
                           next_tokens_tensors: List[Optional[torch.Tensor]],
                           cell_states: Optional[torch.Tensor] = None) -> List[torch.Tensor]:
        """arctic_speculator.py:706-751, step by step through the per-op entry points."""
        for head_index in range(num_predict_tokens):
            states, cell_states = self.generate_states(last_tokens, previous_hidden_states,
                                                       head_index, cell_states)
            previous_hidden_states = states
            if self._head_q is not None and batch_size <= self.fp8_head_max_batch and self.tp_size == 1:
                # :726-727 - the FP8 qhead for batch <= 32 (same rule as inside the fused loop)
                last_tokens = self._ops.lm_head_argmax_fp8(states.flatten(0, 1), self._head_q, self._head_scale,
                                                           self.vocab_size).reshape(batch_size, -1)
            else:
                last_tokens = self.greedy_tokens(states.flatten(0, 1)).reshape(batch_size, -1)
            if next_tokens_tensors[head_index] is None:
                next_tokens_tensors[head_index] = last_tokens
            else:
                next_tokens_tensors[head_index].copy_(last_tokens)
        return next_tokens_tensors

    def generate_proposals(self, input_ids: torch.Tensor, previous_hidden_states: torch.Tensor,
                           num_predict_tokens: int) -> torch.Tensor:
        """arctic_speculator.py:753-866: static buffers, batch padded to ``padding_size(B)``, one
        CUDA graph per (padded batch, S) holding the whole S-step loop (3S+1 kernels, PDL-chained)."""
        if num_predict_tokens > self.max_speculative_tokens:
            raise ValueError(f"Max speculative tokens for model is "
                             f"{self.max_speculative_tokens}, but "
                             f"{num_predict_tokens} were requested")
        B = input_ids.size(0)
        S = num_predict_tokens
        if B == 0:
            return torch.empty(0, S, dtype=torch.long, device=self.device)
        if previous_hidden_states.shape != (B, self.input_hidden_dim):
            raise RuntimeError(f"previous_hidden_states must be [{B}, {self.input_hidden_dim}]")
        if self.tp_size > 1:
            return self._propose_tp(input_ids, previous_hidden_states, B, S)
        use_graph = self.cuda_graph_mode and B <= self.cuda_graph_max_batch_size
        if not use_graph:
            ids = input_ids.to(device=self.device, dtype=torch.long).contiguous()
            hid = previous_hidden_states.to(device=self.device, dtype=self.dtype)
            if hid.stride(1) != 1:
                hid = hid.contiguous()
            out = torch.empty(B, S, dtype=torch.long, device=self.device)
            self._propose_into(ids, hid, out, B, S)
            return out
        return self._replay(input_ids, previous_hidden_states, B, S).clone()   # :862-866

    def _replay(self, input_ids: torch.Tensor, previous_hidden_states: torch.Tensor, B: int,
                S: int) -> torch.Tensor:
        """Copy the inputs (device or host tensors) into the static buffers, replay the graph of the
        padded batch and return a VIEW of the static output buffer (valid until the next call)."""
        bufs = self.static_cuda_buffers
        # inputs that already live in the static buffers (a caller that fills them in place) need no copy
        if input_ids.data_ptr() != bufs["last_tokens"].data_ptr():
            bufs["last_tokens"][:B].copy_(input_ids, non_blocking=True)                        # :625
        if previous_hidden_states.data_ptr() != bufs["previous_hidden_states"].data_ptr():
            bufs["previous_hidden_states"][:B].copy_(previous_hidden_states, non_blocking=True)   # :629
        padded = padding_size(B)                                       # :631
        key = (padded, S)
        g = self.cuda_graphs.get(key)
        out_view = bufs["next_tokens"].view(-1)[:padded * S].view(padded, S)
        if g is None:
            self._get_workspace(self.cuda_graph_max_batch_size, self.max_speculative_tokens)
            # warm-up on a side stream (first-use attribute setting must not happen in capture)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                self._propose_into(bufs["last_tokens"], bufs["previous_hidden_states"], out_view, padded, S)
            torch.cuda.current_stream(self.device).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._propose_into(bufs["last_tokens"], bufs["previous_hidden_states"], out_view, padded, S)
            self.cuda_graphs[key] = g
        g.replay()
        return out_view[:B]


class ArcticProposer:
    """Mirror of ``arctic_inference/vllm/spec_dec/arctic_proposer.py::ArcticProposer.propose``
    (:149-166): host token ids -> device, ``generate_proposals``, device -> host numpy."""

    def __init__(self, model: ArcticLSTMSpeculator) -> None:
        self.model = model
        self.device = model.device
        self.input_hidden_dim = model.input_hidden_dim
        n = model.cuda_graph_max_batch_size
        # pinned staging: ids go host -> static device buffer in ONE async copy, tokens come back
        # in one; no intermediate device tensors (the reference does .to(device) / .cpu(), :160-166)
        self._ids_pinned = torch.empty(n, dtype=torch.long).pin_memory()
        self._out_pinned = torch.empty(n * model.max_speculative_tokens, dtype=torch.long).pin_memory()
        self._plans: Dict[Tuple[int, int], tuple] = {}

    def prepare_hidden_states(self, sample_hidden_states: torch.Tensor, sampled_token_ids: torch.Tensor,
                              num_draft_tokens: Optional[List[int]]) -> torch.Tensor:
        """arctic_proposer.py:113-147: keep, for every request, the hidden state of its last accepted
        token.  ``sampled_token_ids`` [R, max_gen_len] int64 (-1 = rejected / padding),
        ``num_draft_tokens`` = spec_decode_metadata.num_draft_tokens (host list).  One launch."""
        if sample_hidden_states is not None:
            assert sample_hidden_states.shape[-1] == self.input_hidden_dim, \
                (f"hidden_states shape mismatch: {sample_hidden_states.shape[-1]} != {self.input_hidden_dim}. "
                 "Please make sure spec model is trained using the same base model.")
        max_gen_len = sampled_token_ids.shape[-1]
        if max_gen_len == 1:                                              # :134-136
            return sample_hidden_states
        assert num_draft_tokens is not None
        # the reference passes spec_decode_metadata (arctic_proposer.py:113-119) and reads its
        # .num_draft_tokens (:139); a plain host list is accepted as well
        num_draft_tokens = getattr(num_draft_tokens, "num_draft_tokens", num_draft_tokens)
        R = sampled_token_ids.shape[0]
        h = sample_hidden_states
        if h.stride(-1) != 1:
            h = h.contiguous()
        ids = sampled_token_ids.to(device=self.device, dtype=torch.long)
        if ids.stride(-1) != 1:
            ids = ids.contiguous()
        counts = torch.as_tensor(num_draft_tokens, dtype=torch.long) + 1
        offs = (torch.cumsum(counts, 0) - counts).to(self.device, non_blocking=True)   # exclusive cumsum
        out = torch.empty(R, h.shape[1], dtype=h.dtype, device=self.device)
        with torch.cuda.device(self.device):
            _capi.check(_capi.lib().ab200_select_hidden_rows(
                h.data_ptr(), h.stride(0) * h.element_size(), h.shape[0], ids.data_ptr(), ids.stride(0),
                max_gen_len, offs.data_ptr(), R, out.data_ptr(), out.stride(0) * out.element_size(),
                h.shape[1] * h.element_size(), _capi.current_stream_ptr(self.device)), "ab200_select_hidden_rows")
        return out

    def propose(self, context_token_ids, previous_hidden_states: torch.Tensor, num_predict_tokens: int):
        assert num_predict_tokens > 0, \
            f"num_predict_tokens must be greater than 0, got {num_predict_tokens}."
        m = self.model
        S = num_predict_tokens
        # Steady state first: the hidden-state H2D copy is the long pole of the call (0.5 MB over PCIe), so it is
        # enqueued before ANY other host work; everything below overlaps with it.
        plan = self._plans.get((previous_hidden_states.shape[0], S)) if previous_hidden_states.dim() == 2 else None
        if plan is not None and plan[-1] != m._graph_epoch:
            self._plans.clear()              # graphs were invalidated (e.g. qhead switched off): re-plan
            plan = None
        if plan is not None and previous_hidden_states.shape[1] == m.input_hidden_dim:
            hid_dev, ids_dev, ids_pin, ids_np, graph, out_dev, out, out_np, _ = plan
            hid_dev.copy_(previous_hidden_states, non_blocking=True)
            ids = torch.as_tensor(context_token_ids)
            if ids.dim() == 1 and ids.shape[0] == hid_dev.shape[0]:
                if ids.device.type == "cpu":
                    ids_np[:] = ids.numpy()
                    ids_dev.copy_(ids_pin, non_blocking=True)
                else:
                    ids_dev.copy_(ids, non_blocking=True)
                graph.replay()
                out.copy_(out_dev, non_blocking=True)
                torch.cuda.current_stream(self.device).synchronize()
                return out_np.copy()
        ids = torch.as_tensor(context_token_ids)
        B = ids.shape[0]
        fast = (m.cuda_graph_mode and 0 < B <= m.cuda_graph_max_batch_size and S <= m.max_speculative_tokens
                and ids.dim() == 1 and previous_hidden_states.shape == (B, m.input_hidden_dim))
        if not fast:
            input_ids = ids.to(self.device, non_blocking=True)
            if previous_hidden_states.device != self.device:
                previous_hidden_states = previous_hidden_states.to(self.device, non_blocking=True)
            return m.generate_proposals(input_ids=input_ids, previous_hidden_states=previous_hidden_states,
                                        num_predict_tokens=S).cpu().numpy()
        if getattr(m, "tp_size", 1) == 1:
            # Steady state (graph captured): every tensor view is cached per (B, S), so a call is two H2D
            # copies straight into the graph's static buffers (the big one first), one replay, one D2H
            # and a stream sync - no Python-side slicing on the critical path.
            plan = self._plans.get((B, S))
            if plan is None:
                bufs = m.static_cuda_buffers
                m._replay(bufs["last_tokens"][:B], bufs["previous_hidden_states"][:B], B, S)   # capture if needed
                padded = padding_size(B)
                graph = m.cuda_graphs[(padded, S)]
                out_dev = bufs["next_tokens"].view(-1)[:padded * S].view(padded, S)[:B]
                out = self._out_pinned[:B * S].view(B, S)
                plan = (bufs["previous_hidden_states"][:B], bufs["last_tokens"][:B], self._ids_pinned[:B],
                        self._ids_pinned.numpy()[:B], graph, out_dev, out, out.numpy(), m._graph_epoch)
                self._plans[(B, S)] = plan
            hid_dev, ids_dev, ids_pin, ids_np, graph, out_dev, out, out_np, _ = plan
            hid_dev.copy_(previous_hidden_states, non_blocking=True)
            if ids.device.type == "cpu":
                ids_np[:] = ids.numpy()
                ids_dev.copy_(ids_pin, non_blocking=True)
            else:
                ids_dev.copy_(ids, non_blocking=True)
            graph.replay()
            out.copy_(out_dev, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
            return out_np.copy()
        out_dev = m.generate_proposals(ids.to(self.device), previous_hidden_states.to(self.device), S)
        out = self._out_pinned[:B * S].view(B, S)
        out.copy_(out_dev, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return out.numpy().copy()

"""Host-side mirror of the reference's ``ArcticMLPSpeculator`` - also the
``method="sum_rnn"`` configuration of ``ArcticLSTMSpeculator`` with single-stage dims (that class's parameter names
``emb.i.0.weight`` ... and its stacked / list-valued dims are handled by ``sum_rnn_speculator.ArcticSumRNNSpeculator``,
which delegates to this class when every stack has one stage).

Reference: ``arctic_inference/vllm/spec_dec/arctic_speculator.py::ArcticMLPSpeculator``
(:102-401): ``generate_states(last_tokens[B,1], previous_hidden_states[B,1,K], head_index)``
(:265-285), ``generate_token_ids`` (:287-326), ``generate_proposals(input_ids[B],
previous_hidden_states[B,emb_dim], num_predict_tokens) -> [B,S] int64`` (:328-389, ``ValueError``
above ``max_speculative_tokens``), ``load_weights`` (:397-401).  Parameter names / shapes as the
reference registers them: ``emb.i.weight [V_pad64, D]``, ``proj.i.weight [D, emb_dim | D]``,
``ln.i.{weight,bias} [D]``, ``head.i.weight [V_pad64, D]``; with ``tie_weights`` only ``emb.0``,
``ln.0``, ``head.0`` and ``proj.0`` / ``proj.1`` exist (:139-175).

All arithmetic runs in libarctic_b200.so (``ab200_mlp_propose`` and the per-op entry points);
there is no PyTorch or CPU fallback.
"""
from __future__ import annotations

import collections
import ctypes
import math
from typing import Dict, Iterable, List, Optional, Tuple

import torch

from . import _capi
from .speculator import SQRT2, pad_vocab_size, padding_size


class ArcticMLPSpeculator:
    def __init__(self, *, vocab_size: int, emb_dim: int, inner_dim: int, n_predict: int,
                 num_lookahead_tokens: Optional[int] = None, tie_weights: bool = True,
                 max_num_seqs: int = 256, dtype: torch.dtype = torch.bfloat16,
                 device: str | torch.device = "cuda", scale_input: bool = True,
                 enforce_eager: bool = False, emulate_eager_rounding: bool | int = True,
                 alpha_flavor: str = "cuda", quantize_lm_head: Optional[bool] = None,
                 fp8_head_max_batch: int = 32) -> None:
        self._lib = _capi.lib()
        if dtype not in (torch.bfloat16, torch.float32):
            raise ValueError("dtype must be torch.bfloat16 or torch.float32")
        inner_dim = inner_dim if inner_dim != 0 else emb_dim           # :121
        if inner_dim % 8 or emb_dim % 8:
            raise ValueError("inner_dim and emb_dim must be multiples of 8")
        self.device = torch.device(device)
        if self.device.type == "cuda":
            _capi.init_device(self.device)   # function attributes + constant tables, before any capture
        self.dtype = dtype
        self.n_predict = n_predict
        self.vocab_size = vocab_size
        self.emb_dim = emb_dim
        self.input_hidden_dim = emb_dim
        self.inner_dim = inner_dim
        self.max_speculative_tokens = num_lookahead_tokens or n_predict
        if self.max_speculative_tokens > _capi.AB200_MLP_MAX_STAGES:
            raise ValueError(f"at most {_capi.AB200_MLP_MAX_STAGES} speculative tokens are supported")
        self.tie_weights = tie_weights
        if tie_weights and n_predict <= 1:
            raise AssertionError("You cannot tie weights between stages when only 1 exists")   # :133-136
        self.scale_input = scale_input
        self.emulate_eager_rounding = 0 if dtype == torch.float32 else int(emulate_eager_rounding)
        self.state_weight = 0.5 ** (0.5 / n_predict)                   # :213-215
        self.emb_weight = math.sqrt((1 - self.state_weight ** 2) * (inner_dim / 2))
        alpha = torch.tensor(self.emb_weight / self.state_weight, dtype=torch.float32)
        if alpha_flavor == "cpu" and dtype != torch.float32:
            alpha = alpha.to(dtype).to(torch.float32)
        elif alpha_flavor not in ("cpu", "cuda"):
            raise ValueError("alpha_flavor must be 'cuda' or 'cpu'")
        self.alpha = float(alpha)
        self.eps = 1e-6
        # The reference hard-codes quantize_lm_head = True and builds `qhead` for tied weights only
        # (arctic_speculator.py:129, 162-173), used for batch <= 32 (:298-299): None = exactly that.
        fp8_ok = dtype == torch.bfloat16 and inner_dim % 16 == 0 and self.device.type == "cuda"
        self.quantize_lm_head = (fp8_ok and tie_weights) if quantize_lm_head is None \
            else (bool(quantize_lm_head) and fp8_ok)
        self.fp8_head_max_batch = fp8_head_max_batch
        self._graph_epoch = 0

        D, H, V, n = inner_dim, emb_dim, vocab_size, self.max_speculative_tokens
        Vp = pad_vocab_size(V)
        mk = lambda *shape: torch.zeros(*shape, dtype=dtype, device=self.device)
        self._params: Dict[str, torch.Tensor] = collections.OrderedDict()
        n_sets = 1 if tie_weights else n
        n_proj = min(2, n) if tie_weights else n
        for i in range(n_sets):
            self._params[f"emb.{i}.weight"] = mk(Vp, D)
        for i in range(n_proj):
            self._params[f"proj.{i}.weight"] = mk(D, H if i == 0 else D)
        for i in range(n_sets):
            self._params[f"head.{i}.weight"] = mk(Vp, D)
        for i in range(n_sets):
            self._params[f"ln.{i}.weight"] = mk(D)
            self._params[f"ln.{i}.bias"] = mk(D)
        self._head_q: Dict[int, torch.Tensor] = {}
        self._head_scale: Dict[int, torch.Tensor] = {}

        self.cuda_graph_max_batch_size = padding_size(max_num_seqs)
        self.cuda_graph_mode = not enforce_eager
        self.cuda_graphs: Dict[Tuple[int, int], torch.cuda.CUDAGraph] = {}
        nb = self.cuda_graph_max_batch_size
        self.static_cuda_buffers = {
            "last_tokens": torch.zeros(nb, dtype=torch.long, device=self.device),
            "previous_hidden_states": torch.zeros(nb, H, dtype=dtype, device=self.device),
            "next_tokens": torch.zeros(nb, n, dtype=torch.long, device=self.device),
        }
        self._workspace: Optional[torch.Tensor] = None
        self._desc: Optional[_capi.MlpSpeculatorDesc] = None
        if self.quantize_lm_head:
            # allocated ONCE: captured graphs bake these pointers in (re-quantised in place on reload)
            for i in range(n_sets):
                self._head_q[i] = torch.zeros(Vp, D, dtype=torch.uint8, device=self.device)
                self._head_scale[i] = torch.ones(1, dtype=torch.float32, device=self.device)

    @property
    def qhead(self):
        """FP8 heads {index: ([V_pad, D] e4m3 as uint8, scale [1])} or None; assigning None disables
        them (the bf16-oracle recipe `model.qhead = None`)."""
        return None if not self._head_q else {i: (q, self._head_scale[i]) for i, q in self._head_q.items()}

    @qhead.setter
    def qhead(self, value) -> None:
        if value is not None:
            raise ValueError("qhead can only be disabled (set to None); it is rebuilt from head.N.weight")
        self.quantize_lm_head = False
        self._head_q, self._head_scale = {}, {}
        self._desc = None
        self.cuda_graphs.clear()
        self._graph_epoch += 1

    # ---- parameters ---------------------------------------------------------------------------
    def named_parameters(self):
        return list(self._params.items())

    def _idx(self, head_index: int, proj: bool = False) -> int:
        if not self.tie_weights:
            return head_index
        return min(head_index, 1) if proj else 0

    def load_weights(self, weights: Iterable[Tuple[str, torch.Tensor]]) -> None:
        """arctic_speculator.py:397-401 (qhead.* is rebuilt from head.* here, :399-401)."""
        for name, w in weights:
            name = name.replace("speculator.", "")
            if name.startswith("qhead") or name not in self._params:
                continue
            p = self._params[name]
            w = w.to(device=self.device, dtype=self.dtype)
            if tuple(w.shape) != tuple(p.shape):
                if w.dim() == 2 and w.shape[0] <= p.shape[0] and w.shape[1] == p.shape[1] and \
                        (name.startswith("head") or name.startswith("emb")):
                    p.zero_()                     # vocab padding rows stay zero
                    p[:w.shape[0]].copy_(w)
                    continue
                raise ValueError(f"shape mismatch for {name}: {tuple(w.shape)} vs {tuple(p.shape)}")
            p.copy_(w)
        self._desc = None
        if self.quantize_lm_head:
            self._quantize_heads()

    def share_parameters(self, named_tensors: Iterable[Tuple[str, torch.Tensor]]) -> List[str]:
        """Adopt the caller's tensors as parameters WITHOUT copying where they already have the exact shape, dtype,
        device, a dense layout and 16-byte alignment (anything else is copied like ``load_weights``); captured graphs
        are dropped, the FP8 heads re-quantised in place.  Returns the names that were shared."""
        shared, rest = [], []
        for name, t in named_tensors:
            name = name.replace("speculator.", "")
            if name.startswith("qhead") or name not in self._params:
                continue
            p = self._params[name]
            if (tuple(t.shape) == tuple(p.shape) and t.dtype == self.dtype and t.device == p.device
                    and t.is_contiguous() and t.data_ptr() % 16 == 0):
                self._params[name] = t.detach()
                shared.append(name)
            else:
                rest.append((name, t))
        self._desc = None
        self.cuda_graphs.clear()
        self._graph_epoch += 1
        if rest:
            self.load_weights(rest)
        elif self.quantize_lm_head:
            self._quantize_heads()
        return shared

    def _quantize_heads(self) -> None:
        """IN PLACE into the buffers allocated at construction: graphs captured before a second
        load_weights() keep reading valid, freshly quantised heads."""
        for i in sorted(self._head_q):
            head = self._params[f"head.{i}.weight"]
            q, s = self._head_q[i], self._head_scale[i]
            with torch.cuda.device(self.device):
                _capi.check(self._lib.ab200_fp8_quantize_weight(
                    head.data_ptr(), q.data_ptr(), s.data_ptr(), head.shape[0], head.shape[1],
                    head.stride(0), q.stride(0), _capi.current_stream_ptr(self.device)),
                    "ab200_fp8_quantize_weight")

    # ---- native descriptor / workspace -------------------------------------------------------
    def _descriptor(self) -> _capi.MlpSpeculatorDesc:
        if self._desc is None:
            p = self._params
            d = _capi.MlpSpeculatorDesc()
            d.dtype = _capi.torch_dtype_code(self.dtype)
            d.emulate_eager_rounding = int(self.emulate_eager_rounding)
            d.vocab = self.vocab_size
            d.head_rows = p["head.0.weight"].shape[0]
            d.input_hidden_dim = self.emb_dim
            d.inner_dim = self.inner_dim
            d.scale_input = int(self.scale_input)
            d.alpha = self.alpha
            d.eps = self.eps
            for s in range(self.max_speculative_tokens):
                i, ip = self._idx(s), self._idx(s, True)
                d.emb[s] = p[f"emb.{i}.weight"].data_ptr()
                d.proj[s] = p[f"proj.{ip}.weight"].data_ptr()
                d.ln_w[s] = p[f"ln.{i}.weight"].data_ptr()
                d.ln_b[s] = p[f"ln.{i}.bias"].data_ptr()
                d.head[s] = p[f"head.{i}.weight"].data_ptr()
                d.head_q[s] = self._head_q[i].data_ptr() if i in self._head_q else None
                d.head_scale[s] = self._head_scale[i].data_ptr() if i in self._head_scale else None
            d.fp8_head_max_batch = self.fp8_head_max_batch if self._head_q else 0
            self._desc = d
        return self._desc

    def _get_workspace(self, batch: int, steps: int) -> torch.Tensor:
        need = self._lib.ab200_mlp_workspace_bytes(ctypes.byref(self._descriptor()), batch, steps)
        if self._workspace is None or self._workspace.numel() < need:
            if self.cuda_graphs:
                # a batch above the graph limit (eager path): the captured graphs keep pointing at
                # the original workspace, this call gets its own
                return torch.empty(need, dtype=torch.uint8, device=self.device)
            biggest = self._lib.ab200_mlp_workspace_bytes(
                ctypes.byref(self._descriptor()), max(batch, self.cuda_graph_max_batch_size),
                max(steps, self.max_speculative_tokens))
            self._workspace = torch.empty(max(need, biggest), dtype=torch.uint8, device=self.device)
        return self._workspace

    def _propose_into(self, ids: torch.Tensor, hidden: torch.Tensor, out: torch.Tensor, batch: int,
                      steps: int, states: Optional[torch.Tensor] = None) -> None:
        ws = self._get_workspace(batch, steps)
        with torch.cuda.device(self.device):
            rc = self._lib.ab200_mlp_propose(
                ctypes.byref(self._descriptor()), ids.data_ptr(), hidden.data_ptr(), hidden.stride(0),
                batch, steps, out.data_ptr(), states.data_ptr() if states is not None else None,
                ws.data_ptr(), ws.numel(), _capi.current_stream_ptr(self.device))
        _capi.check(rc, "ab200_mlp_propose")

    # ---- reference method boundary ------------------------------------------------------------
    def generate_states(self, last_tokens: torch.Tensor, previous_hidden_states: torch.Tensor,
                        head_index: int) -> torch.Tensor:
        """arctic_speculator.py:265-285, one step through the per-op entry points."""
        lib, p = self._lib, self._params
        B = last_tokens.shape[0]
        D = self.inner_dim
        x = previous_hidden_states.reshape(B, -1)
        if not x.is_contiguous():
            x = x.contiguous()
        code = _capi.torch_dtype_code(self.dtype)
        emu = int(self.emulate_eager_rounding)
        stream = _capi.current_stream_ptr(self.device)
        with torch.cuda.device(self.device):
            if head_index == 0 and self.scale_input:
                y = torch.empty_like(x)
                inv_sqrt2 = float(torch.tensor(1.0, dtype=torch.float32) / torch.tensor(SQRT2, dtype=torch.float32))
                _capi.check(lib.ab200_rms_scale(x.data_ptr(), y.data_ptr(), B, x.shape[1], x.stride(0),
                                                y.stride(0), self.eps, inv_sqrt2, code, int(emu != 0), stream),
                            "ab200_rms_scale")
                x = y
            proj = p[f"proj.{self._idx(head_index, True)}.weight"]
            if x.shape[1] != proj.shape[1]:
                raise RuntimeError(f"hidden width {x.shape[1]} does not match proj weight {tuple(proj.shape)}")
            pre = torch.empty(B, D, dtype=self.dtype, device=self.device)
            _capi.check(lib.ab200_linear(x.data_ptr(), proj.data_ptr(), pre.data_ptr(), B, D, x.shape[1],
                                         x.stride(0), proj.stride(0), pre.stride(0), code, stream), "ab200_linear")
            i = self._idx(head_index)
            state = torch.empty(B, D, dtype=self.dtype, device=self.device)
            toks = last_tokens.reshape(B).to(torch.long).contiguous()
            _capi.check(lib.ab200_sum_mlp_cell(
                pre.data_ptr(), pre.stride(0), p[f"emb.{i}.weight"].data_ptr(), toks.data_ptr(), None,
                p[f"ln.{i}.weight"].data_ptr(), p[f"ln.{i}.bias"].data_ptr(), state.data_ptr(), B, D,
                self.vocab_size, self.alpha, self.eps, code, emu, None, None, None, None, stream),
                "ab200_sum_mlp_cell")
        return state.unsqueeze(1)

    def greedy_tokens(self, states: torch.Tensor, head_index: int = 0) -> torch.Tensor:
        lib, head = self._lib, self._params[f"head.{self._idx(head_index)}.weight"]
        B, D = states.shape
        states = states.contiguous()
        packed = torch.zeros(B + 1, dtype=torch.int64, device=self.device)
        out = torch.empty(B, dtype=torch.long, device=self.device)
        code = _capi.torch_dtype_code(self.dtype)
        ws = torch.empty(B * self.vocab_size, dtype=torch.float32, device=self.device) \
            if self.dtype == torch.float32 else None
        with torch.cuda.device(self.device):
            _capi.check(lib.ab200_head_argmax(states.data_ptr(), head.data_ptr(), B, head.shape[0], D,
                                              self.vocab_size, states.stride(0), head.stride(0),
                                              packed.data_ptr(), out.data_ptr(), 1, packed[B:].data_ptr(), code,
                                              ws.data_ptr() if ws is not None else None,
                                              ws.numel() * 4 if ws is not None else 0,
                                              _capi.current_stream_ptr(self.device)), "ab200_head_argmax")
        return out

    def generate_token_ids(self, batch_size: int, num_predict_tokens: int, last_tokens: torch.Tensor,
                           previous_hidden_states: torch.Tensor,
                           next_tokens_tensors: List[Optional[torch.Tensor]]) -> List[torch.Tensor]:
        """arctic_speculator.py:287-326."""
        for head_index in range(num_predict_tokens):
            states = self.generate_states(last_tokens, previous_hidden_states, head_index)
            previous_hidden_states = states
            last_tokens = self.greedy_tokens(states.flatten(0, 1), head_index).reshape(batch_size, -1)
            if next_tokens_tensors[head_index] is None:
                next_tokens_tensors[head_index] = last_tokens
            else:
                next_tokens_tensors[head_index].copy_(last_tokens)
        return next_tokens_tensors

    def generate_proposals(self, input_ids: torch.Tensor, previous_hidden_states: torch.Tensor,
                           num_predict_tokens: int) -> torch.Tensor:
        """arctic_speculator.py:328-389."""
        if num_predict_tokens > self.max_speculative_tokens:
            raise ValueError(f"Max speculative tokens for model is "
                             f"{self.max_speculative_tokens}, but "
                             f"{num_predict_tokens} were requested")
        B, S = input_ids.size(0), num_predict_tokens
        if B == 0:
            return torch.empty(0, S, dtype=torch.long, device=self.device)
        if previous_hidden_states.shape != (B, self.emb_dim):
            raise RuntimeError(f"previous_hidden_states must be [{B}, {self.emb_dim}]")
        if not (self.cuda_graph_mode and B <= self.cuda_graph_max_batch_size):
            ids = input_ids.to(device=self.device, dtype=torch.long).contiguous()
            hid = previous_hidden_states.to(device=self.device, dtype=self.dtype)
            if hid.stride(1) != 1:
                hid = hid.contiguous()
            out = torch.empty(B, S, dtype=torch.long, device=self.device)
            self._propose_into(ids, hid, out, B, S)
            return out
        return self._replay(input_ids, previous_hidden_states, B, S).clone()

    def _replay(self, input_ids: torch.Tensor, previous_hidden_states: torch.Tensor, B: int,
                S: int) -> torch.Tensor:
        bufs = self.static_cuda_buffers
        if input_ids.data_ptr() != bufs["last_tokens"].data_ptr():
            bufs["last_tokens"][:B].copy_(input_ids, non_blocking=True)
        if previous_hidden_states.data_ptr() != bufs["previous_hidden_states"].data_ptr():
            bufs["previous_hidden_states"][:B].copy_(previous_hidden_states, non_blocking=True)
        padded = padding_size(B)
        key = (padded, S)
        g = self.cuda_graphs.get(key)
        out_view = bufs["next_tokens"].view(-1)[:padded * S].view(padded, S)
        if g is None:
            self._get_workspace(self.cuda_graph_max_batch_size, self.max_speculative_tokens)
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

"""Host-side mirror of ``ArcticLSTMSpeculator`` with ``method = "sum_rnn"`` and LIST-valued dims
(``emb_dim`` / ``proj_dim`` / ``inner_dim`` = "a.b.c" in the HF config) - the multi-stage speculator of
arctic_inference/vllm/spec_dec/arctic_speculator.py:476-536 (construction), :693-704 (``generate_states``),
:706-751 (``generate_token_ids``), :753-866 (``generate_proposals``); TP = 1.

The reference builds three ``nn.Sequential`` stacks per stage set ``i``:
    emb[i]  = Embedding(V, e0)               then per extra stage j:  LN(e_j) -> GELU -> Linear(e_{j-1} -> e_j)
    proj[i] = Linear(H | inner[-1] -> p0)    then per extra stage j:  LN(p_j) -> GELU -> Linear(p_{j-1} -> p_j)
    ln[i]   = LN(i0)                         then per extra stage j:  GELU -> Linear(i_{j-1} -> i_j) -> LN(i_j)
and one step is ``GELU(ln(proj(prev) + alpha * emb(tok)))``.  Because every stacked LayerNorm sits BEFORE the Linear
that produces its width, the reference itself only runs when all emb / proj widths equal ``inner_dim[0]``
(``inner_dim[1:]`` are free); other configurations raise here with that explanation.

Parameter names and shapes are the reference's (``emb.i.0.weight [V_pad64, D]``, ``emb.i.{3j-2}.{weight,bias}``,
``emb.i.{3j}.weight``, ``proj.i.0.weight``, ..., ``ln.i.0.{weight,bias}``, ``ln.i.{3j-1}.weight``,
``ln.i.{3j}.{weight,bias}``, ``head.i.weight [V_pad64, inner[-1]]``).

* One stage per stack: identical arithmetic to ``ArcticMLPSpeculator`` - served by its fused native loop
  (``ab200_mlp_propose``), only the parameter names are mapped.
* Stacked stages: every block of the stacks is one of two native launches - ``ab200_linear`` (tcgen05 / exact-fp32
  GEMM) and ``ab200_sum_mlp_cell`` (``GELU(LN(pre + alpha * table[token]))`` with the eager bf16 rounding points):
    GELU(LN(emb[tok]))      = cell(pre = 0,  table = emb,       token = tok, alpha = 1)
    GELU(LN(x))             = cell(pre = x,  table = one zero row,           alpha = 0)
    GELU(LN(s + alpha * z)) = cell(pre = s,  table = z [B, D],  token = row, alpha)
  so no new arithmetic enters the library and there is no PyTorch or CPU fallback.
"""
from __future__ import annotations

import collections
import math
from typing import Dict, Iterable, List, Optional, Sequence, Tuple, Union

import torch

from . import _capi
from .mlp_speculator import ArcticMLPSpeculator
from .speculator import SQRT2, pad_vocab_size, padding_size

Dims = Union[int, str, Sequence[int]]


def parse_dims(d: Dims) -> List[int]:
    """"4096.4096" (HF config, :424-429), an int, or a sequence of ints."""
    if isinstance(d, str):
        return [int(x) for x in d.split(".")]
    if isinstance(d, int):
        return [d]
    return [int(x) for x in d]


class ArcticSumRNNSpeculator:
    def __init__(self, *, vocab_size: int, input_hidden_dim: int, emb_dim: Dims, proj_dim: Dims, inner_dim: Dims,
                 n_predict: int, num_lookahead_tokens: Optional[int] = None, tie_weights: bool = True,
                 max_num_seqs: int = 256, dtype: torch.dtype = torch.bfloat16,
                 device: str | torch.device = "cuda", scale_input: bool = True, enforce_eager: bool = False,
                 emulate_eager_rounding: bool | int = True, alpha_flavor: str = "cuda",
                 quantize_lm_head: Optional[bool] = None) -> None:
        e, p, n = parse_dims(emb_dim), parse_dims(proj_dim), parse_dims(inner_dim)
        D = n[0]
        if any(x != D for x in e) or any(x != D for x in p):
            raise ValueError(
                f"sum_rnn: every emb_dim {e} and proj_dim {p} entry must equal inner_dim[0] = {D}: the reference "
                "applies the LayerNorm of stage j before the Linear that produces width j and adds emb to proj "
                "(arctic_speculator.py:486-493, :508-519, :701), so it fails on anything else as well")
        if any(x % 8 for x in n) or input_hidden_dim % 8:
            raise ValueError("widths must be multiples of 8")
        self.emb_dim, self.proj_dim, self.inner_dim = e, p, n
        self.input_hidden_dim = input_hidden_dim
        self.vocab_size = vocab_size
        self.n_predict = n_predict
        self.max_speculative_tokens = num_lookahead_tokens or n_predict
        self.tie_weights = tie_weights
        self.scale_input = scale_input
        self.method = "sum_rnn"
        self.device = torch.device(device)
        self.dtype = dtype
        self._fused: Optional[ArcticMLPSpeculator] = None
        if len(e) == len(p) == len(n) == 1:
            # one stage per stack = ArcticMLPSpeculator's arithmetic (:265-285 vs :693-704): the fused native loop
            self._fused = ArcticMLPSpeculator(
                vocab_size=vocab_size, emb_dim=input_hidden_dim, inner_dim=D, n_predict=n_predict,
                num_lookahead_tokens=num_lookahead_tokens, tie_weights=tie_weights, max_num_seqs=max_num_seqs,
                dtype=dtype, device=device, scale_input=scale_input, enforce_eager=enforce_eager,
                emulate_eager_rounding=emulate_eager_rounding, alpha_flavor=alpha_flavor,
                quantize_lm_head=quantize_lm_head)
            return
        if dtype not in (torch.bfloat16, torch.float32):
            raise ValueError("dtype must be torch.bfloat16 or torch.float32")
        if tie_weights and n_predict <= 1:
            raise AssertionError("You cannot tie weights between stages when only 1 exists")
        self._lib = _capi.lib()
        if self.device.type == "cuda":
            _capi.init_device(self.device)
        self.emulate_eager_rounding = 0 if dtype == torch.float32 else int(emulate_eager_rounding)
        self.state_weight = 0.5 ** (0.5 / n_predict)                                    # :575
        self.emb_weight = math.sqrt((1 - self.state_weight ** 2) * (D / 2))             # :576-577 (inner_dim[0])
        alpha = torch.tensor(self.emb_weight / self.state_weight, dtype=torch.float32)
        if alpha_flavor == "cpu" and dtype != torch.float32:
            alpha = alpha.to(dtype).to(torch.float32)
        elif alpha_flavor not in ("cpu", "cuda"):
            raise ValueError("alpha_flavor must be 'cuda' or 'cpu'")
        self.alpha = float(alpha)
        self.eps = 1e-6

        V, H, S = vocab_size, input_hidden_dim, self.max_speculative_tokens
        Vp = pad_vocab_size(V)
        mk = lambda *shape: torch.zeros(*shape, dtype=dtype, device=self.device)
        P: Dict[str, torch.Tensor] = collections.OrderedDict()
        n_sets = 1 if tie_weights else S
        n_proj = min(2, S) if tie_weights else S
        for i in range(n_sets):
            P[f"head.{i}.weight"] = mk(Vp, n[-1])
        for i in range(n_sets):                                                          # :478-493
            P[f"emb.{i}.0.weight"] = mk(Vp, e[0])
            for j in range(1, len(e)):
                P[f"emb.{i}.{3 * j - 2}.weight"] = mk(e[j])
                P[f"emb.{i}.{3 * j - 2}.bias"] = mk(e[j])
                P[f"emb.{i}.{3 * j}.weight"] = mk(e[j], e[j - 1])
        for i in range(n_proj):                                                          # :497-519
            P[f"proj.{i}.0.weight"] = mk(p[0], H if i == 0 else n[-1])
            for j in range(1, len(p)):
                P[f"proj.{i}.{3 * j - 2}.weight"] = mk(p[j])
                P[f"proj.{i}.{3 * j - 2}.bias"] = mk(p[j])
                P[f"proj.{i}.{3 * j}.weight"] = mk(p[j], p[j - 1])
        for i in range(n_sets):                                                          # :521-536
            P[f"ln.{i}.0.weight"] = mk(n[0])
            P[f"ln.{i}.0.bias"] = mk(n[0])
            for j in range(1, len(n)):
                P[f"ln.{i}.{3 * j - 1}.weight"] = mk(n[j], n[j - 1])
                P[f"ln.{i}.{3 * j}.weight"] = mk(n[j])
                P[f"ln.{i}.{3 * j}.bias"] = mk(n[j])
        self._params = P
        self.qhead = None                      # stacked stages: bf16 head at every batch (no FP8 copy is kept)

        self.cuda_graph_max_batch_size = padding_size(max_num_seqs)
        self.cuda_graph_mode = not enforce_eager
        self.cuda_graphs: Dict[Tuple[int, int], torch.cuda.CUDAGraph] = {}
        nb = self.cuda_graph_max_batch_size
        self.static_cuda_buffers = {
            "last_tokens": torch.zeros(nb, dtype=torch.long, device=self.device),
            "previous_hidden_states": torch.zeros(nb, H, dtype=dtype, device=self.device),
            "next_tokens": torch.zeros(nb, S, dtype=torch.long, device=self.device),
        }
        # helpers of the identity / dense-table forms of the cell (see the module docstring)
        self._zero_row = mk(1, max(n))
        self._zero_tok = torch.zeros(nb, dtype=torch.long, device=self.device)
        self._row_ids = torch.arange(nb, dtype=torch.long, device=self.device)

    # ---- parameters ---------------------------------------------------------------------------
    @staticmethod
    def _to_fused_name(name: str) -> Optional[str]:
        """``emb.i.0.weight`` -> ``emb.i.weight`` (single-stage stacks only have Sequential index 0)."""
        parts = name.split(".")
        if parts[0] == "head":
            return name
        if len(parts) == 4 and parts[0] in ("emb", "proj", "ln") and parts[2] == "0":
            return ".".join((parts[0], parts[1], parts[3]))
        return None

    def named_parameters(self):
        if self._fused is None:
            return list(self._params.items())
        out = []
        for k, v in self._fused.named_parameters():
            parts = k.split(".")
            out.append((k if parts[0] == "head" else f"{parts[0]}.{parts[1]}.0.{parts[2]}", v))
        return out

    def load_weights(self, weights: Iterable[Tuple[str, torch.Tensor]]) -> None:
        """:868-902 for this method: names as the reference registers them; ``qhead.*`` is ignored."""
        if self._fused is not None:
            mapped = []
            for name, w in weights:
                m = self._to_fused_name(name.replace("speculator.", ""))
                if m is not None:
                    mapped.append((m, w))
            self._fused.load_weights(mapped)
            return
        for name, w in weights:
            name = name.replace("speculator.", "")
            if name.startswith("qhead") or name not in self._params:
                continue
            p = self._params[name]
            w = w.to(device=self.device, dtype=self.dtype)
            if tuple(w.shape) != tuple(p.shape):
                if w.dim() == 2 and w.shape[0] <= p.shape[0] and w.shape[1] == p.shape[1] and \
                        (name.startswith("head") or name.endswith(".0.weight") and name.startswith("emb")):
                    p.zero_()                     # vocab padding rows stay zero
                    p[:w.shape[0]].copy_(w)
                    continue
                raise ValueError(f"shape mismatch for {name}: {tuple(w.shape)} vs {tuple(p.shape)}")
            p.copy_(w)
        self.cuda_graphs.clear()

    def share_parameters(self, named_tensors: Iterable[Tuple[str, torch.Tensor]]) -> List[str]:
        """Adopt the caller's tensors (the reference's parameter names) WITHOUT copying where shape, dtype, device,
        layout and 16-byte alignment already fit; anything else is copied.  Returns the (reference) names shared."""
        if self._fused is not None:
            back = {}
            mapped = []
            for name, t in named_tensors:
                m = self._to_fused_name(name.replace("speculator.", ""))
                if m is not None:
                    back[m] = name
                    mapped.append((m, t))
            return [back[m] for m in self._fused.share_parameters(mapped)]
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
        self.cuda_graphs.clear()
        if rest:
            self.load_weights(rest)
        return shared

    def _idx(self, head_index: int, proj: bool = False) -> int:
        if not self.tie_weights:
            return head_index
        return (1 if head_index >= 2 else head_index) if proj else 0                     # :659-660

    # ---- the two native building blocks --------------------------------------------------------
    def _linear(self, x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
        B, K = x.shape
        if K != w.shape[1]:
            raise RuntimeError(f"width {K} does not match weight {tuple(w.shape)}")
        out = torch.empty(B, w.shape[0], dtype=self.dtype, device=self.device)
        _capi.check(self._lib.ab200_linear(x.data_ptr(), w.data_ptr(), out.data_ptr(), B, w.shape[0], K,
                                           x.stride(0), w.stride(0), out.stride(0),
                                           _capi.torch_dtype_code(self.dtype), _capi.current_stream_ptr(self.device)),
                    "ab200_linear")
        return out

    def _cell(self, pre: torch.Tensor, table: torch.Tensor, tokens: torch.Tensor, vocab: int, alpha: float,
              ln_w: torch.Tensor, ln_b: torch.Tensor) -> torch.Tensor:
        """GELU(LN(pre + alpha * table[tokens])) with the reference's rounding points (:88-95, nn.GELU)."""
        B, D = pre.shape
        out = torch.empty(B, D, dtype=self.dtype, device=self.device)
        _capi.check(self._lib.ab200_sum_mlp_cell(
            pre.data_ptr(), pre.stride(0), table.data_ptr(), tokens.data_ptr(), None, ln_w.data_ptr(),
            ln_b.data_ptr(), out.data_ptr(), B, D, vocab, alpha, self.eps, _capi.torch_dtype_code(self.dtype),
            int(self.emulate_eager_rounding), None, None, None, None, _capi.current_stream_ptr(self.device)),
            "ab200_sum_mlp_cell")
        return out

    def _ln_gelu(self, x: torch.Tensor, ln_w: torch.Tensor, ln_b: torch.Tensor) -> torch.Tensor:
        return self._cell(x, self._zero_row, self._zero_tok[:x.shape[0]] if x.shape[0] <= self._zero_tok.shape[0]
                          else torch.zeros(x.shape[0], dtype=torch.long, device=self.device), 1, 0.0, ln_w, ln_b)

    # ---- reference method boundary ------------------------------------------------------------
    def generate_states(self, last_tokens: torch.Tensor, previous_hidden_states: torch.Tensor,
                        head_index: int, cell_states: Optional[torch.Tensor] = None) -> torch.Tensor:
        """:648-704, sum_rnn branch (``cell_states`` is accepted and ignored, as there)."""
        if self._fused is not None:
            return self._fused.generate_states(last_tokens, previous_hidden_states, head_index)
        P, lib = self._params, self._lib
        e, p, n = self.emb_dim, self.proj_dim, self.inner_dim
        B = last_tokens.shape[0]
        x = previous_hidden_states.reshape(B, -1)
        if not x.is_contiguous():
            x = x.contiguous()
        toks = last_tokens.reshape(B).to(torch.long).contiguous()
        code = _capi.torch_dtype_code(self.dtype)
        with torch.cuda.device(self.device):
            if head_index == 0 and self.scale_input:                                    # :656-657
                y = torch.empty_like(x)
                inv_sqrt2 = float(torch.tensor(1.0, dtype=torch.float32) / torch.tensor(SQRT2, dtype=torch.float32))
                _capi.check(lib.ab200_rms_scale(x.data_ptr(), y.data_ptr(), B, x.shape[1], x.stride(0), y.stride(0),
                                                self.eps, inv_sqrt2, code, int(self.emulate_eager_rounding != 0),
                                                _capi.current_stream_ptr(self.device)), "ab200_rms_scale")
                x = y
            i, pi = self._idx(head_index), self._idx(head_index, True)
            # emb stack (:478-493): the embedding row itself is consumed inside the cell that follows it
            table, rows, vocab = P[f"emb.{i}.0.weight"], toks, self.vocab_size
            if len(e) > 1:
                zero_pre = torch.zeros(B, e[0], dtype=self.dtype, device=self.device)
                z = self._cell(zero_pre, table, toks, self.vocab_size, 1.0, P[f"emb.{i}.1.weight"], P[f"emb.{i}.1.bias"])
                z = self._linear(z, P[f"emb.{i}.3.weight"])
                for j in range(2, len(e)):
                    z = self._linear(self._ln_gelu(z, P[f"emb.{i}.{3 * j - 2}.weight"], P[f"emb.{i}.{3 * j - 2}.bias"]),
                                     P[f"emb.{i}.{3 * j}.weight"])
                table, vocab = z, B                      # dense [B, D] "table": row r belongs to batch row r
                rows = self._row_ids[:B] if B <= self._row_ids.shape[0] else \
                    torch.arange(B, dtype=torch.long, device=self.device)
            # proj stack (:497-519)
            s = self._linear(x, P[f"proj.{pi}.0.weight"])
            for j in range(1, len(p)):
                s = self._linear(self._ln_gelu(s, P[f"proj.{pi}.{3 * j - 2}.weight"], P[f"proj.{pi}.{3 * j - 2}.bias"]),
                                 P[f"proj.{pi}.{3 * j}.weight"])
            # states.add_(z, alpha) -> ln stack -> activation (:701-702); every LN of the stack is followed by a GELU
            # (the stack's own between stages, the outer self.activation after the last)
            s = self._cell(s, table, rows, vocab, self.alpha, P[f"ln.{i}.0.weight"], P[f"ln.{i}.0.bias"])
            for j in range(1, len(n)):
                s = self._ln_gelu(self._linear(s, P[f"ln.{i}.{3 * j - 1}.weight"]),
                                  P[f"ln.{i}.{3 * j}.weight"], P[f"ln.{i}.{3 * j}.bias"])
        return s.unsqueeze(1)

    def greedy_tokens(self, states: torch.Tensor, head_index: int = 0) -> torch.Tensor:
        """head projection + greedy pick (:726-732), logits never materialised."""
        if self._fused is not None:
            return self._fused.greedy_tokens(states, head_index)
        head = self._params[f"head.{self._idx(head_index)}.weight"]
        B, D = states.shape
        states = states.contiguous()
        packed = torch.zeros(B + 1, dtype=torch.int64, device=self.device)
        out = torch.empty(B, dtype=torch.long, device=self.device)
        ws = torch.empty(B * self.vocab_size, dtype=torch.float32, device=self.device) \
            if self.dtype == torch.float32 else None
        with torch.cuda.device(self.device):
            _capi.check(self._lib.ab200_head_argmax(
                states.data_ptr(), head.data_ptr(), B, head.shape[0], D, self.vocab_size, states.stride(0),
                head.stride(0), packed.data_ptr(), out.data_ptr(), 1, packed[B:].data_ptr(),
                _capi.torch_dtype_code(self.dtype), ws.data_ptr() if ws is not None else None,
                ws.numel() * 4 if ws is not None else 0, _capi.current_stream_ptr(self.device)), "ab200_head_argmax")
        return out

    def generate_token_ids(self, batch_size: int, num_predict_tokens: int, last_tokens: torch.Tensor,
                           previous_hidden_states: torch.Tensor,
                           next_tokens_tensors: List[Optional[torch.Tensor]]) -> List[torch.Tensor]:
        """:706-751."""
        if self._fused is not None:
            return self._fused.generate_token_ids(batch_size, num_predict_tokens, last_tokens, previous_hidden_states,
                                                  next_tokens_tensors)
        for head_index in range(num_predict_tokens):
            states = self.generate_states(last_tokens, previous_hidden_states, head_index)
            previous_hidden_states = states
            last_tokens = self.greedy_tokens(states.flatten(0, 1), head_index).reshape(batch_size, -1)
            if next_tokens_tensors[head_index] is None:
                next_tokens_tensors[head_index] = last_tokens
            else:
                next_tokens_tensors[head_index].copy_(last_tokens)
        return next_tokens_tensors

    def _loop_into(self, ids: torch.Tensor, hid: torch.Tensor, out: torch.Tensor, B: int, S: int) -> None:
        cols: List[Optional[torch.Tensor]] = [out[:, s:s + 1] for s in range(S)]
        self.generate_token_ids(B, S, ids.unsqueeze(1), hid.unsqueeze(1), cols)

    def generate_proposals(self, input_ids: torch.Tensor, previous_hidden_states: torch.Tensor,
                           num_predict_tokens: int) -> torch.Tensor:
        """:753-866: static buffers, batch padded to ``padding_size(B)``, one CUDA graph per (padded batch, S)."""
        if self._fused is not None:
            return self._fused.generate_proposals(input_ids, previous_hidden_states, num_predict_tokens)
        if num_predict_tokens > self.max_speculative_tokens:
            raise ValueError(f"Max speculative tokens for model is "
                             f"{self.max_speculative_tokens}, but "
                             f"{num_predict_tokens} were requested")
        B, S = input_ids.size(0), num_predict_tokens
        if B == 0:
            return torch.empty(0, S, dtype=torch.long, device=self.device)
        if previous_hidden_states.shape != (B, self.input_hidden_dim):
            raise RuntimeError(f"previous_hidden_states must be [{B}, {self.input_hidden_dim}]")
        if not (self.cuda_graph_mode and B <= self.cuda_graph_max_batch_size):
            ids = input_ids.to(device=self.device, dtype=torch.long).contiguous()
            hid = previous_hidden_states.to(device=self.device, dtype=self.dtype).contiguous()
            out = torch.empty(B, S, dtype=torch.long, device=self.device)
            self._loop_into(ids, hid, out, B, S)
            return out
        bufs = self.static_cuda_buffers
        bufs["last_tokens"][:B].copy_(input_ids, non_blocking=True)
        bufs["previous_hidden_states"][:B].copy_(previous_hidden_states, non_blocking=True)
        padded = padding_size(B)
        key = (padded, S)
        out_view = bufs["next_tokens"].view(-1)[:padded * S].view(padded, S)
        g = self.cuda_graphs.get(key)
        if g is None:
            ids_p, hid_p = bufs["last_tokens"][:padded], bufs["previous_hidden_states"][:padded]
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):                     # warm-up outside the capture
                self._loop_into(ids_p, hid_p, out_view, padded, S)
            torch.cuda.current_stream(self.device).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._loop_into(ids_p, hid_p, out_view, padded, S)
            self.cuda_graphs[key] = g
        g.replay()
        return out_view[:B].clone()

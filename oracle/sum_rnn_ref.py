"""TEST INFRASTRUCTURE ONLY - oracle for the ``method="sum_rnn"`` configuration of the reference's
``ArcticLSTMSpeculator`` with dims given as LISTS (``emb_dim`` / ``proj_dim`` / ``inner_dim`` = "a.b.c"), the
"multi-stage" speculator VERDICT r1 listed as missing #7.

Op-for-op PyTorch restatement (same ATen ops, same order) of
/root/reference/arctic_inference/vllm/spec_dec/arctic_speculator.py:
  * construction of the three ``nn.Sequential`` stacks :476-536
      emb[i]  = Embedding(V, e0) , then per extra stage j:  LN(e_j) -> GELU -> Linear(e_{j-1} -> e_j)
      proj[i] = Linear(H | inner[-1] -> p0), then per stage j:  LN(p_j) -> GELU -> Linear(p_{j-1} -> p_j)
      ln[i]   = LN(i0), then per stage j:  GELU -> Linear(i_{j-1} -> i_j) -> LN(i_j)
    (the LayerNorm of stage j is applied BEFORE the Linear that produces width j, so the reference only runs when
    every emb / proj width equals inner_dim[0]; inner_dim[1:] are free - probed in the build container),
  * ``generate_states`` sum_rnn branch :693-704 (index rule :659-660), the shared ``generate_token_ids`` :706-751
    and ``generate_proposals`` :753-866 (TP = 1, qhead disabled).
Parameter names are the reference's: ``emb.i.0.weight [V_pad64, D]``, ``emb.i.{3j-2}.{weight,bias}`` (LN),
``emb.i.{3j}.weight`` (Linear); ``proj.i.0.weight``, ``proj.i.{3j-2}.*``, ``proj.i.{3j}.weight``;
``ln.i.0.{weight,bias}``, ``ln.i.{3j-1}.weight`` (Linear), ``ln.i.{3j}.{weight,bias}`` (LN); ``head.i.weight``.

Parity pinned: yes - bit-exact against golden vectors produced by the unmodified reference class on CPU
(tests/golden/sumrnn_*.npz, generator oracle/gen_golden_sumrnn.py; tests/test_oracle_sumrnn.py).

Nothing under arcticinference_b200/ may import this module.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

from oracle.speculator_ref import SQRT2, pad_vocab, rms_layernorm, state_emb_weights


def parse_dims(d) -> List[int]:
    if isinstance(d, str):
        return [int(x) for x in d.split(".")]
    if isinstance(d, int):
        return [d]
    return [int(x) for x in d]


def make_sumrnn_weights(V: int, H: int, emb_dim, proj_dim, inner_dim, n_stages: int, tie_weights: bool,
                        dtype: torch.dtype, seed: int = 0, device: str = "cpu") -> Dict[str, torch.Tensor]:
    e, p, n = parse_dims(emb_dim), parse_dims(proj_dim), parse_dims(inner_dim)
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    Vp = pad_vocab(V)

    def rn(*shape, mean=0.0, std=0.02):
        t = torch.empty(*shape, dtype=torch.float32).normal_(mean, std, generator=g)
        return t.to(torch.bfloat16).to(dtype)

    def lin(out_f, in_f):
        return rn(out_f, in_f, std=1.0 / (in_f ** 0.5))

    def padded(rows, width):
        t = torch.zeros(Vp, width, dtype=dtype)
        t[:V] = rows
        return t

    w: Dict[str, torch.Tensor] = {}
    n_sets = 1 if tie_weights else n_stages
    n_proj = min(2, n_stages) if tie_weights else n_stages
    for i in range(n_sets):
        w[f"head.{i}.weight"] = padded(rn(V, n[-1]), n[-1])
        w[f"emb.{i}.0.weight"] = padded(rn(V, e[0], std=0.5), e[0])
        for j in range(1, len(e)):
            w[f"emb.{i}.{3 * j - 2}.weight"] = rn(e[j], mean=1.0, std=0.1)
            w[f"emb.{i}.{3 * j - 2}.bias"] = rn(e[j], mean=0.0, std=0.1)
            w[f"emb.{i}.{3 * j}.weight"] = lin(e[j], e[j - 1])
        w[f"ln.{i}.0.weight"] = rn(n[0], mean=1.0, std=0.1)
        w[f"ln.{i}.0.bias"] = rn(n[0], mean=0.0, std=0.1)
        for j in range(1, len(n)):
            w[f"ln.{i}.{3 * j - 1}.weight"] = lin(n[j], n[j - 1])
            w[f"ln.{i}.{3 * j}.weight"] = rn(n[j], mean=1.0, std=0.1)
            w[f"ln.{i}.{3 * j}.bias"] = rn(n[j], mean=0.0, std=0.1)
    for i in range(n_proj):
        w[f"proj.{i}.0.weight"] = lin(p[0], H if i == 0 else n[-1])
        for j in range(1, len(p)):
            w[f"proj.{i}.{3 * j - 2}.weight"] = rn(p[j], mean=1.0, std=0.1)
            w[f"proj.{i}.{3 * j - 2}.bias"] = rn(p[j], mean=0.0, std=0.1)
            w[f"proj.{i}.{3 * j}.weight"] = lin(p[j], p[j - 1])
    return {k: v.to(device) for k, v in w.items()}


class RefSumRNNSpeculator:
    def __init__(self, weights: Dict[str, torch.Tensor], vocab_size: int, n_predict: int, emb_dim, proj_dim,
                 inner_dim, tie_weights: bool, max_speculative_tokens: Optional[int] = None,
                 scale_input: bool = True):
        self.w = weights
        self.vocab_size = vocab_size
        self.n_predict = n_predict
        self.e, self.p, self.n = parse_dims(emb_dim), parse_dims(proj_dim), parse_dims(inner_dim)
        self.tie_weights = tie_weights
        self.max_speculative_tokens = max_speculative_tokens or n_predict
        self.scale_input = scale_input
        self.state_weight, self.emb_weight = state_emb_weights(n_predict, self.n[0])       # :575-577 (inner_dim[0])

    def _emb(self, i: int, tokens: torch.Tensor) -> torch.Tensor:                          # :478-493
        w = self.w
        z = F.embedding(tokens, w[f"emb.{i}.0.weight"])
        for j in range(1, len(self.e)):
            z = rms_layernorm(z, w[f"emb.{i}.{3 * j - 2}.weight"], w[f"emb.{i}.{3 * j - 2}.bias"])
            z = F.linear(F.gelu(z), w[f"emb.{i}.{3 * j}.weight"])
        return z

    def _proj(self, i: int, x: torch.Tensor) -> torch.Tensor:                              # :497-519
        w = self.w
        s = F.linear(x, w[f"proj.{i}.0.weight"])
        for j in range(1, len(self.p)):
            s = rms_layernorm(s, w[f"proj.{i}.{3 * j - 2}.weight"], w[f"proj.{i}.{3 * j - 2}.bias"])
            s = F.linear(F.gelu(s), w[f"proj.{i}.{3 * j}.weight"])
        return s

    def _ln(self, i: int, s: torch.Tensor) -> torch.Tensor:                                # :521-536
        w = self.w
        s = rms_layernorm(s, w[f"ln.{i}.0.weight"], w[f"ln.{i}.0.bias"])
        for j in range(1, len(self.n)):
            s = F.linear(F.gelu(s), w[f"ln.{i}.{3 * j - 1}.weight"])
            s = rms_layernorm(s, w[f"ln.{i}.{3 * j}.weight"], w[f"ln.{i}.{3 * j}.bias"])
        return s

    def generate_states(self, last_tokens, previous_hidden_states, head_index):
        if head_index == 0 and self.scale_input:
            previous_hidden_states = rms_layernorm(previous_hidden_states) / SQRT2          # :656-657
        actual_i = 0 if self.tie_weights else head_index                                    # :659
        actual_proj_i = 1 if self.tie_weights and head_index >= 2 else head_index           # :660
        z = self._emb(actual_i, last_tokens)                                                # :695
        states = self._proj(actual_proj_i, previous_hidden_states)                          # :696
        states.add_(z, alpha=self.emb_weight / self.state_weight)                           # :701
        return F.gelu(self._ln(actual_i, states))                                           # :702

    def logits(self, states2d: torch.Tensor, head_index: int) -> torch.Tensor:
        i = 0 if self.tie_weights else head_index
        return F.linear(states2d, self.w[f"head.{i}.weight"])[..., :self.vocab_size]

    def generate_proposals(self, input_ids, previous_hidden_states, num_predict_tokens):
        if num_predict_tokens > self.max_speculative_tokens:
            raise ValueError(f"Max speculative tokens for model is {self.max_speculative_tokens}, but "
                             f"{num_predict_tokens} were requested")
        B = input_ids.shape[0]
        tok, prev = input_ids.unsqueeze(1), previous_hidden_states.unsqueeze(1)
        out = []
        for s in range(num_predict_tokens):
            prev = self.generate_states(tok, prev, s)
            tok = torch.argmax(self.logits(prev.flatten(0, 1), s), dim=-1).reshape(B, -1)
            out.append(tok)
        return torch.cat(out, dim=1)

"""TEST INFRASTRUCTURE ONLY - oracle for the Arctic MLP speculator / "sum_rnn" draft head
(SURVEY.md 8f rank 3).

Op-for-op PyTorch restatement (same ATen ops, same order, so bf16 rounds where the reference
rounds) of /root/reference/arctic_inference/vllm/spec_dec/arctic_speculator.py:
``ArcticMLPSpeculator.generate_states`` :265-285, ``generate_token_ids`` :287-326,
``generate_proposals`` :328-389 (TP = 1, qhead disabled).  Parameter names follow the reference
(``emb.i.weight [V_pad64, D]``, ``proj.i.weight``, ``ln.i.{weight,bias}``, ``head.i.weight``); with
``tie_weights`` only index 0 exists for emb / ln / head and indices 0, 1 for proj (:139-175).

Parity pinned: yes - bit-exact against golden vectors produced by the unmodified reference class
on CPU (tests/golden/mlp_speculator_*.npz, generator oracle/gen_golden_mlp.py;
tests/test_oracle_mlp.py).

Nothing under arcticinference_b200/ may import this module.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from oracle.speculator_ref import SQRT2, pad_vocab, rms_layernorm, state_emb_weights


def make_mlp_weights(V: int, H: int, D: int, n_stages: int, tie_weights: bool, dtype: torch.dtype,
                     seed: int = 0, device: str = "cpu") -> Dict[str, torch.Tensor]:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    Vp = pad_vocab(V)

    def rn(*shape, mean=0.0, std=0.02):
        t = torch.empty(*shape, dtype=torch.float32).normal_(mean, std, generator=g)
        return t.to(torch.bfloat16).to(dtype)

    def padded(rows):
        t = torch.zeros(Vp, D, dtype=dtype)
        t[:V] = rows
        return t

    w: Dict[str, torch.Tensor] = {}
    n_sets = 1 if tie_weights else n_stages
    n_proj = min(2, n_stages) if tie_weights else n_stages
    for i in range(n_sets):
        w[f"emb.{i}.weight"] = padded(rn(V, D))
        w[f"head.{i}.weight"] = padded(rn(V, D))
        w[f"ln.{i}.weight"] = rn(D, mean=1.0, std=0.1)
        w[f"ln.{i}.bias"] = rn(D, mean=0.0, std=0.1)
    for i in range(n_proj):
        w[f"proj.{i}.weight"] = rn(D, H if i == 0 else D)
    return {k: v.to(device) for k, v in w.items()}


class RefMLPSpeculator:
    def __init__(self, weights: Dict[str, torch.Tensor], vocab_size: int, n_predict: int,
                 tie_weights: bool, max_speculative_tokens: Optional[int] = None,
                 scale_input: bool = True):
        self.w = weights
        self.vocab_size = vocab_size
        self.n_predict = n_predict
        self.tie_weights = tie_weights
        self.max_speculative_tokens = max_speculative_tokens or n_predict
        self.scale_input = scale_input
        self.D = weights["proj.0.weight"].shape[0]
        self.H = weights["proj.0.weight"].shape[1]
        self.state_weight, self.emb_weight = state_emb_weights(n_predict, self.D)

    def _idx(self, head_index: int, proj: bool = False) -> int:
        if not self.tie_weights:
            return head_index
        return min(head_index, 1) if proj else 0        # :139-152 ModuleList([first] + [tied]*(n-1))

    def generate_states(self, last_tokens, previous_hidden_states, head_index):
        w = self.w
        if head_index == 0 and self.scale_input:
            previous_hidden_states = rms_layernorm(previous_hidden_states) / SQRT2        # :271-272
        z = F.embedding(last_tokens, w[f"emb.{self._idx(head_index)}.weight"])            # :274
        states = F.linear(previous_hidden_states, w[f"proj.{self._idx(head_index, True)}.weight"])  # :275
        states.add_(z, alpha=self.emb_weight / self.state_weight)                         # :280
        i = self._idx(head_index)
        return F.gelu(rms_layernorm(states, w[f"ln.{i}.weight"], w[f"ln.{i}.bias"]))       # :282

    def logits(self, states_2d, head_index):
        return F.linear(states_2d, self.w[f"head.{self._idx(head_index)}.weight"])[..., :self.vocab_size]

    def generate_proposals(self, input_ids, previous_hidden_states, num_predict_tokens,
                           trace: Optional[dict] = None) -> torch.Tensor:
        if num_predict_tokens > self.max_speculative_tokens:
            raise ValueError(f"Max speculative tokens for model is "
                             f"{self.max_speculative_tokens}, but "
                             f"{num_predict_tokens} were requested")
        prev = previous_hidden_states.unsqueeze(1)
        last = input_ids.unsqueeze(1)
        B = input_ids.size(0)
        out: List[torch.Tensor] = []
        for s in range(num_predict_tokens):
            states = self.generate_states(last, prev, s)
            prev = states
            logits = self.logits(states.flatten(0, 1), s)
            last = torch.argmax(logits, dim=-1).reshape(B, -1)                             # :302-303
            out.append(last)
            if trace is not None:
                trace[f"state{s}"] = states.flatten(0, 1)
                trace[f"logits{s}"] = logits
        return torch.cat(out, dim=-1)

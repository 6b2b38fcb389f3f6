"""CPU: pins oracle/sum_rnn_ref.py (the ``method = "sum_rnn"`` branch of the reference's ArcticLSTMSpeculator with
list-valued dims) against golden vectors produced by the unmodified reference class (oracle/gen_golden_sumrnn.py),
bit for bit - tokens and the teacher-forced state of every step."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import speculator_ref as O
from oracle import sum_rnn_ref as R

CASES = sorted(os.path.basename(p)[len("sumrnn_"):-4] for p in glob.glob(os.path.join(GOLDEN, "sumrnn_*.npz")))


def load(name):
    z = np.load(os.path.join(GOLDEN, f"sumrnn_{name}.npz"))
    dtype = torch.bfloat16 if name.endswith("bf16") else torch.float32

    def t(key):
        a = torch.from_numpy(z[key])
        return a.view(torch.bfloat16) if a.dtype == torch.int16 else a

    V, H, n_predict, S, B, tie = (int(v) for v in z["meta"])
    emb, proj, inner = (str(x) for x in z["dims"])
    return t, dtype, (V, H, emb, proj, inner, n_predict, S, B, bool(tie))


def test_golden_present():
    assert len(CASES) >= 5 and any("untied" in c for c in CASES) and any("single" in c for c in CASES)


@pytest.mark.parametrize("name", CASES)
def test_op_for_op_matches_reference_bit_exact(name):
    t, dtype, (V, H, emb, proj, inner, n_predict, S, B, tie) = load(name)
    w = R.make_sumrnn_weights(V, H, emb, proj, inner, n_predict, tie, dtype, seed=0)
    ids, hid = O.make_inputs(B, V, H, dtype, seed=1)
    assert torch.equal(ids, t("ids")) and torch.equal(hid, t("hidden"))
    m = R.RefSumRNNSpeculator(w, V, n_predict, emb, proj, inner, tie)
    tokens = t("tokens")
    assert torch.equal(m.generate_proposals(ids, hid, S), tokens)
    prev, tok = hid.unsqueeze(1), ids.unsqueeze(1)
    for s in range(S):
        state = m.generate_states(tok, prev, s)
        assert torch.equal(state.flatten(0, 1), t(f"state{s}")), f"state step {s}"
        tok, prev = tokens[:, s:s + 1], state


def test_single_stage_equals_the_mlp_class_arithmetic():
    """One stage per stack: the LSTM class's sum_rnn branch is ArcticMLPSpeculator's arithmetic under other
    parameter names (emb.i.0 / proj.i.0 / ln.i.0) - which is why the fused MLP path serves it."""
    from oracle import mlp_speculator_ref as M
    V, H, D, n = 500, 48, 32, 3
    w = R.make_sumrnn_weights(V, H, D, D, D, n, True, torch.bfloat16, seed=3)
    mapped = {k.replace(".0.weight", ".weight").replace(".0.bias", ".bias") if not k.startswith("head") else k: v
              for k, v in w.items()}
    ids, hid = O.make_inputs(6, V, H, torch.bfloat16, seed=4)
    a = R.RefSumRNNSpeculator(w, V, n, D, D, D, True).generate_proposals(ids, hid, n)
    b = M.RefMLPSpeculator(mapped, V, n, True).generate_proposals(ids, hid, n)
    assert torch.equal(a, b)


def test_too_many_tokens_raises():
    w = R.make_sumrnn_weights(100, 16, "16.16", "16", "16.32", 2, True, torch.float32)
    m = R.RefSumRNNSpeculator(w, 100, 2, "16.16", "16", "16.32", True)
    ids, hid = O.make_inputs(2, 100, 16, torch.float32)
    with pytest.raises(ValueError):
        m.generate_proposals(ids, hid, 3)

"""CPU: pins oracle/mlp_speculator_ref.py against golden vectors produced by the reference's own
ArcticMLPSpeculator class (oracle/gen_golden_mlp.py), bit for bit - tied and untied weights."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import mlp_speculator_ref as M
from oracle import speculator_ref as O

CASES = sorted(os.path.basename(p)[len("mlp_speculator_"):-4]
               for p in glob.glob(os.path.join(GOLDEN, "mlp_speculator_*.npz")))


def load(name):
    z = np.load(os.path.join(GOLDEN, f"mlp_speculator_{name}.npz"))
    dtype = torch.bfloat16 if name.endswith("bf16") else torch.float32

    def t(key):
        a = torch.from_numpy(z[key])
        return a.view(torch.bfloat16) if a.dtype == torch.int16 else a

    V, H, D, n_predict, S, B, tie = (int(v) for v in z["meta"])
    return t, dtype, (V, H, D, n_predict, S, B, bool(tie))


def test_golden_present():
    assert len(CASES) >= 5 and any("untied" in c for c in CASES)


@pytest.mark.parametrize("name", CASES)
def test_op_for_op_matches_reference_bit_exact(name):
    t, dtype, (V, H, D, n_predict, S, B, tie) = load(name)
    w = M.make_mlp_weights(V, H, D, n_predict, tie, dtype, seed=0)
    ids, hid = O.make_inputs(B, V, H, dtype, seed=1)
    assert torch.equal(ids, t("ids")) and torch.equal(hid, t("hidden"))
    m = M.RefMLPSpeculator(w, V, n_predict, tie)
    trace = {}
    tokens = m.generate_proposals(ids, hid, S, trace=trace)
    assert torch.equal(tokens, t("tokens"))
    for s in range(S):
        assert torch.equal(trace[f"state{s}"], t(f"state{s}")), f"state step {s}"
        assert torch.equal(trace[f"logits{s}"], t(f"logits{s}")), f"logits step {s}"


def test_too_many_tokens_raises():
    w = M.make_mlp_weights(100, 16, 16, 2, True, torch.float32)
    m = M.RefMLPSpeculator(w, 100, 2, True)
    ids, hid = O.make_inputs(2, 100, 16, torch.float32)
    with pytest.raises(ValueError):
        m.generate_proposals(ids, hid, 3)

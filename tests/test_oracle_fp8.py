"""CPU: the FP8-head oracle (oracle/fp8_ref.py) - known-answer conversions, the published
quantisation rules, and golden vectors from vLLM's own fp8 ops run on a B200."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import fp8_ref as F8

GOLD = sorted(glob.glob(os.path.join(GOLDEN, "fp8_vllm_*.npz")))


def test_e4m3_known_answers():
    x = torch.tensor([0.0, 1.0, 448.0, 449.0, 1e9, -1e9, 0.0625 + 0.0625 / 16, 2 ** -9, 2 ** -10, 17.0, 19.0])
    got = F8._sat_cast(x).float().tolist()
    #                0    1    max    sat    sat    -sat   tie->even      min subnormal  underflow  17->16 (tie to even), 19->20
    assert got == [0.0, 1.0, 448.0, 448.0, 448.0, -448.0, 0.0625, 2 ** -9, 0.0, 16.0, 20.0]


def test_weight_quantisation_rule():
    torch.manual_seed(0)
    w = (torch.randn(300, 64) * 0.03).to(torch.bfloat16)
    q, s = F8.quantize_weight(w)
    assert s.shape == (1,) and float(s) == float(w.float().abs().max() / 448.0)
    assert float(q.float().abs().max()) == 448.0                      # the max element maps to +-448
    err = (q.float() * s - w.float()).abs()
    assert (err <= 0.0625 * w.float().abs() + float(s) * 2 ** -10).all()  # 3 mantissa bits


def test_row_quantisation_rule_and_min_scale():
    x = torch.zeros(3, 32, dtype=torch.bfloat16)
    x[1] = torch.linspace(-2, 2, 32).to(torch.bfloat16)
    x[2, 5] = 1e-8
    q, s = F8.quantize_rows(x)
    assert s.shape == (3, 1)
    min32 = float(torch.tensor(F8.MIN_SCALE, dtype=torch.float32))
    assert float(s[0]) == min32 and float(s[2]) == min32                  # clamp for (near-)zero rows
    assert float(s[1]) == float(torch.tensor(2.0) / torch.tensor(448.0)) and float(q[1].float().abs().max()) == 448.0
    assert (q[0].float() == 0).all()


def test_scaled_mm_matches_dequantised_product():
    torch.manual_seed(1)
    w = (torch.randn(128, 64) * 0.02).to(torch.bfloat16)
    x = torch.randn(7, 64).to(torch.bfloat16)
    qw, sw = F8.quantize_weight(w)
    out = F8.fp8_linear(x, qw, sw)
    ref = x.float() @ w.float().t()
    assert out.dtype == torch.bfloat16
    assert ((out.float() - ref).norm() / ref.norm()).item() < 0.06          # fp8 x fp8 noise


def test_selection_rule_batch_32():
    from oracle import speculator_ref as O
    V, H, D = 512, 64, 64
    w = O.make_weights(V, H, D, torch.bfloat16)
    m = F8.RefLSTMSpeculatorFp8(w, V, 3)
    plain = O.RefLSTMSpeculator(w, V, 3)
    x = torch.randn(33, D).to(torch.bfloat16)
    assert torch.equal(m.logits(x), plain.logits(x))                        # batch 33: bf16 head
    assert not torch.equal(m.logits(x[:32]), plain.logits(x[:32]))         # batch 32: fp8 head


@pytest.mark.skipif(not GOLD, reason="fp8 golden fixtures not generated yet")
@pytest.mark.parametrize("path", GOLD, ids=os.path.basename)
def test_oracle_matches_vllm_ops_golden(path):
    """vLLM computes scales with fast-math division (<= 1 ulp from IEEE); quantised codes and
    outputs therefore agree except for a tiny fraction sitting on rounding boundaries."""
    z = np.load(path)
    w = torch.from_numpy(z["w"]).view(torch.bfloat16)
    x = torch.from_numpy(z["x"]).view(torch.bfloat16)
    qw, sw = F8.quantize_weight(w)
    qx, sx = F8.quantize_rows(x)
    ulp = lambda a, b: (a.view(torch.int32) - b.view(torch.int32)).abs().max().item()
    assert ulp(sw, torch.from_numpy(z["sw"]).reshape(1)) <= 1
    assert ulp(sx, torch.from_numpy(z["sx"]).reshape(-1, 1)) <= 1
    eq_w = (qw.view(torch.uint8) == torch.from_numpy(z["qw"])).float().mean().item()
    eq_x = (qx.view(torch.uint8) == torch.from_numpy(z["qx"])).float().mean().item()
    assert eq_w >= 0.999 and eq_x >= 0.999, (eq_w, eq_x)
    # the GEMM itself, on vLLM's own quantised operands and scales: only accumulation-order noise
    out = F8.scaled_mm(torch.from_numpy(z["qx"]).view(F8.F8), torch.from_numpy(z["sx"]).reshape(-1, 1),
                       torch.from_numpy(z["qw"]).view(F8.F8), torch.from_numpy(z["sw"]).reshape(1))
    ref = torch.from_numpy(z["out"]).view(torch.bfloat16)
    assert (out == ref).float().mean().item() >= 0.999
    assert torch.equal(out.float().argmax(-1), ref.float().argmax(-1)) or \
        (out.float().argmax(-1) == ref.float().argmax(-1)).float().mean().item() >= 0.95

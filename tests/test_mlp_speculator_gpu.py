"""GPU: the MLP speculator ("sum_rnn") path - SURVEY 8f rank 3.  Our kernels through the C ABI vs
(i) golden vectors of the reference's own ArcticMLPSpeculator class, (ii) the op-for-op oracle run
on the same CUDA device (identical ATen kernels and rounding points to the reference on GPU)."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import mlp_speculator_ref as M
from oracle import speculator_ref as O

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
CASES = sorted(os.path.basename(p)[len("mlp_speculator_"):-4]
               for p in glob.glob(os.path.join(GOLDEN, "mlp_speculator_*.npz")))


def build(V, H, D, n_predict, tie, dtype, w, max_bs=64, **kw):
    from arcticinference_b200.mlp_speculator import ArcticMLPSpeculator
    m = ArcticMLPSpeculator(vocab_size=V, emb_dim=H, inner_dim=D, n_predict=n_predict, tie_weights=tie,
                            max_num_seqs=max_bs, dtype=dtype, device=DEV, **kw)
    m.load_weights(w.items())
    if "quantize_lm_head" not in kw:
        m.qhead = None      # oracle recipe (the goldens were generated with `model.qhead = None`)
    return m


def relerr(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("name", CASES)
def test_against_reference_golden_vectors(name):
    """CPU-generated goldens: fp32 -> exact tokens, states <= 1e-5; bf16 -> states <= 1e-2
    (the reference's CPU bf16 kernels round differently from CUDA's in alpha / GELU, see DESIGN)."""
    z = np.load(os.path.join(GOLDEN, f"mlp_speculator_{name}.npz"))
    dtype = torch.bfloat16 if name.endswith("bf16") else torch.float32
    V, H, D, n_predict, S, B, tie = (int(v) for v in z["meta"])

    def t(key):
        a = torch.from_numpy(z[key])
        return (a.view(torch.bfloat16) if a.dtype == torch.int16 else a).to(DEV)

    w = M.make_mlp_weights(V, H, D, n_predict, bool(tie), dtype, seed=0, device=DEV)
    m = build(V, H, D, n_predict, bool(tie), dtype, w, alpha_flavor="cpu")
    # teacher-forced by the reference's own tokens / states
    prev, tok = t("hidden").unsqueeze(1), t("ids").unsqueeze(1)
    for s in range(S):
        st = m.generate_states(tok, prev, s)
        tol = 1e-5 if dtype == torch.float32 else 1e-2
        assert relerr(st.flatten(0, 1), t(f"state{s}")) <= tol, f"state step {s}"
        prev, tok = t(f"state{s}").unsqueeze(1), t("tokens")[:, s:s + 1]
    got = m.generate_proposals(t("ids"), t("hidden"), S)
    if dtype == torch.float32:
        assert torch.equal(got, t("tokens"))
    else:
        assert (got == t("tokens")).float().mean().item() >= 0.9


@pytest.mark.parametrize("V,H,D,S,B,tie", [
    (128256, 4096, 4096, 3, 64, True),      # 8B shape
    (32000, 4096, 4096, 4, 48, False),      # untied stages
    (201088, 2880, 2880, 3, 64, True),      # GPT-OSS shape: D not a power of two
])
def test_bf16_agreement_with_cuda_oracle(V, H, D, S, B, tie):
    w = M.make_mlp_weights(V, H, D, S, tie, torch.bfloat16, seed=0, device=DEV)
    ref = M.RefMLPSpeculator(w, V, S, tie)
    m = build(V, H, D, S, tie, torch.bfloat16, w, max_bs=B)
    agree = total = 0
    worst = 0.0
    for seed in range(12):
        ids, hid = O.make_inputs(B, V, H, torch.bfloat16, seed=40 + seed, device=DEV)
        trace = {}
        exp = ref.generate_proposals(ids, hid, S, trace=trace)
        got = m.generate_proposals(ids, hid, S)
        agree += (got == exp).sum().item()
        total += got.numel()
        # teacher-forced states
        prev, tok = hid.unsqueeze(1), ids.unsqueeze(1)
        for s in range(S):
            st = m.generate_states(tok, prev, s).flatten(0, 1)
            worst = max(worst, relerr(st, trace[f"state{s}"]))
            prev, tok = trace[f"state{s}"].unsqueeze(1), exp[:, s:s + 1]
    print(f"mlp {V}x{D} tie={tie}: free-running agreement {agree / total:.5f} on {total} tokens, "
          f"worst teacher-forced state relerr {worst:.2e}")
    assert agree / total >= 0.999
    assert worst <= 1e-2


def test_fp32_exact_tokens_and_states():
    V, H, D, S, B = 5000, 512, 256, 3, 33
    w = M.make_mlp_weights(V, H, D, S, True, torch.float32, seed=2, device=DEV)
    ref = M.RefMLPSpeculator(w, V, S, True)
    m = build(V, H, D, S, True, torch.float32, w, max_bs=B)
    ids, hid = O.make_inputs(B, V, H, torch.float32, seed=3, device=DEV)
    trace = {}
    exp = ref.generate_proposals(ids, hid, S, trace=trace)
    assert torch.equal(m.generate_proposals(ids, hid, S), exp)
    prev, tok = hid.unsqueeze(1), ids.unsqueeze(1)
    for s in range(S):
        st = m.generate_states(tok, prev, s).flatten(0, 1)
        assert relerr(st, trace[f"state{s}"]) <= 1e-5
        prev, tok = trace[f"state{s}"].unsqueeze(1), exp[:, s:s + 1]


@pytest.mark.parametrize("B,D", [(64, 4096), (7, 2880), (5, 40), (3, 20000), (513, 1024)])
def test_mlp_bf16x2_kernel_bit_identical_to_generic(B, D):
    from arcticinference_b200 import _capi
    torch.manual_seed(B + D)
    V = 777
    pre = (torch.randn(B, D, device=DEV) * 1.5).to(torch.bfloat16)
    emb = (torch.randn(V, D, device=DEV) * 0.02).to(torch.bfloat16)
    tok = torch.randint(0, V, (B,), device=DEV)
    lw = (torch.randn(D, device=DEV) * 0.3 + 1).to(torch.bfloat16)
    lb = (torch.randn(D, device=DEV) * 0.8).to(torch.bfloat16)
    outs = []
    for generic in (0, 1):
        state = torch.empty(B, D, device=DEV, dtype=torch.bfloat16)
        # production kernel from the release library; the generic comparator is a debug-flavour switch
        lib = _capi.debug_lib() if generic else _capi.lib()
        prev = lib.ab200_debug_mlp_generic(1) if generic else 0
        try:
            _capi.check(lib.ab200_sum_mlp_cell(pre.data_ptr(), D, emb.data_ptr(), tok.data_ptr(), None,
                                               lw.data_ptr(), lb.data_ptr(), state.data_ptr(), B, D, V, 23.07,
                                               1e-6, _capi.AB200_BF16, 1, None, None, None, None,
                                               torch.cuda.current_stream().cuda_stream))
            torch.cuda.synchronize()
        finally:
            if generic:
                lib.ab200_debug_mlp_generic(prev)
        outs.append(state)
    assert torch.equal(outs[0].view(torch.int16), outs[1].view(torch.int16))


def test_too_many_tokens_raises_and_graph_padding():
    V, H, D, S = 1000, 64, 64, 3
    w = M.make_mlp_weights(V, H, D, S, True, torch.bfloat16, seed=0, device=DEV)
    m = build(V, H, D, S, True, torch.bfloat16, w, max_bs=16)
    ids, hid = O.make_inputs(5, V, H, torch.bfloat16, seed=1, device=DEV)
    with pytest.raises(ValueError):
        m.generate_proposals(ids, hid, S + 1)
    a = m.generate_proposals(ids, hid, S)              # padded to 6, graph
    m2 = build(V, H, D, S, True, torch.bfloat16, w, max_bs=16, enforce_eager=True)
    assert torch.equal(a, m2.generate_proposals(ids, hid, S))
    assert m.generate_proposals(ids[:0], hid[:0], S).shape == (0, S)

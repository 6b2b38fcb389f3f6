"""GPU: ``ArcticLSTMSpeculator(method = "sum_rnn")`` with list-valued dims (VERDICT r1 missing #7) - our class
(ab200_linear + ab200_sum_mlp_cell + ab200_head_argmax through the C ABI; one stage per stack: the fused MLP loop)
vs (i) golden vectors of the unmodified reference class, (ii) the op-for-op oracle on the same CUDA device,
(iii) the unmodified reference class itself on CUDA (six files under baseline/_ref, skip-if-absent)."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import speculator_ref as O
from oracle import sum_rnn_ref as R

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
CASES = sorted(os.path.basename(p)[len("sumrnn_"):-4] for p in glob.glob(os.path.join(GOLDEN, "sumrnn_*.npz")))


def build(V, H, emb, proj, inner, n_predict, tie, dtype, w, max_bs=64, **kw):
    from arcticinference_b200.sum_rnn_speculator import ArcticSumRNNSpeculator
    m = ArcticSumRNNSpeculator(vocab_size=V, input_hidden_dim=H, emb_dim=emb, proj_dim=proj, inner_dim=inner,
                               n_predict=n_predict, tie_weights=tie, max_num_seqs=max_bs, dtype=dtype, device=DEV,
                               quantize_lm_head=False, **kw)
    m.load_weights(w.items())
    return m


def relerr(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("name", CASES)
def test_against_reference_golden_vectors(name):
    """CPU-generated goldens: fp32 -> exact tokens, states <= 1e-5; bf16 -> teacher-forced states <= 2e-2 (the
    reference's CPU bf16 kernels round alpha / GELU / the GEMM sums differently from CUDA's, see DESIGN 2, and a
    three-stage stack chains up to eight bf16 GEMMs per step; the tight bf16 bound is the CUDA-oracle test below)."""
    z = np.load(os.path.join(GOLDEN, f"sumrnn_{name}.npz"))
    dtype = torch.bfloat16 if name.endswith("bf16") else torch.float32
    V, H, n_predict, S, B, tie = (int(v) for v in z["meta"])
    emb, proj, inner = (str(x) for x in z["dims"])

    def t(key):
        a = torch.from_numpy(z[key])
        return (a.view(torch.bfloat16) if a.dtype == torch.int16 else a).to(DEV)

    w = R.make_sumrnn_weights(V, H, emb, proj, inner, n_predict, bool(tie), dtype, seed=0, device=DEV)
    m = build(V, H, emb, proj, inner, n_predict, bool(tie), dtype, w, alpha_flavor="cpu")
    assert {k for k, _ in m.named_parameters()} == set(w)                 # the reference's own parameter names
    prev, tok = t("hidden").unsqueeze(1), t("ids").unsqueeze(1)
    for s in range(S):
        st = m.generate_states(tok, prev, s)
        tol = 1e-5 if dtype == torch.float32 else 2e-2
        assert relerr(st.flatten(0, 1), t(f"state{s}")) <= tol, f"state step {s}"
        prev, tok = t(f"state{s}").unsqueeze(1), t("tokens")[:, s:s + 1]
    got = m.generate_proposals(t("ids"), t("hidden"), S)
    if dtype == torch.float32:
        assert torch.equal(got, t("tokens"))
    else:
        assert (got == t("tokens")).float().mean().item() >= 0.9


@pytest.mark.parametrize("V,H,emb,proj,inner,S,B,tie", [
    (32000, 4096, "1024.1024", "1024.1024", "1024.2048", 3, 64, True),
    (32000, 2048, "512.512.512", "512.512", "512.256.1024", 4, 48, False),     # three / two / three stages, untied
    (128256, 4096, "4096", "4096", "4096", 3, 64, True),                       # one stage each: the fused MLP loop
    (50000, 2880, "2880", "2880.2880", "2880.2880", 3, 33, True),              # only proj / ln stacked, D not a power of 2
])
def test_bf16_agreement_with_cuda_oracle(V, H, emb, proj, inner, S, B, tie):
    w = R.make_sumrnn_weights(V, H, emb, proj, inner, S, tie, torch.bfloat16, seed=0, device=DEV)
    ref = R.RefSumRNNSpeculator(w, V, S, emb, proj, inner, tie)
    m = build(V, H, emb, proj, inner, S, tie, torch.bfloat16, w, max_bs=B)
    agree = total = 0
    worst = 0.0
    with torch.inference_mode():
        for seed in range(12):
            ids, hid = O.make_inputs(B, V, H, torch.bfloat16, seed=70 + seed, device=DEV)
            exp = ref.generate_proposals(ids, hid, S)
            got = m.generate_proposals(ids, hid, S)
            agree += (got == exp).sum().item()
            total += got.numel()
            prev, tok = hid.unsqueeze(1), ids.unsqueeze(1)
            for s in range(S):                                                  # teacher-forced states
                st_ref = ref.generate_states(tok, prev, s)
                worst = max(worst, relerr(m.generate_states(tok, prev, s), st_ref))
                prev, tok = st_ref, exp[:, s:s + 1]
    print(f"sum_rnn {emb}/{proj}/{inner} tie={tie}: free-running agreement {agree / total:.5f} on {total} tokens, "
          f"worst teacher-forced state relerr {worst:.2e}")
    assert agree / total >= 0.999
    assert worst <= 1e-2


def test_fp32_exact_tokens_and_states():
    V, H, emb, proj, inner, S, B = 5000, 512, "256.256", "256.256", "256.128", 3, 33
    w = R.make_sumrnn_weights(V, H, emb, proj, inner, S, True, torch.float32, seed=2, device=DEV)
    ref = R.RefSumRNNSpeculator(w, V, S, emb, proj, inner, True)
    m = build(V, H, emb, proj, inner, S, True, torch.float32, w, max_bs=B)
    ids, hid = O.make_inputs(B, V, H, torch.float32, seed=3, device=DEV)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        exp = ref.generate_proposals(ids, hid, S)
        assert torch.equal(m.generate_proposals(ids, hid, S), exp)
        prev, tok = hid.unsqueeze(1), ids.unsqueeze(1)
        for s in range(S):
            st_ref = ref.generate_states(tok, prev, s)
            assert relerr(m.generate_states(tok, prev, s), st_ref) <= 1e-5
            prev, tok = st_ref, exp[:, s:s + 1]
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def test_contract_graph_equals_eager_and_errors():
    V, H, emb, proj, inner, S = 1000, 64, "64.64", "64", "64.128", 3
    w = R.make_sumrnn_weights(V, H, emb, proj, inner, S, True, torch.bfloat16, seed=0, device=DEV)
    m = build(V, H, emb, proj, inner, S, True, torch.bfloat16, w, max_bs=32)
    e = build(V, H, emb, proj, inner, S, True, torch.bfloat16, w, max_bs=32, enforce_eager=True)
    for B in (1, 5, 17, 32, 40):                       # 40 > max graph batch: eager path of the graph-mode model
        ids, hid = O.make_inputs(B, V, H, torch.bfloat16, seed=B, device=DEV)
        a, b = m.generate_proposals(ids, hid, S), e.generate_proposals(ids, hid, S)
        assert a.shape == (B, S) and a.dtype == torch.long and torch.equal(a, b)
        assert torch.equal(m.generate_proposals(ids, hid, S), a)          # replay
    assert m.generate_proposals(torch.empty(0, dtype=torch.long, device=DEV),
                                torch.empty(0, H, dtype=torch.bfloat16, device=DEV), S).shape == (0, S)
    ids, hid = O.make_inputs(4, V, H, torch.bfloat16, seed=1, device=DEV)
    with pytest.raises(ValueError, match="Max speculative tokens"):
        m.generate_proposals(ids, hid, S + 1)
    from arcticinference_b200.sum_rnn_speculator import ArcticSumRNNSpeculator
    with pytest.raises(ValueError, match="must equal inner_dim"):          # the reference fails on these as well
        ArcticSumRNNSpeculator(vocab_size=V, input_hidden_dim=H, emb_dim="64.96", proj_dim="64", inner_dim="64",
                               n_predict=3, device=DEV)


def test_against_the_unmodified_reference_class_on_cuda():
    from oracle import ref_class
    if ref_class.reference_root() is None:
        pytest.skip("reference sources not shipped (baseline/_ref/arctic_inference missing)")
    from oracle.gen_golden_sumrnn import build_reference
    spec = ref_class.load_reference_module()
    V, H, emb, proj, inner, S, B = 32000, 1024, "512.512", "512.512", "512.1024", 3, 48
    w = R.make_sumrnn_weights(V, H, emb, proj, inner, S, True, torch.bfloat16, seed=0, device=DEV)
    agree = total = 0
    with torch.inference_mode():
        ref = build_reference(spec, V, H, emb, proj, inner, S, True, B, torch.bfloat16, w, device=DEV)
        port = R.RefSumRNNSpeculator(w, V, S, emb, proj, inner, True)
        m = build(V, H, emb, proj, inner, S, True, torch.bfloat16, w, max_bs=B)
        for seed in range(8):
            ids, hid = O.make_inputs(B, V, H, torch.bfloat16, seed=900 + seed, device=DEV)
            exp = ref.generate_proposals(ids, hid, S).clone()
            assert torch.equal(port.generate_proposals(ids, hid, S), exp)     # port == class on this device too
            got = m.generate_proposals(ids, hid, S)
            agree += (got == exp).sum().item()
            total += exp.numel()
    print(f"sum_rnn multi-stage, unmodified reference class on CUDA vs ours: {agree}/{total}")
    assert agree / total >= 0.999


@pytest.mark.parametrize("emb,proj,inner", [("512.512", "512.512", "512.1024"), ("512", "512", "512")])
def test_one_call_drop_in_routes_sum_rnn(emb, proj, inner):
    """integration.patch_lstm_speculator(): the unmodified reference class with method = "sum_rnn" (stacked or single
    dims) answers through the native twin, which shares the module's parameter storage."""
    from oracle import ref_class
    if ref_class.reference_root() is None:
        pytest.skip("reference sources not shipped (baseline/_ref/arctic_inference missing)")
    from arcticinference_b200 import integration
    from oracle.gen_golden_sumrnn import build_reference
    spec = ref_class.load_reference_module()
    cls = spec.ArcticLSTMSpeculator
    V, H, S, B = 32000, 1024, 3, 24
    w = R.make_sumrnn_weights(V, H, emb, proj, inner, S, True, torch.bfloat16, seed=0, device=DEV)
    with torch.inference_mode():
        ref = build_reference(spec, V, H, emb, proj, inner, S, True, B, torch.bfloat16, w, device=DEV)
        ids, hid = O.make_inputs(B, V, H, torch.bfloat16, seed=21, device=DEV)
        before = ref.generate_proposals(ids, hid, S).clone()
        try:
            integration.patch_lstm_speculator(cls)
            after = ref.generate_proposals(ids, hid, S)
            twin = ref._b200_impl
            assert type(twin).__name__ == "ArcticSumRNNSpeculator"
            mine = dict(twin.named_parameters())
            shared = [k for k, p in ref.named_parameters() if mine[k].data_ptr() == p.data_ptr()]
            assert "head.0.weight" in shared and "emb.0.0.weight" in shared and "proj.0.0.weight" in shared
            assert torch.equal(after, twin.generate_proposals(ids, hid, S))
            assert (after == before).float().mean().item() >= 0.999
            with pytest.raises(ValueError, match="Max speculative tokens"):
                ref.generate_proposals(ids, hid, S + 1)
        finally:
            integration.unpatch_lstm_speculator(cls)
        assert torch.equal(ref.generate_proposals(ids, hid, S), before)

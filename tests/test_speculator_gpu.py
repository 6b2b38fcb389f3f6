"""GPU parity for hot path A (Arctic LSTM speculator draft head), through the C ABI.

Oracle: oracle/speculator_ref.py.  The op-for-op restatement is pinned bit-exactly to the
reference class (tests/test_oracle_speculator.py); run on the CUDA device it IS the reference's
deployed eager path (same ATen CUDA kernels, same rounding points), so it is the bf16 agreement
oracle.  The CPU golden vectors (reference class run on CPU) are checked as well.

Tolerances (BASELINE.json north_star): hidden / cell states <= 1e-2 relative in bf16, <= 1e-5 in
fp32; draft tokens >= 99.9 % agreement in bf16, exact in fp32.
"""
import ctypes
import glob
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN
from oracle import speculator_ref as O

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
BF16_STATE_TOL = 1e-2
FP32_STATE_TOL = 1e-5
BF16_TOKEN_AGREEMENT = 0.999


def relerr(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def capi():
    from arcticinference_b200 import _capi
    return _capi


def stream():
    return torch.cuda.current_stream().cuda_stream


def make_model(V, H, D, n_predict, B, dtype, w, **kw):
    from arcticinference_b200.speculator import ArcticLSTMSpeculator
    m = ArcticLSTMSpeculator(vocab_size=V, input_hidden_dim=H, inner_dim=D, n_predict=n_predict,
                             max_num_seqs=B, dtype=dtype, device=DEV, **kw)
    m.load_weights(w.items())
    if "quantize_lm_head" not in kw:
        # the oracle recipe (`model.qhead = None`): bf16 head at every batch size -
        # by default (like the reference, :435, :726-727) batches <= 32 would go through the FP8 qhead
        m.qhead = None
    return m


# ------------------------------------------------------------------------------ kernels, one by one
@pytest.mark.parametrize("rows,cols", [(1, 64), (5, 96), (64, 4096), (33, 2880), (7, 8192), (3, 16384)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_rms_scale(capi, rows, cols, dtype):
    torch.manual_seed(0)
    x = torch.randn(rows, cols, device=DEV).to(dtype)
    y = torch.empty_like(x)
    inv = float(torch.tensor(1.0) / torch.tensor(O.SQRT2))
    capi.check(capi.lib().ab200_rms_scale(x.data_ptr(), y.data_ptr(), rows, cols, cols, cols, 1e-6, inv,
                                          capi.torch_dtype_code(dtype), 1, stream()))
    ref = (O.rms_layernorm(x.unsqueeze(1)) / O.SQRT2).squeeze(1)      # the reference's ops, on CUDA
    if dtype == torch.bfloat16:
        assert (y == ref).float().mean().item() > 0.999               # reduction order may flip a row
        assert relerr(y, ref) < 2e-3
    else:
        torch.testing.assert_close(y, ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("B,D,V", [(1, 64, 100), (5, 128, 1000), (64, 4096, 5000), (33, 2880, 3000),
                                   (16, 8192, 2000)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_fused_cell_against_oracle(capi, B, D, V, dtype):
    """One cell step, teacher-forced: same gate pre-activations, tokens and previous cell."""
    torch.manual_seed(1)
    w = O.make_weights(V, D, D, dtype, seed=3, device=DEV)
    ref = O.RefLSTMSpeculator(w, V, 3)
    x = torch.randn(B, 1, D, device=DEV).to(dtype)
    tok = torch.randint(0, V, (B, 1), device=DEV)
    cprev = (torch.randn(B, 1, D, device=DEV) * 0.5).to(dtype)
    s_ref, c_ref = ref.generate_states(tok, x, 1, cprev)
    gates = F.linear(x.squeeze(1), w["projs.1.weight"])
    state = torch.empty(B, D, device=DEV, dtype=dtype)
    cell = torch.empty(B, D, device=DEV, dtype=dtype)
    alpha = O.emulated_alpha(3, D, dtype, "cuda")
    capi.check(capi.lib().ab200_sum_lstm_cell(
        gates.data_ptr(), 4 * D, w["forget_emb.0.weight"].data_ptr(), tok.data_ptr(), None,
        cprev.data_ptr(), w["cell_ln.0.weight"].data_ptr(), w["cell_ln.0.bias"].data_ptr(),
        w["state_ln.0.weight"].data_ptr(), w["state_ln.0.bias"].data_ptr(), state.data_ptr(),
        cell.data_ptr(), B, D, V, alpha, 1e-6, capi.torch_dtype_code(dtype), 1, None, None, None, None, stream()))
    if dtype == torch.bfloat16:
        assert relerr(state, s_ref.squeeze(1)) < BF16_STATE_TOL
        assert relerr(cell, c_ref.squeeze(1)) < BF16_STATE_TOL
        # DESIGN.md 4.1 claims BIT equality with eager PyTorch (every rounding point emulated, same gate
        # pre-activations): assert exactly that, on the raw bits
        assert torch.equal(state.view(torch.int16), s_ref.squeeze(1).view(torch.int16))
        assert torch.equal(cell.view(torch.int16), c_ref.squeeze(1).view(torch.int16))
    else:
        assert relerr(state, s_ref.squeeze(1)) < FP32_STATE_TOL
        assert relerr(cell, c_ref.squeeze(1)) < FP32_STATE_TOL


def test_fused_cell_without_emulation_stays_in_tolerance(capi):
    B, D, V = 64, 4096, 5000
    w = O.make_weights(V, D, D, torch.bfloat16, seed=3, device=DEV)
    ref = O.RefLSTMSpeculator(w, V, 3)
    torch.manual_seed(2)
    x = torch.randn(B, 1, D, device=DEV).to(torch.bfloat16)
    tok = torch.randint(0, V, (B, 1), device=DEV)
    s_ref, c_ref = ref.generate_states(tok, x, 1, torch.zeros(B, 1, D, device=DEV, dtype=torch.bfloat16))
    gates = F.linear(x.squeeze(1), w["projs.1.weight"])
    state = torch.empty(B, D, device=DEV, dtype=torch.bfloat16)
    cell = torch.empty_like(state)
    capi.check(capi.lib().ab200_sum_lstm_cell(
        gates.data_ptr(), 4 * D, w["forget_emb.0.weight"].data_ptr(), tok.data_ptr(), None, None,
        w["cell_ln.0.weight"].data_ptr(), w["cell_ln.0.bias"].data_ptr(),
        w["state_ln.0.weight"].data_ptr(), w["state_ln.0.bias"].data_ptr(), state.data_ptr(),
        cell.data_ptr(), B, D, V, O.emulated_alpha(3, D, torch.bfloat16, "cuda"), 1e-6,
        capi.AB200_BF16, 0, None, None, None, None, stream()))
    assert relerr(state, s_ref.squeeze(1)) < BF16_STATE_TOL
    assert relerr(cell, c_ref.squeeze(1)) < BF16_STATE_TOL


@pytest.mark.parametrize("M,N,K", [(1, 16, 8), (5, 256, 96), (64, 256, 64), (64, 16384, 4096), (128, 16384, 4096),
                                   (96, 11520, 2880), (129, 4096, 512), (256, 2048, 1024), (17, 1000 // 8 * 8, 72),
                                   (256, 16384, 4096), (257, 4096, 512), (512, 4096, 1024), (700, 2048, 256),
                                   # CTA pairs with two k-blocks per TMA request and an ODD k-block count (45, 3): the
                                   # missing half of the last request is the tensor map's out-of-bounds zero fill
                                   (256, 11520, 2880), (300, 2048, 192), (512, 11520, 2880)])
def test_tcgen05_linear_matches_cublas_and_simt(capi, M, N, K):
    torch.manual_seed(3)
    x = torch.randn(M, K, device=DEV).to(torch.bfloat16)
    w = (torch.randn(N, K, device=DEV) * 0.05).to(torch.bfloat16)
    out = torch.full((M, N), float("nan"), device=DEV, dtype=torch.bfloat16)
    capi.check(capi.lib().ab200_linear(x.data_ptr(), w.data_ptr(), out.data_ptr(), M, N, K, K, K, N,
                                       capi.AB200_BF16, stream()))
    ref = F.linear(x, w)                                               # cuBLAS, what the reference calls
    ref32 = F.linear(x.float(), w.float())
    assert not torch.isnan(out).any()
    # fp32 accumulation in a different order, one bf16 rounding: equal except for rare 1-ulp flips
    assert (out == ref).float().mean().item() > 0.995
    assert ((out.float() - ref32).abs().max() / ref32.abs().max()).item() < 8e-3
    simt = torch.empty_like(out)
    capi.check(capi.lib().ab200_linear_simt_bf16(x.data_ptr(), w.data_ptr(), simt.data_ptr(), M, N, K, K, K,
                                                 N, stream()))
    assert (out == simt).float().mean().item() > 0.995


def test_linear_fp32_exact_mode(capi):
    M, N, K = 37, 1000, 520
    torch.manual_seed(4)
    x = torch.randn(M, K, device=DEV)
    w = torch.randn(N, K, device=DEV) * 0.05
    out = torch.empty(M, N, device=DEV)
    capi.check(capi.lib().ab200_linear(x.data_ptr(), w.data_ptr(), out.data_ptr(), M, N, K, K, K, N,
                                       capi.AB200_F32, stream()))
    ref = F.linear(x.double(), w.double())
    assert relerr(out, ref) < 1e-6


@pytest.mark.parametrize("M,V,K", [(1, 100, 64), (64, 1000, 64), (64, 4096, 512), (5, 2051, 128),
                                   (64, 128256, 4096), (128, 128256, 4096), (200, 32000, 1024),
                                   (64, 201088, 2880), (256, 128256, 4096), (300, 50000, 512),
                                   (512, 32000, 1024)])
def test_fused_head_argmax_matches_torch(capi, M, V, K):
    """Never-materialised logits + packed atomicMax == argmax over cuBLAS bf16 logits[:, :V],
    including the lowest-index tie-break (bf16 logits tie often: SURVEY F5)."""
    torch.manual_seed(5)
    Vp = O.pad_vocab(V)
    x = torch.randn(M, K, device=DEV).to(torch.bfloat16)
    w = torch.zeros(Vp, K, device=DEV, dtype=torch.bfloat16)
    w[:V] = (torch.randn(V, K, device=DEV) * 0.02).to(torch.bfloat16)
    packed = torch.zeros(M, dtype=torch.int64, device=DEV)
    tok = torch.empty(M, dtype=torch.long, device=DEV)
    lib = capi.lib()
    counter = torch.zeros(1, dtype=torch.int32, device=DEV)
    capi.check(lib.ab200_head_argmax(x.data_ptr(), w.data_ptr(), M, Vp, K, V, K, K, packed.data_ptr(),
                                     tok.data_ptr(), 1, counter.data_ptr(), capi.AB200_BF16, None, 0, stream()))
    tok2 = torch.empty_like(tok)
    capi.check(lib.ab200_argmax_unpack(packed.data_ptr(), tok2.data_ptr(), M, 1, stream()))
    assert torch.equal(tok, tok2)          # last-CTA token write == separate unpack of the packed maxima
    logits = F.linear(x, w)[:, :V]
    ref = torch.argmax(logits, -1)
    agree = (tok == ref).float().mean().item()
    # A disagreement is legitimate only where bf16 rounding of two differently-ordered fp32 accumulations
    # can change the winner: our logit and cuBLAS's may each sit one bf16 ulp away from the other's.
    top2 = logits.float().topk(2, dim=-1).values
    best, margin = top2[:, 0], top2[:, 0] - top2[:, 1]
    ulp = torch.exp2(torch.floor(torch.log2(best.abs().clamp_min(1e-30))) - 7)      # bf16: 8 significand bits
    picked = logits.gather(1, tok.clamp(0, V - 1).unsqueeze(1)).squeeze(1).float()
    assert ((picked - best).abs() <= 2 * ulp).all()                    # never further than the rounding noise
    decided = margin > 2 * ulp                                         # rows whose winner rounding cannot flip
    assert torch.equal(tok[decided], ref[decided]), "token differs although the top-2 margin exceeds 2 bf16 ulps"
    assert agree >= 0.98 or M < 16, agree


def test_argmax_tie_break_is_lowest_index(capi):
    """Construct exact ties: duplicated head rows must resolve to the lower token id."""
    M, V, K = 32, 4096, 256
    torch.manual_seed(6)
    x = torch.randn(M, K, device=DEV).to(torch.bfloat16)
    w = (torch.randn(V, K, device=DEV) * 0.02).to(torch.bfloat16)
    ref0 = torch.argmax(F.linear(x, w), -1)
    # copy each row's winning vocab row to a far-away higher index -> exact tie
    w2 = w.clone()
    for r in range(M):
        w2[(int(ref0[r]) + 2000) % V if int(ref0[r]) + 2000 < V else V - 1 - r] = w[ref0[r]]
    ref = torch.argmax(F.linear(x, w2), -1)
    packed = torch.zeros(M, dtype=torch.int64, device=DEV)
    tok = torch.empty(M, dtype=torch.long, device=DEV)
    lib = capi.lib()
    capi.check(lib.ab200_head_argmax(x.data_ptr(), w2.data_ptr(), M, V, K, V, K, K, packed.data_ptr(),
                                     None, 0, None, capi.AB200_BF16, None, 0, stream()))
    capi.check(lib.ab200_argmax_unpack(packed.data_ptr(), tok.data_ptr(), M, 1, stream()))
    assert torch.equal(tok, ref)


def test_argmax_rows_kernel(capi):
    M, N = 9, 5003
    torch.manual_seed(7)
    for dtype in (torch.float32, torch.bfloat16):
        lg = torch.randn(M, N, device=DEV).to(dtype)
        lg[:, 100] = lg.max() + 1
        lg[:, 4000] = lg[:, 100]                                       # tie -> 100
        packed = torch.empty(M, dtype=torch.int64, device=DEV)
        tok = torch.empty(M, dtype=torch.long, device=DEV)
        capi.check(capi.lib().ab200_argmax_rows(lg.data_ptr(), M, N, N, capi.torch_dtype_code(dtype),
                                                packed.data_ptr(), stream()))
        capi.check(capi.lib().ab200_argmax_unpack(packed.data_ptr(), tok.data_ptr(), M, 1, stream()))
        assert torch.equal(tok, torch.argmax(lg, -1)) and (tok == 100).all()


# ------------------------------------------------------------------------------ golden vectors (CPU reference)
GOLD = sorted(glob.glob(os.path.join(GOLDEN, "speculator_*.npz")))


def _load_gold(path):
    z = np.load(path)
    dtype = torch.bfloat16 if path.endswith("bf16.npz") else torch.float32

    def t(key):
        a = torch.from_numpy(z[key])
        return (a.view(torch.bfloat16) if a.dtype == torch.int16 else a).to(DEV)

    return z, t, dtype, tuple(int(v) for v in z["meta"])


@pytest.mark.parametrize("path", [p for p in GOLD if p.endswith("fp32.npz")], ids=os.path.basename)
def test_fp32_against_reference_golden(path):
    """fp32: tokens exactly equal, states within 1e-5 of the reference class's own output."""
    z, t, dtype, (V, H, D, n_predict, S, B) = _load_gold(path)
    w = O.make_weights(V, H, D, dtype, seed=0, device=DEV)
    m = make_model(V, H, D, n_predict, B, dtype, w, enforce_eager=True)
    states = torch.empty(S, B, D, device=DEV)
    cells = torch.empty(S, B, D, device=DEV)
    out = torch.empty(B, S, dtype=torch.long, device=DEV)
    m._propose_into(t("ids"), t("hidden"), out, B, S, states, cells)
    assert torch.equal(out, t("tokens"))
    for s in range(S):
        assert relerr(states[s], t(f"state{s}")) < FP32_STATE_TOL
        assert relerr(cells[s], t(f"cell{s}")) < FP32_STATE_TOL


@pytest.mark.parametrize("path", [p for p in GOLD if p.endswith("bf16.npz")], ids=os.path.basename)
def test_bf16_states_against_reference_golden_teacher_forced(path):
    """bf16 vs the reference class run on CPU, step by step on the reference's own inputs
    (alpha in the CPU flavour = rounded to bf16): hidden and cell within 1e-2 relative."""
    z, t, dtype, (V, H, D, n_predict, S, B) = _load_gold(path)
    w = O.make_weights(V, H, D, dtype, seed=0, device=DEV)
    m = make_model(V, H, D, n_predict, B, dtype, w, enforce_eager=True, alpha_flavor="cpu")
    prev, cell, tok = t("hidden").unsqueeze(1), None, t("ids").unsqueeze(1)
    for s in range(S):
        st, ce = m.generate_states(tok, prev, s, cell)
        assert relerr(st.squeeze(1), t(f"state{s}")) < BF16_STATE_TOL, s
        assert relerr(ce.squeeze(1), t(f"cell{s}")) < BF16_STATE_TOL, s
        prev, cell, tok = t(f"state{s}").unsqueeze(1), t(f"cell{s}").unsqueeze(1), t("tokens")[:, s:s + 1]


# ------------------------------------------------------------------------------ whole loop
def _agreement_run(V, H, D, S, B, seeds, dtype=torch.bfloat16):
    w = O.make_weights(V, H, D, dtype, seed=0, device=DEV)
    ref = O.RefLSTMSpeculator(w, V, S)
    m = make_model(V, H, D, S, B, dtype, w)
    agree = total = 0
    worst_state = worst_cell = 0.0
    per_step = torch.zeros(S)
    for seed in seeds:
        ids, hid = O.make_inputs(B, V, H, dtype, seed=100 + seed, device=DEV)
        trace = []
        exp = ref.generate_proposals(ids, hid, S, trace=trace)
        got = m.generate_proposals(ids, hid, S)
        agree += (got == exp).sum().item()
        per_step += (got == exp).float().sum(0).cpu()
        total += got.numel()
        # teacher-forced states: feed the reference's previous token / state / cell
        prev, cell, tok = hid.unsqueeze(1), None, ids.unsqueeze(1)
        for s in range(S):
            st, ce = m.generate_states(tok, prev, s, cell)
            worst_state = max(worst_state, relerr(st.squeeze(1), trace[s]["state"]))
            worst_cell = max(worst_cell, relerr(ce.squeeze(1), trace[s]["cell"]))
            prev, cell, tok = trace[s]["state"].unsqueeze(1), trace[s]["cell"].unsqueeze(1), trace[s]["tokens"].unsqueeze(1)
    return agree / total, (per_step / (len(seeds) * B)).tolist(), worst_state, worst_cell


def test_bf16_token_agreement_8b_shape():
    """BASELINE configs[1]: Llama-3.1-8B-shaped speculator, bf16, batch 64, 3 draft tokens.
    4096 draft-token rows x 3 steps so that 0.1 % is resolvable (SURVEY 8d numerics)."""
    agree, per_step, ws, wc = _agreement_run(128256, 4096, 4096, 3, 64, range(64))
    print(f"8B bf16: free-running agreement {agree:.5f} per-step {per_step} worst state relerr {ws:.2e} cell {wc:.2e}")
    assert ws < BF16_STATE_TOL and wc < BF16_STATE_TOL
    assert agree >= BF16_TOKEN_AGREEMENT


def test_bf16_token_agreement_gptoss_shape():
    """GPT-OSS-120B-shaped: H = D = 2880 (not a power of two), V = 201 088, batch 64."""
    agree, per_step, ws, wc = _agreement_run(201088, 2880, 2880, 3, 64, range(16))
    print(f"gpt-oss-shape bf16: agreement {agree:.5f} per-step {per_step} state {ws:.2e} cell {wc:.2e}")
    assert ws < BF16_STATE_TOL and wc < BF16_STATE_TOL
    assert agree >= BF16_TOKEN_AGREEMENT


def test_bf16_token_agreement_70b_shape_batch256():
    """BASELINE configs[3] shape: H = 8192, D = 4096, 4 draft tokens, batch 256 (two M passes)."""
    agree, per_step, ws, wc = _agreement_run(128256, 8192, 4096, 4, 256, range(4))
    print(f"70B-shape bf16 B=256: agreement {agree:.5f} per-step {per_step} state {ws:.2e} cell {wc:.2e}")
    assert ws < BF16_STATE_TOL and wc < BF16_STATE_TOL
    assert agree >= BF16_TOKEN_AGREEMENT


def test_fp32_exact_tokens_and_states_vs_cuda_oracle():
    V, H, D, S, B = 32000, 1024, 1024, 3, 48
    w = O.make_weights(V, H, D, torch.float32, seed=0, device=DEV)
    ref = O.RefLSTMSpeculator(w, V, S)
    m = make_model(V, H, D, S, B, torch.float32, w, enforce_eager=True)
    ids, hid = O.make_inputs(B, V, H, torch.float32, seed=9, device=DEV)
    trace = []
    exp = ref.generate_proposals(ids, hid, S, trace=trace)
    states = torch.empty(S, B, D, device=DEV)
    cells = torch.empty(S, B, D, device=DEV)
    out = torch.empty(B, S, dtype=torch.long, device=DEV)
    m._propose_into(ids, hid, out, B, S, states, cells)
    assert torch.equal(out, exp)
    for s in range(S):
        assert relerr(states[s], trace[s]["state"]) < FP32_STATE_TOL
        assert relerr(cells[s], trace[s]["cell"]) < FP32_STATE_TOL


def test_generate_proposals_contract():
    """Reference behaviour of generate_proposals (:753-866): ValueError above
    max_speculative_tokens, batch padding invisible to the caller, graph == eager, shorter S."""
    V, H, D, S = 4096, 512, 256, 3
    w = O.make_weights(V, H, D, torch.bfloat16, seed=0, device=DEV)
    g = make_model(V, H, D, S, 64, torch.bfloat16, w)
    e = make_model(V, H, D, S, 64, torch.bfloat16, w, enforce_eager=True)
    ref = O.RefLSTMSpeculator(w, V, S)
    with pytest.raises(ValueError, match="Max speculative tokens"):
        g.generate_proposals(torch.zeros(2, dtype=torch.long, device=DEV),
                             torch.zeros(2, H, dtype=torch.bfloat16, device=DEV), S + 1)
    for B in (1, 5, 9, 33, 64):
        ids, hid = O.make_inputs(B, V, H, torch.bfloat16, seed=B, device=DEV)
        a, b = g.generate_proposals(ids, hid, S), e.generate_proposals(ids, hid, S)
        assert a.shape == (B, S) and a.dtype == torch.long
        assert torch.equal(a, b)
        assert (a == ref.generate_proposals(ids, hid, S)).float().mean().item() >= 0.97
        assert torch.equal(g.generate_proposals(ids, hid, 2), a[:, :2])   # fewer steps: same prefix
    assert len(g.cuda_graphs) >= 5
    assert g.generate_proposals(torch.zeros(0, dtype=torch.long, device=DEV),
                                torch.zeros(0, H, dtype=torch.bfloat16, device=DEV), S).shape == (0, S)


def test_generate_token_ids_matches_fused_loop():
    V, H, D, S, B = 4096, 512, 512, 3, 16
    w = O.make_weights(V, H, D, torch.bfloat16, seed=0, device=DEV)
    m = make_model(V, H, D, S, B, torch.bfloat16, w, enforce_eager=True)
    ids, hid = O.make_inputs(B, V, H, torch.bfloat16, seed=1, device=DEV)
    fused = m.generate_proposals(ids, hid, S)
    toks = m.generate_token_ids(B, S, ids.unsqueeze(1), hid.unsqueeze(1), [None] * S, None)
    assert torch.equal(torch.cat(toks, -1), fused)


def test_load_weights_fuses_gate_projections():
    """arctic_speculator.py:874-902: forget|input|output|cell -> projs, tied embeddings dropped."""
    V, H, D, S, B = 1000, 128, 64, 3, 4
    w = O.make_weights(V, H, D, torch.bfloat16, seed=0, device=DEV)
    ckpt = {"speculator.head.0.weight": w["head.0.weight"][:V], "forget_emb.0.weight": w["forget_emb.0.weight"],
            "input_emb.0.weight": torch.zeros(1), "cell_emb.0.weight": torch.zeros(1), "output_emb.0.weight": torch.zeros(1)}
    for i in (0, 1):
        f, ii, o, c = w[f"projs.{i}.weight"].chunk(4)
        ckpt.update({f"forget_proj.{i}.weight": f, f"input_proj.{i}.weight": ii,
                     f"output_proj.{i}.weight": o, f"cell_proj.{i}.weight": c})
    for k in ("cell_ln.0.weight", "cell_ln.0.bias", "state_ln.0.weight", "state_ln.0.bias"):
        ckpt[k] = w[k]
    from arcticinference_b200.speculator import ArcticLSTMSpeculator
    m = ArcticLSTMSpeculator(vocab_size=V, input_hidden_dim=H, inner_dim=D, n_predict=S, max_num_seqs=B,
                             device=DEV, enforce_eager=True)
    m.load_weights(ckpt.items())
    m.qhead = None
    for k, p in m.named_parameters():
        assert torch.equal(p, w[k]), k
    ids, hid = O.make_inputs(B, V, H, torch.bfloat16, seed=1, device=DEV)
    assert (m.generate_proposals(ids, hid, S) == O.RefLSTMSpeculator(w, V, S).generate_proposals(ids, hid, S)).all()


def test_proposer_host_round_trip():
    """arctic_proposer.py:149-166: host ids in, host numpy [B, S] out."""
    from arcticinference_b200.speculator import ArcticProposer
    V, H, D, S, B = 4096, 512, 512, 3, 8
    w = O.make_weights(V, H, D, torch.bfloat16, seed=0, device=DEV)
    m = make_model(V, H, D, S, B, torch.bfloat16, w)
    ids, hid = O.make_inputs(B, V, H, torch.bfloat16, seed=1)
    out = ArcticProposer(m).propose(ids.numpy(), hid.pin_memory(), S)
    assert isinstance(out, np.ndarray) and out.shape == (B, S) and out.dtype == np.int64
    assert np.array_equal(out, m.generate_proposals(ids.to(DEV), hid.to(DEV), S).cpu().numpy())

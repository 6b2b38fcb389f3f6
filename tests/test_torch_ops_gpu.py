"""GPU: the ADDITIVE torch ops of hot path A - every op called through
``torch.ops.arctic_inference.*`` against the oracle's ATen op sequence on the same device, plus schema /
dispatch-key / fake-kernel checks.  The reference registers its single op the same way
(csrc/custom_ops/torch_bindings.cpp:5-19: one schema, CUDA key only)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import speculator_ref as O

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


@pytest.fixture(scope="module")
def ops():
    from arcticinference_b200 import py_custom_ops
    return py_custom_ops.load_speculator_ops()


def bits(t):
    return t.view(torch.int16)


def test_speculator_rms_scale_op(ops):
    x = torch.randn(33, 1, 2880, device=DEV).to(torch.bfloat16)
    inv = float(torch.tensor(1.0) / torch.tensor(O.SQRT2))
    y = ops.speculator_rms_scale(x, 1e-6, inv)
    ref = O.rms_layernorm(x) / O.SQRT2
    assert y.shape == x.shape and (bits(y) == bits(ref)).float().mean().item() > 0.999


def test_speculator_linear_op(ops):
    x = torch.randn(64, 1, 512, device=DEV).to(torch.bfloat16)
    w = (torch.randn(2048, 512, device=DEV) * 0.05).to(torch.bfloat16)
    out = ops.speculator_linear(x, w)
    assert out.shape == (64, 1, 2048)
    assert (out == F.linear(x, w)).float().mean().item() > 0.995
    assert torch.equal(ops.linear_bf16(x, w), out)
    with pytest.raises(RuntimeError):
        ops.linear_bf16(x.float(), w.float())
    o32 = ops.speculator_linear(x.float(), w.float())                  # exact-fp32 SIMT mode
    assert torch.allclose(o32, F.linear(x.double(), w.double()).float(), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("B,D,V", [(64, 4096, 5000), (7, 2880, 3000)])
def test_sum_lstm_cell_op_bit_exact(ops, B, D, V):
    w = O.make_weights(V, D, D, torch.bfloat16, seed=3, device=DEV)
    ref = O.RefLSTMSpeculator(w, V, 3)
    torch.manual_seed(B)
    x = torch.randn(B, 1, D, device=DEV).to(torch.bfloat16)
    tok = torch.randint(0, V, (B, 1), device=DEV)
    cprev = (torch.randn(B, 1, D, device=DEV) * 0.5).to(torch.bfloat16)
    s_ref, c_ref = ref.generate_states(tok, x, 1, cprev)
    gates = F.linear(x, w["projs.1.weight"])                           # [B, 1, 4D]: the op flattens leading dims
    state, cell = ops.sum_lstm_cell(gates, w["forget_emb.0.weight"], tok, cprev, w["cell_ln.0.weight"],
                                    w["cell_ln.0.bias"], w["state_ln.0.weight"], w["state_ln.0.bias"],
                                    O.emulated_alpha(3, D, torch.bfloat16, "cuda"), 1e-6, 1)
    assert torch.equal(bits(state), bits(s_ref.squeeze(1))) and torch.equal(bits(cell), bits(c_ref.squeeze(1)))
    # first step: no previous cell state
    s0_ref, c0_ref = ref.generate_states(tok, x, 1, torch.zeros_like(cprev))
    s0, c0 = ops.sum_lstm_cell(gates, w["forget_emb.0.weight"], tok, None, w["cell_ln.0.weight"],
                               w["cell_ln.0.bias"], w["state_ln.0.weight"], w["state_ln.0.bias"],
                               O.emulated_alpha(3, D, torch.bfloat16, "cuda"), 1e-6)
    assert torch.equal(bits(s0), bits(s0_ref.squeeze(1))) and torch.equal(bits(c0), bits(c0_ref.squeeze(1)))


def test_lm_head_argmax_op(ops):
    M, V, K = 64, 32000, 1024
    Vp = O.pad_vocab(V)
    x = torch.randn(M, K, device=DEV).to(torch.bfloat16)
    w = torch.zeros(Vp, K, device=DEV, dtype=torch.bfloat16)
    w[:V] = (torch.randn(V, K, device=DEV) * 0.02).to(torch.bfloat16)
    tok = ops.lm_head_argmax(x, w, V)
    ref = torch.argmax(F.linear(x, w)[:, :V], -1)
    assert tok.dtype == torch.long and tok.shape == (M,) and (tok == ref).float().mean().item() >= 0.98
    # fp32 parity mode: exact
    t32 = ops.lm_head_argmax(x.float(), w.float(), V)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        assert torch.equal(t32, torch.argmax(F.linear(x.float(), w.float())[:, :V], -1))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def test_lm_head_argmax_fp8_op(ops):
    from arcticinference_b200 import _capi
    from oracle import fp8_ref as F8
    M, V, K = 32, 8192, 512
    x = torch.randn(M, K, device=DEV).to(torch.bfloat16)
    w = (torch.randn(V, K, device=DEV) * 0.02).to(torch.bfloat16)
    q = torch.empty(V, K, dtype=torch.uint8, device=DEV)
    s = torch.zeros(1, dtype=torch.float32, device=DEV)
    _capi.check(_capi.lib().ab200_fp8_quantize_weight(w.data_ptr(), q.data_ptr(), s.data_ptr(), V, K, K, K,
                                                      torch.cuda.current_stream().cuda_stream))
    tok = ops.lm_head_argmax_fp8(x, q, s, V)
    logits = F8.fp8_linear(x, *F8.quantize_weight(w))
    picked = logits.gather(1, tok.unsqueeze(1)).squeeze(1).float()
    assert (tok == logits.argmax(-1)).float().mean().item() >= 0.9
    assert ((picked - logits.max(-1).values.float()).abs() <= logits.max(-1).values.float().abs() * 2 ** -6 + 1e-6).all()


def test_lstm_speculator_propose_op_and_graph_capture(ops):
    from arcticinference_b200 import py_custom_ops
    V, H, D, S, B = 4096, 512, 256, 3, 48
    w = O.make_weights(V, H, D, torch.bfloat16, seed=0, device=DEV)
    ids, hid = O.make_inputs(B, V, H, torch.bfloat16, seed=1, device=DEV)
    exp = O.RefLSTMSpeculator(w, V, S).generate_proposals(ids, hid, S)
    nbytes = py_custom_ops.lstm_speculator_workspace_bytes(torch.bfloat16, V, w["head.0.weight"].shape[0], H, D, B, S)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
    out = torch.empty(B, S, dtype=torch.long, device=DEV)
    alpha = O.emulated_alpha(S, D, torch.bfloat16, "cuda")

    def call():
        ops.lstm_speculator_propose(ids, hid, w["forget_emb.0.weight"], w["projs.0.weight"], w["projs.1.weight"],
                                    w["cell_ln.0.weight"], w["cell_ln.0.bias"], w["state_ln.0.weight"],
                                    w["state_ln.0.bias"], w["head.0.weight"], None, None, V, S, alpha, 1e-6, True, 1,
                                    0, ws, out)

    call()
    assert (out == exp).float().mean().item() >= 0.97
    eager = out.clone()
    g = torch.cuda.CUDAGraph()
    out.zero_()
    with torch.cuda.graph(g):
        call()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, eager)
    with pytest.raises(RuntimeError):                                  # workspace too small -> error from the C ABI
        ops.lstm_speculator_propose(ids, hid, w["forget_emb.0.weight"], w["projs.0.weight"], w["projs.1.weight"],
                                    w["cell_ln.0.weight"], w["cell_ln.0.bias"], w["state_ln.0.weight"],
                                    w["state_ln.0.bias"], w["head.0.weight"], None, None, V, S, alpha, 1e-6, True, 1,
                                    0, ws[:256], out)


def test_ops_are_cuda_only_and_traceable(ops):
    x, w = torch.zeros(4, 64), torch.zeros(8, 64)
    with pytest.raises(NotImplementedError):
        ops.speculator_linear(x, w)                                    # no CPU kernel, no fallback
    xm = torch.empty(4, 64, device="meta", dtype=torch.bfloat16)
    wm = torch.empty(256, 64, device="meta", dtype=torch.bfloat16)
    assert ops.speculator_linear(xm, wm).shape == (4, 256)
    assert ops.lm_head_argmax(xm, wm, 250).dtype == torch.long


def test_class_is_routed_through_the_ops(ops, monkeypatch):
    """ArcticLSTMSpeculator calls the registered ops (so a torch-level consumer sees them)."""
    from arcticinference_b200.speculator import ArcticLSTMSpeculator
    V, H, D, S, B = 1000, 128, 64, 3, 4
    w = O.make_weights(V, H, D, torch.bfloat16, seed=0, device=DEV)
    m = ArcticLSTMSpeculator(vocab_size=V, input_hidden_dim=H, inner_dim=D, n_predict=S, max_num_seqs=B,
                             device=DEV, enforce_eager=True)
    m.load_weights(w.items())
    m.qhead = None
    ids, hid = O.make_inputs(B, V, H, torch.bfloat16, seed=1, device=DEV)
    seen = []
    from torch.utils._python_dispatch import TorchDispatchMode

    class Spy(TorchDispatchMode):
        def __torch_dispatch__(self, func, types, args=(), kwargs=None):
            seen.append(str(func))
            return func(*args, **(kwargs or {}))

    with Spy():
        m.generate_proposals(ids, hid, S)
        m.generate_token_ids(B, S, ids.unsqueeze(1), hid.unsqueeze(1), [None] * S, None)
    names = " ".join(seen)
    for op in ("lstm_speculator_propose", "speculator_rms_scale", "speculator_linear", "sum_lstm_cell", "lm_head_argmax"):
        assert f"arctic_inference.{op}" in names, (op, seen)


def test_ops_trace_under_torch_compile_fullgraph(ops):
    """SURVEY 8b row 4 / VERDICT r1 missing #3: the ops must be visible to torch.compile consumers.  Dynamo traces a
    function that mixes ATen ops with ours into ONE graph (fullgraph=True needs the registered fake kernels); the
    "eager" backend then runs the captured graph, i.e. our kernels - no inductor code generation is involved."""
    import torch._dynamo
    V, D, B = 4096, 256, 8
    w = O.make_weights(V, D, D, torch.bfloat16, seed=0, device=DEV)

    def step(x, tok, cprev):
        gates = ops.speculator_linear(x * 1.0, w["projs.1.weight"])
        state, cell = ops.sum_lstm_cell(gates, w["forget_emb.0.weight"], tok, cprev, w["cell_ln.0.weight"],
                                        w["cell_ln.0.bias"], w["state_ln.0.weight"], w["state_ln.0.bias"], 23.0, 1e-6, 1)
        return ops.lm_head_argmax(state, w["head.0.weight"], V) + 0, cell

    x = torch.randn(B, D, device=DEV).to(torch.bfloat16)
    tok = torch.randint(0, V, (B,), device=DEV)
    cprev = torch.zeros(B, D, device=DEV, dtype=torch.bfloat16)
    want_tok, want_cell = step(x, tok, cprev)
    torch._dynamo.reset()
    compiled = torch.compile(step, backend="eager", fullgraph=True)
    got_tok, got_cell = compiled(x, tok, cprev)
    assert torch.equal(got_tok, want_tok) and torch.equal(got_cell.view(torch.int16), want_cell.view(torch.int16))

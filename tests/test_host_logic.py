"""CPU: host-side logic that needs no GPU - batch padding rule, parameter layouts of the speculator
mirrors, weight-index rules, argument validation of the no-fallback entry points."""
import numpy as np
import pytest
import torch

from arcticinference_b200 import _capi
from arcticinference_b200.speculator import pad_vocab_size, padding_size
from oracle import speculator_ref as O


def test_padding_size_matches_reference_rule():
    # arctic_speculator.py:39-44: multiples of pow2ceil(size)//4
    expect = {1: 1, 2: 2, 3: 3, 4: 4, 5: 6, 6: 6, 7: 8, 8: 8, 9: 12, 13: 16, 17: 24, 25: 32, 33: 48, 49: 64, 64: 64,
              65: 96, 100: 128, 129: 192, 200: 256, 257: 384, 512: 512}
    for size, padded in expect.items():
        assert padding_size(size) == padded == O.padding_size(size)
    for size in range(1, 600):
        p = padding_size(size)
        assert p >= size and p < 2 * size + 1


def test_pad_vocab_size():
    assert pad_vocab_size(128256) == 128256 and pad_vocab_size(201088) == 201088
    assert pad_vocab_size(1000) == 1024 and pad_vocab_size(2051) == 2112 and pad_vocab_size(64) == 64


def test_ctypes_struct_layouts_match_the_header():
    # struct sizes as a C compiler lays them out (include/arctic_b200.h)
    import ctypes
    assert ctypes.sizeof(_capi.LstmSpeculatorDesc) == 8 + 4 * 8 + 16 + 8 * 8 + 2 * 8 + 8
    assert ctypes.sizeof(_capi.MlpSpeculatorDesc) == 8 + 4 * 8 + 16 + 7 * 8 * _capi.AB200_MLP_MAX_STAGES + 8
    assert _capi.LstmSpeculatorDesc.head.offset % 8 == 0 and _capi.MlpSpeculatorDesc.emb.offset % 8 == 0


def test_no_cpu_fallback_anywhere():
    """CPU tensors must be refused, not silently computed elsewhere."""
    from arcticinference_b200 import py_custom_ops, swiftkv
    assert py_custom_ops.try_load_torch_library()
    k = torch.zeros(4, 32, dtype=torch.bfloat16)
    c = torch.zeros(2, 2, 4, 2, 8, dtype=torch.bfloat16)
    s = [torch.ones(())]
    with pytest.raises((NotImplementedError, RuntimeError)):
        py_custom_ops.reshape_and_cache_flash_bulk(k, k, [c[0]], [c[1]], torch.zeros(4, dtype=torch.long), "auto", s, s, 2, 8)
    with pytest.raises(NotImplementedError):
        swiftkv.gather_rows([k], torch.zeros(2, dtype=torch.long))


def test_speculator_constructors_validate_without_gpu():
    from arcticinference_b200.mlp_speculator import ArcticMLPSpeculator
    from arcticinference_b200.speculator import ArcticLSTMSpeculator
    with pytest.raises(ValueError):
        ArcticLSTMSpeculator(vocab_size=100, input_hidden_dim=30, inner_dim=32, n_predict=3, device="cpu")     # H % 8
    with pytest.raises(ValueError):
        ArcticLSTMSpeculator(vocab_size=100, input_hidden_dim=32, inner_dim=32, n_predict=3, device="cpu", dtype=torch.float16)
    with pytest.raises(ValueError):
        ArcticLSTMSpeculator(vocab_size=100, input_hidden_dim=32, inner_dim=32, n_predict=3, device="cpu", tp_size=3)  # 128 % 3
    with pytest.raises(AssertionError):
        ArcticMLPSpeculator(vocab_size=100, emb_dim=32, inner_dim=32, n_predict=1, tie_weights=True, device="cpu")
    with pytest.raises(ValueError):
        ArcticMLPSpeculator(vocab_size=100, emb_dim=32, inner_dim=32, n_predict=17, tie_weights=False, device="cpu")


def test_mlp_parameter_names_and_index_rule():
    from arcticinference_b200.mlp_speculator import ArcticMLPSpeculator
    tied = ArcticMLPSpeculator(vocab_size=100, emb_dim=32, inner_dim=16, n_predict=4, tie_weights=True, device="cpu")
    assert [n for n, _ in tied.named_parameters()] == ["emb.0.weight", "proj.0.weight", "proj.1.weight", "head.0.weight",
                                                       "ln.0.weight", "ln.0.bias"]
    shapes = dict((n, tuple(p.shape)) for n, p in tied.named_parameters())
    assert shapes["emb.0.weight"] == (128, 16) and shapes["proj.0.weight"] == (16, 32) and shapes["proj.1.weight"] == (16, 16)
    assert [tied._idx(s) for s in range(4)] == [0, 0, 0, 0] and [tied._idx(s, True) for s in range(4)] == [0, 1, 1, 1]
    untied = ArcticMLPSpeculator(vocab_size=100, emb_dim=32, inner_dim=16, n_predict=3, tie_weights=False, device="cpu")
    assert len(untied.named_parameters()) == 3 * 5
    assert [untied._idx(s, True) for s in range(3)] == [0, 1, 2]


def test_lstm_parameter_layout_and_tp_partition():
    from arcticinference_b200.speculator import ArcticLSTMSpeculator
    m = ArcticLSTMSpeculator(vocab_size=1000, input_hidden_dim=48, inner_dim=32, n_predict=3, device="cpu")
    shapes = dict((n, tuple(p.shape)) for n, p in m.named_parameters())
    assert shapes == {"head.0.weight": (1024, 32), "forget_emb.0.weight": (1000, 32), "projs.0.weight": (128, 48),
                      "projs.1.weight": (128, 32), "cell_ln.0.weight": (32,), "cell_ln.0.bias": (32,),
                      "state_ln.0.weight": (32,), "state_ln.0.bias": (32,)}
    r1 = ArcticLSTMSpeculator(vocab_size=1000, input_hidden_dim=48, inner_dim=32, n_predict=3, device="cpu", tp_size=4, tp_rank=1)
    assert dict(r1.named_parameters())["head.0.weight"].shape == (256, 32)
    full = torch.arange(1024 * 32, dtype=torch.float32).reshape(1024, 32).to(torch.bfloat16)
    r1.load_weights([("head.0.weight", full)])
    assert torch.equal(dict(r1.named_parameters())["head.0.weight"], full[256:512])
    # raw checkpoints: gate projections fused in the order f, i, o, c (arctic_speculator.py:878-892)
    parts = {f"{g}_proj.1.weight": torch.full((32, 32), float(i)) for i, g in enumerate(("forget", "input", "output", "cell"))}
    m.load_weights(list(parts.items()) + [("input_emb.0.weight", torch.zeros(1))])
    w = dict(m.named_parameters())["projs.1.weight"].float()
    assert [w[i * 32, 0].item() for i in range(4)] == [0.0, 1.0, 2.0, 3.0]


def test_share_parameters_adopts_storage_without_copying():
    """The drop-in twin next to the reference module must not hold a second copy of the weights."""
    from arcticinference_b200.speculator import ArcticLSTMSpeculator
    m = ArcticLSTMSpeculator(vocab_size=1000, input_hidden_dim=48, inner_dim=32, n_predict=3, device="cpu",
                             allocate_parameters=False)
    assert all(p.numel() == 0 for _, p in m.named_parameters())
    ts = {"head.0.weight": torch.randn(1024, 32).bfloat16(), "forget_emb.0.weight": torch.randn(1000, 32).bfloat16(),
          "projs.0.weight": torch.randn(128, 48).bfloat16(), "projs.1.weight": torch.randn(128, 32).bfloat16(),
          "cell_ln.0.weight": torch.randn(32).bfloat16(), "cell_ln.0.bias": torch.randn(32).bfloat16(),
          "state_ln.0.weight": torch.randn(32).bfloat16(), "state_ln.0.bias": torch.randn(64).bfloat16()[::2]}   # strided
    shared = m.share_parameters(ts.items())
    p = dict(m.named_parameters())
    assert "head.0.weight" in shared and p["head.0.weight"].data_ptr() == ts["head.0.weight"].data_ptr()
    assert "state_ln.0.bias" not in shared and torch.equal(p["state_ln.0.bias"], ts["state_ln.0.bias"])     # copied
    ts["projs.1.weight"].zero_()
    assert not p["projs.1.weight"].any()                                  # later in-place updates are visible
    with pytest.raises(ValueError, match="no tensor given"):
        ArcticLSTMSpeculator(vocab_size=1000, input_hidden_dim=48, inner_dim=32, n_predict=3, device="cpu",
                             allocate_parameters=False).share_parameters([("head.0.weight", ts["head.0.weight"])])


def test_reference_arm_is_silent_on_other_ranks():
    """bench.py --impl reference under torchrun (N > 1): rank 0 alone runs and prints; every other rank exits 0 without
    work and without a byte on stdout (the driver reads ONE JSON line per arm)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "1"], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout == ""


def test_sum_rnn_parameter_names_and_validation():
    """ArcticSumRNNSpeculator registers exactly the parameter names / shapes of the reference's
    ArcticLSTMSpeculator(method="sum_rnn") Sequential stacks (arctic_speculator.py:476-536) - compared with the oracle's
    weight factory, which the goldens pin to the unmodified class - and rejects what the reference cannot run."""
    from arcticinference_b200.sum_rnn_speculator import ArcticSumRNNSpeculator, parse_dims
    from oracle import sum_rnn_ref as R
    assert parse_dims("64.128") == [64, 128] and parse_dims(64) == [64] and parse_dims((8, 16)) == [8, 16]
    for emb, proj, inner, tie in (("64.64", "64.64", "64.128", True), ("128.128.128", "128.128", "128.64.256", False),
                                  ("64", "64", "64", True), ("64", "64.64", "64", False)):
        m = ArcticSumRNNSpeculator(vocab_size=1000, input_hidden_dim=96, emb_dim=emb, proj_dim=proj, inner_dim=inner,
                                   n_predict=3, tie_weights=tie, device="cpu", max_num_seqs=8)
        w = R.make_sumrnn_weights(1000, 96, emb, proj, inner, 3, tie, torch.bfloat16)
        assert {k: tuple(v.shape) for k, v in m.named_parameters()} == {k: tuple(v.shape) for k, v in w.items()}
        m.load_weights(w.items())
        got = dict(m.named_parameters())
        assert all(torch.equal(got[k], w[k]) for k in w)
    with pytest.raises(ValueError, match="must equal inner_dim"):
        ArcticSumRNNSpeculator(vocab_size=1000, input_hidden_dim=96, emb_dim="64.96", proj_dim="64", inner_dim="64",
                               n_predict=3, device="cpu")
    with pytest.raises(ValueError, match="must equal inner_dim"):
        ArcticSumRNNSpeculator(vocab_size=1000, input_hidden_dim=96, emb_dim="64", proj_dim="32", inner_dim="64",
                               n_predict=3, device="cpu")

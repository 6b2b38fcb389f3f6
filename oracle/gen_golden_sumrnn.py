"""TEST INFRASTRUCTURE ONLY - generates tests/golden/sumrnn_*.npz by running the UNMODIFIED reference class
ArcticLSTMSpeculator with ``method = "sum_rnn"`` and list-valued dims
(/root/reference/arctic_inference/vllm/spec_dec/arctic_speculator.py:476-536, :693-704) on CPU in the build
container, through the four import stubs of oracle/ref_class.py (qhead disabled).

    python oracle/gen_golden_sumrnn.py        # build container only
"""
import contextlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_class  # noqa: E402
from oracle import speculator_ref as O  # noqa: E402
from oracle import sum_rnn_ref as R  # noqa: E402
from oracle.gen_golden import to_np  # noqa: E402


class _C:
    pass


def build_reference(spec, V, H, emb, proj, inner, n_predict, tie, max_bs, dtype, weights, device="cpu"):
    hf = _C()
    hf.n_predict = n_predict
    hf.vocab_size = V
    hf.input_hidden_dim = H
    hf.emb_dim, hf.proj_dim, hf.inner_dim = emb, proj, inner          # "a.b" strings, as in the HF config
    hf.num_lookahead_tokens = n_predict
    hf.tie_weights = tie
    hf.tie_lstm_embs = True
    hf.scale_input = True
    hf.method = "sum_rnn"
    vc = _C()
    vc.model_config = _C()
    vc.model_config.hf_config = hf
    vc.model_config.enforce_eager = True
    vc.scheduler_config = _C()
    vc.scheduler_config.max_num_seqs = max_bs
    old = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        with torch.device(device), contextlib.redirect_stdout(sys.stderr):   # the class prints while constructing
            model = spec.ArcticLSTMSpeculator(vllm_config=vc)
    finally:
        torch.set_default_dtype(old)
    model.qhead = None
    model = model.to(device)
    params = {k: p for k, p in model.named_parameters() if not k.startswith("qhead")}
    assert set(params) == set(weights), (sorted(params), sorted(weights))
    for k, p in params.items():
        assert p.shape == weights[k].shape, (k, p.shape, weights[k].shape)
        p.data.copy_(weights[k])
    return model


CASES = {
    # name: (V, H, emb, proj, inner, n_predict, S, B, tie, dtype)
    "single_tied_bf16": (1000, 96, "64", "64", "64", 3, 3, 5, True, torch.bfloat16),          # the LSTM class's naming, one stage
    "multi_tied_fp32": (1000, 96, "64.64", "64.64", "64.128", 3, 3, 5, True, torch.float32),
    "multi_tied_bf16": (1000, 96, "64.64", "64.64", "64.128", 3, 3, 7, True, torch.bfloat16),
    "multi_untied_bf16": (2051, 192, "128.128.128", "128.128", "128.64.256", 3, 3, 9, False, torch.bfloat16),
    "mixed_tied_bf16": (1000, 256, "256", "256.256", "256.256", 4, 4, 33, True, torch.bfloat16),  # only proj / ln stacked
}


def main():
    spec = ref_class.load_reference_module()
    out_dir = os.path.join(ROOT, "tests", "golden")
    with torch.inference_mode():
        for name, (V, H, emb, proj, inner, n_predict, S, B, tie, dtype) in CASES.items():
            w = R.make_sumrnn_weights(V, H, emb, proj, inner, n_predict, tie, dtype, seed=0)
            ids, hid = O.make_inputs(B, V, H, dtype, seed=1)
            model = build_reference(spec, V, H, emb, proj, inner, n_predict, tie, B, dtype, w)
            tokens = model.generate_proposals(ids, hid, S).clone()
            port = R.RefSumRNNSpeculator(w, V, n_predict, emb, proj, inner, tie)
            assert torch.equal(port.generate_proposals(ids, hid, S), tokens), name
            prev, tok = hid.unsqueeze(1), ids.unsqueeze(1)
            rec = {"ids": ids.numpy(), "hidden": to_np(hid), "tokens": tokens.numpy(),
                   "meta": np.array([V, H, n_predict, S, B, int(tie)], dtype=np.int64),
                   "dims": np.array([emb, proj, inner])}
            for s in range(S):
                state = model.generate_states(tok, prev, s, None)
                rec[f"state{s}"] = to_np(state.flatten(0, 1))
                tok, prev = tokens[:, s:s + 1], state
            path = os.path.join(out_dir, f"sumrnn_{name}.npz")
            np.savez_compressed(path, **rec)
            print(name, "tokens[0] =", tokens[0].tolist(), "->", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

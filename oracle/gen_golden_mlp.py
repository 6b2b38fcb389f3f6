"""TEST INFRASTRUCTURE ONLY - generates tests/golden/mlp_speculator_*.npz by running the
UNMODIFIED reference class ArcticMLPSpeculator
(/root/reference/arctic_inference/vllm/spec_dec/arctic_speculator.py:102-401) on CPU in the build
container, through the same four import stubs as oracle/gen_golden.py (qhead disabled).

    python oracle/gen_golden_mlp.py        # build container only
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import mlp_speculator_ref as M  # noqa: E402
from oracle import speculator_ref as O  # noqa: E402
from oracle.gen_golden import _C, load_reference_module, to_np  # noqa: E402


def build_reference(spec, V, H, D, n_predict, tie, max_bs, dtype, weights):
    hf = _C()
    hf.n_predict = n_predict
    hf.vocab_size = V
    hf.emb_dim = H
    hf.inner_dim = D
    hf.num_lookahead_tokens = n_predict
    hf.tie_weights = tie
    hf.scale_input = True
    vc = _C()
    vc.model_config = _C()
    vc.model_config.hf_config = hf
    vc.model_config.enforce_eager = True
    vc.scheduler_config = _C()
    vc.scheduler_config.max_num_seqs = max_bs
    torch.set_default_dtype(dtype)
    model = spec.ArcticMLPSpeculator(vllm_config=vc)
    model.qhead = None
    params = {k: p for k, p in model.named_parameters() if not k.startswith("qhead")}
    assert set(params) == set(weights), (sorted(params), sorted(weights))
    for k, p in params.items():
        assert p.shape == weights[k].shape, (k, p.shape, weights[k].shape)
        p.data.copy_(weights[k])
    return model


CASES = {
    # name: (V, H, D, n_predict, S, B, tie, dtype)
    "tied_fp32": (1000, 96, 64, 3, 3, 5, True, torch.float32),
    "tied_bf16": (1000, 96, 64, 3, 3, 5, True, torch.bfloat16),
    "untied_bf16": (2051, 192, 128, 4, 4, 9, False, torch.bfloat16),
    "untied_fp32": (2051, 192, 128, 4, 4, 9, False, torch.float32),
    "pow2_tied_bf16": (1024, 256, 256, 3, 3, 64, True, torch.bfloat16),
}


def main():
    spec = load_reference_module()
    out_dir = os.path.join(ROOT, "tests", "golden")
    with torch.inference_mode():
        for name, (V, H, D, n_predict, S, B, tie, dtype) in CASES.items():
            w = M.make_mlp_weights(V, H, D, n_predict, tie, dtype, seed=0)
            ids, hid = O.make_inputs(B, V, H, dtype, seed=1)
            model = build_reference(spec, V, H, D, n_predict, tie, B, dtype, w)
            tokens = model.generate_proposals(ids, hid, S).clone()
            prev = hid.unsqueeze(1)
            tok = ids.unsqueeze(1)
            rec = {"ids": ids.numpy(), "hidden": to_np(hid), "tokens": tokens.numpy(),
                   "meta": np.array([V, H, D, n_predict, S, B, int(tie)], dtype=np.int64)}
            for s in range(S):
                state = model.generate_states(tok, prev, s)
                logits = model.logits_processor(model.head[s], state.flatten(0, 1))
                tok = torch.argmax(logits, dim=-1).reshape(B, -1)
                prev = state
                rec[f"state{s}"] = to_np(state.flatten(0, 1))
                rec[f"logits{s}"] = to_np(logits)
                assert torch.equal(tok.flatten(), tokens[:, s])
            path = os.path.join(out_dir, f"mlp_speculator_{name}.npz")
            np.savez_compressed(path, **rec)
            print(name, "tokens[0] =", tokens[0].tolist(), "->", path, os.path.getsize(path), "bytes")
    torch.set_default_dtype(torch.float32)


if __name__ == "__main__":
    main()

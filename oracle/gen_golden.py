"""TEST INFRASTRUCTURE ONLY - generates tests/golden/speculator_*.npz by running the
UNMODIFIED reference class (/root/reference/arctic_inference/vllm/spec_dec/arctic_speculator.py,
ArcticLSTMSpeculator) on CPU in the build container.

The reference does not import against the installed vllm 0.22, so four stubs are installed
first: a fake ``vllm.model_executor.layers.sampler``, a fake
``arctic_inference.vllm.spec_dec.fp8``, ``VLLM_USE_V1 = True`` and world-size-1 TP/SP groups.
``model.qhead = None`` forces the unquantised head for every batch size (the FP8 qhead's
arithmetic lives in vllm 0.9.2, which is absent: parity unpinned, out of scope).

Run (build container only; /root/reference does not exist on the GPU box):
    python oracle/gen_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import speculator_ref as O  # noqa: E402

from oracle.ref_class import _C, build_reference, load_reference_module  # noqa: E402,F401  (shared with the GPU test)


def to_np(t: torch.Tensor):
    if t.dtype == torch.bfloat16:
        return t.contiguous().view(torch.int16).numpy()
    return t.contiguous().numpy()


CASES = {
    # name: (V, H, D, n_predict, S, B, dtype)
    "tiny_fp32": (1000, 96, 64, 3, 3, 5, torch.float32),
    "tiny_bf16": (1000, 96, 64, 3, 3, 5, torch.bfloat16),
    "rect_fp32": (2051, 192, 128, 4, 4, 9, torch.float32),
    "rect_bf16": (2051, 192, 128, 4, 4, 9, torch.bfloat16),
    "wide_bf16": (3000, 512, 512, 3, 3, 33, torch.bfloat16),
    # power-of-two dims and a batch that is a multiple of the CPU vector width: no ATen-CPU
    # scalar-tail quirks, so the explicit-rounding emulation is bit-comparable here
    "pow2_bf16": (1024, 256, 256, 3, 3, 64, torch.bfloat16),
}


def main():
    spec = load_reference_module()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    with torch.inference_mode():
        for name, (V, H, D, n_predict, S, B, dtype) in CASES.items():
            w = O.make_weights(V, H, D, dtype, seed=0)
            ids, hid = O.make_inputs(B, V, H, dtype, seed=1)
            model = build_reference(spec, V, H, D, n_predict, B, dtype, w)
            tokens = model.generate_proposals(ids, hid, S).clone()
            # per-step trace through the reference's own methods (teacher-forced by its own tokens)
            prev = hid.unsqueeze(1)
            cell = torch.zeros(B, 1, D, dtype=dtype)
            tok = ids.unsqueeze(1)
            rec = {"ids": ids.numpy(), "hidden": to_np(hid), "tokens": tokens.numpy(),
                   "meta": np.array([V, H, D, n_predict, S, B], dtype=np.int64),
                   "ln0": to_np((model.ln0(hid.unsqueeze(1)) / spec.SQRT2).flatten(0, 1)),
                   "weights_checksum": np.array(
                       [float(sum(v.double().sum() for v in w.values()))])}
            for s in range(S):
                state, cell = model.generate_states(tok, prev, s, cell)
                logits = model.logits_processor(model.head[s], state.flatten(0, 1))
                tok = torch.argmax(logits, dim=-1).reshape(B, -1)
                prev = state
                rec[f"state{s}"] = to_np(state.flatten(0, 1))
                rec[f"cell{s}"] = to_np(cell.flatten(0, 1))
                rec[f"logits{s}"] = to_np(logits)
                assert torch.equal(tok.flatten(), tokens[:, s])
            path = os.path.join(out_dir, f"speculator_{name}.npz")
            np.savez_compressed(path, **rec)
            print(name, "tokens[0] =", tokens[0].tolist(), "->", path,
                  os.path.getsize(path), "bytes")
    torch.set_default_dtype(torch.float32)


if __name__ == "__main__":
    main()

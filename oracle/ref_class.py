"""TEST INFRASTRUCTURE ONLY - loads the UNMODIFIED reference class
(arctic_inference/vllm/spec_dec/arctic_speculator.py::ArcticLSTMSpeculator / ArcticMLPSpeculator)
so that it can serve as the oracle: on CPU in the build container (golden vectors, oracle/gen_golden*.py)
and on CUDA on the GPU box (tests/test_reference_class_gpu.py, bench.py's gpu_torch_baseline leg).

The reference does not import against the installed vllm 0.22, so four stubs are installed first
(SURVEY.md Appendix C): a fake ``vllm.model_executor.layers.sampler``, a fake
``arctic_inference.vllm.spec_dec.fp8``, ``VLLM_USE_V1 = True`` and world-size-1 TP/SP groups.
``model.qhead = None`` forces the unquantised head for every batch size (the FP8 qhead's arithmetic
lives in vllm 0.9.2, which is absent: parity unpinned).

Where the sources come from: ``/root/reference`` in the build container; on the GPU box (where that
mount does not exist) a copy of the SIX files the class needs under the git-ignored
``baseline/_ref/arctic_inference`` - written by ``ship_reference()`` (called from
``__graft_entry__.build()`` while /root/reference is mounted).  Reference sources never enter the
tracked tree.
"""
from __future__ import annotations

import contextlib
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_MOUNT = "/root/reference"
REF_SHIPPED = os.path.join(ROOT, "baseline", "_ref")

# the files ArcticLSTMSpeculator's arithmetic needs (notes)
REF_FILES = [
    "arctic_inference/__init__.py",
    "arctic_inference/vllm/__init__.py",
    "arctic_inference/vllm/spec_dec/__init__.py",
    "arctic_inference/vllm/spec_dec/arctic_speculator.py",
    "arctic_inference/vllm/spec_dec/logits_processor_opt.py",
    "arctic_inference/vllm/spec_dec/vocab_parallel_embedding.py",
]


def ship_reference() -> bool:
    """Copy REF_FILES from the read-only mount to baseline/_ref/ (git-ignored, travels with gpurun).
    Returns False when the mount is absent (GPU box: nothing to do, the copy is already there)."""
    if not os.path.isdir(os.path.join(REF_MOUNT, "arctic_inference")):
        return False
    for rel in REF_FILES:
        dst = os.path.join(REF_SHIPPED, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF_MOUNT, rel), dst)
    return True


def reference_root() -> str | None:
    for root in (REF_MOUNT, REF_SHIPPED):
        if all(os.path.exists(os.path.join(root, rel)) for rel in REF_FILES):
            return root
    return None


_SPEC = None


def load_reference_module(root: str | None = None):
    """Import the reference's arctic_speculator module (unmodified source) behind the four stubs.
    Raises RuntimeError when the sources are not available."""
    global _SPEC
    if _SPEC is not None:
        return _SPEC
    root = root or reference_root()
    if root is None:
        raise RuntimeError("reference sources not found (neither /root/reference nor baseline/_ref)")
    if root not in sys.path:
        sys.path.insert(0, root)
    m = types.ModuleType("vllm.model_executor.layers.sampler")          # stub 1 (module removed in vllm 0.22)
    m.SamplerOutput = type("SamplerOutput", (), {})
    m.get_sampler = lambda: None
    sys.modules[m.__name__] = m
    f8 = types.ModuleType("arctic_inference.vllm.spec_dec.fp8")         # stub 2 (real file needs vllm 0.9.2)

    class Fp8ConfigWithEmbedding:
        def __init__(self, *a, **k):
            pass

        def get_quant_method(self, layer, prefix=""):
            return None

        def get_name(self):
            return "fp8"

    class OriginalFp8LinearMethod:
        def __init__(self, quant_config=None):
            pass

    f8.Fp8ConfigWithEmbedding = Fp8ConfigWithEmbedding
    f8.OriginalFp8LinearMethod = OriginalFp8LinearMethod
    sys.modules[f8.__name__] = f8
    with contextlib.redirect_stdout(sys.stderr):                        # importing vLLM may print; stdout is the caller's
        import vllm.envs as envs                                         # stub 3
        import vllm.distributed.parallel_state as ps                     # stub 4
    envs.VLLM_USE_V1 = True

    class G:
        world_size = 1
        rank = 0
        rank_in_group = 0

        def all_reduce(self, x):
            return x

        def all_gather(self, x, dim=-1):
            return x

    ps._TP = G()
    ps._SP = G()
    with contextlib.redirect_stdout(sys.stderr):
        import arctic_inference.vllm.spec_dec.arctic_speculator as spec
    _SPEC = spec
    return spec


class _C:
    pass


def build_reference(spec, V, H, D, n_predict, max_bs, dtype, weights, device="cpu"):
    """Instantiate the reference's ArcticLSTMSpeculator (eager, qhead = None) on `device` and fill its
    parameters from `weights` (name -> tensor, the reference's own parameter names)."""
    import torch
    hf = _C()
    hf.n_predict = n_predict
    hf.vocab_size = V
    hf.input_hidden_dim = H
    hf.inner_dim = hf.emb_dim = hf.proj_dim = str(D)
    hf.num_lookahead_tokens = n_predict
    hf.tie_weights = hf.tie_lstm_embs = hf.scale_input = True
    hf.method = "sum_lstm"
    vc = _C()
    vc.model_config = _C()
    vc.model_config.hf_config = hf
    vc.model_config.enforce_eager = True
    vc.scheduler_config = _C()
    vc.scheduler_config.max_num_seqs = max_bs
    old_dtype = torch.get_default_dtype()
    torch.set_default_dtype(dtype)                   # the class allocates with the default dtype / device (:592-605)
    try:
        # the class prints while constructing ("CREATED NEXT PREVIOUS HIDDEN STATES", :607): keep the caller's stdout clean
        with torch.device(device), contextlib.redirect_stdout(sys.stderr):
            model = spec.ArcticLSTMSpeculator(vllm_config=vc)
    finally:
        torch.set_default_dtype(old_dtype)
    model.qhead = None
    model = model.to(device)
    params = dict(model.named_parameters())
    assert set(params) == set(weights), (sorted(params), sorted(weights))
    for k, p in params.items():
        assert p.shape == weights[k].shape, (k, p.shape, weights[k].shape)
        p.data.copy_(weights[k])
    return model

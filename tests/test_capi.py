"""CPU: the C-ABI library builds for sm_100a, loads, exports every symbol include/arctic_b200.h
declares, and the drop-in torch op is registered with the reference's schema (no compute here)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


@pytest.fixture(scope="module")
def built_lib():
    from arcticinference_b200 import _build
    return _build.build()


@pytest.fixture(scope="module")
def built_debug_lib():
    from arcticinference_b200 import _build
    return _build.build(debug=True)


def _header_symbols(debug=False):
    """Symbols include/arctic_b200.h declares: the release set, or (debug=True) the extra ones inside
    its #ifdef AB200_DEBUG blocks."""
    text = open(os.path.join(ROOT, "include", "arctic_b200.h")).read()
    dbg_blocks = re.findall(r"#ifdef AB200_DEBUG(.*?)#endif /\* AB200_DEBUG \*/", text, flags=re.S)
    pat = r"AB200_API [\w\* ]+? (ab200_\w+)\("
    dbg = sorted({m for b in dbg_blocks for m in re.findall(pat, b)})
    if debug:
        return dbg
    return sorted(set(re.findall(pat, text)) - set(dbg))


def test_header_declares_the_path():
    syms = _header_symbols()
    for required in ("ab200_kv_write_bulk", "ab200_sum_lstm_cell", "ab200_rms_scale", "ab200_linear",
                     "ab200_head_argmax", "ab200_lstm_propose", "ab200_last_error"):
        assert required in syms


def test_library_exports_every_declared_symbol(built_lib):
    handle = ctypes.CDLL(built_lib)
    for s in _header_symbols():
        assert hasattr(handle, s), f"{s} declared in include/arctic_b200.h but not exported"


def test_release_library_has_no_debug_hooks_and_reads_no_environment(built_lib):
    """tuning switches and search kernels must not ship in the product library."""
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True).stdout
    exported = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert exported and not [e for e in exported if "debug" in e], exported
    assert sorted(e for e in exported if e.startswith("ab200_")) == _header_symbols()
    csrc = os.path.join(ROOT, "arcticinference_b200", "csrc")
    for f in os.listdir(csrc):
        assert "getenv" not in open(os.path.join(csrc, f)).read(), f"{f} reads the environment"


def test_debug_library_exports_the_hooks(built_debug_lib):
    handle = ctypes.CDLL(built_debug_lib)
    dbg = _header_symbols(debug=True)
    assert "ab200_debug_set_tunable" in dbg and "ab200_debug_unary_bf16" in dbg
    for s in _header_symbols() + dbg:
        assert hasattr(handle, s), s


def test_ctypes_table_matches_header(built_lib, built_debug_lib):
    from arcticinference_b200 import _capi
    assert sorted(_capi.SIGNATURES) == _header_symbols()
    assert sorted(_capi.DEBUG_SIGNATURES) == _header_symbols(debug=True)
    lib = _capi.lib()
    assert lib.ab200_abi_version() == 5
    assert lib.ab200_last_error() == b""
    dbg = _capi.debug_lib()
    assert dbg.ab200_debug_set_tunable(b"no_such_tunable", 1) == _capi.AB200_E_INVALID
    assert dbg.ab200_debug_set_tunable(b"max_stages", 0) == 0


def test_descriptor_struct_layout():
    """ctypes mirror of struct ab200_lstm_speculator: 4+4+8*4+4+4+4(+4 pad)+8*8 + 8+8+4(+4 pad)."""
    from arcticinference_b200 import _capi
    assert ctypes.sizeof(_capi.LstmSpeculatorDesc) == 144
    assert _capi.LstmSpeculatorDesc.forget_emb.offset == 56


def test_block_n_heuristic(built_lib):
    from arcticinference_b200 import _capi
    lib = _capi.lib()
    # 16384 gate outputs on 148 SMs -> 147 tiles of 112; padded 128256 vocab -> 729 tiles of 176
    assert lib.ab200_gemm_block_n(16384, 148) == 112
    assert lib.ab200_gemm_block_n(128256, 148) == 176
    for n in (16, 100, 1000, 4096, 201088):
        bn = lib.ab200_gemm_block_n(n, 148)
        assert bn % 16 == 0 and 16 <= bn <= 256


def test_op_registered_with_reference_schema(built_lib):
    from arcticinference_b200 import py_custom_ops
    assert py_custom_ops.try_load_torch_library() is True
    assert py_custom_ops.try_load_torch_library() is True  # idempotent
    schema = str(torch.ops.arctic_inference.reshape_and_cache_flash_bulk.default._schema)
    # csrc/custom_ops/torch_bindings.cpp:6-16
    assert schema == ("arctic_inference::reshape_and_cache_flash_bulk(Tensor keys, Tensor values, "
                      "Tensor(c!)[] key_caches, Tensor(d!)[] value_caches, Tensor slot_mapping, "
                      "str kv_cache_dtype, Tensor(e)[] k_scales, Tensor(f)[] v_scales, "
                      "int num_heads, int head_size) -> ()")


def test_op_has_no_cpu_implementation(built_lib):
    """torch_bindings.cpp:17-18 registers the CUDA key only; so do we - CPU tensors must not be
    silently handled by a fallback."""
    from arcticinference_b200 import py_custom_ops
    py_custom_ops.try_load_torch_library()
    k = torch.zeros(2, 32)
    with pytest.raises((NotImplementedError, RuntimeError)):
        py_custom_ops.reshape_and_cache_flash_bulk(k, k, [torch.zeros(1, 2, 32)], [torch.zeros(1, 2, 32)],
                                                   torch.zeros(2, dtype=torch.long), "auto",
                                                   [torch.ones(())], [torch.ones(())], 2, 16)


def test_try_load_is_soft_when_library_missing(monkeypatch, built_lib):
    """py_custom_ops.py:24-37: never raise, return False."""
    from arcticinference_b200 import _capi, py_custom_ops
    monkeypatch.setattr(py_custom_ops, "_registered", False)
    monkeypatch.setattr(_capi, "_lib", None)
    monkeypatch.setattr(_capi, "LIB_PATH", "/nonexistent/libarctic_b200.so")
    assert py_custom_ops.try_load_torch_library() is False


def test_speculator_fails_loudly_without_library(monkeypatch):
    from arcticinference_b200 import _capi
    from arcticinference_b200.speculator import ArcticLSTMSpeculator
    monkeypatch.setattr(_capi, "_lib", None)
    monkeypatch.setattr(_capi, "LIB_PATH", "/nonexistent/libarctic_b200.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        ArcticLSTMSpeculator(vocab_size=64, input_hidden_dim=64, inner_dim=64, n_predict=3, device="cpu")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "arcticinference_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_host_helpers():
    from arcticinference_b200.speculator import padding_size, pad_vocab_size
    assert [padding_size(n) for n in (1, 5, 9, 33, 64, 65, 257)] == [1, 6, 12, 48, 64, 96, 384]
    assert pad_vocab_size(128256) == 128256 and pad_vocab_size(1000) == 1024


def test_speculator_ops_registered_cuda_only(built_lib):
    """additive path-A ops in the arctic_inference namespace (the reference has none, F1)."""
    from arcticinference_b200 import py_custom_ops
    ns = py_custom_ops.load_speculator_ops()
    for name in ("speculator_rms_scale", "speculator_linear", "linear_bf16", "sum_lstm_cell", "lm_head_argmax",
                 "lm_head_argmax_fp8", "lstm_speculator_propose"):
        schema = str(getattr(ns, name).default._schema)
        assert schema.startswith(f"arctic_inference::{name}("), schema
    assert "Tensor(a!) workspace, Tensor(b!) tokens_out" in str(ns.lstm_speculator_propose.default._schema)
    with pytest.raises(NotImplementedError):                  # CUDA key only: CPU tensors are not silently handled
        ns.lm_head_argmax(torch.zeros(2, 8), torch.zeros(16, 8), 16)
    meta = torch.empty(3, 1, 8, device="meta", dtype=torch.bfloat16)
    assert ns.speculator_linear(meta, torch.empty(32, 8, device="meta", dtype=torch.bfloat16)).shape == (3, 1, 32)

"""GPU: FP8 KV-cache bulk write through the reference-facing torch op.
Bit-exact on every cache byte against (i) goldens written by vLLM's per-layer op, (ii) the oracle
restatement at larger shapes, (iii) vLLM's op live (the reference's fallback path,
llama_swiftkv.py:629-656) on the same inputs."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import kv_write_fp8_ref as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SRC = {"torch.bfloat16": torch.bfloat16, "torch.float16": torch.float16, "torch.float32": torch.float32}
CASES = sorted(os.path.basename(p)[len("kvfp8_"):-4] for p in glob.glob(os.path.join(GOLDEN, "kvfp8_*.npz")))


@pytest.fixture(scope="module")
def ops():
    from arcticinference_b200 import py_custom_ops
    assert py_custom_ops.try_load_torch_library()
    return py_custom_ops


@pytest.mark.parametrize("name", CASES)
def test_against_vllm_goldens(ops, name):
    z = np.load(os.path.join(GOLDEN, f"kvfp8_{name}.npz"))
    L, T, Hkv, Dh, bs, nb = (int(v) for v in z["meta"])
    dt = SRC[str(z["src_dtype"])]
    keys = torch.from_numpy(z["keys"]).view(dt).to(DEV)
    values = torch.from_numpy(z["values"]).view(dt).to(DEV)
    caches = [torch.from_numpy(z[f"init{l}"].copy()).to(DEV) for l in range(L)]
    ks = [torch.tensor(float(s), device=DEV) for s in z["k_scales"]]
    vs = [torch.tensor(float(s), device=DEV) for s in z["v_scales"]]
    ops.reshape_and_cache_flash_bulk(keys, values, [c[0] for c in caches], [c[1] for c in caches],
                                     torch.from_numpy(z["slots"]).to(DEV), str(z["kv_cache_dtype"]), ks, vs, Hkv, Dh)
    torch.cuda.synchronize()
    for l in range(L):
        assert np.array_equal(caches[l].cpu().numpy(), z[f"final{l}"]), f"layer {l}"


def make(L, T, Hkv, Dh, bs, dt, layout, std, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    n = Hkv * Dh
    nb = 2 * ((T + bs - 1) // bs)
    keys = (torch.randn(T, L * n, generator=g) * std).to(dt)
    values = (torch.randn(T, L * n, generator=g) * std).to(dt)
    keys[0, :6] = torch.tensor([0.0, -0.0, 1e-9, 1e30, -1e30, float("inf")]).to(dt)
    slots = torch.randperm(nb * bs, generator=g)[:T].to(torch.int64)
    slots[torch.rand(T, generator=g) < 0.05] = -1
    shape = (2, nb, bs, Hkv, Dh) if layout == "flash" else (nb, 2, bs, Hkv, Dh)
    init = [torch.randint(0, 256, shape, generator=g, dtype=torch.uint8) for _ in range(L)]
    ks = [torch.tensor(0.3 + 0.21 * l) for l in range(L)]
    vs = [torch.tensor(0.02 * (l + 1)) for l in range(L)]
    return keys, values, slots, init, ks, vs


def views(caches, layout):
    if layout == "flash":
        return [c[0] for c in caches], [c[1] for c in caches]
    return [c[:, 0] for c in caches], [c[:, 1] for c in caches]


@pytest.mark.parametrize("L,T,Hkv,Dh,bs,dt,layout,kvd,std", [
    (16, 1024, 8, 128, 16, torch.bfloat16, "flash", "fp8", 1.0),
    (18, 300, 8, 64, 16, torch.bfloat16, "flashinfer", "fp8_e4m3", 4.0),
    (4, 257, 8, 128, 16, torch.float16, "flash", "fp8", 0.2),
    (3, 100, 2, 64, 8, torch.float32, "flashinfer", "fp8_e5m2", 30.0),
    (70, 64, 8, 128, 16, torch.bfloat16, "flash", "fp8", 1.0),          # > 64 layers: two launches
])
def test_against_oracle_restatement(ops, L, T, Hkv, Dh, bs, dt, layout, kvd, std):
    keys, values, slots, init, ks, vs = make(L, T, Hkv, Dh, bs, dt, layout, std, seed=L * 7 + T)
    ref = [c.clone() for c in init]
    rk, rv = views(ref, layout)
    F.kv_write_fp8(keys, values, rk, rv, slots, kvd, ks, vs, Hkv, Dh)
    dev = [c.clone().to(DEV) for c in init]
    dk, dv = views(dev, layout)
    ops.reshape_and_cache_flash_bulk(keys.to(DEV), values.to(DEV), dk, dv, slots.to(DEV), kvd,
                                     [s.to(DEV) for s in ks], [s.to(DEV) for s in vs], Hkv, Dh)
    torch.cuda.synchronize()
    for l in range(L):
        got, exp = dev[l].cpu(), ref[l]
        if not torch.equal(got, exp):
            bad = (got != exp).nonzero()
            raise AssertionError(f"layer {l}: {bad.shape[0]} bytes differ, first at {bad[0].tolist()}: "
                                 f"{got[tuple(bad[0])].item()} vs {exp[tuple(bad[0])].item()}")


def test_against_vllm_op_live(ops):
    vops = pytest.importorskip("vllm._custom_ops")
    L, T, Hkv, Dh, bs = 6, 2048, 8, 128, 16
    keys, values, slots, init, ks, vs = make(L, T, Hkv, Dh, bs, torch.bfloat16, "flash", 1.5, seed=99)
    n = Hkv * Dh
    a = [c.clone().to(DEV) for c in init]
    b = [c.clone().to(DEV) for c in init]
    dk, dv, ds = keys.to(DEV), values.to(DEV), slots.to(DEV)
    dks, dvs = [s.to(DEV) for s in ks], [s.to(DEV) for s in vs]
    try:
        for l in range(L):
            vops.reshape_and_cache_flash(dk[:, l * n:(l + 1) * n].reshape(T, Hkv, Dh).contiguous(),
                                         dv[:, l * n:(l + 1) * n].reshape(T, Hkv, Dh).contiguous(),
                                         b[l][0], b[l][1], ds, "fp8", dks[l], dvs[l])
    except Exception as e:  # vllm._C not loadable on this box
        pytest.skip(f"vLLM cache op unavailable: {e}")
    ops.reshape_and_cache_flash_bulk(dk, dv, [c[0] for c in a], [c[1] for c in a], ds, "fp8", dks, dvs, Hkv, Dh)
    torch.cuda.synchronize()
    for l in range(L):
        assert torch.equal(a[l], b[l]), f"layer {l}"


def test_division_shortcut_exact_for_all_bf16_values(ops):
    """q = RN(x / s) through the reciprocal + fma correction: every bf16 input x, several scales,
    compared with torch's IEEE division on the device (via a 1-token-per-value write)."""
    bits = torch.arange(65536, dtype=torch.int32).to(torch.int16).view(torch.bfloat16)
    finite = bits.float().isfinite()
    x = bits[finite]
    T = x.numel()
    Hkv, Dh, bs = 1, 8, 16
    keys = x.reshape(T, 1).repeat(1, 8).contiguous().to(DEV)
    nb = (T + bs - 1) // bs
    for s in (1.0, 0.3, 0.0123, 7.77, 1e-3, 123.456):
        cache = torch.zeros(2, nb, bs, Hkv, Dh, dtype=torch.uint8, device=DEV)
        sc = torch.tensor(s, device=DEV)
        slots = torch.arange(T, device=DEV)
        ops.reshape_and_cache_flash_bulk(keys, keys, [cache[0]], [cache[1]], slots, "fp8", [sc], [sc], Hkv, Dh)
        exp = F.quantize(x.to(DEV), sc, "fp8")
        got = cache[0].reshape(-1, 8)[:T, 0]
        assert torch.equal(got, exp), f"scale {s}: {(got != exp).sum().item()} of {T} differ"


def test_errors(ops):
    keys = torch.zeros(4, 2 * 12, dtype=torch.bfloat16, device=DEV)        # n = 12: not a multiple of 8
    cache = torch.zeros(2, 2, 16, 1, 12, dtype=torch.uint8, device=DEV)
    sc = torch.ones((), device=DEV)
    with pytest.raises(RuntimeError):
        ops.reshape_and_cache_flash_bulk(keys, keys, [cache[0], cache[0]], [cache[1], cache[1]],
                                         torch.arange(4, device=DEV), "fp8", [sc, sc], [sc, sc], 1, 12)
    with pytest.raises(RuntimeError):
        ops.reshape_and_cache_flash_bulk(keys, keys, [cache[0]], [cache[1]], torch.arange(4, device=DEV),
                                         "int4", [sc], [sc], 1, 12)

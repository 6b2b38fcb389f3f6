"""GPU parity for hot path B: the sm_100a bulk KV scatter, called through the reference-facing
boundary (torch op -> ctypes -> C ABI), must be BIT-EXACT against the oracle on every byte of
every cache, including untouched slots."""
import glob
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT
from oracle import kv_write_ref as K

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TDT = {("bf16"): (np.uint16, torch.bfloat16), ("fp16"): (np.uint16, torch.float16),
       ("fp32"): (np.uint32, torch.float32)}


@pytest.fixture(scope="module")
def ops():
    from arcticinference_b200 import py_custom_ops
    assert py_custom_ops.try_load_torch_library()
    return py_custom_ops


def _to_dev(arr, wdt, tdt):
    signed = np.int16 if wdt == np.uint16 else np.int32
    base = arr
    t = torch.from_numpy(np.ascontiguousarray(base).view(signed)).to(DEV).view(tdt)
    return t


def _run_case(ops, L, T, Hkv, Dh, bs, dt, layout, slots="perm", frac_pad=0.1, extra=0, pad=0, seed=0,
              oracle=K.kv_write_c):
    wdt, tdt = TDT[dt]
    keys, values, sm, nb = K.make_kv_inputs(L, T, Hkv, Dh, bs, wdt, slots, frac_pad, seed, extra, pad)
    backing, kc, vc = K.make_caches(L, nb, bs, Hkv, Dh, wdt, layout, "random", seed)
    n = Hkv * Dh
    # device copies keep the (possibly padded) source row stride
    kfull = _to_dev(keys.base if pad else keys, wdt, tdt)
    vfull = _to_dev(values.base if pad else values, wdt, tdt)
    tk, tv = kfull[:, :L * n], vfull[:, :L * n]
    tb = [_to_dev(b, wdt, tdt) for b in backing]
    if layout == "flash":
        tkc, tvc = [b[0] for b in tb], [b[1] for b in tb]
    else:
        tkc, tvc = [b[:, 0] for b in tb], [b[:, 1] for b in tb]
    scales = [torch.ones((), device=DEV) for _ in range(L)]
    ops.reshape_and_cache_flash_bulk(tk, tv, tkc, tvc, torch.from_numpy(sm).to(DEV), "auto", scales,
                                     scales, Hkv, Dh)
    torch.cuda.synchronize()
    oracle(keys, values, kc, vc, sm, Hkv, Dh)
    signed_t = torch.int16 if wdt == np.uint16 else torch.int32
    for l in range(L):
        got = tb[l].view(signed_t).cpu().numpy().view(wdt)
        assert np.array_equal(got, backing[l]), f"layer {l} differs"


@pytest.mark.parametrize("layout", ["flash", "flashinfer"])
@pytest.mark.parametrize("dt", ["bf16", "fp16", "fp32"])
def test_small_cases_bit_exact(ops, layout, dt):
    _run_case(ops, 4, 37, 2, 16, 16, dt, layout)
    _run_case(ops, 3, 50, 4, 32, 8, dt, layout, seed=1)
    _run_case(ops, 16, 129, 8, 128, 16, dt, layout, seed=2)           # cfg3 geometry, few tokens


@pytest.mark.parametrize("layout", ["flash", "flashinfer"])
def test_ragged_and_edge_cases(ops, layout):
    _run_case(ops, 2, 13, 3, 20, 4, "bf16", layout, frac_pad=0.0)     # n*e = 120 B: 8-byte vectors
    _run_case(ops, 2, 13, 1, 7, 4, "bf16", layout)                    # n*e = 14 B: 2-byte path
    _run_case(ops, 2, 13, 1, 7, 4, "fp32", layout)                    # 28 B rows: 4-byte path
    _run_case(ops, 3, 21, 2, 64, 16, "bf16", layout, extra=5, pad=24)  # strided source, extra rows
    _run_case(ops, 1, 1, 1, 8, 1, "bf16", layout, frac_pad=0.0)       # single token / block_size 1
    _run_case(ops, 5, 64, 8, 64, 16, "fp16", layout, slots="sequential", frac_pad=0.0)
    _run_case(ops, 18, 257, 8, 64, 16, "bf16", layout, seed=3)        # cfg5 geometry (n = 512)
    _run_case(ops, 40, 33, 8, 128, 16, "bf16", layout, seed=4)        # 70B-shape layer count


def test_more_layers_than_one_pointer_table(ops):
    _run_case(ops, 131, 9, 2, 16, 4, "bf16", "flash")                 # > 128 layers: two launches


def test_all_tokens_padded_and_empty(ops):
    _run_case(ops, 2, 9, 2, 8, 4, "bf16", "flash", frac_pad=1.1)      # every slot = -1
    _run_case(ops, 2, 0, 2, 8, 4, "bf16", "flash")                    # no tokens
    # zero layers: silent return (kernels.cu:99-101)
    k = torch.zeros(4, 16, device=DEV, dtype=torch.bfloat16)
    ops.reshape_and_cache_flash_bulk(k, k, [], [], torch.zeros(4, dtype=torch.long, device=DEV),
                                     "auto", [], [], 2, 8)


def test_numpy_oracle_agrees_too(ops):
    _run_case(ops, 3, 50, 4, 32, 8, "bf16", "flash", oracle=K.kv_write_numpy)


def test_reference_unit_test_shape(ops):
    """The reference's own (vacuous) unit test geometry, with a real comparison: 2-D caches
    [T, H*Dh], fp32, L=4, T=2, H=2, Dh=16 (tests/unit_tests/test_custom_ops.py:56-90)."""
    L, T, H, Dh = 4, 2, 2, 16
    n = H * Dh
    g = torch.Generator(device="cpu").manual_seed(0)
    keys = torch.randn(T, L * n, generator=g).to(DEV)
    values = torch.randn(T, L * n, generator=g).to(DEV)
    kc = [torch.randn(T, n, generator=g).to(DEV) for _ in range(L)]
    vc = [torch.randn(T, n, generator=g).to(DEV) for _ in range(L)]
    sm = torch.tensor([1, 0], device=DEV)
    sc = [torch.tensor(0.1, device=DEV) for _ in range(L)]
    ops.reshape_and_cache_flash_bulk(keys, values, kc, vc, sm, "auto", sc, sc, H, Dh)
    for l in range(L):
        assert torch.equal(kc[l][1], keys[0, l * n:(l + 1) * n])
        assert torch.equal(kc[l][0], keys[1, l * n:(l + 1) * n])
        assert torch.equal(vc[l][1], values[0, l * n:(l + 1) * n])


def test_error_behaviour(ops):
    k = torch.zeros(4, 32, device=DEV, dtype=torch.bfloat16)
    c = torch.zeros(2, 4, 2, 8, device=DEV, dtype=torch.bfloat16)
    sm = torch.zeros(4, dtype=torch.long, device=DEV)
    s = [torch.ones((), device=DEV)] * 2
    with pytest.raises(RuntimeError):   # list lengths differ (TORCH_CHECK kernels.cu:103-106)
        ops.reshape_and_cache_flash_bulk(k, k, [c, c], [c], sm, "auto", s, s, 2, 8)
    with pytest.raises(RuntimeError):   # unknown kv cache dtype (quant_utils.cuh:567)
        ops.reshape_and_cache_flash_bulk(k, k, [c, c], [c, c], sm, "int4", s, s, 2, 8)
    with pytest.raises(RuntimeError, match="fp8"):   # fp8 needs a one-byte cache dtype
        ops.reshape_and_cache_flash_bulk(k, k, [c, c], [c, c], sm, "fp8", s, s, 2, 8)
    with pytest.raises(RuntimeError):   # int32 slot mapping (data_ptr<int64_t>() throws)
        ops.reshape_and_cache_flash_bulk(k, k, [c, c], [c, c], sm.int(), "auto", s, s, 2, 8)


GOLD = sorted(glob.glob(os.path.join(GOLDEN, "kv_write_*.npz")))


@pytest.mark.skipif(not GOLD, reason="kv golden fixtures not generated yet")
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_against_reference_kernel_golden(ops, path):
    """Golden vectors written by the reference's own kernel on a B200."""
    z = np.load(path)
    L, T, Hkv, Dh, bs, nb, er, rp = (int(v) for v in z["meta"])
    layout = str(z["layout"])
    wdt = z["keys"].dtype.type
    tdt = torch.float32 if wdt == np.uint32 else (torch.float16 if "fp16" in path else torch.bfloat16)
    tb = [_to_dev(z[f"init{l}"], wdt, tdt) for l in range(L)]
    if layout == "flash":
        kc, vc = [b[0] for b in tb], [b[1] for b in tb]
    else:
        kc, vc = [b[:, 0] for b in tb], [b[:, 1] for b in tb]
    sc = [torch.ones((), device=DEV) for _ in range(L)]
    ops.reshape_and_cache_flash_bulk(_to_dev(z["keys"], wdt, tdt), _to_dev(z["values"], wdt, tdt), kc, vc,
                                     torch.from_numpy(z["slots"]).to(DEV), "auto", sc, sc, Hkv, Dh)
    signed_t = torch.int16 if wdt == np.uint16 else torch.int32
    for l in range(L):
        assert np.array_equal(tb[l].view(signed_t).cpu().numpy().view(wdt), z[f"final{l}"])


def _full_size_case(ops, L, T, Hkv, Dh, layout, seed):
    """BASELINE-size check through size-independent properties: (1) gather-back round trip - every
    valid token's row is found at its slot in every layer; (2) untouched slots keep their bytes
    (checksum over the complement); (3) idempotence - a second call changes nothing."""
    bs, n = 16, Hkv * Dh
    nb = 2 * ((T + bs - 1) // bs)
    g = torch.Generator(device=DEV).manual_seed(seed)
    keys = torch.randn(T, L * n, device=DEV, dtype=torch.bfloat16, generator=g)
    vals = torch.randn(T, L * n, device=DEV, dtype=torch.bfloat16, generator=g)
    shape = (2, nb, bs, Hkv, Dh) if layout == "flash" else (nb, 2, bs, Hkv, Dh)
    kv = [torch.randn(shape, device=DEV, dtype=torch.bfloat16, generator=g) for _ in range(L)]
    kc = [x[0] if layout == "flash" else x[:, 0] for x in kv]
    vc = [x[1] if layout == "flash" else x[:, 1] for x in kv]
    slots = torch.randperm(nb * bs, device=DEV, generator=g)[:T].to(torch.int64)
    slots[::97] = -1
    valid = slots >= 0
    s = slots[valid]
    untouched = torch.ones(nb * bs, dtype=torch.bool, device=DEV)
    untouched[s] = False
    ub, uo = untouched.nonzero().squeeze(1) // bs, untouched.nonzero().squeeze(1) % bs
    before = [(kc[l][ub, uo].view(torch.int16).to(torch.int64).sum().item(),
               vc[l][ub, uo].view(torch.int16).to(torch.int64).sum().item()) for l in range(L)]
    sc = [torch.ones((), device=DEV) for _ in range(L)]
    ops.reshape_and_cache_flash_bulk(keys, vals, kc, vc, slots, "auto", sc, sc, Hkv, Dh)
    torch.cuda.synchronize()
    for l in range(L):
        assert torch.equal(kc[l][s // bs, s % bs].reshape(-1, n), keys[valid, l * n:(l + 1) * n])
        assert torch.equal(vc[l][s // bs, s % bs].reshape(-1, n), vals[valid, l * n:(l + 1) * n])
        after = (kc[l][ub, uo].view(torch.int16).to(torch.int64).sum().item(),
                 vc[l][ub, uo].view(torch.int16).to(torch.int64).sum().item())
        assert after == before[l]
    snap = [x.clone() for x in kv[:2]]
    ops.reshape_and_cache_flash_bulk(keys, vals, kc, vc, slots, "auto", sc, sc, Hkv, Dh)
    assert all(torch.equal(a, b) for a, b in zip(snap, kv[:2]))


@pytest.mark.parametrize("layout", ["flash", "flashinfer"])
def test_cfg3_full_size(ops, layout):
    _full_size_case(ops, 16, 8192, 8, 128, layout, seed=0)            # BASELINE configs[2]


def test_cfg5_shape_full_size(ops):
    _full_size_case(ops, 18, 32768, 8, 64, "flash", seed=1)           # GPT-OSS-shaped, 32K tokens


def test_more_than_65535_tokens(ops):
    """The reference's grid(L, T) cannot launch T >= 65 536; tokens sit on grid.x here."""
    _full_size_case(ops, 2, 70000, 2, 64, "flash", seed=2)


def test_graph_capturable(ops):
    """No per-call H2D copies (the reference does four blocking ones, kernels.cu:135-146)."""
    L, T, Hkv, Dh, bs = 4, 64, 2, 64, 16
    n = Hkv * Dh
    keys = torch.randn(T, L * n, device=DEV, dtype=torch.bfloat16)
    vals = torch.randn(T, L * n, device=DEV, dtype=torch.bfloat16)
    kv = [torch.zeros(2, 8, bs, Hkv, Dh, device=DEV, dtype=torch.bfloat16) for _ in range(L)]
    kc, vc = [x[0] for x in kv], [x[1] for x in kv]
    slots = torch.randperm(8 * bs, device=DEV)[:T]
    sc = [torch.ones((), device=DEV) for _ in range(L)]
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        ops.reshape_and_cache_flash_bulk(keys, vals, kc, vc, slots, "auto", sc, sc, Hkv, Dh)
    for x in kv:
        x.zero_()
    g.replay()
    torch.cuda.synchronize()
    for l in range(L):
        assert torch.equal(kc[l][slots // bs, slots % bs].reshape(-1, n), keys[:, l * n:(l + 1) * n])


REF_SO = os.path.join(ROOT, "oracle", "_ref", "ref_custom_ops.so")
_REF_SCRIPT = r'''
import sys, torch
so, out, L, T, Hkv, Dh, bs, layout = sys.argv[1], sys.argv[2], *map(int, sys.argv[3:8]), sys.argv[8]
torch.ops.load_library(so)
dev = "cuda:0"; n = Hkv * Dh; nb = 2 * ((T + bs - 1) // bs)
g = torch.Generator(device=dev).manual_seed(7)
keys = torch.randn(T, L * n, device=dev, dtype=torch.bfloat16, generator=g)
vals = torch.randn(T, L * n, device=dev, dtype=torch.bfloat16, generator=g)
shape = (2, nb, bs, Hkv, Dh) if layout == "flash" else (nb, 2, bs, Hkv, Dh)
kv = [torch.randn(shape, device=dev, dtype=torch.bfloat16, generator=g) for _ in range(L)]
kc = [x[0] if layout == "flash" else x[:, 0] for x in kv]
vc = [x[1] if layout == "flash" else x[:, 1] for x in kv]
slots = torch.randperm(nb * bs, device=dev, generator=g)[:T].to(torch.int64); slots[::13] = -1
sc = [torch.ones((), device=dev) for _ in range(L)]
if out != "-":
    torch.ops.arctic_inference.reshape_and_cache_flash_bulk(keys, vals, kc, vc, slots, "auto", sc, sc, Hkv, Dh)
    torch.cuda.synchronize()
    torch.save([x.cpu() for x in kv], out)
'''


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="reference kernel not built (oracle/build_ref.py)")
@pytest.mark.parametrize("layout", ["flash", "flashinfer"])
def test_bit_exact_against_reference_kernel_process(ops, layout, tmp_path):
    """The reference's own kernel (compiled for sm_100a from /root/reference sources into
    oracle/_ref) run in a SEPARATE process - it registers the same op name - on the same seeded
    inputs; every cache byte must match."""
    L, T, Hkv, Dh, bs = 6, 1000, 8, 128, 16
    out = str(tmp_path / "ref.pt")
    subprocess.check_call([sys.executable, "-c", _REF_SCRIPT, REF_SO, out, str(L), str(T), str(Hkv),
                           str(Dh), str(bs), layout], cwd=str(tmp_path))
    ref = torch.load(out)
    # regenerate the same inputs in this process
    n = Hkv * Dh; nb = 2 * ((T + bs - 1) // bs)
    g = torch.Generator(device=DEV).manual_seed(7)
    keys = torch.randn(T, L * n, device=DEV, dtype=torch.bfloat16, generator=g)
    vals = torch.randn(T, L * n, device=DEV, dtype=torch.bfloat16, generator=g)
    shape = (2, nb, bs, Hkv, Dh) if layout == "flash" else (nb, 2, bs, Hkv, Dh)
    kv = [torch.randn(shape, device=DEV, dtype=torch.bfloat16, generator=g) for _ in range(L)]
    kc = [x[0] if layout == "flash" else x[:, 0] for x in kv]
    vc = [x[1] if layout == "flash" else x[:, 1] for x in kv]
    slots = torch.randperm(nb * bs, device=DEV, generator=g)[:T].to(torch.int64); slots[::13] = -1
    sc = [torch.ones((), device=DEV) for _ in range(L)]
    ops.reshape_and_cache_flash_bulk(keys, vals, kc, vc, slots, "auto", sc, sc, Hkv, Dh)
    torch.cuda.synchronize()
    for l in range(L):
        assert torch.equal(kv[l].cpu().view(torch.int16), ref[l].view(torch.int16)), f"layer {l}"

"""GPU: seeded random-shape sweeps of every kernel against its oracle / comparator - catches tail,
alignment and tiling bugs that hand-picked shapes miss.  Bit-exact where the path is integer / byte
work or exact emulation; GEMMs against torch within one bf16 ulp of accumulation-order noise."""
import random

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import kv_write_fp8_ref as F8
from oracle import kv_write_ref as K
from test_cell_gpu import run_cell
from test_kv_write_gpu import _run_case

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    from arcticinference_b200 import py_custom_ops
    assert py_custom_ops.try_load_torch_library()
    return py_custom_ops


@pytest.fixture(scope="module")
def capi():
    from arcticinference_b200 import _capi
    return _capi


def test_kv_write_random_geometries(ops):
    rng = random.Random(1234)
    for i in range(40):
        L = rng.choice([1, 2, 3, 5, 16, 33])
        T = rng.choice([1, 2, 7, 31, 64, 200, 513])
        Hkv = rng.choice([1, 2, 3, 8])
        Dh = rng.choice([4, 7, 8, 20, 64, 128])
        bs = rng.choice([1, 4, 16, 32])
        dt = rng.choice(["bf16", "fp16", "fp32"])
        layout = rng.choice(["flash", "flashinfer"])
        pad = rng.choice([0, 0, 8, 24])
        _run_case(ops, L, T, Hkv, Dh, bs, dt, layout, slots=rng.choice(["perm", "sequential"]),
                  frac_pad=rng.choice([0.0, 0.05, 0.3]), extra=rng.choice([0, 3]), pad=pad, seed=100 + i)


def test_kv_write_fp8_random_geometries(ops):
    rng = random.Random(77)
    for i in range(25):
        L = rng.choice([1, 2, 5, 16, 65])
        T = rng.choice([1, 3, 40, 257])
        Hkv = rng.choice([1, 2, 8])
        Dh = rng.choice([8, 16, 64, 128])
        bs = rng.choice([1, 8, 16])
        dt = rng.choice([torch.bfloat16, torch.float16, torch.float32])
        kvd = rng.choice(["fp8", "fp8_e4m3", "fp8_e5m2"])
        layout = rng.choice(["flash", "flashinfer"])
        n = Hkv * Dh
        g = torch.Generator(device="cpu").manual_seed(500 + i)
        nb = 2 * ((T + bs - 1) // bs)
        keys = (torch.randn(T, L * n, generator=g) * rng.choice([0.01, 1.0, 50.0])).to(dt)
        values = (torch.randn(T, L * n, generator=g) * 2).to(dt)
        slots = torch.randperm(nb * bs, generator=g)[:T].to(torch.int64)
        slots[torch.rand(T, generator=g) < 0.1] = -1
        shape = (2, nb, bs, Hkv, Dh) if layout == "flash" else (nb, 2, bs, Hkv, Dh)
        init = [torch.randint(0, 256, shape, generator=g, dtype=torch.uint8) for _ in range(L)]
        ks = [torch.tensor(0.05 + 0.3 * rng.random()) for _ in range(L)]
        vs = [torch.tensor(0.01 + 0.1 * rng.random()) for _ in range(L)]
        view = (lambda cs: ([c[0] for c in cs], [c[1] for c in cs])) if layout == "flash" else \
            (lambda cs: ([c[:, 0] for c in cs], [c[:, 1] for c in cs]))
        ref = [c.clone() for c in init]
        F8.kv_write_fp8(keys, values, *view(ref), slots, kvd, ks, vs, Hkv, Dh)
        dev = [c.clone().to(DEV) for c in init]
        ops.reshape_and_cache_flash_bulk(keys.to(DEV), values.to(DEV), *view(dev), slots.to(DEV), kvd,
                                         [s.to(DEV) for s in ks], [s.to(DEV) for s in vs], Hkv, Dh)
        torch.cuda.synchronize()
        for l in range(L):
            assert torch.equal(dev[l].cpu(), ref[l]), f"case {i} layer {l}: L={L} T={T} n={n} {dt} {kvd} {layout}"


def test_cell_random_shapes_fast_equals_generic(capi):
    rng = random.Random(5)
    for i in range(16):
        B = rng.choice([1, 3, 17, 64, 161, 300, 1030, 4100])
        D = 8 * rng.choice([1, 5, 32, 45, 128, 360, 512, 513, 1024, 1100])
        if B * D > 4100 * 1024:
            B = 64
        torch.manual_seed(i)
        V = 333
        gates = (torch.randn(B, 4 * D, device=DEV) * rng.choice([0.5, 1.3, 5.0])).to(torch.bfloat16)
        emb = (torch.randn(V, D, device=DEV) * 0.05).to(torch.bfloat16)
        tok = torch.randint(0, V, (B,), device=DEV)
        cprev = (torch.randn(B, D, device=DEV)).to(torch.bfloat16) if rng.random() < 0.8 else None
        ln = [(torch.randn(D, device=DEV) * 0.3 + (1.0 if k % 2 == 0 else 0.0)).to(torch.bfloat16) for k in range(4)]
        fp8 = D % 16 == 0
        a = run_cell(capi, False, gates, emb, tok, cprev, ln, 19.5, want_fp8=fp8)
        b = run_cell(capi, True, gates, emb, tok, cprev, ln, 19.5, want_fp8=fp8)
        assert torch.equal(a[0].view(torch.int16), b[0].view(torch.int16)), f"state B={B} D={D}"
        assert torch.equal(a[1].view(torch.int16), b[1].view(torch.int16)), f"cell B={B} D={D}"
        if fp8:
            assert torch.equal(a[2], b[2]) and torch.equal(a[3], b[3])


def test_gemm_and_argmax_random_shapes(capi):
    rng = random.Random(9)
    lib = capi.lib()
    st = torch.cuda.current_stream().cuda_stream
    for i in range(24):
        M = rng.choice([1, 2, 9, 33, 64, 65, 100, 128, 129, 200, 256, 257, 300])
        N = 8 * rng.choice([2, 13, 64, 100, 257, 1000, 2503])
        Kd = 8 * rng.choice([1, 9, 64, 100, 360])
        torch.manual_seed(i)
        x = torch.randn(M, Kd, device=DEV).to(torch.bfloat16)
        w = (torch.randn(N, Kd, device=DEV) * 0.05).to(torch.bfloat16)
        out = torch.full((M, N), float("nan"), device=DEV, dtype=torch.bfloat16)
        capi.check(lib.ab200_linear(x.data_ptr(), w.data_ptr(), out.data_ptr(), M, N, Kd, Kd, Kd, N, capi.AB200_BF16, st))
        ref = F.linear(x, w)
        assert not torch.isnan(out).any(), f"M={M} N={N} K={Kd}"
        assert (out == ref).float().mean().item() > 0.99, f"M={M} N={N} K={Kd}"
        assert ((out.float() - ref.float()).abs().max() / ref.float().abs().max().clamp_min(1e-6)).item() < 1.6e-2
        vocab = N - rng.choice([0, 1, 5])
        packed = torch.zeros(M, dtype=torch.int64, device=DEV)
        tok = torch.empty(M, dtype=torch.long, device=DEV)
        counter = torch.zeros(1, dtype=torch.int32, device=DEV)
        capi.check(lib.ab200_head_argmax(x.data_ptr(), w.data_ptr(), M, N, Kd, vocab, Kd, Kd, packed.data_ptr(),
                                         tok.data_ptr(), 1, counter.data_ptr(), capi.AB200_BF16, None, 0, st))
        logits = ref[:, :vocab]
        picked = logits.gather(1, tok.clamp(0, vocab - 1).unsqueeze(1)).squeeze(1).float()
        best = logits.max(-1).values.float()
        assert (tok >= 0).all() and (tok < vocab).all()
        assert ((picked - best).abs() <= best.abs() * 2 ** -7 + 1e-6).all(), f"argmax M={M} N={N} K={Kd}"


def test_gather_random_shapes():
    from arcticinference_b200.swiftkv import gather_rows
    rng = random.Random(3)
    for i in range(20):
        T = rng.choice([1, 5, 64, 300])
        m = rng.choice([0, 1, 7, 64])
        dts = [torch.bfloat16, torch.float32, torch.int64, torch.uint8, torch.float16]
        tensors = []
        for _ in range(rng.choice([1, 2, 5, 8, 9])):
            dt = rng.choice(dts)
            cols = rng.choice([1, 3, 8, 100, 1025])
            t = torch.randint(0, 100, (T, cols) if rng.random() < 0.8 else (T,), device=DEV).to(dt)
            tensors.append(t)
        idx = torch.randint(0, T, (m,), device=DEV)
        got = gather_rows(tensors, idx)
        for a, t in zip(got, tensors):
            assert torch.equal(a, t.index_select(0, idx))

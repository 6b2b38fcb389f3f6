"""Dev script: first contact of every kernel with a real B200 (prints diagnostics, no asserts)."""
import os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.nn.functional as F
from arcticinference_b200 import _capi, py_custom_ops
from arcticinference_b200.speculator import ArcticLSTMSpeculator
from oracle import speculator_ref as O

dev = torch.device("cuda:0")
lib = _capi.lib()
st = lambda: torch.cuda.current_stream().cuda_stream


def section(name):
    def deco(fn):
        print(f"==== {name}", flush=True)
        try:
            fn()
            torch.cuda.synchronize()
        except Exception:
            traceback.print_exc()
        return fn
    return deco


@section("gemm store small")
def _():
    for (M, N, K) in [(64, 256, 64), (64, 256, 256), (5, 256, 96), (64, 16384, 4096), (128, 16384, 4096), (200, 4096, 512)]:
        torch.manual_seed(0)
        x = torch.randn(M, K, device=dev).to(torch.bfloat16)
        w = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
        out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        rc = lib.ab200_linear(x.data_ptr(), w.data_ptr(), out.data_ptr(), M, N, K, K, K, N, _capi.AB200_BF16, st())
        torch.cuda.synchronize()
        ref = F.linear(x, w)
        ref32 = F.linear(x.float(), w.float())
        eq = (out == ref).float().mean().item()
        err = ((out.float() - ref32).abs().max() / ref32.abs().max()).item()
        print(f"M={M} N={N} K={K} rc={rc} bn={lib.ab200_gemm_block_n(N,0)} eq_vs_cublas={eq:.5f} maxrelerr_vs_fp32={err:.2e}", flush=True)
        if eq < 0.9:
            print(" out[0,:8]", out[0, :8].tolist()); print(" ref[0,:8]", ref[0, :8].tolist())


@section("gemm argmax")
def _():
    for (M, V, K) in [(64, 1000, 64), (64, 4096, 512), (64, 128256, 4096), (128, 128256, 4096)]:
        torch.manual_seed(1)
        Vp = (V + 63) // 64 * 64
        x = torch.randn(M, K, device=dev).to(torch.bfloat16)
        w = torch.zeros(Vp, K, device=dev, dtype=torch.bfloat16)
        w[:V] = (torch.randn(V, K, device=dev) * 0.02).to(torch.bfloat16)
        packed = torch.zeros(M, dtype=torch.int64, device=dev)
        tok = torch.empty(M, dtype=torch.long, device=dev)
        rc = lib.ab200_head_argmax(x.data_ptr(), w.data_ptr(), M, Vp, K, V, K, K, packed.data_ptr(), None, 0, None, _capi.AB200_BF16, None, 0, st())
        rc2 = lib.ab200_argmax_unpack(packed.data_ptr(), tok.data_ptr(), M, 1, st())
        torch.cuda.synchronize()
        ref = torch.argmax(F.linear(x, w)[:, :V], -1)
        print(f"M={M} V={V} K={K} rc={rc},{rc2} agree={(tok == ref).float().mean().item():.4f}", flush=True)


@section("cell + rms")
def _():
    V, H, D, B = 4096, 512, 512, 32
    w = O.make_weights(V, H, D, torch.bfloat16, device=dev)
    ids, hid = O.make_inputs(B, V, H, torch.bfloat16, device=dev)
    ref = O.RefLSTMSpeculator(w, V, 3)
    m = ArcticLSTMSpeculator(vocab_size=V, input_hidden_dim=H, inner_dim=D, n_predict=3, max_num_seqs=B, device=dev)
    m.load_weights(w.items())
    cell0 = torch.zeros(B, 1, D, device=dev, dtype=torch.bfloat16)
    s_ref, c_ref = ref.generate_states(ids.unsqueeze(1), hid.unsqueeze(1), 0, cell0)
    s, c = m.generate_states(ids.unsqueeze(1), hid.unsqueeze(1), 0, cell0)
    print("state eq", (s == s_ref).float().mean().item(), "cell eq", (c == c_ref).float().mean().item(),
          "relerr", ((s.float() - s_ref.float()).norm() / s_ref.float().norm()).item())
    got = m.generate_proposals(ids, hid, 3)
    exp = ref.generate_proposals(ids, hid, 3)
    print("tokens agree", (got == exp).float().mean().item())
    print(got[:4].tolist(), exp[:4].tolist())


@section("8B shape timing")
def _():
    V, H, D, B, S = 128256, 4096, 4096, 64, 3
    w = O.make_weights(V, H, D, torch.bfloat16, device=dev)
    ids, hid = O.make_inputs(B, V, H, torch.bfloat16, device=dev)
    m = ArcticLSTMSpeculator(vocab_size=V, input_hidden_dim=H, inner_dim=D, n_predict=S, max_num_seqs=B, device=dev)
    m.load_weights(w.items())
    ref = O.RefLSTMSpeculator(w, V, S)
    got = m.generate_proposals(ids, hid, S)
    exp = ref.generate_proposals(ids, hid, S)
    print("tokens agree 8B", (got == exp).float().mean().item())
    for _ in range(3):
        m.generate_proposals(ids, hid, S)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        m.generate_proposals(ids, hid, S)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"ours graph: {ms*1e3:.1f} us/proposal -> {B*S/ms*1e3:.0f} draft tok/s")
    for _ in range(2):
        ref.generate_proposals(ids, hid, S)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5):
        ref.generate_proposals(ids, hid, S)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"torch eager restatement: {ms*1e3:.1f} us/proposal -> {B*S/ms*1e3:.0f} draft tok/s")


@section("kv write cfg3")
def _():
    assert py_custom_ops.try_load_torch_library()
    L, T, Hkv, Dh, bs = 16, 8192, 8, 128, 16
    n = Hkv * Dh; nb = 2 * T // bs
    g = torch.Generator(device=dev); g.manual_seed(0)
    keys = torch.randn(T, L * n, device=dev, dtype=torch.bfloat16, generator=g)
    vals = torch.randn(T, L * n, device=dev, dtype=torch.bfloat16, generator=g)
    kv = [torch.zeros(2, nb, bs, Hkv, Dh, device=dev, dtype=torch.bfloat16) for _ in range(L)]
    kc = [x[0] for x in kv]; vc = [x[1] for x in kv]
    slots = torch.randperm(nb * bs, device=dev, generator=g)[:T].to(torch.int64)
    slots[::97] = -1
    sc = [torch.ones((), device=dev) for _ in range(L)]
    f = lambda: py_custom_ops.reshape_and_cache_flash_bulk(keys, vals, kc, vc, slots, "auto", sc, sc, Hkv, Dh)
    f(); torch.cuda.synchronize()
    valid = slots >= 0; s = slots[valid]
    ok = all(torch.equal(kc[l][s // bs, s % bs].reshape(-1, n), keys[valid, l * n:(l + 1) * n]) and
             torch.equal(vc[l][s // bs, s % bs].reshape(-1, n), vals[valid, l * n:(l + 1) * n]) for l in range(L))
    print("roundtrip ok", ok)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(10):
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort()
    byt = int(valid.sum()) * L * n * 2 * 4 + 8 * T
    print(f"kv write cfg3: median {ts[5]*1e3:.1f} us min {ts[0]*1e3:.1f} us -> {byt/ts[5]/1e6:.0f} GB/s (median)")
print("DONE")

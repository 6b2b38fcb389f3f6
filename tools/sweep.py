"""Batch sweeps of every kernel on the hot path against its roofline (B200, one GPU).

    python tools/sweep.py [--out gpurun_out/sweep.json]

For each kernel: CUDA-event time (median of N, L2 flushed with a 256 MiB write between launches),
algorithmic bytes / flops per launch, GB/s against the measured HBM copy peak
and TFLOP/s against the measured cuBLAS bf16 peak (MEASURED_PEAKS.json).  cuBLAS / ATen timings of
the same op are printed beside ours as the library baseline.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from arcticinference_b200 import _capi  # noqa: E402

lib = _capi.lib()
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
try:
    PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1680.0, "bf16_tflops_sustained": 1414.0}
HBM, TF = PEAKS["hbm_gbs"], PEAKS["bf16_tflops"]


def timeit(fn, n=15):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e3  # us


def ck(rc, what):
    _capi.check(rc, what)


def sweep_gemm(rows, V, D, K, batches, fp8_batches):
    Wh = (torch.randn(V, D, device=dev) * 0.02).to(torch.bfloat16)
    Wg = (torch.randn(4 * D, K, device=dev) * 0.02).to(torch.bfloat16)
    Wq = torch.empty(V, D, dtype=torch.uint8, device=dev)
    sw = torch.zeros(1, dtype=torch.float32, device=dev)
    fp8_ok = D % 16 == 0
    if fp8_ok:
        ck(lib.ab200_fp8_quantize_weight(Wh.data_ptr(), Wq.data_ptr(), sw.data_ptr(), V, D, D, D, st), "qw")
    for M in batches:
        x = torch.randn(M, D, device=dev).to(torch.bfloat16)
        xg = torch.randn(M, K, device=dev).to(torch.bfloat16)
        packed = torch.zeros(M + 1, dtype=torch.int64, device=dev)
        tok = torch.empty(M, dtype=torch.int64, device=dev)
        gates = torch.empty(M, 4 * D, dtype=torch.bfloat16, device=dev)

        def head():
            ck(lib.ab200_head_argmax(x.data_ptr(), Wh.data_ptr(), M, V, D, V, D, D, packed.data_ptr(),
                                     tok.data_ptr() if M <= 512 else None, 1,
                                     packed[M:].data_ptr() if M <= 512 else None, 2, None, 0, st), "head")
            packed.zero_()

        def gate():
            ck(lib.ab200_linear(xg.data_ptr(), Wg.data_ptr(), gates.data_ptr(), M, 4 * D, K, K, K, 4 * D, 2, st), "gate")

        th = timeit(head)
        tz = timeit(lambda: packed.zero_())
        th -= tz
        tg = timeit(gate)
        tcb = timeit(lambda: torch.nn.functional.linear(x, Wh).argmax(-1))
        tcg = timeit(lambda: torch.nn.functional.linear(xg, Wg))
        hb = V * D * 2 + M * D * 2 + 8 * M
        gb = 4 * D * K * 2 + M * K * 2 + M * 4 * D * 2
        hf, gf = 2.0 * M * D * V, 2.0 * M * K * 4 * D
        r = {"kernel": "head_argmax+gate", "V": V, "D": D, "K": K, "B": M,
             "head_us": th, "head_gbs": hb / th / 1e3, "head_hbm_frac": hb / th / 1e3 / HBM,
             "head_tflops": hf / th / 1e6, "head_tensor_frac_measured": hf / th / 1e6 / TF,
             "head_tensor_frac_nominal": hf / th / 1e6 / 2250.0,
             "gate_us": tg, "gate_gbs": gb / tg / 1e3, "gate_hbm_frac": gb / tg / 1e3 / HBM,
             "gate_tflops": gf / tg / 1e6, "gate_tensor_frac_measured": gf / tg / 1e6 / TF,
             "aten_linear_argmax_us": tcb, "cublas_gate_us": tcg}
        rows.append(r)
        print(f"[gemm V={V} D={D}] B={M:4d} head {th:7.1f} us {r['head_gbs']:5.0f} GB/s ({r['head_hbm_frac']:.2f} hbm) "
              f"{r['head_tflops']:6.0f} TF/s ({r['head_tensor_frac_measured']:.2f} of cuBLAS peak) | "
              f"gate {tg:6.1f} us {r['gate_gbs']:5.0f} GB/s {r['gate_tflops']:5.0f} TF/s | "
              f"ATen linear+argmax {tcb:7.1f} us, cuBLAS gate {tcg:6.1f} us", flush=True)
    if not fp8_ok:
        return
    for M in fp8_batches:
        x = torch.randn(M, D, device=dev).to(torch.bfloat16)
        xq = torch.empty(M, D, dtype=torch.uint8, device=dev)
        sx = torch.empty(M, dtype=torch.float32, device=dev)
        ck(lib.ab200_fp8_quantize_rows(x.data_ptr(), xq.data_ptr(), sx.data_ptr(), M, D, D, D, st), "qx")
        packed = torch.zeros(M + 1, dtype=torch.int64, device=dev)
        tok = torch.empty(M, dtype=torch.int64, device=dev)

        def head8():
            ck(lib.ab200_head_argmax_fp8(xq.data_ptr(), sx.data_ptr(), Wq.data_ptr(), sw.data_ptr(), M, V, D, V,
                                         D, D, packed.data_ptr(), tok.data_ptr(), 1, packed[M:].data_ptr(), st), "h8")
            packed.zero_()

        t8 = timeit(head8) - timeit(lambda: packed.zero_())
        b8 = V * D + M * D + 12 * M
        rows.append({"kernel": "head_argmax_fp8", "V": V, "D": D, "B": M, "us": t8, "gbs": b8 / t8 / 1e3,
                     "hbm_frac": b8 / t8 / 1e3 / HBM})
        print(f"[fp8 head V={V} D={D}] B={M:3d} {t8:7.1f} us {b8 / t8 / 1e3:5.0f} GB/s ({b8 / t8 / 1e3 / HBM:.2f} hbm)", flush=True)


def sweep_cell(rows, D, V, batches):
    emb = (torch.randn(V, D, device=dev) * 0.02).to(torch.bfloat16)
    ln = [(torch.randn(D, device=dev) * 0.1 + (1 if i % 2 == 0 else 0)).to(torch.bfloat16) for i in range(4)]
    for B in batches:
        gates = torch.randn(B, 4 * D, device=dev).to(torch.bfloat16)
        toks = torch.randint(0, V, (B,), device=dev)
        cprev = torch.randn(B, D, device=dev).to(torch.bfloat16)
        state = torch.empty(B, D, device=dev, dtype=torch.bfloat16)
        cell = torch.empty_like(state)
        x = torch.randn(B, D, device=dev).to(torch.bfloat16)
        y = torch.empty_like(x)
        for emu in (1, 2, 0):
            t = timeit(lambda: ck(lib.ab200_sum_lstm_cell(
                gates.data_ptr(), 4 * D, emb.data_ptr(), toks.data_ptr(), None, cprev.data_ptr(), ln[0].data_ptr(),
                ln[1].data_ptr(), ln[2].data_ptr(), ln[3].data_ptr(), state.data_ptr(), cell.data_ptr(), B, D, V,
                23.07, 1e-6, 2, emu, None, None, None, None, st), "cell"))
            nbytes = B * D * 2 * 8 + 4 * D * 2 + 8 * B
            rows.append({"kernel": "sum_lstm_cell", "emulate_eager_rounding": emu, "D": D, "B": B, "us": t,
                         "gbs": nbytes / t / 1e3, "hbm_frac": nbytes / t / 1e3 / HBM})
            print(f"[cell D={D} emu={emu}] B={B:6d} {t:8.1f} us {nbytes / t / 1e3:6.0f} GB/s ({nbytes / t / 1e3 / HBM:.2f} hbm)", flush=True)
        t = timeit(lambda: ck(lib.ab200_rms_scale(x.data_ptr(), y.data_ptr(), B, D, D, D, 1e-6, 0.70710678, 2, 1, st), "rms"))
        nbytes = 2 * B * D * 2
        rows.append({"kernel": "rms_scale", "D": D, "B": B, "us": t, "gbs": nbytes / t / 1e3,
                     "hbm_frac": nbytes / t / 1e3 / HBM})
        print(f"[rms  D={D}] B={B:6d} {t:8.1f} us {nbytes / t / 1e3:6.0f} GB/s ({nbytes / t / 1e3 / HBM:.2f} hbm)", flush=True)


def sweep_kv(rows):
    from arcticinference_b200 import py_custom_ops
    assert py_custom_ops.try_load_torch_library()
    cases = [("cfg3 8B", 16, 8192, 8, 128, "flash", 0.0), ("cfg3 8B flashinfer", 16, 8192, 8, 128, "flashinfer", 0.0),
             ("cfg3 8B 1% padded", 16, 8192, 8, 128, "flash", 0.01), ("cfg5 gpt-oss 32K", 18, 32768, 8, 64, "flash", 0.0),
             ("70B-shape", 40, 8192, 8, 128, "flash", 0.0), ("decode-like T=256", 16, 256, 8, 128, "flash", 0.0),
             ("T=1024", 16, 1024, 8, 128, "flash", 0.0)]
    for name, L, T, Hkv, Dh, layout, frac_pad in cases:
        n, bs = Hkv * Dh, 16
        nb = 2 * T // bs
        keys = torch.randn(T, L * n, device=dev, dtype=torch.bfloat16)
        vals = torch.randn(T, L * n, device=dev, dtype=torch.bfloat16)
        if layout == "flash":
            kv = [torch.zeros(2, nb, bs, Hkv, Dh, device=dev, dtype=torch.bfloat16) for _ in range(L)]
            kc, vc = [x[0] for x in kv], [x[1] for x in kv]
        else:
            kv = [torch.zeros(nb, 2, bs, Hkv, Dh, device=dev, dtype=torch.bfloat16) for _ in range(L)]
            kc, vc = [x[:, 0] for x in kv], [x[:, 1] for x in kv]
        slots = torch.randperm(nb * bs, device=dev)[:T].to(torch.int64)
        if frac_pad:
            slots[torch.rand(T, device=dev) < frac_pad] = -1
        valid = int((slots >= 0).sum())
        sc = [torch.ones((), device=dev) for _ in range(L)]
        t = timeit(lambda: py_custom_ops.reshape_and_cache_flash_bulk(keys, vals, kc, vc, slots, "auto", sc, sc, Hkv, Dh))
        alg = valid * L * n * 2 * 4 + 8 * T
        rows.append({"kernel": "kv_write_bulk", "case": name, "L": L, "T": T, "n": n, "layout": layout, "us": t,
                     "gbs": alg / t / 1e3, "hbm_frac": alg / t / 1e3 / HBM, "frac_of_8TBs": alg / t / 1e3 / 8000})
        print(f"[kv {name}] {t:8.1f} us {alg / t / 1e3:6.0f} GB/s ({alg / t / 1e3 / HBM:.2f} of measured peak, "
              f"{alg / t / 1e3 / 8000:.2f} of 8 TB/s)  [includes ~host launch path of the torch op]", flush=True)
        if layout == "flash" and T >= 8192:
            kv8 = [torch.zeros(2, nb, bs, Hkv, Dh, device=dev, dtype=torch.uint8) for _ in range(L)]
            kc8, vc8 = [x[0] for x in kv8], [x[1] for x in kv8]
            sc8 = [torch.full((), 0.05 * (i + 1), device=dev) for i in range(L)]
            t8 = timeit(lambda: py_custom_ops.reshape_and_cache_flash_bulk(keys, vals, kc8, vc8, slots, "fp8", sc8, sc8, Hkv, Dh))
            alg8 = valid * L * n * 2 * 3 + 8 * T
            rows.append({"kernel": "kv_write_bulk_fp8", "case": name, "L": L, "T": T, "n": n, "us": t8,
                         "gbs": alg8 / t8 / 1e3, "hbm_frac": alg8 / t8 / 1e3 / HBM})
            print(f"[kv fp8 {name}] {t8:8.1f} us {alg8 / t8 / 1e3:6.0f} GB/s ({alg8 / t8 / 1e3 / HBM:.2f} of measured peak)", flush=True)
            del kv8, kc8, vc8
        del kv, kc, vc, keys, vals


def sweep_gather(rows):
    """SwiftKV select compaction (llama_swiftkv.py:669-685): five index_selects in one launch."""
    from arcticinference_b200 import swiftkv
    T, Hd, Ln = 8192, 4096, 16 * 1024
    g = torch.Generator(device=dev).manual_seed(0)
    hid = torch.randn(T, Hd, device=dev, generator=g).to(torch.bfloat16)
    res = torch.randn(T, Hd, device=dev, generator=g).to(torch.bfloat16)
    pos = torch.arange(T, device=dev)
    k = torch.randn(T, Ln, device=dev, generator=g).to(torch.bfloat16)
    v = torch.randn(T, Ln, device=dev, generator=g).to(torch.bfloat16)
    for m in (64, 256, 2048, 4096):
        idx = torch.randperm(T, device=dev, generator=g)[:m].sort().values
        tensors = [hid, res, pos, k, v]
        outs = [torch.empty((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev) for t in tensors]
        t_ours = timeit(lambda: swiftkv.gather_rows(tensors, idx, outs))
        t_aten = timeit(lambda: [torch.index_select(t, 0, idx, out=o) for t, o in zip(tensors, outs)])
        alg = 2 * m * sum(t[0].numel() * t.element_size() if t.dim() > 1 else t.element_size() for t in tensors) + 8 * m
        rows.append({"kernel": "gather_rows", "T": T, "selected": m, "us": t_ours, "gbs": alg / t_ours / 1e3,
                     "hbm_frac": alg / t_ours / 1e3 / HBM, "aten_5x_index_select_us": t_aten})
        print(f"[gather T={T}] selected={m:5d} {t_ours:8.1f} us {alg / t_ours / 1e3:6.0f} GB/s ({alg / t_ours / 1e3 / HBM:.2f} hbm) | "
              f"5 x torch.index_select {t_aten:8.1f} us", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.json"))
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    rows = []
    only = set(a.only.split(",")) if a.only else None
    if not only or "gemm" in only:
        sweep_gemm(rows, 128256, 4096, 4096, [1, 8, 32, 64, 128, 256, 512], [1, 8, 16, 32])
    if only and "gemmbig" in only:
        sweep_gemm(rows, 128256, 4096, 4096, [160, 192, 256, 384, 512], [])
    if not only or "gemm70" in only:
        sweep_gemm(rows, 128256, 8192, 8192, [64, 128, 256], [32])
    if not only or "gemmoss" in only:
        sweep_gemm(rows, 201088, 2880, 2880, [64, 128, 256, 512], [32])
    if not only or "cell" in only:
        sweep_cell(rows, 4096, 128256, [1, 64, 256, 1024, 4096, 16384])
    if not only or "kv" in only:
        sweep_kv(rows)
    if not only or "gather" in only:
        sweep_gather(rows)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump({"peaks": PEAKS, "rows": rows}, open(a.out, "w"), indent=1)
    print("wrote", a.out)


if __name__ == "__main__":
    main()

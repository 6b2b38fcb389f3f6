"""TEST INFRASTRUCTURE ONLY - produces tests/golden/kv_write_*.npz by running the REFERENCE's own
kernel (oracle/_ref/ref_custom_ops.so, built by oracle/build_ref.py from
/root/reference/csrc/custom_ops/{kernels.cu,torch_bindings.cpp}) on a B200.

    gpurun -- python oracle/gen_golden_kv_gpu.py      # writes gpurun_out/golden/kv_write_*.npz
    cp gpurun_out/golden/kv_write_*.npz tests/golden/

Each file holds the inputs (as raw 16/32-bit words), the initial cache backing buffers (random,
so that untouched bytes are checked too) and the backing buffers after the reference call.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import kv_write_ref as K  # noqa: E402

# name: (L, T, Hkv, Dh, bs, word dtype, torch dtype, layout, slots, frac_pad, extra_rows, row_pad)
CASES = {
    "flash_bf16": (4, 37, 2, 16, 16, np.uint16, torch.bfloat16, "flash", "perm", 0.1, 0, 0),
    "flashinfer_bf16": (3, 50, 4, 32, 8, np.uint16, torch.bfloat16, "flashinfer", "perm", 0.1, 0, 0),
    "flash_fp16_seq": (2, 40, 4, 32, 16, np.uint16, torch.float16, "flash", "sequential", 0.0, 0, 0),
    "flash_fp32": (2, 19, 2, 16, 16, np.uint32, torch.float32, "flash", "perm", 0.1, 0, 0),
    "strided_src_bf16": (3, 21, 2, 64, 16, np.uint16, torch.bfloat16, "flash", "perm", 0.05, 5, 24),
    "odd_n_bf16": (2, 13, 3, 20, 4, np.uint16, torch.bfloat16, "flashinfer", "perm", 0.0, 0, 0),
}


def main():
    torch.ops.load_library(os.path.join(HERE, "_ref", "ref_custom_ops.so"))
    op = torch.ops.arctic_inference.reshape_and_cache_flash_bulk
    out = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out, exist_ok=True)
    dev = "cuda"
    for name, (L, T, Hkv, Dh, bs, wdt, tdt, layout, slots, fp, er, rp) in CASES.items():
        keys, values, sm, nb = K.make_kv_inputs(L, T, Hkv, Dh, bs, wdt, slots, fp, seed=3,
                                                extra_rows=er, row_pad=rp)
        backing, _, _ = K.make_caches(L, nb, bs, Hkv, Dh, wdt, layout, fill="random", seed=3)
        signed = np.int16 if wdt == np.uint16 else np.int32
        # keep the source row stride (keys is a column-slice of a wider array when row_pad > 0)
        kfull = torch.from_numpy(np.ascontiguousarray(keys.base if keys.base is not None else keys)
                                 .view(signed)).to(dev).view(tdt)
        vfull = torch.from_numpy(np.ascontiguousarray(values.base if values.base is not None else values)
                                 .view(signed)).to(dev).view(tdt)
        n = Hkv * Dh
        tk, tv = kfull[:, :L * n], vfull[:, :L * n]
        tb = [torch.from_numpy(b.view(signed)).to(dev).view(tdt) for b in backing]
        if layout == "flash":
            kc, vc = [b[0] for b in tb], [b[1] for b in tb]
        else:
            kc, vc = [b[:, 0] for b in tb], [b[:, 1] for b in tb]
        scales = [torch.tensor(1.0, device=dev) for _ in range(L)]
        op(tk, tv, kc, vc, torch.from_numpy(sm).to(dev), "auto", scales, scales, Hkv, Dh)
        torch.cuda.synchronize()
        rec = {"meta": np.array([L, T, Hkv, Dh, bs, nb, er, rp], dtype=np.int64),
               "layout": np.array(layout), "slots": sm,
               "keys": np.ascontiguousarray(keys), "values": np.ascontiguousarray(values)}
        for l in range(L):
            rec[f"init{l}"] = backing[l]
            rec[f"final{l}"] = tb[l].view(torch.int16 if wdt == np.uint16 else torch.int32).cpu().numpy().view(wdt)
        path = os.path.join(out, f"kv_write_{name}.npz")
        np.savez_compressed(path, **rec)
        print(name, "->", path, os.path.getsize(path))


if __name__ == "__main__":
    main()

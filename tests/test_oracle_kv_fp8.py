"""CPU: pins oracle/kv_write_fp8_ref.py against golden vectors written by vLLM's own per-layer op
(reshape_and_cache_flash with an fp8 cache) on a B200 - every byte of every cache."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import kv_write_fp8_ref as F

CASES = sorted(os.path.basename(p)[len("kvfp8_"):-4] for p in glob.glob(os.path.join(GOLDEN, "kvfp8_*.npz")))
SRC = {"torch.bfloat16": torch.bfloat16, "torch.float16": torch.float16, "torch.float32": torch.float32}


def load(name):
    z = np.load(os.path.join(GOLDEN, f"kvfp8_{name}.npz"))
    L, T, Hkv, Dh, bs, nb = (int(v) for v in z["meta"])
    dt = SRC[str(z["src_dtype"])]
    keys = torch.from_numpy(z["keys"]).view(dt)
    values = torch.from_numpy(z["values"]).view(dt)
    return z, (L, T, Hkv, Dh, bs, nb), dt, keys, values, str(z["kv_cache_dtype"])


def test_golden_present():
    assert len(CASES) >= 4 and any(c.startswith("e5m2") for c in CASES)


@pytest.mark.parametrize("name", CASES)
def test_restatement_matches_vllm_op_bit_exact(name):
    z, (L, T, Hkv, Dh, bs, nb), dt, keys, values, kvd = load(name)
    caches = [torch.from_numpy(z[f"init{l}"].copy()) for l in range(L)]
    ks = [torch.tensor(float(s)) for s in z["k_scales"]]
    vs = [torch.tensor(float(s)) for s in z["v_scales"]]
    F.kv_write_fp8(keys, values, [c[0] for c in caches], [c[1] for c in caches], torch.from_numpy(z["slots"]),
                   kvd, ks, vs, Hkv, Dh)
    for l in range(L):
        assert np.array_equal(caches[l].numpy(), z[f"final{l}"]), f"layer {l}"
        assert not np.array_equal(z[f"init{l}"], z[f"final{l}"])


def test_quantize_specials():
    x = torch.tensor([0.0, -0.0, 1e-9, 1e6, -1e6, 0.4375, 0.46875, float("nan")])
    q = F.quantize(x, torch.tensor(1.0), "fp8")
    # saturation to +-448, signed zero kept, tiny values flush to zero, grid points stay put
    assert q[3].item() == 0x7E and q[4].item() == 0xFE and q[1].item() == 0x80 and q[2].item() == 0x00
    assert torch.tensor([q[5].item(), q[6].item()], dtype=torch.uint8).view(torch.float8_e4m3fn).float().tolist() == [0.4375, 0.46875]

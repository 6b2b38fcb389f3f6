"""CPU: the fused-rotary oracle (oracle/kv_rope_ref.py) is pinned against golden vectors written by vLLM's own
rotary_embedding CUDA op on a B200 (oracle/gen_golden_rope_gpu.py -> tests/golden/rope_vllm_*.npz): the "fp32"
contract reproduces the op BIT FOR BIT, the "storage" contract (forward_native in the storage dtype) is the other
function.  Also cross-checks the torch scatter used for the cache write against the numpy oracle of the bulk op
(oracle/kv_write_ref.py, itself pinned to the reference kernel's goldens)."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import kv_rope_ref as R
from oracle import kv_write_ref as K

CASES = sorted(glob.glob(os.path.join(GOLDEN, "rope_vllm_*.npz")))


def _load(path):
    z = np.load(path)
    T, Hkv, Dh, rot, neox, max_pos = (int(v) for v in z["meta"])
    dtype = torch.float16 if path.endswith("f16.npz") and "bf16" not in os.path.basename(path) else torch.bfloat16
    t = lambda a: torch.from_numpy(a).view(dtype)
    return z, t, (T, Hkv, Dh, rot, bool(neox), max_pos), dtype


def test_goldens_present():
    assert len(CASES) >= 4


@pytest.mark.parametrize("path", CASES, ids=os.path.basename)
def test_fp32_contract_reproduces_vllm_cuda_op_bit_for_bit(path):
    z, t, (T, Hkv, Dh, rot, neox, max_pos), dtype = _load(path)
    k, want = t(z["k"]), t(z["k_rot"])
    # the golden stores the cos/sin rows of its positions: rebuild a cache that holds them at those positions
    pos = torch.from_numpy(z["positions"])
    cache = torch.zeros(max_pos, rot, dtype=dtype)
    cache[pos] = t(z["cos_sin_rows"])
    assert torch.equal(cache[pos], R.make_cos_sin_cache(max_pos, rot, dtype)[pos])      # same cache construction
    got = R.rope_torch(pos, k, cache, Dh, rot, neox, "fp32")
    assert torch.equal(got.view(torch.int16), want.view(torch.int16))
    native = R.rope_torch(pos, k, cache, Dh, rot, neox, "storage")
    assert not torch.equal(native.view(torch.int16), want.view(torch.int16))            # forward_native is NOT the op
    rel = ((native.float() - want.float()).norm() / want.float().norm()).item()
    assert rel < 2.0 ** -(7 if dtype == torch.bfloat16 else 10)
    if rot < Dh:                                                                        # pass-through part untouched
        assert torch.equal(got.view(T, Hkv, Dh)[..., rot:], k.view(T, Hkv, Dh)[..., rot:])


def test_reference_sequence_and_cache_scatter_match_the_bulk_oracle():
    L, T, Hkv, Dh, bs = 3, 37, 2, 64, 16
    n = Hkv * Dh
    g = torch.Generator().manual_seed(0)
    kv_list = [torch.randn(T, 2 * n, generator=g).to(torch.bfloat16) for _ in range(L)]
    pos = torch.randint(0, 512, (T,), generator=g)
    cs = R.make_cos_sin_cache(512, Dh, torch.bfloat16)
    k_states, v_states = R.reference_sequence(kv_list, pos, cs, Hkv, Dh, Dh, True)
    assert k_states.shape == (T, L * n) and torch.equal(v_states[:, n:2 * n], kv_list[1][:, n:])   # V is a copy
    keys = k_states.view(torch.int16).numpy().view(np.uint16)
    vals = v_states.view(torch.int16).numpy().view(np.uint16)
    _, _, sm, nb = K.make_kv_inputs(L, T, Hkv, Dh, bs, np.uint16, "perm", 0.1, seed=3)
    backing, kc, vc = K.make_caches(L, nb, bs, Hkv, Dh, np.uint16, "flash", "random", seed=3)
    tb = [torch.from_numpy(b.view(np.int16).copy()).view(torch.bfloat16) for b in backing]
    R.write_caches(k_states, v_states, [b[0] for b in tb], [b[1] for b in tb], torch.from_numpy(sm), Hkv, Dh)
    K.kv_write_numpy(keys, vals, kc, vc, sm, Hkv, Dh)
    for l in range(L):
        assert np.array_equal(tb[l].view(torch.int16).numpy().view(np.uint16), backing[l])

"""CPU: the two restatements of the reference's bulk KV kernel (numpy scatter, plain-C element
loop) agree with each other on every edge case, and with golden vectors produced on a B200 by the
reference's own kernel (tests/golden/kv_write_*.npz, generator oracle/gen_golden_kv_gpu.py)."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import kv_write_ref as K

GOLD = sorted(glob.glob(os.path.join(GOLDEN, "kv_write_*.npz")))


def _views(backing, layout):
    if layout == "flash":
        return [b[0] for b in backing], [b[1] for b in backing]
    return [b[:, 0] for b in backing], [b[:, 1] for b in backing]


@pytest.mark.parametrize("layout", ["flash", "flashinfer"])
@pytest.mark.parametrize("wdt", [np.uint16, np.uint32])
@pytest.mark.parametrize("L,T,Hkv,Dh,bs,extra,pad", [
    (4, 37, 2, 16, 16, 0, 0), (1, 1, 1, 8, 1, 0, 0), (3, 50, 4, 32, 8, 3, 8), (2, 13, 3, 20, 4, 0, 0),
    (5, 0, 2, 16, 16, 0, 0),
])
def test_numpy_and_c_agree(layout, wdt, L, T, Hkv, Dh, bs, extra, pad):
    keys, values, sm, nb = K.make_kv_inputs(L, T, Hkv, Dh, bs, wdt, "perm", 0.1, seed=1,
                                            extra_rows=extra, row_pad=pad)
    b1, kc1, vc1 = K.make_caches(L, nb, bs, Hkv, Dh, wdt, layout, "random", 1)
    b2, kc2, vc2 = K.make_caches(L, nb, bs, Hkv, Dh, wdt, layout, "random", 1)
    init = [b.copy() for b in b1]
    K.kv_write_numpy(keys, values, kc1, vc1, sm, Hkv, Dh)
    K.kv_write_c(keys, values, kc2, vc2, sm, Hkv, Dh)
    for a, b in zip(b1, b2):
        assert np.array_equal(a, b)
    # untouched slots keep their bytes, written slots hold the source rows
    n = Hkv * Dh
    valid = sm >= 0
    for l in range(L):
        kc = kc1[l].reshape(nb * bs, n) if layout == "flash" else None
        if kc is not None and T:
            assert np.array_equal(kc[sm[valid]], keys[:T][valid, l * n:(l + 1) * n])
            mask = np.ones(nb * bs, bool)
            mask[sm[valid]] = False
            assert np.array_equal(kc[mask], init[l][0].reshape(nb * bs, n)[mask])


def test_zero_layers_is_noop():
    keys, values, sm, nb = K.make_kv_inputs(1, 4, 1, 8, 4)
    K.kv_write_numpy(keys, values, [], [], sm, 1, 8)
    K.kv_write_c(keys, values, [], [], sm, 1, 8)


def test_all_padded_tokens_write_nothing():
    keys, values, sm, nb = K.make_kv_inputs(2, 9, 2, 8, 4, frac_pad=0.0)
    sm[:] = -1
    b, kc, vc = K.make_caches(2, nb, 4, 2, 8, fill="random")
    init = [x.copy() for x in b]
    K.kv_write_c(keys, values, kc, vc, sm, 2, 8)
    assert all(np.array_equal(x, y) for x, y in zip(b, init))


def test_block_stride_mismatch_raises():
    keys, values, sm, nb = K.make_kv_inputs(1, 4, 1, 8, 4)
    b, kc, _ = K.make_caches(1, nb, 4, 1, 8, layout="flash")
    b2, _, vc2 = K.make_caches(1, nb, 4, 1, 8, layout="flashinfer")
    with pytest.raises(RuntimeError):
        K.kv_write_numpy(keys, values, kc, vc2, sm, 1, 8)


def test_reference_unit_test_shape():
    """tests/unit_tests/test_custom_ops.py:56-90 passes 2-D caches [T, H*Dh]: block_size becomes
    size(1) = H*Dh and block_stride = stride(0) = H*Dh (kernels.cu:109-113).  With slots < T the
    write lands in row `slot`."""
    L, T, H, Dh = 4, 2, 2, 16
    n = H * Dh
    rng = np.random.default_rng(0)
    keys = rng.integers(0, 2 ** 32 - 1, size=(T, L * n), dtype=np.uint32)
    values = rng.integers(0, 2 ** 32 - 1, size=(T, L * n), dtype=np.uint32)
    kc = [np.zeros((T, n), np.uint32) for _ in range(L)]
    vc = [np.zeros((T, n), np.uint32) for _ in range(L)]
    sm = np.array([1, 0], dtype=np.int64)
    K.kv_write_c(keys, values, kc, vc, sm, H, Dh)
    for l in range(L):
        assert np.array_equal(kc[l][1], keys[0, l * n:(l + 1) * n])
        assert np.array_equal(vc[l][0], values[1, l * n:(l + 1) * n])


def test_algorithmic_bytes_cfg3():
    # cfg3 = 8192 tokens x 16 layers x n=1024 bf16
    assert K.algorithmic_bytes(8192, 8192, 16, 1024, 2) == 1_073_807_360


@pytest.mark.skipif(not GOLD, reason="kv golden fixtures not generated yet")
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_oracle_matches_reference_kernel_golden(path, impl):
    z = np.load(path)
    L, T, Hkv, Dh, bs, nb, er, rp = (int(v) for v in z["meta"])
    layout = str(z["layout"])
    backing = [z[f"init{l}"].copy() for l in range(L)]
    kc, vc = _views(backing, layout)
    fn = K.kv_write_numpy if impl == "numpy" else K.kv_write_c
    fn(z["keys"], z["values"], kc, vc, z["slots"], Hkv, Dh)
    for l in range(L):
        assert np.array_equal(backing[l], z[f"final{l}"]), f"layer {l}"

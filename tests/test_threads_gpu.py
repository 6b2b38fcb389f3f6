"""GPU: the library is safe to drive from several host threads on several streams:
per-thread error string, launch counter, trace state and tensor-map cache; per-device one-time initialisation
under a mutex.  Two threads hammer the eager draft loop (different shapes, own streams) while a third writes KV
caches; every result must equal the single-threaded one."""
import threading

import pytest
import torch

from oracle import speculator_ref as O

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def test_two_threads_two_streams_plus_kv_writer():
    from arcticinference_b200 import _capi, py_custom_ops
    from arcticinference_b200.speculator import ArcticLSTMSpeculator
    assert py_custom_ops.try_load_torch_library()
    shapes = [(4096, 512, 512, 3, 48), (3000, 256, 384, 4, 33)]
    models, inputs, expect = [], [], []
    for V, H, D, S, B in shapes:
        w = O.make_weights(V, H, D, torch.bfloat16, seed=0, device=DEV)
        m = ArcticLSTMSpeculator(vocab_size=V, input_hidden_dim=H, inner_dim=D, n_predict=S, max_num_seqs=B,
                                 device=DEV, enforce_eager=True)        # eager: every call goes through the C ABI
        m.load_weights(w.items())
        m.qhead = None
        ids, hid = O.make_inputs(B, V, H, torch.bfloat16, seed=1, device=DEV)
        models.append(m); inputs.append((ids, hid, S)); expect.append(m.generate_proposals(ids, hid, S).clone())
    # KV writer inputs
    L, T, Hkv, Dh, bs = 4, 512, 8, 128, 16
    n, nb = Hkv * Dh, 2 * T // bs
    keys = torch.randn(T, L * n, device=DEV).to(torch.bfloat16)
    vals = torch.randn(T, L * n, device=DEV).to(torch.bfloat16)
    slots = torch.randperm(nb * bs, device=DEV)[:T].to(torch.int64)
    sc = [torch.ones((), device=DEV) for _ in range(L)]
    torch.cuda.synchronize()
    errors, launches = [], {}

    def spec_worker(i):
        try:
            stream = torch.cuda.Stream(device=DEV)
            ids, hid, S = inputs[i]
            n0 = _capi.lib().ab200_launch_count()
            with torch.cuda.stream(stream):
                for it in range(60):
                    got = models[i].generate_proposals(ids, hid, S)
                    if it % 10 == 9:
                        stream.synchronize()
                        if not torch.equal(got, expect[i]):
                            errors.append(f"thread {i} iteration {it}: tokens differ")
                stream.synchronize()
            launches[i] = _capi.lib().ab200_launch_count() - n0      # per-thread counter
            # an error raised on this thread is reported on this thread
            rc = _capi.lib().ab200_rms_scale(None, None, 1, 8, 8, 8, 1e-6, 1.0, 2, 1, None)
            if rc == 0 or b"rms_scale" not in _capi.lib().ab200_last_error():
                errors.append(f"thread {i}: error string not thread-local")
        except Exception as e:  # noqa: BLE001
            errors.append(f"thread {i}: {type(e).__name__}: {e}")

    def kv_worker():
        try:
            stream = torch.cuda.Stream(device=DEV)
            with torch.cuda.stream(stream):
                for it in range(40):
                    kv = [torch.zeros(2, nb, bs, Hkv, Dh, device=DEV, dtype=torch.bfloat16) for _ in range(L)]
                    py_custom_ops.reshape_and_cache_flash_bulk(keys, vals, [x[0] for x in kv], [x[1] for x in kv], slots,
                                                               "auto", sc, sc, Hkv, Dh)
                    if it % 10 == 9:
                        stream.synchronize()
                        for l in range(L):
                            if not torch.equal(kv[l][0][slots // bs, slots % bs].reshape(T, n), keys[:, l * n:(l + 1) * n]):
                                errors.append(f"kv writer iteration {it} layer {l}: cache differs")
                stream.synchronize()
        except Exception as e:  # noqa: BLE001
            errors.append(f"kv writer: {type(e).__name__}: {e}")

    threads = [threading.Thread(target=spec_worker, args=(0,)), threading.Thread(target=spec_worker, args=(1,)),
               threading.Thread(target=kv_worker)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not any(t.is_alive() for t in threads), "a worker hung"
    assert not errors, errors
    # 60 proposals x (1 + 3 S) launches each, counted per thread
    for i, (_, _, _, S, _) in enumerate(shapes):
        assert launches[i] == 60 * (1 + 3 * S), (i, launches[i])
    assert _capi.lib().ab200_last_error() == b"" or b"rms_scale" not in _capi.lib().ab200_last_error()


def test_first_use_inside_capture_is_refused_not_corrupted():
    """ab200_init() at construction makes capture safe; the C ABI itself refuses a FIRST use of a table-reading
    kernel inside a stream capture instead of recording a table fill into the graph."""
    from arcticinference_b200 import _capi
    lib = _capi.lib()
    _capi.init_device(DEV)          # idempotent
    _capi.init_device(DEV)
    g = torch.cuda.CUDAGraph()
    x = torch.randn(4, 64, device=DEV).to(torch.bfloat16)
    y = torch.empty_like(x)
    with torch.cuda.graph(g):       # already initialised: capturing works
        _capi.check(lib.ab200_rms_scale(x.data_ptr(), y.data_ptr(), 4, 64, 64, 64, 1e-6, 1.0, _capi.AB200_BF16, 1,
                                        torch.cuda.current_stream().cuda_stream))
    g.replay()
    torch.cuda.synchronize()
    assert torch.isfinite(y.float()).all()

"""GPU: no kernel writes outside its output.  compute-sanitizer is closed on this pool, so the outputs of the
kernels that compute addresses from tile / slot / row arithmetic are carved out of the MIDDLE of a canary-filled
arena and the margins are checked byte for byte afterwards (ragged shapes on purpose)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
PAD = 1 << 16          # canary bytes on either side


class Arena:
    def __init__(self, nbytes):
        self.buf = torch.full((nbytes + 2 * PAD,), 0xA5, dtype=torch.uint8, device=DEV)
        self.nbytes = nbytes

    def view(self, dtype, shape):
        return self.buf[PAD:PAD + self.nbytes].view(dtype).view(shape)

    def intact(self):
        return bool((self.buf[:PAD] == 0xA5).all() and (self.buf[PAD + self.nbytes:] == 0xA5).all())


@pytest.mark.parametrize("M,N,K", [(5, 1000, 72), (64, 16384, 4096), (96, 11520, 2880), (129, 4104, 520), (300, 4096, 512),
                                   (512, 2048, 1024), (700, 1032, 256)])
def test_linear_stays_inside_its_output(M, N, K):
    from arcticinference_b200 import _capi
    x = torch.randn(M, K, device=DEV).to(torch.bfloat16)
    w = (torch.randn(N, K, device=DEV) * 0.05).to(torch.bfloat16)
    arena = Arena(M * N * 2)
    out = arena.view(torch.bfloat16, (M, N))
    _capi.check(_capi.lib().ab200_linear(x.data_ptr(), w.data_ptr(), out.data_ptr(), M, N, K, K, K, N, _capi.AB200_BF16,
                                         torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert arena.intact()
    assert (out == torch.nn.functional.linear(x, w)).float().mean().item() > 0.99


@pytest.mark.parametrize("M,V,K", [(7, 1000, 64), (64, 50000, 512), (300, 32000, 256), (512, 9999 // 8 * 8, 128)])
def test_head_argmax_stays_inside_its_outputs(M, V, K):
    from arcticinference_b200 import _capi
    Vp = (V + 63) // 64 * 64
    x = torch.randn(M, K, device=DEV).to(torch.bfloat16)
    w = torch.zeros(Vp, K, device=DEV, dtype=torch.bfloat16)
    w[:V] = (torch.randn(V, K, device=DEV) * 0.02).to(torch.bfloat16)
    a_packed, a_tok = Arena((M + 1) * 8), Arena(M * 8)
    packed = a_packed.view(torch.int64, (M + 1,))
    packed.zero_()
    tok = a_tok.view(torch.int64, (M,))
    _capi.check(_capi.lib().ab200_head_argmax(x.data_ptr(), w.data_ptr(), M, Vp, K, V, K, K, packed.data_ptr(), tok.data_ptr(), 1,
                                              packed[M:].data_ptr(), _capi.AB200_BF16, None, 0,
                                              torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert a_packed.intact() and a_tok.intact()
    assert ((tok >= 0) & (tok < V)).all()


def test_rope_write_stays_inside_caches_and_selected_rows():
    from arcticinference_b200 import swiftkv
    from oracle import kv_rope_ref as R
    L, T, Hkv, Dh, bs, nb = 5, 333, 8, 64, 16, 42                     # exactly enough blocks: the last slot is the last row
    n = Hkv * Dh
    kv_list = [torch.randn(T, 2 * n, device=DEV).to(torch.bfloat16) for _ in range(L)]
    arenas = [Arena(2 * nb * bs * n * 2) for _ in range(L)]
    caches = [a.view(torch.bfloat16, (2, nb, bs, Hkv, Dh)) for a in arenas]
    slots = torch.randperm(nb * bs, device=DEV)[:T].to(torch.int64)
    slots[0], slots[1] = nb * bs - 1, 0                               # first and last cache row are written
    pos = torch.randint(0, 1024, (T,), device=DEV)
    cs = R.make_cos_sin_cache(1024, Dh, torch.bfloat16, device=DEV)
    sel = torch.tensor([0, T - 1, -1, 17], device=DEV)
    a_k, a_v = Arena(4 * L * n * 2), Arena(4 * L * n * 2)
    sc = [torch.ones((), device=DEV) for _ in range(L)]
    swiftkv.kv_rope_write_bulk(kv_list, pos, cs, [c[0] for c in caches], [c[1] for c in caches], slots, "auto", sc, sc, Hkv, Dh,
                               logits_indices=sel, k_states_out=a_k.view(torch.bfloat16, (4, L * n)),
                               v_states_out=a_v.view(torch.bfloat16, (4, L * n)))
    torch.cuda.synchronize()
    assert all(a.intact() for a in arenas) and a_k.intact() and a_v.intact()


def test_cell_and_rms_stay_inside_their_outputs():
    from arcticinference_b200 import _capi
    lib, st = _capi.lib(), torch.cuda.current_stream().cuda_stream
    B, D, V = 37, 2880, 999
    gates = torch.randn(B, 4 * D, device=DEV).to(torch.bfloat16)
    emb = (torch.randn(V, D, device=DEV) * 0.02).to(torch.bfloat16)
    tok = torch.randint(0, V, (B,), device=DEV)
    ln = [torch.randn(D, device=DEV).to(torch.bfloat16) for _ in range(4)]
    a_s, a_c, a_q, a_sc = Arena(B * D * 2), Arena(B * D * 2), Arena(B * D), Arena(B * 4)
    _capi.check(lib.ab200_sum_lstm_cell(gates.data_ptr(), 4 * D, emb.data_ptr(), tok.data_ptr(), None, None, ln[0].data_ptr(),
                                        ln[1].data_ptr(), ln[2].data_ptr(), ln[3].data_ptr(),
                                        a_s.view(torch.bfloat16, (B, D)).data_ptr(), a_c.view(torch.bfloat16, (B, D)).data_ptr(),
                                        B, D, V, 23.0, 1e-6, _capi.AB200_BF16, 1, None, None,
                                        a_q.view(torch.uint8, (B, D)).data_ptr(), a_sc.view(torch.float32, (B,)).data_ptr(), st))
    x = torch.randn(B, D, device=DEV).to(torch.bfloat16)
    a_y = Arena(B * D * 2)
    _capi.check(lib.ab200_rms_scale(x.data_ptr(), a_y.view(torch.bfloat16, (B, D)).data_ptr(), B, D, D, D, 1e-6, 0.7071,
                                    _capi.AB200_BF16, 1, st))
    torch.cuda.synchronize()
    assert all(a.intact() for a in (a_s, a_c, a_q, a_sc, a_y))

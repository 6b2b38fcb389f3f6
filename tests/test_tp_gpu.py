"""GPU: vocab-sharded (tensor-parallel) head of the LSTM speculator - SURVEY 8f rank 4.
The reference merges per-rank top-1s with topk + all_gather + argmax + gather
(arctic_speculator.py:733-744); here every rank turns its shard-local packed arg-max into a key and
ONE all_reduce(MAX) over int64 combines them.  On a single GPU the ranks are simulated (each rank's
model holds its own head shard; the test double max-combines their keys exactly as NCCL would);
`tools/tp_nccl_check.py` runs the same comparison over real NCCL on 2+ GPUs."""
import pytest
import torch

from oracle import speculator_ref as O

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


class SimGroup:
    """Stands in for the NCCL group: all ranks' states are identical in TP (everything but the head is
    replicated), so rank 0 can ask every other rank's model for its keys and MAX them elementwise."""

    def __init__(self):
        self.others = []

    def all_reduce_max(self, keys, states):
        for m in self.others:
            keys = torch.maximum(keys, m.local_argmax_keys(states))
        return keys


def reference_tp_tokens(logits_full, tp, per):
    """arctic_speculator.py:733-744 verbatim on per-rank logit slices (padding columns included)."""
    B = logits_full.shape[0]
    vals_l, idx_l = [], []
    for r in range(tp):
        logits = logits_full[:, r * per:(r + 1) * per]
        vals, indices = torch.topk(logits, 1, dim=-1)
        indices = indices + r * logits.shape[-1]
        vals_l.append(vals.to(torch.float64))
        idx_l.append(indices)
    vals = torch.cat(vals_l, dim=-1)
    indices = torch.cat(idx_l, dim=-1)
    argidx = torch.argmax(vals, -1).reshape(B, -1)
    return torch.gather(indices, -1, argidx).flatten()


@pytest.mark.parametrize("V,H,D,tp,dtype", [(128256, 4096, 4096, 8, torch.bfloat16), (32000, 1024, 1024, 2, torch.bfloat16),
                                            (5000, 512, 256, 4, torch.float32), (2051, 192, 128, 2, torch.bfloat16)])
def test_tp_proposals_equal_single_rank(V, H, D, tp, dtype):
    from arcticinference_b200.speculator import ArcticLSTMSpeculator, pad_vocab_size
    S, B = 3, 48
    w = O.make_weights(V, H, D, dtype, seed=0, device=DEV)
    single = ArcticLSTMSpeculator(vocab_size=V, input_hidden_dim=H, inner_dim=D, n_predict=S, max_num_seqs=B,
                                  dtype=dtype, device=DEV, enforce_eager=True)
    single.load_weights(w.items())
    group = SimGroup()
    ranks = [ArcticLSTMSpeculator(vocab_size=V, input_hidden_dim=H, inner_dim=D, n_predict=S, max_num_seqs=B, dtype=dtype,
                                  device=DEV, enforce_eager=True, tp_size=tp, tp_rank=r, tp_group=group) for r in range(tp)]
    for m in ranks:
        m.load_weights(w.items())                                   # every rank keeps its vocab partition
        assert dict(m.named_parameters())["head.0.weight"].shape[0] == pad_vocab_size(V) // tp
    group.others = ranks[1:]
    for seed in range(4):
        ids, hid = O.make_inputs(B, V, H, dtype, seed=90 + seed, device=DEV)
        exp = single.generate_proposals(ids, hid, S)
        got = ranks[0].generate_proposals(ids, hid, S)
        if pad_vocab_size(V) == V:
            assert torch.equal(got, exp)        # no padding rows anywhere: identical to TP = 1, bit for bit
        else:
            # padding rows (zero logits) take part under TP in the reference; they can only win when every
            # real logit is negative - compare against the reference's own merge on materialised logits
            assert (got == exp).float().mean().item() > 0.9
    # one step against the reference's merge sequence on the same states
    states = (torch.randn(B, D, device=DEV) * 0.3).to(dtype)
    Vp = pad_vocab_size(V)
    logits = torch.nn.functional.linear(states, w["head.0.weight"])            # [B, Vp], padding rows zero
    ref = reference_tp_tokens(logits, tp, Vp // tp)
    got = ranks[0].greedy_tokens(states)
    if dtype == torch.float32:
        picked = logits.gather(1, got.unsqueeze(1)).squeeze(1)
        assert torch.allclose(picked, logits.gather(1, ref.unsqueeze(1)).squeeze(1), rtol=1e-5, atol=1e-6)
    else:
        assert (got == ref).float().mean().item() >= 0.97


def test_shard_key_order_and_ties():
    from arcticinference_b200 import _capi
    lib = _capi.lib()
    # two shards report the same bf16 logit for one row: the lower GLOBAL index must win; an empty shard loses
    def key(bits_hi, idx, off):
        p = torch.tensor([(bits_hi << 32) | (0xFFFFFFFF - idx)], dtype=torch.int64, device=DEV) if bits_hi else \
            torch.zeros(1, dtype=torch.int64, device=DEV)
        _capi.check(lib.ab200_argmax_shard_key(p.data_ptr(), 1, off, torch.cuda.current_stream().cuda_stream))
        return p
    a = key(0x0F800000, 5, 0)          # shard 0, local 5
    b = key(0x0F800000, 2, 1000)       # shard 1, local 2 -> global 1002, same value
    c = key(0, 0, 2000)                # empty shard
    win = torch.maximum(torch.maximum(a, b), c)
    tok = torch.empty(1, dtype=torch.long, device=DEV)
    _capi.check(lib.ab200_argmax_unpack(win.data_ptr(), tok.data_ptr(), 1, 1, torch.cuda.current_stream().cuda_stream))
    assert tok.item() == 5
    d = key(0x0F900000, 7, 1000)       # larger value in shard 1 wins regardless of index
    win = torch.maximum(a, d)
    _capi.check(lib.ab200_argmax_unpack(win.data_ptr(), tok.data_ptr(), 1, 1, torch.cuda.current_stream().cuda_stream))
    assert tok.item() == 1007

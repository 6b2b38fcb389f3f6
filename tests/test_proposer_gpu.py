"""GPU: drafter glue - ArcticProposer.prepare_hidden_states against the
reference's torch op sequence (arctic_proposer.py:137-147), and propose() end to end with host
inputs against generate_proposals."""
import numpy as np
import pytest
import torch

from oracle import speculator_ref as O

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def reference_prepare(sample_hidden_states, sampled_token_ids, num_draft_tokens):
    """arctic_proposer.py:133-147, verbatim op sequence."""
    max_gen_len = sampled_token_ids.shape[-1]
    if max_gen_len == 1:
        return sample_hidden_states
    valid_mask = sampled_token_ids != -1
    gen_lens = valid_mask.sum(dim=1)
    num_sampled_tokens = torch.tensor(np.array(num_draft_tokens), device=gen_lens.device) + 1
    hidden_states_idx = (gen_lens - 1) + torch.cumsum(num_sampled_tokens, 0) - num_sampled_tokens
    return sample_hidden_states[hidden_states_idx]


@pytest.fixture(scope="module")
def proposer():
    from arcticinference_b200.speculator import ArcticLSTMSpeculator, ArcticProposer
    V, H, D, S = 4096, 512, 512, 3
    w = O.make_weights(V, H, D, torch.bfloat16, seed=0, device=DEV)
    m = ArcticLSTMSpeculator(vocab_size=V, input_hidden_dim=H, inner_dim=D, n_predict=S, max_num_seqs=32, device=DEV)
    m.load_weights(w.items())
    return ArcticProposer(m), w, (V, H, D, S)


@pytest.mark.parametrize("R,G", [(1, 4), (7, 4), (32, 6), (5, 2)])
def test_prepare_hidden_states(proposer, R, G):
    p, _, (V, H, D, S) = proposer
    g = torch.Generator(device="cpu").manual_seed(R * 10 + G)
    num_draft = [int(x) for x in torch.randint(0, G, (R,), generator=g)]          # drafts per request (<= G-1)
    total = sum(n + 1 for n in num_draft)
    hidden = torch.randn(total, H, generator=g).to(torch.bfloat16).to(DEV)
    ids = torch.full((R, G), -1, dtype=torch.long)
    for r, n in enumerate(num_draft):
        acc = int(torch.randint(1, n + 2, (1,), generator=g))                     # 1..n+1 accepted tokens
        ids[r, :acc] = torch.randint(0, V, (acc,), generator=g)
    ids = ids.to(DEV)
    got = p.prepare_hidden_states(hidden, ids, num_draft)
    exp = reference_prepare(hidden, ids, num_draft)
    assert torch.equal(got, exp)


def test_prepare_hidden_states_single_token_passthrough(proposer):
    p, _, (V, H, D, S) = proposer
    hidden = torch.randn(5, H, device=DEV).to(torch.bfloat16)
    ids = torch.randint(0, V, (5, 1), device=DEV)
    assert p.prepare_hidden_states(hidden, ids, None) is hidden
    with pytest.raises(AssertionError):
        p.prepare_hidden_states(hidden[:, :8], ids, None)


@pytest.mark.parametrize("B", [1, 5, 32, 40])
def test_propose_host_inputs_equals_generate_proposals(proposer, B):
    p, w, (V, H, D, S) = proposer
    ids, hid = O.make_inputs(B, V, H, torch.bfloat16, seed=77 + B)
    exp = p.model.generate_proposals(ids.to(DEV), hid.to(DEV), S).cpu().numpy()
    got_pinned = p.propose(ids.numpy(), hid.pin_memory(), S)
    got_dev = p.propose(ids.numpy(), hid.to(DEV), S)
    assert isinstance(got_pinned, np.ndarray) and got_pinned.shape == (B, S) and got_pinned.dtype == np.int64
    assert np.array_equal(got_pinned, exp) and np.array_equal(got_dev, exp)
    again = p.propose(ids.numpy(), hid.to(DEV), S)           # the returned array must not alias the staging buffer
    assert np.array_equal(got_dev, again) and got_dev is not again

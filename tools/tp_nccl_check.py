"""Real NCCL check of the tensor-parallel head (run with torchrun on >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/tp_nccl_check.py

Every rank holds its vocab partition of the head; proposals over NCCL (one MAX all-reduce of [B] int64
per step) must equal the single-rank result bit for bit; reports time per proposal (eager, TP path)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from arcticinference_b200 import replicas
from arcticinference_b200.speculator import ArcticLSTMSpeculator
from oracle import speculator_ref as O

rank, local_rank, world = replicas.env_rank_world()
dev = torch.device(f"cuda:{local_rank}")
torch.cuda.set_device(dev)
replicas.init_from_env("nccl", dev)
V, H, D, S, B = 128256, 4096, 4096, 3, 64
w = O.make_weights(V, H, D, torch.bfloat16, seed=0, device=dev)
tp = ArcticLSTMSpeculator(vocab_size=V, input_hidden_dim=H, inner_dim=D, n_predict=S, max_num_seqs=B, device=dev,
                          enforce_eager=True, tp_size=world, tp_rank=rank, tp_group=None)
tp.load_weights(w.items())
single = ArcticLSTMSpeculator(vocab_size=V, input_hidden_dim=H, inner_dim=D, n_predict=S, max_num_seqs=B, device=dev,
                              enforce_eager=True)
single.load_weights(w.items())
ok = True
for seed in range(6):
    ids, hid = O.make_inputs(B, V, H, torch.bfloat16, seed=300 + seed, device=dev)
    got, exp = tp.generate_proposals(ids, hid, S), single.generate_proposals(ids, hid, S)
    ok &= bool(torch.equal(got, exp))
for _ in range(5):
    tp.generate_proposals(ids, hid, S)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    tp.generate_proposals(ids, hid, S)
e1.record(); torch.cuda.synchronize()
flag = torch.tensor([int(ok)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"TP={world} over NCCL: tokens equal to TP=1 on all ranks: {bool(flag.item())}; "
          f"{e0.elapsed_time(e1) / 50 * 1e3:.1f} us per proposal (eager per-op path, head bytes / {world} per rank)")
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)

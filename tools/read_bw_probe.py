import torch
dev = torch.device("cuda:0")
n = 1 << 30  # 4 GB float32
x = torch.empty(n, dtype=torch.float32, device=dev).normal_()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, fn in (("sum_f32", lambda: x.sum()), ("max_f32", lambda: x.max()), ("sum_bf16view", lambda: x.view(torch.bfloat16).sum(dtype=torch.float32)),
                 ("copy_half", lambda: x[: n // 2].copy_(x[n // 2:]))):
    ts = []
    for _ in range(6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = min(ts[1:])
    bytes_ = 4 * n if not name.startswith("copy") else 4 * n  # copy: read half + write half = 4n bytes total
    print(f"{name}: {t:.3f} ms -> {bytes_ / t / 1e6:.0f} GB/s")

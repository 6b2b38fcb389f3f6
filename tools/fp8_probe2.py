import torch
from vllm import _custom_ops as ops
dev = "cuda"; torch.manual_seed(1); FP8_MAX = 448.0
def cast(x): return x.clamp(-FP8_MAX, FP8_MAX).to(torch.float8_e4m3fn)
inv448 = torch.tensor(1.0 / 448.0, dtype=torch.float32, device=dev)
ok = {"w_div": 0, "w_mul": 0, "x_div": 0, "x_mul": 0}; n = 0
wd = wm = 0.0
for seed in range(20):
    torch.manual_seed(seed)
    w = (torch.randn(2048, 512, device=dev) * (0.01 + 0.01 * seed)).to(torch.bfloat16)
    x = (torch.randn(64, 512, device=dev) * (0.5 + seed)).to(torch.bfloat16)
    qw, sw = ops.scaled_fp8_quant(w, scale=None)
    amax = w.float().abs().max()
    ok["w_div"] += int(torch.equal(sw.flatten(), (amax / FP8_MAX).flatten())); ok["w_mul"] += int(torch.equal(sw.flatten(), (amax * inv448).flatten()))
    wd += (cast(w.float() / sw).view(torch.uint8) == qw.view(torch.uint8)).float().mean().item()
    wm += (cast(w.float() * (1.0 / sw)).view(torch.uint8) == qw.view(torch.uint8)).float().mean().item()
    qx, sx = ops.scaled_fp8_quant(x, scale=None, use_per_token_if_dynamic=True)
    at = x.float().abs().amax(-1, keepdim=True)
    ok["x_div"] += int(torch.equal(sx, at / FP8_MAX)); ok["x_mul"] += int(torch.equal(sx, at * inv448))
    n += 1
print(ok, "of", n, "weight conv div", wd / n, "mulinv", wm / n)

"""Dev: learn the exact arithmetic of vLLM's fp8 ops (the reference qhead's backend) on the GPU box."""
import torch
from vllm import _custom_ops as ops
dev = "cuda"
torch.manual_seed(0)
FP8_MAX = 448.0
def cast(x):  # clamp + RN-even cast
    return x.clamp(-FP8_MAX, FP8_MAX).to(torch.float8_e4m3fn)
V, D, M = 4096, 1024, 48
w = (torch.randn(V, D, device=dev) * 0.02).to(torch.bfloat16)
x = torch.randn(M, D, device=dev).to(torch.bfloat16)
# --- per-tensor weight quant
qw, sw = ops.scaled_fp8_quant(w, scale=None)
amax = w.float().abs().max()
print("w scale == amax/448:", float(sw), float(amax / FP8_MAX), torch.equal(sw.flatten(), (amax / FP8_MAX).flatten()))
for name, cand in {"div": cast(w.float() / sw), "mulinv": cast(w.float() * (1.0 / sw))}.items():
    print("  weight", name, (cand.view(torch.uint8) == qw.view(torch.uint8)).float().mean().item())
# --- per-token activation quant
qx, sx = ops.scaled_fp8_quant(x, scale=None, use_per_token_if_dynamic=True)
amax_t = x.float().abs().amax(-1, keepdim=True)
min_sf = 1.0 / (FP8_MAX * 512.0)
cand_s = torch.clamp(amax_t / FP8_MAX, min=min_sf)
print("x scales equal:", torch.equal(sx, cand_s), (sx - cand_s).abs().max().item(), sx.shape)
for name, cand in {"div": cast(x.float() / sx), "mulinv": cast(x.float() * (1.0 / sx))}.items():
    print("  act", name, (cand.view(torch.uint8) == qx.view(torch.uint8)).float().mean().item())
# --- scaled mm
out = ops.cutlass_scaled_mm(qx, qw.t(), scale_a=sx, scale_b=sw, out_dtype=torch.bfloat16)
acc = qx.float() @ qw.float().t()
acc64 = (qx.double() @ qw.double().t())
for name, cand in {"sa*(sb*acc)": sx * (sw * acc), "(sa*sb)*acc": (sx * sw) * acc, "(acc*sa)*sb": (acc * sx) * sw,
                   "f64": (acc64 * sx.double() * sw.double())}.items():
    c = cand.to(torch.bfloat16)
    print("  mm", name, (c == out).float().mean().item())
print("argmax agree (sa*(sb*acc)):", ((sx * (sw * acc)).to(torch.bfloat16).argmax(-1) == out.argmax(-1)).float().mean().item())
# zero row / tiny row behaviour
xz = torch.zeros(2, D, device=dev, dtype=torch.bfloat16); xz[1, 3] = 1e-7
qz, sz = ops.scaled_fp8_quant(xz, scale=None, use_per_token_if_dynamic=True)
print("zero-row scale", sz.flatten().tolist(), "min_sf", min_sf, qz.float().abs().max().item())

"""Dev (GPU box): the CTA-pair GEMM with ONE vs TWO k-blocks per TMA request (debug flavour, tunables pair_gate_kbs /
pair_head_kbs): bit-equality of the outputs and stand-alone device time [entry -> last exit, first -> last data].
    timeout 300 python tools/pair_kbs_probe.py"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from arcticinference_b200 import _capi

lib = _capi.debug_lib()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(384 << 20, dtype=torch.uint8, device=dev)
I64MAX = torch.iinfo(torch.int64).max


def tune(**kw):
    for k, v in kw.items():
        assert lib.ab200_debug_set_tunable(k.encode(), int(v)) == 0, k


def run(M, N, K, head, cl4, iters=8):
    tune(pair_gate_kbs=cl4 + 1, pair_head_kbs=cl4 + 1)
    g = torch.Generator(device=dev).manual_seed(M + N + K)
    x = torch.randn(M, K, device=dev, generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev, generator=g) * 0.02).to(torch.bfloat16)
    out = torch.zeros(M, N, dtype=torch.bfloat16, device=dev)
    packed = torch.zeros(M + 1, dtype=torch.int64, device=dev)
    toks = torch.zeros(M, dtype=torch.int64, device=dev)
    trace = torch.zeros(1, 8, dtype=torch.int64, device=dev)

    def launch():
        if head:
            packed.zero_()
            rc = lib.ab200_head_argmax(x.data_ptr(), w.data_ptr(), M, N, K, N, K, K, packed.data_ptr(), toks.data_ptr(), 1,
                                       packed[M:].data_ptr(), 2, None, 0, st)
        else:
            rc = lib.ab200_linear(x.data_ptr(), w.data_ptr(), out.data_ptr(), M, N, K, K, K, N, 2, st)
        _capi.check(rc, "gemm", lib)

    launch()
    torch.cuda.synchronize()
    res = toks.clone() if head else out.clone()
    ts = []
    for _ in range(iters):
        flush.zero_()
        trace.zero_(); trace[:, [0, 1, 3, 5]] = I64MAX
        torch.cuda.synchronize()
        lib.ab200_set_trace_buffer(trace.data_ptr(), 1)
        launch()
        torch.cuda.synchronize()
        lib.ab200_set_trace_buffer(None, 0)
        t = trace[0].cpu().tolist()
        ts.append(((t[2] - t[0]) / 1e3, (t[4] - t[3]) / 1e3))
    ts.sort()
    return res, ts[len(ts) // 2]


lib.ab200_init(st)
for (M, N, K, head) in ((256, 16384, 4096, False), (200, 16384 + 96, 4096, False), (512, 16384, 4096, False),
                        (256, 11520, 2880, False), (512, 11520, 2880, False), (256, 16384, 8192, False),
                        (256, 128256, 4096, True), (256, 201088, 2880, True), (129, 32000, 1024, True),
                        (256, 128256, 8192, True), (130, 5000, 192, True), (300, 4096, 64, False)):
    a, ta = run(M, N, K, head, 0)
    b, tb = run(M, N, K, head, 1)
    same = bool(torch.equal(a, b))
    ref_ok = None
    if not head and M * N <= 512 * 16480:
        g = torch.Generator(device=dev).manual_seed(M + N + K)
        x = torch.randn(M, K, device=dev, generator=g).to(torch.bfloat16)
        w = (torch.randn(N, K, device=dev, generator=g) * 0.02).to(torch.bfloat16)
        ref = torch.nn.functional.linear(x.float(), w.float())
        ref_ok = float((b.float() - ref).abs().max())
    print(json.dumps({"M": M, "N": N, "K": K, "head": head, "us_kbs1": ta, "us_kbs2": tb, "bit_equal": same,
                      "max_abs_err_vs_fp32": ref_ok}), flush=True)
tune(pair_gate_kbs=0, pair_head_kbs=0)

"""Dev/profiling (GPU box): where a stand-alone weight-streaming GEMM spends its time, and a sweep of the launch
tunables of the DEBUG library flavour (libarctic_b200_dbg.so, ab200_debug_set_tunable).

For every configuration: CUDA-event time behind a spin kernel with an L2 flush in between ("standalone"), and the
device-side trace of the same launch: entry -> first operand bytes landed -> last bytes landed -> first / last CTA exit.
    python tools/gemm_tune.py [gate|head|all] > gpurun_out/gemm_tune.json
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from arcticinference_b200 import _capi

lib = _capi.debug_lib()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(384 << 20, dtype=torch.uint8, device=dev)
I64MAX = torch.iinfo(torch.int64).max
what = sys.argv[1] if len(sys.argv) > 1 else "all"


def tune(**kw):
    for k, v in kw.items():
        assert lib.ab200_debug_set_tunable(k.encode(), int(v)) == 0, k


def run_case(M, N, K, head, iters=12, warm_x=False, **tun):
    tune(**tun)
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16)
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
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

    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    ev, tr = [], []
    for _ in range(iters):
        flush.zero_()
        if head:
            packed.zero_()
        trace.zero_(); trace[:, [0, 1, 3, 5]] = I64MAX
        if warm_x:
            x.add_(0)      # the producer of x has just written it: L2-hot, as in the draft loop
        torch.cuda.synchronize()
        lib.ab200_set_trace_buffer(trace.data_ptr(), 1)
        torch.cuda._sleep(200000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if head:
            rc = lib.ab200_head_argmax(x.data_ptr(), w.data_ptr(), M, N, K, N, K, K, packed.data_ptr(), toks.data_ptr(), 1,
                                       packed[M:].data_ptr(), 2, None, 0, st)
        else:
            rc = lib.ab200_linear(x.data_ptr(), w.data_ptr(), out.data_ptr(), M, N, K, K, K, N, 2, st)
        e1.record()
        torch.cuda.synchronize()
        lib.ab200_set_trace_buffer(None, 0)
        t = trace[0].cpu().tolist()
        ev.append(e0.elapsed_time(e1) * 1e3)
        tr.append([(t[j] - t[0]) / 1e3 for j in (1, 3, 4, 5, 2)])
    ev.sort()
    med = lambda j: sorted(r[j] for r in tr)[len(tr) // 2]
    nbytes = N * K * 2 + M * K * 2 + (0 if head else M * N * 2)
    dev_us = med(4)
    res = {"M": M, "N": N, "K": K, "kind": "head" if head else "gate", "warm_x": warm_x, **tun, "event_us": ev[len(ev) // 2],
           "device_us(entry->last exit)": dev_us, "dep_ok": med(0), "first_data": med(1), "last_data": med(2),
           "first_exit": med(3), "stream_us": med(2) - med(1), "tail_us": dev_us - med(2),
           "gbs_device": nbytes / dev_us / 1e3, "gbs_streaming": N * K * 2 / max(med(2) - med(1), 1e-3) / 1e3}
    print(json.dumps(res), flush=True)
    return res


DEFAULTS = dict(gate_block_n=0, head_block_n=0, max_stages=0, pre_stages=-1, umma_m64=1, gate_kbs=2, head_kbs=1)
rows = []
if what in ("gate", "all"):
    for kbs in (1, 2):
        for (M, N, K) in ((64, 16384, 4096), (128, 16384, 4096), (64, 11520, 2880), (128, 11520, 2880), (64, 16384, 8192),
                          (64, 32768, 8192), (32, 16384, 4096)):
            rows.append(run_case(M, N, K, False, warm_x=True, **{**DEFAULTS, "gate_kbs": kbs}))
    for bn in (56, 64, 80, 96, 112, 128):
        rows.append(run_case(64, 16384, 4096, False, warm_x=True, **{**DEFAULTS, "gate_block_n": bn}))
    for stg in (3, 4):
        rows.append(run_case(64, 16384, 4096, False, warm_x=True, **{**DEFAULTS, "max_stages": stg}))
    for bn in (40, 48, 64, 80, 96):
        rows.append(run_case(64, 11520, 2880, False, warm_x=True, **{**DEFAULTS, "gate_block_n": bn}))
if what in ("pair",):
    # CTA-pair kernel (129..512 rows): where the cases that lose to cuBLAS spend their time
    for (M, N, K) in ((256, 16384, 4096), (512, 16384, 4096), (256, 11520, 2880), (512, 11520, 2880), (256, 32768, 8192)):
        rows.append(run_case(M, N, K, False, warm_x=True, **DEFAULTS))
        for bn in (96, 128, 192, 256):
            rows.append(run_case(M, N, K, False, warm_x=True, **{**DEFAULTS, "gate_block_n": bn}))
if what in ("pairk",):
    # CTA-pair kernel with two k-blocks per request: tile width / stage count sweep, head and gate shapes
    K2 = {**DEFAULTS, "pair_gate_kbs": 2, "pair_head_kbs": 2}
    for (M, N, K, head) in ((256, 128256, 4096, True), (256, 201088, 2880, True), (200, 128256, 4096, True)):
        rows.append(run_case(M, N, K, head, warm_x=True, **{**DEFAULTS, "pair_head_kbs": 1}))
        for bn in (0, 160, 192, 224, 256):
            rows.append(run_case(M, N, K, head, warm_x=True, **{**K2, "head_block_n": bn}))
    for (M, N, K) in ((256, 16384, 4096), (512, 16384, 4096), (256, 11520, 2880), (512, 11520, 2880), (256, 32768, 8192)):
        rows.append(run_case(M, N, K, False, warm_x=True, **{**DEFAULTS, "pair_gate_kbs": 1}))
        for bn in (0, 128, 160, 192, 224, 256):
            rows.append(run_case(M, N, K, False, warm_x=True, **{**K2, "gate_block_n": bn}))
    tune(pair_gate_kbs=0, pair_head_kbs=0)
if what in ("pairk3",):
    for kbs in (2, 3, 4):
        K2 = {**DEFAULTS, "pair_gate_kbs": kbs, "pair_head_kbs": kbs}
        for (M, N, K, head) in ((256, 128256, 4096, True), (256, 201088, 2880, True), (256, 16384, 4096, False),
                                (512, 16384, 4096, False), (256, 11520, 2880, False), (512, 11520, 2880, False)):
            rows.append(run_case(M, N, K, head, warm_x=True, **K2))
            if kbs == 3 and head:
                for bn in (192, 224):
                    rows.append(run_case(M, N, K, head, warm_x=True, **{**K2, "head_block_n": bn}))
    tune(pair_gate_kbs=0, pair_head_kbs=0)
if what in ("single",):
    # 129..512 rows: CTA pairs (default) vs the single-CTA kernel with 2 m-tiles (gemm_2cta = 0), gate shapes
    for (M, N, K) in ((256, 11520, 2880), (512, 11520, 2880), (256, 16384, 4096), (512, 16384, 4096), (200, 11520, 2880)):
        for two in (1, 0):
            rows.append(run_case(M, N, K, False, warm_x=True, **{**DEFAULTS, "gemm_2cta": two}))
    tune(gemm_2cta=1)
if what in ("bound",):
    # what bounds the CTA-pair kernel: the same launches with the x loads (1), the W loads (2) or both (3) left out
    # after the first ring fill (debug flavour only; results are garbage, the timing is the point)
    for (M, N, K, head) in ((256, 128256, 4096, True), (256, 16384, 4096, False), (512, 16384, 4096, False),
                            (512, 11520, 2880, False)):
        for skip in (0, 1, 2, 3):
            rows.append(run_case(M, N, K, head, warm_x=True, **{**DEFAULTS, "dbg_skip": skip}))
    tune(dbg_skip=0)
if what in ("head", "all"):
    for kbs in (1, 2):
        rows.append(run_case(64, 128256, 4096, True, warm_x=True, **{**DEFAULTS, "head_kbs": kbs}))
        rows.append(run_case(128, 128256, 4096, True, warm_x=True, **{**DEFAULTS, "head_kbs": kbs}))
    for bn in (112, 144, 160, 176):
        rows.append(run_case(64, 128256, 4096, True, warm_x=True, **{**DEFAULTS, "head_kbs": 2, "head_block_n": bn}))
tune(**DEFAULTS)
# cuBLAS reference for the gate shapes, same timing method
for (M, N, K) in ((64, 16384, 4096), (128, 16384, 4096), (64, 11520, 2880), (256, 16384, 4096), (512, 16384, 4096),
                  (256, 11520, 2880), (512, 11520, 2880)):
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16)
    for _ in range(3):
        torch.nn.functional.linear(x, w)
    ts = []
    for _ in range(12):
        flush.zero_(); torch.cuda.synchronize(); torch.cuda._sleep(200000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.nn.functional.linear(x, w); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print(json.dumps({"cublas": True, "M": M, "N": N, "K": K, "event_us": ts[len(ts) // 2]}), flush=True)

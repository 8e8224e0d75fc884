"""Per-layer timing of the planes conv kernels at VGG-11 shapes (CUDA events, not under a profiler).
python tools/planes_bench.py [batch]  -> one line per layer and pass: ms, algorithmic TF/s, fraction of the 3-pass MMA floor."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clsurvey_b200 import _capi
from clsurvey_b200._capi import call

_capi.lib()
dev = "cuda"
S = lambda: torch.cuda.current_stream().cuda_stream
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200
ONCE = "once" in sys.argv            # one call per kernel (for ncu)
LAYERS = [(32, 64, 128), (16, 128, 256), (16, 256, 256), (8, 256, 512), (8, 512, 512), (4, 512, 512), (4, 512, 512)]


def planes(*shape):
    x = torch.randn(*shape, device=dev)
    hi = x.bfloat16()
    lo = (x - hi.float()).bfloat16()
    return hi.view(torch.int16), lo.view(torch.int16)


def timeit(fn, reps=20):
    if ONCE:
        fn()
        torch.cuda.synchronize()
        return 1.0
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


tot = {"fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0}
flops_tot = 0.0
for H, C, K in LAYERS:
    W = H
    x = planes(N, H, W, C)
    dy = planes(N, H, W, K)
    y = planes(N, H, W, K)
    dx = planes(N, H, W, C)
    w = torch.randn(K, C, 3, 3, device=dev) * 0.05
    b = torch.zeros(K, device=dev)
    wf = planes(K, 9, C)
    wt = planes(C, 9, K)
    call("clb_planes_weights", w.data_ptr(), wf[0].data_ptr(), wf[1].data_ptr(), wt[0].data_ptr(), wt[1].data_ptr(), K, C, S())
    ws_bytes = _capi.lib().clb_planes_conv_wgrad_ws(N, H, W, C, K)
    ws = torch.zeros(ws_bytes // 4 + 4, device=dev)
    dw = torch.zeros(K, C, 3, 3, device=dev)
    db = torch.zeros(K, device=dev)
    fl = 2.0 * N * H * W * C * K * 9
    flops_tot += fl
    floor_ms = 3 * fl / (2 * 4096 * 148 * 1.965e9) * 1e3       # 3 MMA passes at 4096 MAC/clk/SM, 148 SMs, 1965 MHz
    t = {}
    t["fwd"] = timeit(lambda: call("clb_planes_conv_fwd", x[0].data_ptr(), x[1].data_ptr(), wf[0].data_ptr(), wf[1].data_ptr(), b.data_ptr(),
                                   y[0].data_ptr(), y[1].data_ptr(), N, H, W, C, K, 1, S()))
    t["dgrad"] = timeit(lambda: call("clb_planes_conv_dgrad", dy[0].data_ptr(), dy[1].data_ptr(), wt[0].data_ptr(), wt[1].data_ptr(), x[0].data_ptr(),
                                     dx[0].data_ptr(), dx[1].data_ptr(), N, H, W, C, K, S()))
    t["wgrad"] = timeit(lambda: call("clb_planes_conv_wgrad", x[0].data_ptr(), x[1].data_ptr(), dy[0].data_ptr(), dy[1].data_ptr(), dw.data_ptr(),
                                     db.data_ptr(), ws.data_ptr(), ws_bytes, N, H, W, C, K, 0, 0, 0.0, 0.0, S()))
    t["wprep"] = timeit(lambda: call("clb_planes_weights", w.data_ptr(), wf[0].data_ptr(), wf[1].data_ptr(), wt[0].data_ptr(), wt[1].data_ptr(), K, C, S()))
    for k in ("fwd", "dgrad", "wgrad"):
        tot[k] += t[k]
    print("N%d %2dx%-2d %3d->%3d  " % (N, H, W, C, K) + "  ".join("%s %.3f ms %5.0f TF/s %4.0f%%" % (k, t[k], fl / t[k] / 1e9, 100 * floor_ms / t[k])
                                                                for k in ("fwd", "dgrad", "wgrad")) + "  wprep %.3f ms" % t["wprep"], flush=True)
print("total fwd %.3f dgrad %.3f wgrad %.3f ms = %.3f ms; %.1f GF x3 passes -> %.0f TF/s algorithmic" % (
    tot["fwd"], tot["dgrad"], tot["wgrad"], sum(tot.values()), flops_tot / 1e9, 3 * flops_tot / sum(tot.values()) / 1e9))

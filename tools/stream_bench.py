"""Achieved HBM bandwidth of every streaming kernel of the hot path (algorithmic bytes of SURVEY.md 8d / CUDA-event time),
on an AlexNet-sized flat buffer (P = 57,085,780; GEM: P = 57,823,240, 10 tasks) -- larger than the 126 MB L2.
Writes one JSON object per kernel; `frac` is against the measured copy peak of MEASURED_PEAKS.json."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clsurvey_b200 import _capi

_capi.lib()
S = lambda: torch.cuda.current_stream().cuda_stream
peak = 6650.0
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p))["hbm_gbs"]


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def report(name, bytes_per_param, P, ms, extra=None):
    gbs = bytes_per_param * P / (ms * 1e-3) / 1e9
    d = {"kernel": name, "bytes_per_param": bytes_per_param, "params": P, "ms": ms, "achieved_gbs": gbs, "peak_gbs": peak,
         "frac": gbs / peak}
    if extra:
        d.update(extra)
    print(json.dumps(d), flush=True)


P = 57085780
g = torch.Generator(device="cuda").manual_seed(1)
th, gr, om, ts, bf, w = [torch.randn(P, device="cuda", generator=g) for _ in range(6)]
om.abs_()
n_pen = P - 81940            # all but the 20-way head
report("sgd_penalty_kernel (EWC/MAS step)", 28, P, timeit(lambda: _capi.call(
    "clb_sgd_penalty_step", th.data_ptr(), gr.data_ptr(), om.data_ptr(), ts.data_ptr(), bf.data_ptr(), P, n_pen, 6.0, 1e-3,
    0.9, 0.0, 1.0, 0, S())))
report("sgd_penalty_kernel (plain SGD-momentum)", 20, P, timeit(lambda: _capi.call(
    "clb_sgd_penalty_step", th.data_ptr(), gr.data_ptr(), 0, 0, bf.data_ptr(), P, 0, 0.0, 1e-3, 0.9, 0.0, 1.0, 0, S())))
report("si_step_kernel", 36, P, timeit(lambda: _capi.call(
    "clb_si_step", th.data_ptr(), gr.data_ptr(), om.data_ptr(), ts.data_ptr(), bf.data_ptr(), w.data_ptr(), P, 6.0, 1e-3, 0.9,
    0.0, 1.0, 0, S())))
report("fisher_kernel", 12, P, timeit(lambda: _capi.call("clb_fisher_accum", om.data_ptr(), gr.data_ptr(), 8000.0, P, S())))
report("mas_kernel", 12, P, timeit(lambda: _capi.call("clb_mas_accum", om.data_ptr(), gr.data_ptr(), 600.0, 800.0, P, S())))
report("si_consolidate_kernel", 28, P, timeit(lambda: _capi.call(
    "clb_si_consolidate", om.data_ptr(), w.data_ptr(), th.data_ptr(), ts.data_ptr(), 1e-3, P, S())))
del th, om, ts, bf, w
Pg, T = 57823240, 10
G = torch.randn(T, Pg, device="cuda", generator=g)
cur = torch.randn(Pg, device="cuda", generator=g)
dots = torch.zeros(16, dtype=torch.float64, device="cuda")
gram = torch.zeros(256, dtype=torch.float64, device="cuda")
v = torch.full((16,), 0.5, dtype=torch.float64, device="cuda")
viol = torch.ones(1, dtype=torch.int32, device="cuda")
for k in (1, 3, 9):
    idx = torch.arange(k, dtype=torch.int32, device="cuda")
    ms = timeit(lambda: _capi.call("clb_gem_dots_gram", cur.data_ptr(), G.data_ptr(), Pg, Pg, idx.data_ptr(), k, dots.data_ptr(),
                                   gram.data_ptr(), S()), reps=10)
    report("gem_dots_gram_kernel<k=%d>" % k, 4 * (k + 1), Pg, ms, {"k": k})
    ms = timeit(lambda: _capi.call("clb_gem_project", cur.data_ptr(), G.data_ptr(), Pg, Pg, idx.data_ptr(), k, v.data_ptr(),
                                   viol.data_ptr(), S()), reps=10)
    report("gem_project_kernel<k=%d>" % k, 4 * (k + 2), Pg, ms, {"k": k})
k = 9
idx = torch.arange(k, dtype=torch.int32, device="cuda")
dots.zero_(); gram.zero_()
_capi.call("clb_gem_dots_gram", cur.data_ptr(), G.data_ptr(), Pg, Pg, idx.data_ptr(), k, dots.data_ptr(), gram.data_ptr(), S())
ms = timeit(lambda: _capi.call("clb_gem_solve_qp", dots.data_ptr(), gram.data_ptr(), k, 1.0, 1e-3, v.data_ptr(), viol.data_ptr(), S()))
print(json.dumps({"kernel": "gem_qp_kernel<k=9> (512 active sets, one CTA)", "ms": ms}), flush=True)

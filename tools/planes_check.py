"""GPU bring-up check of the planes pipeline (csrc/clb_planes_*.cu): every C-ABI entry point against torch CPU fp64 on small
and full-size shapes.  `python tools/planes_check.py [quick]` prints one line per case and PLANES_CHECK_OK at the end."""
import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from clsurvey_b200 import _capi
from clsurvey_b200._capi import call

_capi.lib()
dev = "cuda"
S = lambda: torch.cuda.current_stream().cuda_stream
FAIL = []


def to_planes(x):
    """fp32 tensor (any shape) -> (hi, lo) int16-typed bf16 bit planes on the GPU."""
    hi = x.bfloat16()
    lo = (x - hi.float()).bfloat16()
    return hi.view(torch.int16).contiguous().to(dev), lo.view(torch.int16).contiguous().to(dev)


def from_planes(hi, lo):
    return hi.view(torch.bfloat16).float().cpu() + lo.view(torch.bfloat16).float().cpu()


def quant(x):
    hi = x.bfloat16().float()
    return hi + (x - hi).bfloat16().float()


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def report(name, err, tol):
    ok = err <= tol
    print("%-58s err %.3e  (tol %.0e) %s" % (name, err, tol, "ok" if ok else "FAIL"), flush=True)
    if not ok:
        FAIL.append(name)


def empty_planes(*shape):
    return torch.zeros(shape, dtype=torch.int16, device=dev), torch.zeros(shape, dtype=torch.int16, device=dev)


def weight_planes(w):
    K, C = w.shape[:2]
    wf = empty_planes(K, 9, C)
    wt = empty_planes(C, 9, K)
    call("clb_planes_weights", w.to(dev).data_ptr(), wf[0].data_ptr(), wf[1].data_ptr(), wt[0].data_ptr(), wt[1].data_ptr(), K, C, S())
    return wf, wt


def check_weights():
    g = torch.Generator().manual_seed(1)
    w = torch.randn(128, 64, 3, 3, generator=g)
    wf, wt = weight_planes(w)
    ref_f = quant(w).permute(0, 2, 3, 1).reshape(128, 9, 64)
    ref_t = quant(w).flip(2, 3).permute(1, 2, 3, 0).reshape(64, 9, 128)
    report("weights planes fwd layout", rel(from_planes(*wf), ref_f), 0)
    report("weights planes dgrad layout", rel(from_planes(*wt), ref_t), 0)


def check_conv(N, H, C, K, seed, relu=1):
    g = torch.Generator().manual_seed(seed)
    W = H
    x = torch.randn(N, C, H, W, generator=g)
    x = quant(torch.relu(x) if seed % 2 else x)
    w = torch.randn(K, C, 3, 3, generator=g) * (2.0 / (9 * C)) ** 0.5
    b = torch.randn(K, generator=g) * 0.1
    tag = "N%d %dx%d %d->%d" % (N, H, W, C, K)
    xp = to_planes(x.permute(0, 2, 3, 1).contiguous())
    wf, wt = weight_planes(w)
    wq = quant(w)
    # ---- forward
    y = empty_planes(N, H, W, K)
    bd = b.to(dev)
    call("clb_planes_conv_fwd", xp[0].data_ptr(), xp[1].data_ptr(), wf[0].data_ptr(), wf[1].data_ptr(), bd.data_ptr(), y[0].data_ptr(),
         y[1].data_ptr(), N, H, W, C, K, relu, S())
    torch.cuda.synchronize()
    ref = F.conv2d(x.double(), wq.double(), b.double(), padding=1)
    if relu:
        ref = torch.relu(ref)
    report("conv fwd   " + tag, rel(from_planes(*y).permute(0, 3, 1, 2), ref), 2e-5)
    # ---- dgrad (with and without mask)
    dy = quant(torch.randn(N, K, H, W, generator=g))
    dyp = to_planes(dy.permute(0, 2, 3, 1).contiguous())
    dx = empty_planes(N, H, W, C)
    call("clb_planes_conv_dgrad", dyp[0].data_ptr(), dyp[1].data_ptr(), wt[0].data_ptr(), wt[1].data_ptr(), 0, dx[0].data_ptr(),
         dx[1].data_ptr(), N, H, W, C, K, S())
    torch.cuda.synchronize()
    ref_dx = F.conv_transpose2d(dy.double(), wq.double(), padding=1)
    report("conv dgrad " + tag, rel(from_planes(*dx).permute(0, 3, 1, 2), ref_dx), 2e-5)
    call("clb_planes_conv_dgrad", dyp[0].data_ptr(), dyp[1].data_ptr(), wt[0].data_ptr(), wt[1].data_ptr(), xp[0].data_ptr(),
         dx[0].data_ptr(), dx[1].data_ptr(), N, H, W, C, K, S())
    torch.cuda.synchronize()
    report("conv dgrad+mask " + tag, rel(from_planes(*dx).permute(0, 3, 1, 2), ref_dx * (x > 0)), 2e-5)
    # ---- wgrad (+ bias, + fused Fisher)
    ws_bytes = _capi.lib().clb_planes_conv_wgrad_ws(N, H, W, C, K)
    ws = torch.zeros(ws_bytes // 4 + 4, device=dev)
    dw = torch.zeros(K, C, 3, 3, device=dev)
    db = torch.zeros(K, device=dev)
    om = torch.full((K, C, 3, 3), 0.5, device=dev)
    call("clb_planes_conv_wgrad", xp[0].data_ptr(), xp[1].data_ptr(), dyp[0].data_ptr(), dyp[1].data_ptr(), dw.data_ptr(), db.data_ptr(),
         ws.data_ptr(), ws_bytes, N, H, W, C, K, 1, om.data_ptr(), 8000.0, 0.0, S())
    torch.cuda.synchronize()
    xd = x.double().requires_grad_(False)
    ref_dw = torch.nn.grad.conv2d_weight(xd, (K, C, 3, 3), dy.double(), padding=1)
    report("conv wgrad " + tag, rel(dw.cpu(), ref_dw), 2e-5)
    report("conv dbias " + tag, rel(db.cpu(), dy.double().sum((0, 2, 3))), 2e-5)
    report("fused fisher " + tag, rel(om.cpu(), 0.5 + dw.cpu() ** 2 / 8000.0), 1e-6)


def check_pools():
    g = torch.Generator().manual_seed(3)
    N, C, H, W = 3, 128, 8, 8
    x = quant(torch.relu(torch.randn(N, C, H, W, generator=g)))
    x[0, :, 0:2, 0:2] = 0.0                                           # all-zero windows: gradient must be masked
    xp = to_planes(x.permute(0, 2, 3, 1).contiguous())
    y = empty_planes(N, H // 2, W // 2, C)
    am = torch.zeros(N, H // 2, W // 2, C, dtype=torch.uint8, device=dev)
    call("clb_planes_pool_fwd", xp[0].data_ptr(), xp[1].data_ptr(), y[0].data_ptr(), y[1].data_ptr(), 0, am.data_ptr(), N, H, W, C, S())
    ref, idx = F.max_pool2d(x, 2, 2, return_indices=True)
    report("pool fwd planes", rel(from_planes(*y).permute(0, 3, 1, 2), ref), 0)
    yf = torch.zeros(N, C, H // 2, W // 2, device=dev)
    call("clb_planes_pool_fwd", xp[0].data_ptr(), xp[1].data_ptr(), 0, 0, yf.data_ptr(), am.data_ptr(), N, H, W, C, S())
    report("pool fwd planes -> fp32 NCHW", rel(yf.cpu(), ref), 0)
    # backward: reference = maxpool backward followed by the ReLU mask of x
    dy = quant(torch.randn(N, C, H // 2, W // 2, generator=g))
    xr = x.clone().requires_grad_(True)
    F.max_pool2d(xr, 2, 2).backward(dy)
    ref_dx = xr.grad * (x > 0)
    dyp = to_planes(dy.permute(0, 2, 3, 1).contiguous())
    dx = empty_planes(N, H, W, C)
    call("clb_planes_pool_bwd", dyp[0].data_ptr(), dyp[1].data_ptr(), 0, y[0].data_ptr(), 0, am.data_ptr(), dx[0].data_ptr(),
         dx[1].data_ptr(), N, H, W, C, S())
    report("pool bwd planes", rel(from_planes(*dx).permute(0, 3, 1, 2), ref_dx), 0)
    dyd = dy.to(dev)
    call("clb_planes_pool_bwd", 0, 0, dyd.data_ptr(), 0, yf.data_ptr(), am.data_ptr(), dx[0].data_ptr(), dx[1].data_ptr(), N, H, W, C, S())
    report("pool bwd fp32 NCHW -> planes", rel(from_planes(*dx).permute(0, 3, 1, 2), ref_dx), 0)
    # NCHW fp32 ends (first layer): 64 channels, 16x16
    N, C, H, W = 2, 64, 16, 16
    x = torch.relu(torch.randn(N, C, H, W, generator=g))
    xd = x.to(dev)
    y = empty_planes(N, H // 2, W // 2, C)
    am = torch.zeros(N, H // 2, W // 2, C, dtype=torch.uint8, device=dev)
    call("clb_planes_pool_fwd_nchw", xd.data_ptr(), y[0].data_ptr(), y[1].data_ptr(), am.data_ptr(), N, C, H, W, S())
    ref = F.max_pool2d(x, 2, 2)
    report("pool fwd fp32 NCHW -> planes", rel(from_planes(*y).permute(0, 3, 1, 2), quant(ref)), 0)
    dy = quant(torch.randn(N, C, H // 2, W // 2, generator=g))
    dyp = to_planes(dy.permute(0, 2, 3, 1).contiguous())
    xr = x.clone().requires_grad_(True)
    F.max_pool2d(xr, 2, 2).backward(dy)
    ref_dx = xr.grad * (x > 0)
    dxf = torch.zeros(N, C, H, W, device=dev)
    call("clb_planes_pool_bwd_nchw", dyp[0].data_ptr(), dyp[1].data_ptr(), y[0].data_ptr(), am.data_ptr(), dxf.data_ptr(), N, C, H, W, S())
    report("pool bwd planes -> fp32 NCHW", rel(dxf.cpu(), ref_dx), 0)


def check_first():
    g = torch.Generator().manual_seed(11)
    for N, H in ((3, 16), (5, 64)):
        W, K = H, 64
        x = torch.randn(N, 3, H, W, generator=g)
        w = torch.randn(K, 3, 3, 3, generator=g) * 0.2
        b = torch.randn(K, generator=g) * 0.1
        xd, wd, bd = x.to(dev), w.to(dev), b.to(dev)
        y = empty_planes(N, H // 2, W // 2, K)
        am = torch.zeros(N, H // 2, W // 2, K, dtype=torch.uint8, device=dev)
        call("clb_planes_conv1_pool_fwd", xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), y[0].data_ptr(), y[1].data_ptr(), am.data_ptr(),
             N, 3, H, W, K, S())
        torch.cuda.synchronize()
        xr = x.double().requires_grad_(False)
        wr = w.double().requires_grad_(True)
        br = b.double().requires_grad_(True)
        act = torch.relu(F.conv2d(xr, wr, br, padding=1))
        pooled, idx = F.max_pool2d(act, 2, 2, return_indices=True)
        got = from_planes(*y).permute(0, 3, 1, 2)
        report("conv1+relu+pool fwd N%d %dx%d" % (N, H, W), rel(got, pooled.detach()), 2e-5)
        # arg-max agreement (window-local index) wherever the pooled value is > 0 and not a near-tie
        hh, ww = idx // W, idx % W
        loc = ((hh % 2) * 2 + (ww % 2)).permute(0, 2, 3, 1)
        agree = ((am.cpu().long() == loc) | (pooled.permute(0, 2, 3, 1) <= 0)).float().mean().item()
        report("conv1+relu+pool argmax agreement (1 - frac)", 1.0 - agree, 1e-3)
        dp = quant(torch.randn(N, K, H // 2, W // 2, generator=g))
        pooled.backward(dp.double())
        dpp = to_planes(dp.permute(0, 2, 3, 1).contiguous())
        ws_bytes = _capi.lib().clb_planes_conv1_ws()
        ws = torch.zeros(ws_bytes // 4 + 4, device=dev)
        dw = torch.zeros(K, 3, 3, 3, device=dev)
        db = torch.zeros(K, device=dev)
        call("clb_planes_conv1_pool_bwd", xd.data_ptr(), dpp[0].data_ptr(), dpp[1].data_ptr(), y[0].data_ptr(), am.data_ptr(), dw.data_ptr(),
             db.data_ptr(), ws.data_ptr(), ws_bytes, N, 3, H, W, K, S())
        torch.cuda.synchronize()
        report("conv1 fused bwd dW N%d %dx%d" % (N, H, W), rel(dw.cpu(), wr.grad), 2e-4)
        report("conv1 fused bwd db N%d %dx%d" % (N, H, W), rel(db.cpu(), br.grad), 2e-4)


def main():
    quick = "quick" in sys.argv
    cases = [(3, 32, 64, 128, 1), (5, 16, 128, 256, 2), (9, 8, 256, 512, 3), (25, 4, 512, 512, 4), (2, 16, 64, 64, 5)]
    if not quick:
        cases += [(200, 32, 64, 128, 6), (200, 4, 512, 512, 7), (200, 8, 512, 512, 8), (25, 16, 256, 256, 9)]
    for fn, args in [(check_weights, ()), (check_pools, ()), (check_first, ())] + [(check_conv, c) for c in cases]:
        try:
            fn(*args)
        except Exception:
            traceback.print_exc()
            FAIL.append("%s%r raised" % (fn.__name__, args))
            try:
                torch.cuda.synchronize()
            except Exception:
                print("CUDA context is dead, stopping", flush=True)
                break
    print("PLANES_CHECK_OK" if not FAIL else "PLANES_CHECK_FAILED: %s" % FAIL, flush=True)


if __name__ == "__main__":
    main()

"""Diagnostic for the tcgen05 conv path: runs fwd / dgrad / wgrad in TF32x1 and TF32x3 against torch CPU and prints
error structure (not a test; used while bringing the kernel up on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from clsurvey_b200 import _capi

S = lambda: torch.cuda.current_stream().cuda_stream


def rel(a, b):
    return ((a.cpu().double() - b.double()).abs().max() / b.double().abs().max()).item()


def run(shape, mode, structured=False):
    N, C, H, W, K = shape
    g = torch.Generator().manual_seed(1)
    if structured:
        x = torch.zeros(N, C, H, W); x[:, 0, :, :] = 1.0
        w = torch.zeros(K, C, 3, 3); w[:, 0, 1, 1] = torch.arange(K).float() + 1
    else:
        x = torch.randn(N, C, H, W, generator=g)
        w = torch.randn(K, C, 3, 3, generator=g) / (9 * C) ** 0.5
    b = torch.randn(K, generator=g)
    xr, wr = x.clone().requires_grad_(), w.clone().requires_grad_()
    y_ref = F.conv2d(xr, wr, b, padding=1)
    dy = torch.randn(y_ref.shape, generator=g)
    y_ref.backward(dy)
    _capi.call("clb_set_matmul_mode", mode)
    xd, wd, bd, dyd = x.cuda(), w.cuda(), b.cuda(), dy.cuda()
    y = torch.zeros_like(y_ref).cuda()
    wws = torch.empty(2 * max(w.numel(), K * 32, C * 32) + 8, device="cuda")
    _capi.call("clb_conv2d_fwd", xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), y.data_ptr(), wws.data_ptr(), N, C, H, W, K, 3, 3, 1, 1, 0, S())
    torch.cuda.synchronize()
    e = rel(y, y_ref.detach())
    print("shape", shape, "mode", mode, "structured", structured, "fwd rel err %.3e" % e, flush=True)
    if e > 1e-2:
        d = (y.cpu() - y_ref.detach())
        print("  y[0,0,0,:8]", y[0, 0, 0, :8].cpu().tolist())
        print("  ref        ", y_ref[0, 0, 0, :8].tolist())
        print("  y[0,:8,0,0]", y[0, :8, 0, 0].cpu().tolist())
        print("  ref        ", y_ref[0, :8, 0, 0].tolist())
        bad = (d.abs() > 1e-2 * y_ref.abs().max()).float()
        print("  bad frac %.3f; per-channel bad (first 16):" % bad.mean().item(), bad.mean(dim=(0, 2, 3))[:16].tolist())
        print("  per-row bad (first img, first 16 pixels rows):", bad[0].mean(dim=0).flatten()[:16].tolist())
    ws_bytes = _capi.lib().clb_conv2d_wgrad_ws(N, C, H, W, K, 3, 3, 1, 1)
    ws = torch.empty(ws_bytes // 4 + 4, device="cuda")
    dw, db = torch.zeros_like(w).cuda(), torch.zeros(K).cuda()
    _capi.call("clb_conv2d_wgrad", xd.data_ptr(), dyd.data_ptr(), dw.data_ptr(), db.data_ptr(), ws.data_ptr(), ws.numel() * 4, N, C, H, W, K, 3, 3, 1, 1, S())
    torch.cuda.synchronize()
    print("   wgrad rel err %.3e" % rel(dw, wr.grad), flush=True)
    dx = torch.zeros_like(x).cuda()
    _capi.call("clb_conv2d_dgrad", dyd.data_ptr(), wd.data_ptr(), dx.data_ptr(), wws.data_ptr(), N, C, H, W, K, 3, 3, 1, 1, S())
    torch.cuda.synchronize()
    print("   dgrad rel err %.3e" % rel(dx, xr.grad), flush=True)
    _capi.call("clb_set_matmul_mode", 3)


if __name__ == "__main__":
    _capi.lib()
    if len(sys.argv) > 1 and sys.argv[1] == "tma":
        run((2, 64, 8, 8, 128), 1)
        print("tc_debug tma done")
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "first":
        run((2, 3, 16, 16, 8), 1)
        print("tc_debug first done")
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "bf16":          # mode 3: bf16 hi/lo split for fwd / dgrad (C % 64 == 0)
        run((2, 64, 8, 8, 128), 3, structured=True)
        for shp in [(2, 64, 8, 8, 128), (9, 64, 32, 32, 128), (7, 256, 8, 8, 512), (3, 128, 16, 16, 64), (25, 512, 4, 4, 512)]:
            run(shp, 3)
            run(shp, 1)
        print("tc_debug bf16 done")
        sys.exit(0)
    run((2, 32, 8, 8, 128), 2, structured=True)
    run((2, 32, 8, 8, 128), 2)
    run((2, 32, 8, 8, 128), 1)
    run((9, 64, 32, 32, 128), 1)
    run((7, 256, 8, 8, 512), 1)
    run((3, 128, 16, 16, 64), 1)
    run((25, 512, 4, 4, 512), 1)
    run((5, 3, 64, 64, 64), 1)          # first layer: small-C forward (one padded K block) + generic wgrad
    run((3, 3, 16, 16, 8), 1)
    print("tc_debug done")

"""GPU: every C-ABI kernel against the oracle (oracle/restate.py formulas, torch-CPU fp32 layer arithmetic).

Tolerances (normalised max error, tests/util.rel_err):
  streaming kernels (same op order as the reference, fp32): 1e-6
  layer kernels, fp32 SIMT path: 2e-5 (summation order differs from MKL-DNN)
  layer kernels, tcgen05 TF32x3 path: 1e-4 (north_star tolerance)
  integer results (arg-max, #correct, violation mask): exact
"""
import ctypes

import numpy as np
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import qp as oqp
from oracle import restate
from tests.util import rel_err

pytestmark = pytest.mark.gpu
DEFAULT_MODE = int(os.environ.get("CLB_MM_MODE", "3"))      # the library default; tests that switch modes restore it


@pytest.fixture(scope="module")
def capi():
    from clsurvey_b200 import _capi
    _capi.lib()
    return _capi


def dev(t):
    return t.cuda().contiguous()


def S():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("n,n_pen", [(1, 1), (7, 5), (4096, 4096), (100003, 99001), (1 << 20, 0)])
@pytest.mark.parametrize("first,wd", [(1, 0.0), (0, 5e-4)])
def test_sgd_penalty_step(capi, n, n_pen, first, wd):
    g = torch.Generator().manual_seed(n + first)
    th, gr, om, ts, bf = [torch.randn(n, generator=g) for _ in range(5)]
    om = om.abs()
    lam, lr = 3.0, 0.01
    ref_t, ref_b = th.clone(), bf.clone()
    t1, b1 = restate.penalised_sgd_step(th[:n_pen], gr[:n_pen], om[:n_pen], ts[:n_pen], None if first else bf[:n_pen], lam, lr, wd)
    t2, b2 = restate.penalised_sgd_step(th[n_pen:], gr[n_pen:], None, None, None if first else bf[n_pen:], 0.0, lr, wd)
    ref_t, ref_b = torch.cat([t1, t2]), torch.cat([b1, b2])
    dth, dg, dom, dts, dbf = map(dev, (th, gr, om, ts, bf))
    capi.call("clb_sgd_penalty_step", dth.data_ptr(), dg.data_ptr(), dom.data_ptr(), dts.data_ptr(), dbf.data_ptr(), n,
              n_pen, 2 * lam, lr, 0.9, wd, 1.0, first, S())
    assert rel_err(dth, ref_t) <= 1e-6 and rel_err(dbf, ref_b) <= 1e-6


@pytest.mark.parametrize("n", [3, 1000, 262147])
@pytest.mark.parametrize("first", [1, 0])
def test_si_step_and_consolidate(capi, n, first):
    g = torch.Generator().manual_seed(n)
    th, gr, om, ts, bf, w = [torch.randn(n, generator=g) for _ in range(6)]
    om = om.abs()
    t, b, wn = restate.si_step(th, gr, om, ts, None if first else bf, w, 2.0, 0.01, 1e-4)
    d = list(map(dev, (th, gr, om, ts, bf, w)))
    capi.call("clb_si_step", *[x.data_ptr() for x in d], n, 4.0, 0.01, 0.9, 1e-4, 1.0, first, S())
    assert rel_err(d[0], t) <= 1e-6 and rel_err(d[4], b) <= 1e-6 and rel_err(d[5], wn) <= 1e-6
    o2, w2, ts2 = restate.si_consolidate(om, w, th, ts)
    dom, dw, dth, dts = map(dev, (om, w, th, ts))
    capi.call("clb_si_consolidate", dom.data_ptr(), dw.data_ptr(), dth.data_ptr(), dts.data_ptr(), 1e-3, n, S())
    assert rel_err(dom, o2) <= 1e-6 and float(dw.abs().max()) == 0.0 and torch.equal(dts.cpu(), th)


@pytest.mark.parametrize("n", [5, 4096, 1000003])
def test_fisher_mas_axpby(capi, n):
    g = torch.Generator().manual_seed(n)
    om, gr = torch.randn(n, generator=g).abs(), torch.randn(n, generator=g)
    dom, dg = dev(om), dev(gr)
    capi.call("clb_fisher_accum", dom.data_ptr(), dg.data_ptr(), 8000.0, n, S())
    assert rel_err(dom, om + gr ** 2 / 8000) <= 1e-6
    dom = dev(om)
    capi.call("clb_mas_accum", dom.data_ptr(), dg.data_ptr(), 3 * 200.0, 4 * 200.0, n, S())
    assert rel_err(dom, (om * 600 + gr.abs()) / 800) <= 1e-6
    dom = dev(om)
    capi.call("clb_axpby", dom.data_ptr(), dom.data_ptr(), dg.data_ptr(), 1.0, n, S())
    assert rel_err(dom, om + gr) <= 1e-7


CONVS = [  # N, C, H, W, K, R, stride, pad
    (3, 3, 16, 16, 8, 3, 1, 1),        # golden tiny first layer (C=3)
    (5, 8, 8, 8, 16, 3, 1, 1),
    (2, 16, 4, 4, 16, 3, 1, 1),
    (4, 3, 64, 64, 64, 11, 4, 2),      # AlexNet conv1
    (3, 64, 7, 7, 192, 5, 1, 2),       # AlexNet conv2
    (2, 192, 3, 3, 384, 3, 1, 1),      # AlexNet conv3
    (9, 64, 32, 32, 128, 3, 1, 1),     # VGG-11 conv2 (ragged batch)
    (7, 256, 8, 8, 512, 3, 1, 1),      # VGG-11 conv5
    (200, 512, 4, 4, 512, 3, 1, 1),    # VGG-11 conv7 at the benchmark batch
]


@pytest.mark.parametrize("shape", CONVS)
@pytest.mark.parametrize("mode", [0, 1, 3])
def test_conv2d_fwd_bwd(capi, shape, mode):
    N, C, H, W, K, R, stride, pad = shape
    capi.call("clb_set_matmul_mode", mode)
    tol = 2e-5 if mode == 0 else 1e-4
    try:
        g = torch.Generator().manual_seed(sum(shape))
        x = torch.randn(N, C, H, W, generator=g)
        w = torch.randn(K, C, R, R, generator=g) / (C * R * R) ** 0.5
        b = torch.randn(K, generator=g)
        xr, wr, br = x.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
        y_ref = F.relu(F.conv2d(xr, wr, br, stride=stride, padding=pad))
        dy = torch.randn(y_ref.shape, generator=g) * (y_ref > 0)
        y_ref.backward(dy)
        P, Q = y_ref.shape[2:]
        dx_, dw_, db_ = dev(torch.zeros_like(x)), dev(torch.zeros_like(w)), dev(torch.zeros_like(b))
        y_ = dev(torch.zeros_like(y_ref))
        dxc, dwc, dyc = dev(x), dev(w), dev(dy)
        dbc = dev(b)
        wws = torch.empty(2 * max(w.numel(), K * 32, C * 32) + 8, device="cuda")
        capi.call("clb_conv2d_fwd", dxc.data_ptr(), dwc.data_ptr(), dbc.data_ptr(), y_.data_ptr(), wws.data_ptr(), N, C, H,
                  W, K, R, R, stride, pad, 1, S())
        assert rel_err(y_, y_ref) <= tol
        ws_bytes = capi.lib().clb_conv2d_wgrad_ws(N, C, H, W, K, R, R, stride, pad)
        ws = torch.empty(ws_bytes // 4 + 4, device="cuda")
        capi.call("clb_conv2d_wgrad", dxc.data_ptr(), dyc.data_ptr(), dw_.data_ptr(), db_.data_ptr(), ws.data_ptr(),
                  ws.numel() * 4, N, C, H, W, K, R, R, stride, pad, S())
        assert rel_err(dw_, wr.grad) <= tol, "wgrad"
        assert rel_err(db_, br.grad) <= tol, "bias grad"
        wt = torch.empty(2 * max(w.numel(), K * 32, C * 32) + 8, device="cuda")
        capi.call("clb_conv2d_dgrad", dyc.data_ptr(), dwc.data_ptr(), dx_.data_ptr(), wt.data_ptr(), N, C, H, W, K, R, R,
                  stride, pad, S())
        assert rel_err(dx_, xr.grad) <= tol, "dgrad"
    finally:
        capi.call("clb_set_matmul_mode", DEFAULT_MODE)


@pytest.mark.parametrize("M,inf,outf", [(16, 64, 32), (200, 2048, 512), (37, 512, 20), (5, 9216, 4096), (200, 4096, 20)])
@pytest.mark.parametrize("mode", [0, 1, 3])
def test_linear_fwd_bwd(capi, M, inf, outf, mode):
    capi.call("clb_set_matmul_mode", mode)
    tol = 2e-5 if mode == 0 else 1e-4
    try:
        g = torch.Generator().manual_seed(M + inf + outf)
        x = torch.randn(M, inf, generator=g)
        w = torch.randn(outf, inf, generator=g) / inf ** 0.5
        b = torch.randn(outf, generator=g)
        xr, wr, br = x.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
        y_ref = F.relu(F.linear(xr, wr, br))
        dy = torch.randn(y_ref.shape, generator=g) * (y_ref > 0)
        y_ref.backward(dy)
        dx, dw, dyc, db = dev(x), dev(w), dev(dy), dev(b)
        y_ = torch.zeros(M, outf, device="cuda")
        gx, gw, gb = torch.zeros(M, inf, device="cuda"), torch.zeros(outf, inf, device="cuda"), torch.zeros(outf, device="cuda")
        ws = torch.empty(capi.lib().clb_linear_ws(M, inf, outf) // 4 + 4, device="cuda")
        wsb = ws.numel() * 4
        capi.call("clb_linear_fwd", dx.data_ptr(), dw.data_ptr(), db.data_ptr(), y_.data_ptr(), ws.data_ptr(), wsb, M, inf,
                  outf, 1, S())
        assert rel_err(y_, y_ref) <= tol
        capi.call("clb_linear_wgrad", dx.data_ptr(), dyc.data_ptr(), gw.data_ptr(), gb.data_ptr(), ws.data_ptr(), wsb, M, inf,
                  outf, S())
        capi.call("clb_linear_dgrad", dyc.data_ptr(), dw.data_ptr(), gx.data_ptr(), ws.data_ptr(), wsb, M, inf, outf, S())
        assert rel_err(gw, wr.grad) <= tol and rel_err(gb, br.grad) <= tol and rel_err(gx, xr.grad) <= tol
    finally:
        capi.call("clb_set_matmul_mode", DEFAULT_MODE)


@pytest.mark.parametrize("N,C,H,W,k,s", [(3, 8, 16, 16, 2, 2), (2, 64, 15, 15, 3, 2), (4, 5, 7, 7, 3, 2), (2, 3, 9, 9, 2, 2)])
def test_maxpool_relu(capi, N, C, H, W, k, s):
    g = torch.Generator().manual_seed(N * C * H)
    pre = torch.randn(N, C, H, W, generator=g)
    pre[0, 0, :4, :4] = 0.5                                   # ties inside windows: first max must win
    x = F.relu(pre).requires_grad_()
    y_ref = F.max_pool2d(x, k, s)
    dy = torch.randn(y_ref.shape, generator=g)
    y_ref.backward(dy)
    dx_relu = x.grad * (x > 0)
    dxd, y_ = dev(x.detach()), dev(torch.zeros_like(y_ref))
    am = torch.zeros(y_ref.numel(), dtype=torch.uint8, device="cuda")
    capi.call("clb_maxpool_fwd", dxd.data_ptr(), y_.data_ptr(), am.data_ptr(), N, C, H, W, k, s, S())
    assert torch.equal(y_.cpu(), y_ref.detach())
    gx = torch.zeros(N, C, H, W, device="cuda")
    dyd = dev(dy)
    capi.call("clb_maxpool_bwd", dyd.data_ptr(), am.data_ptr(), 0, gx.data_ptr(), N, C, H, W, k, s, S())
    assert rel_err(gx, x.grad) <= 1e-6
    capi.call("clb_maxpool_bwd", dyd.data_ptr(), am.data_ptr(), dxd.data_ptr(), gx.data_ptr(), N, C, H, W, k, s, S())
    assert rel_err(gx, dx_relu) <= 1e-6
    d2 = dev(dy.new_ones(x.shape))
    capi.call("clb_relu_bwd", d2.data_ptr(), dxd.data_ptr(), d2.data_ptr(), x.numel(), S())
    assert torch.equal(d2.cpu(), (x > 0).float())


def test_avgpool_and_mask(capi):
    x = torch.randn(3, 4, 1, 1).requires_grad_()
    y = F.adaptive_avg_pool2d(x, (6, 6))
    dy = torch.randn_like(y)
    y.backward(dy)
    y_, gx = torch.zeros(3, 4, 6, 6, device="cuda"), torch.zeros(3, 4, 1, 1, device="cuda")
    xd, dyd = dev(x.detach()), dev(dy)
    capi.call("clb_adaptive_avgpool_fwd", xd.data_ptr(), y_.data_ptr(), 3, 4, 1, 1, 6, 6, S())
    capi.call("clb_adaptive_avgpool_bwd", dyd.data_ptr(), gx.data_ptr(), 3, 4, 1, 1, 6, 6, S())
    assert rel_err(y_, y) <= 1e-7 and rel_err(gx, x.grad) <= 1e-6
    x2 = torch.randn(2, 3, 13, 13).requires_grad_()
    y2 = F.adaptive_avg_pool2d(x2, (6, 6))
    y2.backward(torch.ones_like(y2))
    y2_, gx2 = torch.zeros(2, 3, 6, 6, device="cuda"), torch.zeros(2, 3, 13, 13, device="cuda")
    x2d, o2d = dev(x2.detach()), dev(torch.ones_like(y2))
    capi.call("clb_adaptive_avgpool_fwd", x2d.data_ptr(), y2_.data_ptr(), 2, 3, 13, 13, 6, 6, S())
    capi.call("clb_adaptive_avgpool_bwd", o2d.data_ptr(), gx2.data_ptr(), 2, 3, 13, 13, 6, 6, S())
    assert rel_err(y2_, y2) <= 1e-6 and rel_err(gx2, x2.grad) <= 1e-6
    a, m = torch.randn(7, 33), (torch.rand(33) > 0.5).float() * 2
    out = torch.zeros(7, 33, device="cuda")
    ad, md = dev(a), dev(m)
    capi.call("clb_mask_mul", ad.data_ptr(), md.data_ptr(), out.data_ptr(), 7, 33, 1, S())
    assert torch.equal(out.cpu(), a * m)


@pytest.mark.parametrize("B,ld,off,nc", [(16, 5, 0, 5), (200, 20, 0, 20), (33, 200, 40, 20), (1, 15, 10, 5)])
def test_softmax_loss_modes(capi, B, ld, off, nc):
    g = torch.Generator().manual_seed(B + ld)
    z = (torch.randn(B, ld, generator=g) * 3).requires_grad_()
    y = torch.randint(0, nc, (B,), generator=g)
    zs = z[:, off:off + nc]
    for mode, loss_ref in ((0, restate.loss_mean_ce(zs, y)), (1, restate.loss_sum_nll(zs, y)), (2, restate.loss_sum_sq(zs))):
        z.grad = None
        loss_ref.backward(retain_graph=True)
        loss = torch.zeros(1, device="cuda")
        corr = torch.zeros(1, dtype=torch.int32, device="cuda")
        dz = torch.full((B, ld), 7.0, device="cuda")
        zd, yd = dev(z.detach()), dev(y)                      # keep references: the allocator recycles temporaries
        capi.call("clb_softmax_loss", zd.data_ptr(), ld, off, nc, yd.data_ptr(), B, mode, float(B),
                  loss.data_ptr(), corr.data_ptr(), dz.data_ptr(), S())
        assert abs(loss.item() - loss_ref.item()) <= 2e-6 * abs(loss_ref.item())
        assert rel_err(dz, z.grad) <= 2e-6
        assert corr.item() == restate.num_correct(zs, y)


@pytest.mark.parametrize("k", [1, 2, 5, 9])
def test_gem_dots_gram_qp_project(capi, k):
    P, ld, n_tasks = 100003, 100004, 10
    g = torch.Generator().manual_seed(k)
    G = torch.randn(n_tasks, ld, generator=g)
    G[:, P:] = 0
    cur = torch.randn(ld, generator=g)
    cur[P:] = 0
    G[1] = -0.5 * cur + 0.5 * G[1]                              # force a violation with task 1
    prev = list(range(1, k + 1)) if k < 9 else list(range(0, 9))
    idx = torch.tensor(prev, dtype=torch.int32, device="cuda")
    dG, dcur = dev(G), dev(cur)
    dots = torch.zeros(16, dtype=torch.float64, device="cuda")
    gram = torch.zeros(256, dtype=torch.float64, device="cuda")
    v = torch.zeros(16, dtype=torch.float64, device="cuda")
    viol = torch.zeros(1, dtype=torch.int32, device="cuda")
    capi.call("clb_gem_dots_gram", dcur.data_ptr(), dG.data_ptr(), ld, P, idx.data_ptr(), k, dots.data_ptr(),
              gram.data_ptr(), S())
    M = G[prev].double()
    ref_dots, ref_gram = M @ cur.double(), M @ M.T
    # products are summed in short fp32 runs that are flushed into fp64 accumulators (clb_gem.cu): ~1e-10 of the exact fp64
    # value, five orders below the fp32 torch.mm of the reference (gem.py:157-158)
    assert rel_err(dots[:k], ref_dots) <= 1e-8 and rel_err(gram[:k * k].view(k, k), ref_gram) <= 1e-8
    capi.call("clb_gem_solve_qp", dots.data_ptr(), gram.data_ptr(), k, 0.5, 1e-3, v.data_ptr(), viol.data_ptr(), S())
    assert viol.item() == int((ref_dots < 0).sum())              # violation mask: exact
    x_ref, v_ref = oqp.project2cone2(cur.numpy(), G[prev].numpy(), 0.5)
    assert np.abs(v[:k].cpu().numpy() - v_ref).max() <= 1e-7 * max(1.0, np.abs(v_ref).max())
    capi.call("clb_gem_project", dcur.data_ptr(), dG.data_ptr(), ld, P, idx.data_ptr(), k, v.data_ptr(), viol.data_ptr(), S())
    assert rel_err(dcur[:P], torch.from_numpy(x_ref)[:P]) <= 1e-6
    # no violation -> v = 0 and g untouched
    dots2 = torch.ones(16, dtype=torch.float64, device="cuda")
    capi.call("clb_gem_solve_qp", dots2.data_ptr(), gram.data_ptr(), k, 0.5, 1e-3, v.data_ptr(), viol.data_ptr(), S())
    before = dcur.clone()
    capi.call("clb_gem_project", dcur.data_ptr(), dG.data_ptr(), ld, P, idx.data_ptr(), k, v.data_ptr(), viol.data_ptr(), S())
    assert viol.item() == 0 and torch.equal(before, dcur)

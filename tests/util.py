"""Shared helpers for the tests (model builders for the golden fixtures, comparison metrics)."""
import os

import torch

from clsurvey_b200.models import VGGSlim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
TINY_CFG = [8, "M", 16, "M", 16, 16, "M"]
NCLS, BS = 5, 16


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def tiny_model(state=None, dropout=False, num_classes=NCLS):
    m = VGGSlim(TINY_CFG, num_classes, 16 * 2 * 2, 32, 32, dropout=dropout)
    if state is not None:
        m.load_state_dict(state)
    return m


WIDE_CFG = [64, "M", 64, 128, "M"]       # tests/golden/wide.pt: 8x8 inputs, batch 2 (oracle/gen_golden.py gen_wide)
WIDE_HW, WIDE_BS = 8, 2


def wide_model(state=None, num_classes=NCLS):
    m = VGGSlim(WIDE_CFG, num_classes, 128 * 2 * 2, 32, 32)
    if state is not None:
        m.load_state_dict(state)
    return m


def loaders(data, bs=BS):
    xt, yt, xv, yv = data
    mk = lambda x, y: torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=bs, shuffle=False)
    return {"train": mk(xt, yt), "val": mk(xv, yv)}, {"train": len(xt), "val": len(xv)}


def batches(x, y, bs=BS):
    return [(x[i:i + bs], y[i:i + bs]) for i in range(0, len(x), bs)]


def rel_err(a, b):
    """max |a-b| / max |b|  (tensor-wise normalised max error; the tolerance metric used throughout)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    denom = b.abs().max().item()
    return (a - b).abs().max().item() / (denom if denom > 0 else 1.0)


def assert_close_state(sd_a, sd_b, tol, what=""):
    for k in sd_b:
        e = rel_err(sd_a[k], sd_b[k])
        assert e <= tol, "%s %s: rel err %.3e > %.1e" % (what, k, e, tol)

"""CPU (no GPU needed): the C-ABI library builds, loads and exports every symbol include/clb.h declares; host-side
logic (QP solver host twin, model definitions, sharding helpers); world_size-2 gloo test of the data-parallel scheme."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from tests.util import ROOT, load_golden


@pytest.fixture(scope="module")
def capi():
    import __graft_entry__ as ge
    ge.build()
    from clsurvey_b200 import _capi
    return _capi


def test_header_symbols_exported(capi):
    hdr = open(os.path.join(ROOT, "include", "clb.h")).read()
    declared = set(re.findall(r"\b(clb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    lib = ctypes.CDLL(capi.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), "libclb.so does not export %s" % name
        assert name in capi.SIGNATURES, "no ctypes signature for %s" % name
    assert capi.lib().clb_version() >= 100


def test_invalid_arguments_fail_loudly(capi):
    with pytest.raises(capi.ClbError):
        capi.call("clb_fisher_accum", 0, 0, 8000.0, 10, 0)
    assert b"invalid argument" in capi.lib().clb_last_error()
    with pytest.raises(capi.ClbError):
        capi.call("clb_set_matmul_mode", 99)


def test_engine_refuses_without_gpu(capi):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from clsurvey_b200.engine import Engine
    from clsurvey_b200.models import make_vgg
    with pytest.raises(capi.ClbError):
        Engine(make_vgg("small_VGG9_cl_128_128"), (3, 64, 64), 4)


def test_host_qp_matches_known_answers(capi):
    for c in load_golden("qp"):
        M, g = c["M"].numpy(), c["g"].numpy()
        k = M.shape[0]
        gram = np.ascontiguousarray(M @ M.T)
        dots = np.ascontiguousarray(M @ g)
        v = np.zeros(k)
        viol = ctypes.c_int(0)
        capi.call("clb_gem_solve_qp_host", dots.ctypes.data, gram.ctypes.data, k, float(c["margin"]), 1e-3,
                  v.ctypes.data, ctypes.addressof(viol))
        assert viol.value == int((dots < 0).sum())
        if viol.value:
            ref = c["v"].numpy()
            assert np.abs(v - ref).max() <= 1e-8 * max(1.0, np.abs(ref).max())
        else:
            assert (v == 0).all()


def test_model_definitions_match_reference_sizes():
    from clsurvey_b200.models import make_alexnet, make_vgg
    n = lambda m: sum(p.numel() for p in m.parameters())
    assert n(make_vgg("VGG11_cl_512_512")) == 10542484
    assert n(make_vgg("small_VGG9_cl_128_128")) == 615380
    assert n(make_vgg("wide_VGG9_cl_512_512")) == 8968596
    assert n(make_alexnet(20)) == 57085780
    m = make_vgg("small_VGG9_cl_128_128")
    assert m(torch.zeros(2, 3, 64, 64)).shape == (2, 20)
    assert list(m.state_dict())[:2] == ["features.0.weight", "features.0.bias"]


def test_shard_helpers():
    from clsurvey_b200 import dist
    for n in (200, 201, 7, 1):
        for w in (1, 2, 4, 8):
            spans = [dist.shard_rows(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert dist.shard_batches(40, 8, 3) == [3, 11, 19, 27, 35]
    assert sorted(sum((dist.shard_batches(40, 8, r) for r in range(8)), [])) == list(range(40))


def test_trainer_quirks_on_host():
    from clsurvey_b200.methods.trainers import set_lr
    opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=0.1)
    assert set_lr(opt, 0.1, 10)[2] is True and set_lr(opt, 0.1, 11)[2] is False        # others stop at count > 10
    assert set_lr(opt, 0.1, 10, stop_ge=True)[2] is False                               # SI stops at count >= 10
    _, lr, _ = set_lr(opt, 0.1, 5)
    assert abs(lr - 0.01) < 1e-12 and abs(opt.param_groups[0]["lr"] - 0.01) < 1e-12


DP_WORKER = r'''
import os, sys, torch
sys.path.insert(0, %r)
import torch.nn.functional as F
from clsurvey_b200 import dist
from tests.util import tiny_model, load_golden
dist.init("gloo")
assert dist.world_size() == 2
f = load_golden("finetune")["wd0"]
m = tiny_model(f["init"])
x, y = f["data"][0][:16], f["data"][1][:16]
lo, hi = dist.shard_rows(16)
# each rank: gradient of the GLOBAL-batch mean CE restricted to its rows (sum-CE over the shard / global B)
loss = F.cross_entropy(m(x[lo:hi]), y[lo:hi], reduction="sum") / 16
loss.backward()
flat = torch.cat([p.grad.reshape(-1) for p in m.parameters()])
dist.allreduce_flat(flat)
m2 = tiny_model(f["init"])
F.cross_entropy(m2(x), y).backward()
ref = torch.cat([p.grad.reshape(-1) for p in m2.parameters()])
err = (flat - ref).abs().max().item() / ref.abs().max().item()
assert err < 1e-5, err
# importance pass: whole batches dealt round robin, squared batch gradients summed over ranks
mine = dist.shard_batches(4)
acc = torch.zeros_like(ref)
for b in mine:
    m3 = tiny_model(f["init"]); m3.eval()
    xb, yb = f["data"][0][b*16:(b+1)*16], f["data"][1][b*16:(b+1)*16]
    F.cross_entropy(m3(xb), yb, reduction="sum").backward()
    acc += torch.cat([p.grad.reshape(-1) for p in m3.parameters()]) ** 2 / 64
dist.allreduce_flat(acc)
from oracle import restate
m4 = tiny_model(f["init"])
bl = [(f["data"][0][b*16:(b+1)*16], f["data"][1][b*16:(b+1)*16]) for b in range(4)]
om = torch.cat([o.reshape(-1) for o in restate.fisher_pass(m4, bl, 64)])
err2 = (acc - om).abs().max().item() / om.abs().max().item()
assert err2 < 1e-5, err2
# MAS omega pass with RAGGED batches (16, 16, 16, 8), sharded by whole batches: the reference's running mean
# (train_MAS.py:163-177: omega <- (omega*b*n_b + |g|)/((b+1)*n_b) with n_b the CURRENT batch size) unrolls to
# (1/n_batches) * sum_b |g_b|/n_b, which is what the sharded pass of methods/MAS/train_MAS.py accumulates
xs, ys = f["data"][0][:56], f["data"][1][:56]
rag = [(xs[i:i + 16], ys[i:i + 16]) for i in range(0, 56, 16)]
acc = torch.zeros_like(ref)
for b in dist.shard_batches(len(rag)):
    m5 = tiny_model(f["init"]); m5.eval()
    (m5(rag[b][0]) ** 2).sum().backward()
    g = torch.cat([p.grad.reshape(-1) for p in m5.parameters()])
    n_b = float(rag[b][0].shape[0])
    acc = (acc * n_b + g.abs()) / n_b
dist.allreduce_flat(acc)
acc = acc / len(rag)
om = torch.cat([o.reshape(-1) for o in restate.mas_pass(tiny_model(f["init"]), rag)])
err3 = (acc - om).abs().max().item() / om.abs().max().item()
assert err3 < 1e-5, err3
tot = dist.allreduce_scalars([1.0, float(dist.rank())])
assert tot == [2.0, 1.0]
# replicas: one seed for all ranks, rank 0's buffers everywhere
torch.manual_seed(100 + dist.rank())
s1 = dist.shared_seed()
t = torch.tensor([float(s1 %% 1000)])
dist.allreduce_flat(t)
assert t.item() == 2.0 * (s1 %% 1000)
buf = torch.full((8,), float(dist.rank() + 1))
dist.broadcast_flat(buf)
assert buf.eq(1.0).all()
dist.shutdown()
print("DP_OK", dist.rank() if False else os.environ["RANK"])
'''


def test_data_parallel_scheme_gloo_world2(tmp_path):
    script = tmp_path / "dp_worker.py"
    script.write_text(DP_WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29731", OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29731", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("DP_OK") == 2


def test_method_plugin_surface():
    """The Method classes expose what the reference framework reads (src/framework/main.py:109-111,
    framework_train.py:24,186, lr_grid_train.py:113) and parse()/set_hyperparams() behave like method.py:35-78,238-274."""
    from clsurvey_b200.methods import method as M
    for name, cls in (("EWC", M.EWC), ("MAS", M.MAS), ("SI", M.SI), ("GEM", M.GEM), ("finetuning", M.Finetune)):
        m = M.parse(name)
        assert isinstance(m, cls) and isinstance(m, M.Method)
        for attr in ("name", "eval_name", "category", "extra_hyperparams_count", "hyperparams"):
            assert hasattr(m, attr)
        assert callable(m.get_output) and callable(m.inference_eval) and callable(m.grid_train)
    assert M.EWC.hyperparams["lambda"] == 400 and M.MAS.hyperparams["lambda"] == 3 and M.SI.hyperparams["lambda"] == 400
    assert M.GEM.hyperparams["margin"] == 1 and M.GEM.static_hyperparams["mem_per_task"] == 1024
    assert M.GEM.wrap_first_task_model and M.Finetune.start_scratch and M.Finetune.no_framework
    assert M.EWC.category == M.Category.MODEL_BASED and M.GEM.category == M.Category.REHEARSAL_BASED
    with pytest.raises(NotImplementedError):
        M.parse("LWF")
    g = M.GEM()
    g.hyperparams = type(g.hyperparams)(g.hyperparams)
    g.static_hyperparams = type(g.static_hyperparams)(g.static_hyperparams)
    M.set_hyperparams(g, "0.5")
    M.set_hyperparams(g, "256", static_params=True)
    assert g.hyperparams["margin"] == 0.5 and g.static_hyperparams["mem_per_task"] == 256.0


def test_gem_ring_buffer_host_logic_matches_reference_fixture():
    """fill_buffer arithmetic (gem.py:322-345) is pure host logic: replay the fixture's key/label stream through the
    mirror's bookkeeping and compare mem_cnt, labels and exemplar keys bit for bit (no GPU needed)."""
    from clsurvey_b200.methods.rehearsal.model import common
    g = load_golden("gem")
    n_tasks, n_mem = g["n_tasks"], g["n_mem"]
    mem = common.RehearsalMemory(n_tasks, n_mem, (3, 16, 16))
    labels = torch.zeros(n_tasks, n_mem, dtype=torch.long)
    mem_cnt, si = 0, 0
    for t, (x, y) in enumerate(g["data"]):
        for b in range(3):
            st = g["steps"][si]
            si += 1
            yb = y[b * 16:(b + 1) * 16]
            endcnt = min(mem_cnt + yb.size(0), n_mem)
            eff = endcnt - mem_cnt
            mem.exemplars[t][mem_cnt:endcnt] = list(st["keys"][:eff])
            labels[t, mem_cnt:endcnt] = yb[:eff]
            mem_cnt += eff
            if mem_cnt == n_mem:
                mem_cnt = 0
            assert mem_cnt == st["mem_cnt"]
    assert torch.equal(labels, g["memory_labels"])
    for t in range(n_tasks):
        assert mem[t] == g["exemplars"][t]
    assert common.compute_offsets(0, [5, 10, 15]) == (0, 5) and common.compute_offsets(2, [5, 10, 15]) == (10, 15)

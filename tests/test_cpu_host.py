"""Host-side logic of the engine and the trainers that needs no GPU: the data-parallel bucket plan, the checkpoint writer and
its host-move helper, the local_only switch of dist."""
import os
import types

import torch
import torch.nn as nn

from clsurvey_b200 import dist as cdist
from clsurvey_b200.engine import Engine
from clsurvey_b200.methods import trainers
from clsurvey_b200.models import make_vgg


def _layout(model):
    """ops / offsets / total as Engine lays them out (slots rounded up to 4 elements), without a device."""
    ops, offsets, off, k = [], [], 0, 0
    for m in list(model.features) + list(model.classifier):
        if isinstance(m, (nn.Conv2d, nn.Linear)):
            ops.append({"kind": "conv" if isinstance(m, nn.Conv2d) else "linear", "w": k, "b": k + 1})
            for p in (m.weight, m.bias):
                offsets.append(off)
                off += (p.numel() + 3) // 4 * 4
            k += 2
        else:
            ops.append({"kind": "other"})
    return types.SimpleNamespace(ops=ops, offsets=offsets, total=off, _dp_cut_cache=None, _dp_buckets_cache=None)


def test_dp_bucket_plan_vgg11():
    """Engine._dp_cut / _dp_buckets on VGG-11's layout: head = conv1-4 (<= 10 % of the parameters); the tail in three
    buckets, back to front, contiguous, each starting at a layer boundary (train_EWC.py:187: one backward pass fills them in
    that order)."""
    st = _layout(make_vgg("VGG11_cl_512_512"))
    st._dp_cut = lambda: Engine._dp_cut(st)
    i, off = Engine._dp_cut(st)
    assert off == 960896 and st.ops[i]["kind"] == "conv"
    b = Engine._dp_buckets(st)
    assert [x[1] for x in b] == [6860672, 4500864, 960896]
    end = st.total
    for bi, bo, be in b:
        assert be == end and bo % 4 == 0 and st.offsets[st.ops[bi]["w"]] == bo
        end = bo
    assert end == off
    # backward visits ops back to front: a bucket's op index is reached after every op behind it
    assert [x[0] for x in b] == sorted((x[0] for x in b), reverse=True)


def test_dp_bucket_plan_without_cut():
    """A net whose first layer already holds > 10 % of the parameters has no head/tail split and therefore no buckets."""
    net = types.SimpleNamespace(features=[nn.Conv2d(3, 64, 3)], classifier=[nn.Linear(64, 10)])
    st = _layout(net)
    st._dp_cut = lambda: Engine._dp_cut(st)
    assert Engine._dp_cut(st) == (0, 0) and Engine._dp_buckets(st) == []


def test_local_only_switch():
    saved = cdist._state["world"]
    try:
        cdist._state["world"] = 4
        assert cdist.is_distributed()
        with cdist.local_only():
            assert not cdist.is_distributed() and cdist.world_size() == 1
        assert cdist.world_size() == 4
    finally:
        cdist._state["world"] = saved


def test_to_host_inplace_keeps_identities():
    """reg_params is keyed by the Parameters themselves (main_EWC.py:160-173): the host move must not re-create them."""
    m = nn.Sequential(nn.Linear(4, 3), nn.ReLU(), nn.Linear(3, 2))
    params = list(m.parameters())
    m.reg_params = {p: {"omega": torch.ones_like(p), "init_val": p.detach().clone()} for p in params[:-2]}
    m.reg_params["lambda"] = 3.0
    ckpt = {"model": m, "state_dict": m.state_dict(), "optimizer": {"state": {0: {"momentum_buffer": torch.zeros(3, 4)}}}, "epoch": 1}
    trainers._to_host_inplace(ckpt)
    assert list(m.parameters()) == params and all(p in m.reg_params for p in params[:-2])
    assert all(not t.is_cuda for t in m.state_dict().values())


def test_checkpoint_writer_files_complete_on_wait(tmp_path):
    """train_EWC.py:207-227: best_model.pth.tar is the pickled module, epoch.pth.tar a dict; both must be loadable the moment
    train_model returns (writer.wait()).  A newer snapshot of the same path replaces an older pending one."""
    w = trainers._CheckpointWriter()
    m = nn.Linear(8, 4)
    for k in range(3):
        with torch.no_grad():
            m.weight.fill_(float(k))
        w.submit(m, os.path.join(tmp_path, "best_model.pth.tar"))
    w.submit({"epoch": 3, "model": m, "state_dict": m.state_dict()}, os.path.join(tmp_path, "epoch.pth.tar"))
    with torch.no_grad():
        m.weight.fill_(99.0)                       # later training steps must not leak into the snapshots
    w.wait()
    best = torch.load(os.path.join(tmp_path, "best_model.pth.tar"), weights_only=False)
    assert isinstance(best, nn.Linear) and float(best.weight.detach()[0, 0]) == 2.0
    ep = torch.load(os.path.join(tmp_path, "epoch.pth.tar"), weights_only=False)
    assert ep["epoch"] == 3 and float(ep["state_dict"]["weight"][0, 0]) == 2.0

"""GPU: the reference-named per-method entry points end to end (load pickled model -> importance pass -> new head ->
training -> pickled best model with reg_params), i.e. what `Method.train` / `Method.grid_train` call
(src/methods/method.py:674-685, 708-716, 737-750, 1006-1025, 383-413).  Checks the plumbing the framework relies on:
return types, side files, that the pickled model reloads with its reg_params keyed by its own parameters, and that a
reloaded model continues (task 3) -- plus CUDA-graph replay == eager execution, bit for bit."""
import os
import types

import pytest
import torch
import torch.nn as nn

from tests.util import BS, NCLS, load_golden, loaders, rel_err, tiny_model

pytestmark = pytest.mark.gpu


class DS(torch.utils.data.TensorDataset):
    classes = list(range(NCLS))


def _task(seed, n=64):
    g = torch.Generator().manual_seed(seed)
    return DS(torch.randn(n, 3, 16, 16, generator=g), torch.randint(0, NCLS, (n,), generator=g))


def _save_task(tmp_path, name, seed):
    p = str(tmp_path / (name + ".pth"))
    torch.save({"train": _task(seed), "val": _task(seed + 500, 32)}, p)
    return p


def _first_model(tmp_path):
    torch.manual_seed(7)
    m = tiny_model()
    p = str(tmp_path / "first_task_model.pth.tar")
    torch.save(m, p)
    return p


@pytest.mark.parametrize("which", ["EWC", "MAS", "SI"])
def test_method_train_two_tasks(tmp_path, which):
    from clsurvey_b200.methods import method as M
    meth = M.parse(which)
    model_path = _first_model(tmp_path)
    for task in (2, 3):
        exp_dir = str(tmp_path / ("task_%d" % task))
        manager = types.SimpleNamespace(current_task_dataset_path=_save_task(tmp_path, "t%d" % task, 10 + task),
                                        previous_task_model_path=model_path, heuristic_exp_dir=exp_dir,
                                        reg_sets=[_save_task(tmp_path, "prev%d" % task, 9 + task)])
        args = types.SimpleNamespace(data_dir=None, batch_size=BS, num_epochs=2, lr=0.05, weight_decay=0.0,
                                     saving_freq=1, init_model_path=None)
        model, acc = meth.train(args, manager, {"lambda": 2.0})
        assert 0.0 <= acc <= 1.0
        best = os.path.join(exp_dir, "best_model.pth.tar")
        assert os.path.isfile(best) and os.path.isfile(os.path.join(exp_dir, "epoch.pth.tar"))
        assert os.path.isfile(os.path.join(exp_dir, "preprocess_time.pth.tar"))
        re = torch.load(best, weights_only=False)
        assert hasattr(re, "reg_params") and re.reg_params["lambda"] == 2.0
        n_reg = sum(1 for p in re.parameters() if p in re.reg_params)
        assert n_reg >= len(list(re.parameters())) - 2            # everything but (possibly) the fresh head
        for p in re.parameters():
            if p in re.reg_params:
                assert re.reg_params[p]["omega"].shape == p.shape and torch.isfinite(re.reg_params[p]["omega"]).all()
        model_path = best


def test_finetune_grid_train(tmp_path):
    from clsurvey_b200.methods import method as M
    manager = types.SimpleNamespace(current_task_dataset_path=_save_task(tmp_path, "t2", 21),
                                    previous_task_model_path=_first_model(tmp_path),
                                    gridsearch_exp_dir=str(tmp_path / "grid"))
    args = types.SimpleNamespace(batch_size=BS, num_epochs=2, weight_decay=0.0, saving_freq=1)
    model, acc = M.Finetune.grid_train(args, manager, 0.05)
    assert os.path.isfile(os.path.join(manager.gridsearch_exp_dir, "best_model.pth.tar")) and 0 <= acc <= 1


def test_gem_postprocess_then_train(tmp_path):
    from clsurvey_b200.methods.rehearsal import main_rehearsal, train_rehearsal
    base = _first_model(tmp_path)
    n_tasks, nc = 3, [NCLS] * 3
    wrapped = str(tmp_path / "gem_task1.pth.tar")
    common = dict(weight_decay=0.0, task_name="t", n_outputs=sum(nc), method="gem", n_memories=24, n_epochs=2,
                  memory_strength=0.5, n_tasks=n_tasks, batch_size=BS, lr=0.05, finetune=False)
    r = main_rehearsal.main(dict(common, task_count=1, prev_model_path=base, save_path=wrapped, is_scratch_model=True,
                                 postprocess=True, dataset_path=_save_task(tmp_path, "g1", 31)), nc)
    assert r == (None, None) and os.path.isfile(wrapped)
    out2 = str(tmp_path / "gem_task2")
    os.makedirs(out2)
    model, acc = main_rehearsal.main(dict(common, task_count=2, prev_model_path=wrapped, save_path=out2,
                                          is_scratch_model=False, postprocess=False,
                                          dataset_path=_save_task(tmp_path, "g2", 32)), nc)
    assert 0 <= acc <= 1 and os.path.isfile(os.path.join(out2, "best_model.pth.tar"))
    assert model.observed_tasks == [0, 1]
    assert len(train_rehearsal.LAST_RUN["violations"]) == 2 * (64 // BS)
    re = torch.load(os.path.join(out2, "best_model.pth.tar"), weights_only=False)
    assert re.memory_labels.shape == (n_tasks, 24) and re.observed_tasks == [0, 1]


def test_cuda_graph_replay_equals_eager(tmp_path, monkeypatch):
    """The trainer replays a captured graph from the third step of a configuration on; results must be bit-identical
    to fully eager execution."""
    from clsurvey_b200.engine import Engine
    from clsurvey_b200.methods import trainers
    from clsurvey_b200.methods.Finetune import train_SGD
    from clsurvey_b200.methods.optim import SGD
    f = load_golden("finetune")["wd5e-4"]
    out = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("CLB_CUDA_GRAPH", flag)
        m = tiny_model(f["init"])
        Engine(m, (3, 16, 16), BS)
        ld, sizes = loaders(f["data"])
        opt = SGD(m.parameters(), f["lr"], momentum=0.9, weight_decay=f["wd"])
        m, best = train_SGD.train_model(m, nn.CrossEntropyLoss(), opt, f["lr"], ld, sizes, True, f["epochs"],
                                        exp_dir=str(tmp_path), resume="", save_models_mode=False)
        out[flag] = (best, list(trainers.LAST_RUN["batch_losses"]), {k: v.clone() for k, v in m.state_dict().items()})
    assert out["0"][0] == out["1"][0] and out["0"][1] == out["1"][1]
    for k in out["0"][2]:
        assert torch.equal(out["0"][2][k], out["1"][2][k]), k


def test_evaluation_forward_path(tmp_path):
    """f-1: `test_model` / `get_output_def` / `inference_eval` / `get_prev_heads` (src/framework/inference.py:8-87,
    method.py:230-235,1066-1087, utils.py:235-262) through the engine: overall accuracy AND the per-class counters equal
    the torch-CPU evaluation of the same model with the same head -- for a head passed directly and for a head fetched by
    path from the model of an earlier task while a later model (different head, same trunk) is evaluated."""
    import copy
    from clsurvey_b200.framework import inference
    from clsurvey_b200.methods import method as M
    torch.manual_seed(3)
    model = tiny_model()
    with torch.no_grad():
        for m in model.classifier:
            if hasattr(m, "weight"):
                m.weight.mul_(30.0)                      # spread the logits so that arg-max is not degenerate
    ref = copy.deepcopy(model)
    ds = {"train": _task(76, 32), "val": _task(78, 32), "test": _task(77, 96)}
    head = copy.deepcopy(model.classifier._modules["4"])
    acc = inference.test_model(M.EWC(), model, ds, 0, target_head=[head], batch_size=BS, subset="test", per_class_stats=True)
    ref.eval()
    x, y = ds["test"].tensors
    with torch.no_grad():
        pred = ref(x).argmax(1)
    assert abs(acc - 100.0 * (pred == y).float().mean().item()) < 1e-9
    correct, total = inference.test_model.last_class_stats
    for c in range(NCLS):
        assert total[c] == float((y == c).sum()) and correct[c] == float(((pred == y) & (y == c)).sum())   # exact counters
    # a later model (new head) evaluated on the old task with the old model's head, both given as paths
    p1 = str(tmp_path / "task1_model.pth.tar")
    torch.save(ref, p1)
    later = copy.deepcopy(ref)
    torch.manual_seed(9)
    later.classifier._modules["4"] = nn.Linear(32, NCLS)
    p2 = str(tmp_path / "task2_model.pth.tar")
    torch.save({"model": later}, p2)                      # epoch.pth.tar-style dict: both forms load (method.py:1069-1070)
    dpath = str(tmp_path / "task1_data.pth")
    torch.save(ds, dpath)
    args = types.SimpleNamespace(eval_model_path=p2, head_paths=[p1], dset_path=dpath, test_set="test", batch_size=BS,
                                 eval_dset_idx=0)
    manager = types.SimpleNamespace(method=M.EWC())
    acc2 = M.EWC.inference_eval(args, manager)
    assert abs(acc2 - acc) < 1e-9                         # same trunk + the task-1 head = the task-1 accuracy
    with torch.no_grad():
        later.eval()
        acc_wrong_head = 100.0 * (later(x).argmax(1) == y).float().mean().item()
    assert abs(acc_wrong_head - acc) > 1e-9               # (the fresh head alone would give something else)


def test_two_heads_round_trip_do_not_alias():
    """Heads of the same shape swapped in and out of one model (get_prev_heads / test_model, utils.py:235-262,
    inference.py:56-87): a head that leaves the model must keep its own values, and coming back must restore its logits."""
    from clsurvey_b200.engine import get_engine
    torch.manual_seed(3)
    m = tiny_model()
    g = torch.Generator().manual_seed(4)
    x = torch.randn(BS, 3, 16, 16, generator=g).cuda()
    head_a = m.classifier._modules["4"]
    head_b = nn.Linear(32, NCLS)
    wb = head_b.weight.detach().clone()
    eng = get_engine(m, (3, 16, 16), BS)
    wa = head_a.weight.detach().clone().cpu()
    logits_a = eng.forward(x).clone()
    m.classifier._modules["4"] = head_b
    eng = get_engine(m)
    logits_b = eng.forward(x).clone()
    assert not torch.equal(logits_a, logits_b)
    assert torch.equal(head_a.weight.detach().cpu(), wa), "the head that left the model was overwritten"
    m.classifier._modules["4"] = head_a
    eng = get_engine(m)
    assert torch.equal(eng.forward(x), logits_a)
    assert torch.equal(head_b.weight.detach().cpu(), wb)
    m.classifier._modules["4"] = head_b
    assert torch.equal(get_engine(m).forward(x), logits_b)


def test_lr_grid_concurrent_replicas_match_sequential(tmp_path):
    """f-4: the learning-rate grid (src/framework/lr_grid_train.py:9-160) with its nodes trained as concurrent processes gives
    the accuracies, the best lr and the files of the node-by-node loop: every node is seeded by its iteration index
    (lr_grid_train.py:73,77), so where and when it runs does not matter."""
    from clsurvey_b200.framework import lr_grid_train as G
    from clsurvey_b200.methods import method as M
    from clsurvey_b200.utilities import utils
    lrs = [0.05, 0.01]
    common = dict(current_task_dataset_path=_save_task(tmp_path, "t2", 21), previous_task_model_path=_first_model(tmp_path))
    args = types.SimpleNamespace(lrs=lrs, finetune_iterations=2, task_counter=2, batch_size=BS, num_epochs=2, weight_decay=0.0,
                                 saving_freq=1)
    manager = types.SimpleNamespace(method=M.Finetune(), parent_exp_dir=str(tmp_path / "par"), **common)
    best_lr, best_acc = G.lr_grid_single_task(args, manager, save_models_mode="all", gpus=1)
    ck = torch.load(os.path.join(manager.ft_parent_exp_dir, "grid_checkpoint.pth"), weights_only=False)["processed_lrs"]
    # sequential reference: the same nodes one after the other in this process
    seq = {}
    for lr in lrs:
        for it in range(2):
            utils.set_random(it)
            m2 = types.SimpleNamespace(method=M.Finetune(), gridsearch_exp_dir=str(tmp_path / ("seq_%g_%d" % (lr, it))), **common)
            _, acc = M.Finetune.grid_train(args, m2, lr)
            seq[(lr, it)] = acc
    for lr in lrs:
        assert ck[lr]["acc"] == [seq[(lr, 0)], seq[(lr, 1)]], (lr, ck[lr], seq)
    avg = {lr: (seq[(lr, 0)] + seq[(lr, 1)]) / 2 for lr in lrs}
    exp_best = max(lrs, key=lambda l: (avg[l], -lrs.index(l)))           # first lr wins ties (strict '>' in the reference)
    assert best_lr == exp_best and abs(best_acc - avg[exp_best]) < 1e-12
    assert os.path.isfile(os.path.join(manager.best_exp_grid_node_dirname, "best_model.pth.tar"))
    assert manager.previous_task_model_path == os.path.join(manager.best_exp_grid_node_dirname, "best_model.pth.tar")
    # resumed grid: nothing left to train, same answer from the checkpoint
    manager2 = types.SimpleNamespace(method=M.Finetune(), parent_exp_dir=str(tmp_path / "par"), **common)
    assert G.lr_grid_single_task(args, manager2, save_models_mode="all", gpus=1) == (best_lr, best_acc)

"""GPU: the engine behind the reference-named entry points against the golden fixtures (= outputs of the
UNMODIFIED reference on CPU, oracle/gen_golden.py).  Same inputs, same seeds, same call sequence.

Tolerances: per-batch loss 1e-5 relative; parameters / omega 1e-4 normalised max error (north_star: fp32 loss and
importance weights within 1e-4); SI's w 2e-3 (ill-conditioned in the reference's own arithmetic, see
tests/test_oracle_golden.py); best_acc, #correct, ring-buffer state and violation counts exact.
"""
import types

import numpy as np
import pytest
import torch
import torch.nn as nn

from tests.util import BS, NCLS, load_golden, loaders, rel_err, tiny_model

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _dsets(x, y):
    ds = torch.utils.data.TensorDataset(x, y)
    ds.classes = list(range(NCLS))
    return {"train": ds, "val": ds}


def _train_losses(ref_losses, epochs, nb_t, nb_v):
    return [l for e in range(epochs) for l in ref_losses[e * (nb_t + nb_v): e * (nb_t + nb_v) + nb_t]]


def test_finetune_train_model(tmp_path):
    from clsurvey_b200.engine import Engine
    from clsurvey_b200.methods import trainers
    from clsurvey_b200.methods.Finetune import train_SGD
    from clsurvey_b200.methods.optim import SGD
    for tag, f in load_golden("finetune").items():
        m = tiny_model(f["init"])
        Engine(m, (3, 16, 16), BS)
        ld, sizes = loaders(f["data"])
        opt = SGD(m.parameters(), f["lr"], momentum=0.9, weight_decay=f["wd"])
        m, best = train_SGD.train_model(m, nn.CrossEntropyLoss(), opt, f["lr"], ld, sizes, True, f["epochs"],
                                        exp_dir=str(tmp_path), resume="", save_models_mode=False)
        assert best == f["best_acc"]
        ref = _train_losses(f["losses"], f["epochs"], len(ld["train"]), len(ld["val"]))
        assert np.allclose(trainers.LAST_RUN["batch_losses"], ref, rtol=1e-5, atol=0)
        sd = m.state_dict()
        for k, v in f["final"].items():
            assert rel_err(sd[k], v) <= TOL, (tag, k)


def _penalty(which, tmp_path):
    if which == "ewc":
        from clsurvey_b200.methods.EWC import main_EWC as M, train_EWC as T
        accumulate = lambda model, ds: M.accumulate_EWC_weights(None, [ds], model, BS)
    else:
        from clsurvey_b200.methods.MAS import main_MAS as M, train_MAS as T
        accumulate = lambda model, ds: M.accumulate_objective_based_weights(None, [ds], model, BS, "L2", "train")
    from clsurvey_b200.engine import get_engine
    from clsurvey_b200.methods import trainers
    g = load_golden(which)
    m = tiny_model(g["init"])
    get_engine(m, (3, 16, 16), BS)
    for r, rnd in enumerate(g["rounds"]):
        xp, yp = rnd["prev_data"]
        m = accumulate(m, _dsets(xp, yp))
        m.reg_params["lambda"] = rnd["lam"]
        named = dict(m.named_parameters())
        for n, ref in rnd["reg_after_pass"].items():
            rp = m.reg_params[named[n]]
            assert rel_err(rp["omega"], ref["omega"]) <= TOL, (which, r, n)
            assert rel_err(rp["init_val"], ref["init_val"]) <= TOL, (which, r, n)
        m.classifier._modules["4"] = nn.Linear(32, NCLS)
        m.classifier._modules["4"].load_state_dict(rnd["new_head"])
        get_engine(m).bind(m)
        ld, sizes = loaders(rnd["data"])
        opt = T.Weight_Regularized_SGD(m.parameters(), rnd["lr"], momentum=0.9, weight_decay=rnd["wd"])
        m, best = T.train_model(m, nn.CrossEntropyLoss(), opt, rnd["lr"], ld, sizes, True, rnd["epochs"],
                                exp_dir=str(tmp_path), resume="")
        assert best == rnd["best_acc"]
        ref = _train_losses(rnd["losses"], rnd["epochs"], len(ld["train"]), len(ld["val"]))
        assert np.allclose(trainers.LAST_RUN["batch_losses"], ref, rtol=1e-5, atol=0)
        sd = m.state_dict()
        for k, v in rnd["final"].items():
            assert rel_err(sd[k], v) <= TOL, (which, r, k)


def test_ewc_fisher_and_penalised_training(tmp_path):
    _penalty("ewc", tmp_path)


def test_mas_omega_and_penalised_training(tmp_path):
    _penalty("mas", tmp_path)


def test_si_path_integral(tmp_path):
    from clsurvey_b200.engine import get_engine
    from clsurvey_b200.methods import trainers
    from clsurvey_b200.methods.SI import train_SI as T
    g = load_golden("si")
    m = tiny_model(g["init"])
    get_engine(m, (3, 16, 16), BS)
    for r, rnd in enumerate(g["rounds"]):
        if r == 0:
            reg = T.initialize_reg_params(m)
        else:
            m.classifier._modules["4"] = nn.Linear(32, NCLS)
            m.classifier._modules["4"].load_state_dict(rnd["head"])
            get_engine(m).bind(m)
            params = list(m.parameters())
            m.reg_params.pop(params[-1], None)
            m.reg_params.pop(params[-2], None)
            reg = T.update_reg_params(m)
        reg["lambda"] = rnd["lam"]
        m.reg_params = reg
        named = dict(m.named_parameters())
        for n, ref in rnd["reg_before"].items():
            for key, tol in (("omega", 2e-3), ("w", 2e-3), ("init_val", TOL)):
                assert rel_err(reg[named[n]][key], ref[key]) <= tol, (r, n, key)
        ld, sizes = loaders(rnd["data"])
        opt = T.Elastic_SGD(m.parameters(), rnd["lr"], momentum=0.9, weight_decay=0.0)
        m, best = T.train_model(m, nn.CrossEntropyLoss(), opt, rnd["lr"], ld, sizes, True, rnd["epochs"],
                                exp_dir=str(tmp_path), resume="")
        assert best == rnd["best_acc"]
        assert len(trainers.LAST_RUN["batch_losses"]) == (rnd["epochs"] + 1) * len(ld["train"])
        sd = m.state_dict()
        for k, v in rnd["final"].items():
            assert rel_err(sd[k], v) <= TOL, (r, k)
        for n, ref in rnd["reg_after"].items():
            assert rel_err(m.reg_params[named[n]]["w"], ref["w"]) <= 2e-3, (r, n)


def test_gem_observe(tmp_path):
    from clsurvey_b200.methods.rehearsal.model import gem as G
    g = load_golden("gem")
    base = tiny_model(g["init"], dropout=True)
    args = types.SimpleNamespace(prev_model_path=base, n_memories=g["n_mem"], lr=g["lr"], weight_decay=0.0,
                                 memory_strength=g["margin"], batch_size=g["bs"], nc_per_task=[NCLS] * g["n_tasks"],
                                 input_shape=(3, 16, 16), shuffle_memory=False)
    net = G.Net(0, NCLS * g["n_tasks"], g["n_tasks"], args)
    net.net.load_state_dict(g["wrapped_init"])
    si = 0
    for t, (x, y) in enumerate(g["data"]):
        for b in range(3):
            st = g["steps"][si]
            si += 1
            xb, yb = x[b * 16:(b + 1) * 16], y[b * 16:(b + 1) * 16]
            net.forced_masks = st["masks"]
            loss, corr, stats = net.observe(xb, t, yb, st["keys"], args)
            v = stats["projected_grads"][0]
            assert int(v.item() if torch.is_tensor(v) else v) == st["violations"], si     # violation count: exact
            assert net.mem_cnt == st["mem_cnt"]
            assert abs(loss.item() - st["loss"]) <= 1e-5 * abs(st["loss"])
            assert int(corr.item()) == st["correct"]
            flat = torch.cat([p.data.reshape(-1) for p in net.net.parameters()])
            assert rel_err(flat, st["params"]) <= TOL, si
    assert torch.equal(net.memory_labels, g["memory_labels"])                              # ring buffer: bit-exact
    for t in range(g["n_tasks"]):
        assert net.memory_data[t] == g["exemplars"][t]

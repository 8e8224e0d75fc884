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

from tests.util import BS, NCLS, WIDE_BS, WIDE_HW, load_golden, loaders, rel_err, tiny_model, wide_model

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


# ---------------------------------------------------------------------------------------------- ragged tails / epoch protocol
def _one_penalty_round(which, r, m, tmp_path, bs, hw, n_train_batches, n_val_batches, loss_rtol=1e-5):
    """importance pass on the previous task's data, fresh head, penalised train_model -- one round of a fixture"""
    if which == "ewc":
        from clsurvey_b200.methods.EWC import main_EWC as M, train_EWC as T
        accumulate = lambda model, ds: M.accumulate_EWC_weights(None, [ds], model, bs)
    else:
        from clsurvey_b200.methods.MAS import main_MAS as M, train_MAS as T
        accumulate = lambda model, ds: M.accumulate_objective_based_weights(None, [ds], model, bs, "L2", "train")
    from clsurvey_b200.engine import get_engine
    from clsurvey_b200.methods import trainers
    get_engine(m, (3, hw, hw), bs)
    xp, yp = r["prev_data"]
    m = accumulate(m, _dsets(xp, yp))
    m.reg_params["lambda"] = r["lam"]
    named = dict(m.named_parameters())
    for n, ref in r["reg_after_pass"].items():
        assert rel_err(m.reg_params[named[n]]["omega"], ref["omega"]) <= TOL, (which, n, "omega")
        assert rel_err(m.reg_params[named[n]]["init_val"], ref["init_val"]) <= TOL, (which, n, "init_val")
    m.classifier._modules["4"] = nn.Linear(32, NCLS)
    m.classifier._modules["4"].load_state_dict(r["new_head"])
    get_engine(m).bind(m)
    ld, sizes = loaders(r["data"], bs)
    opt = T.Weight_Regularized_SGD(m.parameters(), r["lr"], momentum=0.9, weight_decay=r["wd"])
    m, best = T.train_model(m, nn.CrossEntropyLoss(), opt, r["lr"], ld, sizes, True, r["epochs"], exp_dir=str(tmp_path), resume="")
    assert best == r["best_acc"], (which, best, r["best_acc"])
    ref = _train_losses(r["losses"], r["epochs"], n_train_batches, n_val_batches)
    assert np.allclose(trainers.LAST_RUN["batch_losses"], ref, rtol=loss_rtol, atol=0), which
    for k, v in r["final"].items():
        assert rel_err(m.state_dict()[k], v) <= TOL, (which, k)
    return m


def test_ragged_tails(tmp_path):
    """Dataset sizes that leave ragged last batches (importance pass 56 = 16,16,16,8 -- MAS's running mean uses the CURRENT
    batch size, train_MAS.py:168-173 --, training 41 = 16,16,9, validation 18 = 16,2): tests/golden/ragged.pt."""
    from clsurvey_b200.engine import Engine
    from clsurvey_b200.methods import trainers
    from clsurvey_b200.methods.Finetune import train_SGD
    from clsurvey_b200.methods.optim import SGD
    g = load_golden("ragged")
    f = g["finetune"]
    m = tiny_model(f["init"])
    Engine(m, (3, 16, 16), BS)
    ld, sizes = loaders(f["data"])
    opt = SGD(m.parameters(), f["lr"], momentum=0.9, weight_decay=f["wd"])
    m, best = train_SGD.train_model(m, nn.CrossEntropyLoss(), opt, f["lr"], ld, sizes, True, f["epochs"], exp_dir=str(tmp_path),
                                    resume="", save_models_mode=False)
    assert best == f["best_acc"]
    assert np.allclose(trainers.LAST_RUN["batch_losses"], _train_losses(f["losses"], f["epochs"], 3, 2), rtol=1e-5, atol=0)
    for k, v in f["final"].items():
        assert rel_err(m.state_dict()[k], v) <= TOL, ("finetune", k)
    for which in ("ewc", "mas"):
        _one_penalty_round(which, g[which], tiny_model(g[which]["init"]), tmp_path, BS, 16, 3, 2)


@pytest.mark.parametrize("which", ["sgd", "ewc", "si", "sgd_short", "diverge_ewc", "diverge_si", "diverge_sgd"])
def test_epoch_protocol(which, tmp_path):
    """Long runs without improvement: lr cut at val_beat_counts == 5, stop at > 10 (Finetune train_SGD.py:10-30, EWC
    train_EWC.py:89-101) resp. >= 10 with the num_epochs + 1 range (SI train_SI.py:129-141,182); divergence: EWC / SI abort
    when the epoch loss exceeds 1e4 or is NaN (train_EWC.py:204-205, train_SI.py:242-244), Finetune keeps going."""
    from clsurvey_b200.engine import get_engine
    from clsurvey_b200.methods import trainers
    from clsurvey_b200.methods.EWC import train_EWC as TE
    from clsurvey_b200.methods.Finetune import train_SGD
    from clsurvey_b200.methods.SI import train_SI as TI
    from clsurvey_b200.methods.optim import SGD
    sch = load_golden("schedule")
    ld, sizes = loaders(sch["data"])
    r = sch[which]
    m = tiny_model(sch["init"])
    get_engine(m, (3, 16, 16), BS)
    tmp = str(tmp_path)
    if which.endswith("sgd") or which == "sgd_short":
        opt = SGD(m.parameters(), r["lr"], momentum=0.9, weight_decay=0.0)
        m, best = train_SGD.train_model(m, nn.CrossEntropyLoss(), opt, r["lr"], ld, sizes, True, r["epochs"], exp_dir=tmp,
                                        resume="", save_models_mode=False)
    elif which.endswith("ewc"):
        m.reg_params = {p: dict(omega=torch.ones_like(p), init_val=p.data.clone()) for p in m.parameters()}
        m.reg_params["lambda"] = 1.0
        opt = TE.Weight_Regularized_SGD(m.parameters(), r["lr"], momentum=0.9, weight_decay=0.0)
        m, best = TE.train_model(m, nn.CrossEntropyLoss(), opt, r["lr"], ld, sizes, True, r["epochs"], exp_dir=tmp, resume="")
    else:
        reg = TI.initialize_reg_params(m)
        reg["lambda"] = 1.0
        m.reg_params = reg
        opt = TI.Elastic_SGD(m.parameters(), r["lr"], momentum=0.9, weight_decay=0.0)
        m, best = TI.train_model(m, nn.CrossEntropyLoss(), opt, r["lr"], ld, sizes, True, r["epochs"], exp_dir=tmp, resume="")
    n_calls = sum(len(ld[phase]) for _, phase, _, _ in trainers.LAST_RUN["epochs"])
    assert n_calls == r["n_criterion_calls"], (which, n_calls, r["n_criterion_calls"])
    assert abs(opt.param_groups[0]["lr"] - r["final_lr"]) <= 1e-12 * r["final_lr"], (which, opt.param_groups[0]["lr"])
    assert best == r["best_acc"], (which, best, r["best_acc"])


# ---------------------------------------------------------------------------------------------- >= 64 channels (planes kernels)
def test_wide_ewc_through_planes_kernels(tmp_path):
    """tests/golden/wide.pt: the UNMODIFIED reference's EWC round (diag_fisher + Weight_Regularized_SGD training) on a 64 /
    64 / 128-channel VGGSlim -- every conv runs on the kernels bench.py times (fused first layer, TMA-fed tcgen05 convs)."""
    from clsurvey_b200.engine import get_engine
    r = load_golden("wide")["ewc"]
    m = wide_model(r["init"])
    eng = get_engine(m, (3, WIDE_HW, WIDE_HW), WIDE_BS)
    assert eng.ops[0].get("fused_first") and sum(1 for op in eng.ops if op.get("planes")) == 2, "planes pipeline not active"
    _one_penalty_round("ewc", r, m, tmp_path, WIDE_BS, WIDE_HW, 2, 1, loss_rtol=TOL)    # tensor-core layers: north_star 1e-4


def test_wide_si_through_planes_kernels(tmp_path):
    from clsurvey_b200.engine import get_engine
    from clsurvey_b200.methods import trainers
    from clsurvey_b200.methods.SI import train_SI as T
    r = load_golden("wide")["si"]
    m = wide_model(r["init"])
    eng = get_engine(m, (3, WIDE_HW, WIDE_HW), WIDE_BS)
    assert eng.ops[0].get("fused_first") and sum(1 for op in eng.ops if op.get("planes")) == 2, "planes pipeline not active"
    reg = T.initialize_reg_params(m)
    reg["lambda"] = r["lam"]
    m.reg_params = reg
    named = dict(m.named_parameters())
    with torch.no_grad():
        for n, ref in r["reg_before"].items():
            reg[named[n]]["omega"].copy_(ref["omega"])
            reg[named[n]]["init_val"].copy_(ref["init_val"])
    ld, sizes = loaders(r["data"], WIDE_BS)
    opt = T.Elastic_SGD(m.parameters(), r["lr"], momentum=0.9, weight_decay=0.0)
    m, best = T.train_model(m, nn.CrossEntropyLoss(), opt, r["lr"], ld, sizes, True, r["epochs"], exp_dir=str(tmp_path), resume="")
    assert best == r["best_acc"]
    assert len(trainers.LAST_RUN["batch_losses"]) == (r["epochs"] + 1) * len(ld["train"])
    ref_losses = [l for e in range(r["epochs"] + 1) for l in r["losses"][e * 3: e * 3 + 2]]
    assert np.allclose(trainers.LAST_RUN["batch_losses"], ref_losses, rtol=TOL, atol=0)
    for k, v in r["final"].items():
        assert rel_err(m.state_dict()[k], v) <= TOL, k
    for n, ref in r["reg_after"].items():
        assert rel_err(m.reg_params[named[n]]["w"], ref["w"]) <= 2e-3, n


# ---------------------------------------------------------------------------------------------- IMM
def test_imm_precision_and_merge(tmp_path):
    """tests/golden/imm.pt (methods/IMM/merge.py run unmodified): mode-IMM precision with labels sampled from the host
    generator (same draws as the reference: the multinomial stream is replayed), the mode merge bit-for-bit op order, and the
    reference's mean-merge quirk (an unchanged copy of the last model)."""
    from clsurvey_b200.engine import get_engine
    from clsurvey_b200.methods.IMM import merge as MG
    g = load_golden("imm")
    models, precisions = [], []
    for t, state in enumerate(g["states"]):
        m = tiny_model(state)
        get_engine(m, (3, 16, 16), BS)
        (xt, yt), (xv, yv) = g["data"][t]
        mk = lambda x, y: torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=BS, shuffle=False)
        torch.manual_seed(g["seeds"][t])
        prec = MG.diag_fisher(m, {"train": mk(xt, yt), "val": mk(xv, yv)}, exclude_params=g["head_names"])
        assert set(prec) == set(g["precisions"][t])
        for n, v in g["precisions"][t].items():
            assert rel_err(prec[n], v) <= TOL, (t, n)
        models.append(m)
        precisions.append(g["precisions"][t])              # merge from the reference's precisions: isolates the merge kernel
    sums = [precisions[0]]
    for t in (1, 2):
        sums.append({n: sums[-1][n] + precisions[t][n] for n in precisions[t]})
    for i, upto in enumerate((1, 2)):
        mean = MG.IMM_merge_models(models, upto, g["head_names"], mean_mode=True).state_dict()
        mode = MG.IMM_merge_models(models, upto, g["head_names"], precision=precisions, sum_precision=sums[upto],
                                   mean_mode=False).state_dict()
        for k in g["merged_mean"][i]:
            assert torch.equal(mean[k].cpu(), g["merged_mean"][i][k]), ("mean", upto, k)
            assert rel_err(mode[k], g["merged_mode"][i][k]) <= 1e-6, ("mode", upto, k)


def test_imm_l2_transfer_is_the_penalty_step_with_unit_omega(tmp_path):
    """train_L2transfer.py:35-100 == train_EWC.py:46-84 with omega = 1 (main_L2transfer.py:41,58): the L2-transfer entry
    point reproduces an EWC-style run of the oracle with unit omega, including its fresh head being registered."""
    import copy
    from oracle import restate
    from clsurvey_b200.methods import method as M
    from clsurvey_b200.methods.IMM import main_L2transfer as L
    g = load_golden("finetune")["wd0"]
    torch.manual_seed(5)
    base = tiny_model(g["init"])
    mp = str(tmp_path / "prev.pth.tar")
    torch.save(base, mp)
    xt, yt, xv, yv = g["data"]

    class DS(torch.utils.data.TensorDataset):
        classes = list(range(NCLS))
    dsets = {"train": DS(xt, yt), "val": DS(xv, yv)}
    torch.manual_seed(77)
    model, acc = L.fine_tune_l2transfer(dsets, mp, str(tmp_path / "exp"), batch_size=BS, num_epochs=2, lr=0.05, reg_lambda=0.5)
    # oracle: same head draw, omega = 1 everywhere, theta* = theta at the start, same shuffled batches
    ref = tiny_model(g["init"])
    torch.manual_seed(77)
    ref.classifier._modules["4"] = nn.Linear(32, NCLS)
    reg = [dict(omega=torch.ones_like(p), init_val=p.data.clone()) for p in ref.parameters()]
    tr = restate.Trainer(ref, "penalty", 0.05, reg=reg, lam=0.5)
    mk = lambda x, y: torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=BS, shuffle=True)
    best, _, _ = tr.train_model({"train": mk(xt, yt), "val": mk(xv, yv)}, {"train": len(xt), "val": len(xv)}, 2)
    assert acc == best
    for k, v in ref.state_dict().items():
        assert rel_err(model.state_dict()[k], v) <= TOL, k
    assert isinstance(M.parse("modeIMM"), M.IMM) and M.parse("meanIMM").mode == "mean"

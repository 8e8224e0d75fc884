"""GPU: the ragged-tail and long-schedule golden fixtures (tests/golden/{ragged,schedule}.pt, written at the end of round 1
after the GPU budget was spent) through the reference-named entry points.  Run once on a B200 (`python tools/ragged_parity.py`), then move the
body into tests/test_gpu_golden.py::test_ragged_tails -- it mirrors test_finetune_train_model / _penalty there."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn as nn
from clsurvey_b200.engine import Engine, get_engine
from clsurvey_b200.methods import trainers
from clsurvey_b200.methods.Finetune import train_SGD
from clsurvey_b200.methods.optim import SGD
from tests.util import BS, NCLS, load_golden, loaders, rel_err, tiny_model

TOL = 1e-4
tmp = tempfile.mkdtemp()
g = load_golden("ragged")


def train_losses(ref, epochs, nb_t, nb_v):
    return [l for e in range(epochs) for l in ref[e * (nb_t + nb_v): e * (nb_t + nb_v) + nb_t]]


f = g["finetune"]
m = tiny_model(f["init"])
Engine(m, (3, 16, 16), BS)
ld, sizes = loaders(f["data"])
opt = SGD(m.parameters(), f["lr"], momentum=0.9, weight_decay=f["wd"])
m, best = train_SGD.train_model(m, nn.CrossEntropyLoss(), opt, f["lr"], ld, sizes, True, f["epochs"], exp_dir=tmp, resume="",
                                save_models_mode=False)
assert best == f["best_acc"], (best, f["best_acc"])
assert np.allclose(trainers.LAST_RUN["batch_losses"], train_losses(f["losses"], f["epochs"], 3, 2), rtol=1e-5, atol=0)
for k, v in f["final"].items():
    assert rel_err(m.state_dict()[k], v) <= TOL, ("finetune", k)
print("finetune ragged ok")


class DS(torch.utils.data.TensorDataset):
    classes = list(range(NCLS))


for which in ("ewc", "mas"):
    if which == "ewc":
        from clsurvey_b200.methods.EWC import main_EWC as M, train_EWC as T
        accumulate = lambda model, ds: M.accumulate_EWC_weights(None, [ds], model, BS)
    else:
        from clsurvey_b200.methods.MAS import main_MAS as M, train_MAS as T
        accumulate = lambda model, ds: M.accumulate_objective_based_weights(None, [ds], model, BS, "L2", "train")
    r = g[which]
    m = tiny_model(r["init"])
    get_engine(m, (3, 16, 16), BS)
    xp, yp = r["prev_data"]
    ds = DS(xp, yp)
    m = accumulate(m, {"train": ds, "val": ds})
    m.reg_params["lambda"] = r["lam"]
    named = dict(m.named_parameters())
    for n, ref in r["reg_after_pass"].items():
        assert rel_err(m.reg_params[named[n]]["omega"], ref["omega"]) <= TOL, (which, n)
    m.classifier._modules["4"] = nn.Linear(32, NCLS)
    m.classifier._modules["4"].load_state_dict(r["new_head"])
    get_engine(m).bind(m)
    ld, sizes = loaders(r["data"])
    opt = T.Weight_Regularized_SGD(m.parameters(), r["lr"], momentum=0.9, weight_decay=r["wd"])
    m, best = T.train_model(m, nn.CrossEntropyLoss(), opt, r["lr"], ld, sizes, True, r["epochs"], exp_dir=tmp, resume="")
    assert best == r["best_acc"], (which, best, r["best_acc"])
    assert np.allclose(trainers.LAST_RUN["batch_losses"], train_losses(r["losses"], r["epochs"], 3, 2), rtol=1e-5, atol=0)
    for k, v in r["final"].items():
        assert rel_err(m.state_dict()[k], v) <= TOL, (which, k)
    print(which, "ragged ok")
# ---- long-run epoch protocol (tests/golden/schedule.pt): lr cut at count 5, stop at > 10 (SI >= 10)
from clsurvey_b200.methods.EWC import train_EWC as TE
from clsurvey_b200.methods.SI import train_SI as TI
sch = load_golden("schedule")
ld, sizes = loaders(sch["data"])
per_epoch = len(ld["train"]) + len(ld["val"])
for which in ("sgd", "ewc", "si", "sgd_short", "diverge_ewc", "diverge_si", "diverge_sgd"):
    r = sch[which]
    m = tiny_model(sch["init"])
    get_engine(m, (3, 16, 16), BS)
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
    print(which, "schedule ok:", n_calls // per_epoch, "epochs")
print("RAGGED_PARITY_OK")

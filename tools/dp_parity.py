"""Data-parallel parity on real GPUs (run under torchrun, world >= 2): the golden fixtures must be reproduced when the
mini-batches are sharded over ranks (one NCCL all-reduce of the flat gradient per step; Fisher / MAS passes sharded by
whole batches with one all-reduce of omega).  Every rank checks its own replica."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn as nn
from clsurvey_b200 import dist as cdist
from clsurvey_b200.engine import Engine, get_engine
from clsurvey_b200.methods import trainers
from clsurvey_b200.methods.EWC import main_EWC, train_EWC
from clsurvey_b200.methods.Finetune import train_SGD
from clsurvey_b200.methods.MAS import main_MAS
from clsurvey_b200.methods.optim import SGD
from tests.util import BS, NCLS, load_golden, loaders, rel_err, tiny_model

cdist.init()
rank, world = cdist.rank(), cdist.world_size()
assert world >= 2, "run with torchrun --nproc-per-node >= 2"
tmp = tempfile.mkdtemp()

f = load_golden("finetune")["wd5e-4"]
m = tiny_model(f["init"])
Engine(m, (3, 16, 16), BS)
ld, sizes = loaders(f["data"])
opt = SGD(m.parameters(), f["lr"], momentum=0.9, weight_decay=f["wd"])
m, best = train_SGD.train_model(m, nn.CrossEntropyLoss(), opt, f["lr"], ld, sizes, True, f["epochs"], exp_dir=tmp, resume="",
                                save_models_mode=False)
assert best == f["best_acc"], (best, f["best_acc"])
nb_t, nb_v = len(ld["train"]), len(ld["val"])
ref = [l for e in range(f["epochs"]) for l in f["losses"][e * (nb_t + nb_v): e * (nb_t + nb_v) + nb_t]]
assert np.allclose(trainers.LAST_RUN["batch_losses"], ref, rtol=1e-5, atol=0), "finetune losses"
for k, v in f["final"].items():
    assert rel_err(m.state_dict()[k], v) <= 1e-4, ("finetune", k)


class DS(torch.utils.data.TensorDataset):
    classes = list(range(NCLS))


for which, mod in (("ewc", main_EWC), ("mas", main_MAS)):
    g = load_golden(which)
    m = tiny_model(g["init"])
    get_engine(m, (3, 16, 16), BS)
    rnd = g["rounds"][0]
    xp, yp = rnd["prev_data"]
    ds = {"train": DS(xp, yp)}
    if which == "ewc":
        m = mod.accumulate_EWC_weights(None, [ds], m, BS)
    else:
        m = mod.accumulate_objective_based_weights(None, [ds], m, BS, "L2", "train")
    named = dict(m.named_parameters())
    for n, r in rnd["reg_after_pass"].items():
        assert rel_err(m.reg_params[named[n]]["omega"], r["omega"]) <= 1e-4, (which, n)
    m.reg_params["lambda"] = rnd["lam"]
    m.classifier._modules["4"] = nn.Linear(32, NCLS)
    m.classifier._modules["4"].load_state_dict(rnd["new_head"])
    get_engine(m).bind(m)
    ld, sizes = loaders(rnd["data"])
    opt = train_EWC.Weight_Regularized_SGD(m.parameters(), rnd["lr"], momentum=0.9, weight_decay=rnd["wd"])
    m, best = train_EWC.train_model(m, nn.CrossEntropyLoss(), opt, rnd["lr"], ld, sizes, True, rnd["epochs"], exp_dir=tmp, resume="")
    assert best == rnd["best_acc"]
    for k, v in rnd["final"].items():
        assert rel_err(m.state_dict()[k], v) <= 1e-4, (which, k)
# replicas must be bit-identical (deterministic all-reduce + identical fused update)
flat = torch.cat([p.data.reshape(-1) for p in m.parameters()])
gathered = [torch.zeros_like(flat) for _ in range(world)]
import torch.distributed as td
td.all_gather(gathered, flat)
assert all(torch.equal(gathered[0], t) for t in gathered), "replicas diverged"
print("DP_PARITY_OK rank %d of %d" % (rank, world), flush=True)
cdist.shutdown()

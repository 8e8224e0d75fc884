"""BASELINE.json configs C2-C5 at full size through the reference-named entry points (synthetic TinyImagenet-shaped data:
8000 train / 2000 val images of 64x64 per task, batch 200), one short run each; prints one JSON line per config.

  C2  EWC, AlexNet      : Fisher pass over the previous task (40 batches) + 1 training epoch with the EWC penalty
  C3  MAS, VGG-11       : omega pass (40 batches) + 1 training epoch
  C4  SI,  VGG-11       : 1 training epoch with the path-integral step, then consolidation
  C5  GEM, AlexNet      : task 1 memory fill (256 exemplars), then 20 observe() steps of task 2 (1 memory pass + QP each)
"""
import json, os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
from clsurvey_b200 import _capi
from clsurvey_b200.engine import get_engine
from clsurvey_b200.models import make_alexnet, make_vgg

from clsurvey_b200 import dist as cdist
_capi.lib()
cdist.init()                                  # under torchrun: data-parallel (rows of every mini-batch / whole importance batches)
dev = torch.device("cuda", torch.cuda.current_device())
_print = print


def print(*a, **k):                           # one line per config, from rank 0
    if cdist.rank() == 0:
        _print(*a, **k)


class DS(torch.utils.data.TensorDataset):
    classes = list(range(20))


def task(seed, n):
    g = torch.Generator().manual_seed(seed)
    return DS(torch.randn(n, 3, 64, 64, generator=g).to(dev), torch.randint(0, 20, (n,), generator=g).to(dev))


def dsets(t):
    return {"train": task(7 + t, 8000), "val": task(1007 + t, 2000)}


def loaders(d, bs=200):
    from clsurvey_b200.data import CachedLoader                 # the device-resident task cache the entry points use (8f-2)
    return {k: CachedLoader(v, bs, shuffle=False, device=dev) for k, v in d.items()}


def timed(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return r, time.perf_counter() - t0


def penalty_config(tag, which, model):
    if which == "ewc":
        from clsurvey_b200.methods.EWC import main_EWC as M, train_EWC as T
        acc = lambda m, d: M.accumulate_EWC_weights(None, [d], m, 200)
    else:
        from clsurvey_b200.methods.MAS import main_MAS as M, train_MAS as T
        acc = lambda m, d: M.accumulate_objective_based_weights(None, [d], m, 200, "L2", "train")
    torch.manual_seed(7)
    get_engine(model, (3, 64, 64), 200)
    model.eval()
    d1, d2 = dsets(1), dsets(2)
    model, _ = timed(lambda: acc(model, d1))                                  # warm-up (lazy inits, allocations)
    del model.reg_params
    model, t_imp = timed(lambda: acc(model, d1))
    model.reg_params["lambda"] = 400.0 if which == "ewc" else 3.0
    last = str(len(model.classifier._modules) - 1)
    model.classifier._modules[last] = nn.Linear(model.classifier._modules[last].in_features, 20)
    eng = get_engine(model)
    eng.dropout_rng = "device"
    opt = T.Weight_Regularized_SGD(model.parameters(), 0.001, momentum=0.9)
    ld = loaders(d2)
    import tempfile
    tmp = tempfile.mkdtemp()
    (_, acc1), t_ep = timed(lambda: T.train_model(model, nn.CrossEntropyLoss(), opt, 0.001, ld, {"train": 8000, "val": 2000}, True, 2, tmp + "/", ""))
    from clsurvey_b200.methods import trainers
    tr_s = trainers.LAST_RUN["train_seconds"]
    print(json.dumps({"config": tag, "importance_pass_images_per_s": 8000 / t_imp, "importance_pass_s": t_imp,
                      "train_images_per_s": trainers.LAST_RUN["train_images"] / tr_s, "two_epochs_incl_val_and_checkpoints_s": t_ep,
                      "n_gpus": cdist.world_size()}), flush=True)


def si_config():
    from clsurvey_b200.methods.SI import train_SI as T
    from clsurvey_b200.methods import trainers
    torch.manual_seed(7)
    model = make_vgg("VGG11_cl_512_512")
    get_engine(model, (3, 64, 64), 200)
    model.reg_params = T.initialize_reg_params(model)
    model.reg_params["lambda"] = 400.0
    opt = T.Elastic_SGD(model.parameters(), 0.001, momentum=0.9)
    import tempfile
    tmp = tempfile.mkdtemp()
    (_, acc), t_ep = timed(lambda: T.train_model(model, nn.CrossEntropyLoss(), opt, 0.001, loaders(dsets(1)), {"train": 8000, "val": 2000}, True, 1, tmp + "/", ""))
    tr_s = trainers.LAST_RUN["train_seconds"]
    _, t_c = timed(lambda: T.update_reg_params(model))
    print(json.dumps({"config": "C4 SI VGG-11", "train_images_per_s": trainers.LAST_RUN["train_images"] / tr_s,
                      "epochs_run": len([e for e in trainers.LAST_RUN["epochs"] if e[1] == "train"]), "consolidation_s": t_c,
                      "n_gpus": cdist.world_size()}), flush=True)


def gem_config():
    from clsurvey_b200.methods.rehearsal.model import gem as G
    torch.manual_seed(7)
    base = make_alexnet(20)
    args = types.SimpleNamespace(prev_model_path=base, n_memories=256, lr=0.001, weight_decay=0.0, memory_strength=1.0,
                                 batch_size=200, nc_per_task=[20] * 10, input_shape=(3, 64, 64), shuffle_memory=True)
    net = G.Net(0, 200, 10, args)
    net._engine().dropout_rng = "device"
    d1, d2 = task(8, 400), task(9, 4200)
    x1, y1 = d1.tensors
    for b in range(2):
        net.observe(x1[b * 200:(b + 1) * 200], 0, y1[b * 200:(b + 1) * 200], list(range(b * 200, (b + 1) * 200)), args)
    x2, y2 = d2.tensors
    net.observe(x2[:200], 1, y2[:200], list(range(200)), args)               # warm-up of the task-2 path
    viol = []
    def run():
        for b in range(1, 21):
            _, _, st = net.observe(x2[b * 200:(b + 1) * 200], 1, y2[b * 200:(b + 1) * 200], list(range(200)), args)
            viol.append(st["projected_grads"][0])
    _, t = timed(run)
    print(json.dumps({"config": "C5 GEM AlexNet n_mem=256 (task 2: 1 memory pass of 200+56 + current batch + dots/Gram/QP/projection + step)",
                      "observe_steps": 20, "s_per_observe": t / 20, "train_images_per_s": 200 * 20 / t,
                      "violations": [int(v.item()) if torch.is_tensor(v) else int(v) for v in viol]}), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["C2", "C3", "C4", "C5"]
    if "C2" in which: penalty_config("C2 EWC AlexNet", "ewc", make_alexnet(20))
    if "C3" in which: penalty_config("C3 MAS VGG-11", "mas", make_vgg("VGG11_cl_512_512"))
    if "C4" in which: si_config()
    if "C5" in which: gem_config()
    cdist.shutdown()

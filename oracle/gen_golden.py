"""Generate tests/golden/*.pt by RUNNING THE UNMODIFIED REFERENCE on CPU (TEST INFRASTRUCTURE ONLY).

    python -m oracle.gen_golden            # needs /root/reference (authoring container only)

The reference ships no tests / golden vectors, so the fixtures are made here
from the reference's own functions, imported through oracle/refshim.py:

  finetune : methods.Finetune.train_SGD.train_model                       (a1)
  ewc      : methods.EWC.main_EWC.accumulate_EWC_weights (diag_fisher) +
             methods.EWC.train_EWC.{Weight_Regularized_SGD, train_model}   (a4-a7)
  mas      : methods.MAS.main_MAS.accumulate_objective_based_weights +
             methods.MAS.train_MAS.{Weight_Regularized_SGD, train_model}   (a6, a8, a9)
  si       : methods.SI.train_SI.{initialize_reg_params, Elastic_SGD, train_model,
             update_reg_params}                                            (a10-a12)
  gem      : methods.rehearsal.model.gem.Net.observe                       (a13-a16)
  qp       : known-answer vectors for project2cone2's QP (oracle/qp.py, cross-checked with scipy)
  ragged   : finetune / EWC / MAS again with dataset sizes that leave ragged last batches (56, 41, 18 at bs 16)
  schedule : the epoch protocol over 30 epochs without improvement (lr cut at count 5, stop at > 10; SI: >= 10)
  wide     : EWC and SI again on a >= 64-channel net (data seeds searched for decision margins, see MarginMonitor)
  imm      : methods.IMM.merge.{diag_fisher, IMM_merge_models} (mode-IMM precision with sampled labels, mean / mode merge)

Inputs are synthetic (torch.Generator seeds recorded in each fixture), the seed protocol is the
reference's utils.set_random(7).  The model is a reference VGGSlim with a small extra config
entry ('tiny': 3 conv stages) so that fixtures stay < 1 MB; the code path is identical to
small_VGG9 / 11normal.
"""
import contextlib
import copy
import io
import os
import sys
import tempfile
import types

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
TINY_CFG = [8, "M", 16, "M", 16, 16, "M"]
IN_HW, NCLS, NTRAIN, NVAL, BS = 16, 5, 64, 32, 16


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def make_task(seed, n):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 3, IN_HW, IN_HW, generator=g)
    y = torch.randint(0, NCLS, (n,), generator=g)
    return x, y


def loaders_for(task_seed):
    xt, yt = make_task(task_seed, NTRAIN)
    xv, yv = make_task(task_seed + 1000, NVAL)
    mk = lambda x, y: torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=BS, shuffle=False)
    return {"train": mk(xt, yt), "val": mk(xv, yv)}, {"train": NTRAIN, "val": NVAL}, (xt, yt, xv, yv)


class RecCE(nn.Module):
    """criterion wrapper that records every batch loss (the criterion is an argument of train_model)."""

    def __init__(self):
        super().__init__()
        self.ce = nn.CrossEntropyLoss()
        self.losses = []

    def forward(self, out, y):
        l = self.ce(out, y)
        self.losses.append(float(l.item()))
        return l


def new_model(dropout=False):
    import models.VGGSlim as V
    import utilities.utils as U
    V.cfg["tiny"] = TINY_CFG
    U.set_random(7)
    return V.VGGSlim(config="tiny", num_classes=NCLS, classifier_inputdim=16 * 2 * 2,
                     classifier_dim1=32, classifier_dim2=32, dropout=dropout)


def sd(model):
    return {k: v.detach().clone() for k, v in model.state_dict().items()}


def reg_dump(model, keys=("omega", "init_val", "w")):
    out = {}
    for n, p in model.named_parameters():
        if p in model.reg_params:
            out[n] = {k: model.reg_params[p][k].detach().clone() for k in keys if k in model.reg_params[p]}
    return out


def save_dsets(path, x, y):
    ds = torch.utils.data.TensorDataset(x, y)
    ds.classes = list(range(NCLS))
    torch.save({"train": ds, "val": ds}, path)


def gen_finetune(tmp):
    import torch.optim as optim
    import methods.Finetune.train_SGD as T
    out = {}
    for tag, wd in (("wd0", 0.0), ("wd5e-4", 5e-4)):
        model = new_model()
        init = sd(model)
        loaders, sizes, data = loaders_for(11)
        crit = RecCE()
        opt = optim.SGD(model.parameters(), 0.05, momentum=0.9, weight_decay=wd)
        with quiet():
            model, best = T.train_model(model, crit, opt, 0.05, loaders, sizes, False, 3, exp_dir=tmp, resume="",
                                        save_models_mode=False)
        out[tag] = dict(init=init, final=sd(model), best_acc=float(best), losses=crit.losses, lr=0.05, wd=wd,
                        epochs=3, data=data)
    torch.save(out, os.path.join(GOLDEN, "finetune.pt"))


def _penalty_method(tmp, which):
    """EWC / MAS: importance pass on task A data, new head, penalised training on task B; then repeat (accumulate)."""
    if which == "ewc":
        import methods.EWC.main_EWC as M
        import methods.EWC.train_EWC as T
        accumulate = lambda model, path: M.accumulate_EWC_weights(None, [path], model, BS)
    else:
        import methods.MAS.main_MAS as M
        import methods.MAS.train_MAS as T
        accumulate = lambda model, path: M.accumulate_objective_based_weights(None, [path], model, BS, "L2",
                                                                               test_set="train")
    model = new_model()
    out = dict(init=sd(model), rounds=[])
    lam = 50.0 if which == "ewc" else 3.0
    for rnd, (seed_prev, seed_cur) in enumerate(((21, 22), (22, 23))):
        xp, yp = make_task(seed_prev, NTRAIN)
        dpath = os.path.join(tmp, "%s_prev_%d.pth" % (which, rnd))
        save_dsets(dpath, xp, yp)
        with quiet():
            model = accumulate(model, dpath)
        model.reg_params["lambda"] = lam
        omega_after_pass = reg_dump(model, ("omega", "init_val"))
        torch.manual_seed(100 + rnd)                       # fresh-head init drawn from the host generator
        model.classifier._modules["4"] = nn.Linear(32, NCLS)
        head = {k: v.clone() for k, v in model.classifier._modules["4"].state_dict().items()}
        loaders, sizes, data = loaders_for(seed_cur)
        crit = RecCE()
        opt = T.Weight_Regularized_SGD(model.parameters(), 0.05, momentum=0.9, weight_decay=1e-4 if rnd else 0.0)
        with quiet():
            model, best = T.train_model(model, crit, opt, 0.05, loaders, sizes, False, 2, exp_dir=tmp + "/", resume="")
        out["rounds"].append(dict(prev_data=(xp, yp), data=data, lam=lam, lr=0.05, wd=1e-4 if rnd else 0.0, epochs=2,
                                  reg_after_pass=omega_after_pass, new_head=head, final=sd(model),
                                  best_acc=float(best), losses=crit.losses))
    torch.save(out, os.path.join(GOLDEN, which + ".pt"))


def gen_si(tmp):
    import methods.SI.train_SI as T
    model = new_model()
    out = dict(init=sd(model), rounds=[])
    lam = 2.0
    for rnd, seed in enumerate((31, 32)):
        if rnd == 0:
            with quiet():
                reg = T.initialize_reg_params(model)
        else:
            torch.manual_seed(200 + rnd)
            model.classifier._modules["4"] = nn.Linear(32, NCLS)
            params = list(model.parameters())
            model.reg_params.pop(params[-1], None)
            model.reg_params.pop(params[-2], None)
            with quiet():
                reg = T.update_reg_params(model)
        reg["lambda"] = lam
        model.reg_params = reg
        head = {k: v.clone() for k, v in model.classifier._modules["4"].state_dict().items()}
        reg_before = reg_dump(model)
        loaders, sizes, data = loaders_for(seed)
        crit = RecCE()
        opt = T.Elastic_SGD(model.parameters(), 0.05, momentum=0.9, weight_decay=0.0)
        with quiet():
            model, best = T.train_model(model, crit, opt, 0.05, loaders, sizes, False, 2, exp_dir=tmp + "/", resume="")
        out["rounds"].append(dict(data=data, lam=lam, lr=0.05, epochs=2, head=head, reg_before=reg_before,
                                  reg_after=reg_dump(model), final=sd(model), best_acc=float(best),
                                  losses=crit.losses))
    torch.save(out, os.path.join(GOLDEN, "si.pt"))


def gen_gem(tmp):
    refshim.patch_gem_memory()
    import methods.rehearsal.model.gem as G
    n_tasks, n_mem, bs = 3, 24, 16
    base = new_model(dropout=True)
    init = sd(base)
    mpath = os.path.join(tmp, "gem_prev.pth")
    torch.save(base, mpath)
    args = types.SimpleNamespace(prev_model_path=mpath, cuda=False, n_memories=n_mem, lr=0.05, weight_decay=0.0,
                                 memory_strength=0.5, batch_size=bs, nc_per_task=[NCLS] * n_tasks,
                                 task_imgfolders={"train": types.SimpleNamespace(transform=None)})
    with quiet():
        net = G.Net(0, NCLS * n_tasks, n_tasks, args)
    wrapped_init = sd(net.net)
    steps = []
    store = refshim.KeyedTensorStore.store
    store.clear()
    datas = []
    torch.manual_seed(77)                                   # dropout unit masks come from the host generator
    for t in range(n_tasks):
        x, y = make_task(41 + t, 48)
        datas.append((x, y))
        for b in range(3):
            xb, yb = x[b * bs:(b + 1) * bs], y[b * bs:(b + 1) * bs]
            keys = [t * 1000 + b * bs + i for i in range(bs)]
            for k, xi in zip(keys, xb):
                store[k] = xi
            rng_state = torch.get_rng_state()
            with quiet():
                loss, corr, stats = net.observe(xb, t, yb, keys, args)
            steps.append(dict(t=t, keys=keys, loss=float(loss.item()), correct=int(corr.item()),
                              violations=int(stats["projected_grads"][0]), mem_cnt=net.mem_cnt,
                              masks={k: v.clone() for k, v in net.dropout_masks.items()},
                              rng_state=rng_state,
                              grad_col=net.grads[:, t].clone(),
                              params=torch.cat([p.data.reshape(-1) for p in net.parameters()]).clone()))
    out = dict(init=init, wrapped_init=wrapped_init, data=datas, steps=steps, n_tasks=n_tasks, n_mem=n_mem, bs=bs,
               lr=0.05, margin=0.5, memory_labels=net.memory_labels.clone(),
               exemplars={t: list(net.memory_data[t]) for t in range(n_tasks)} if net.memory_data is not None else {},
               grads=net.grads.clone(), final=sd(net.net))
    torch.save(out, os.path.join(GOLDEN, "gem.pt"))


def gen_ragged(tmp):
    """Ragged tails: a 56-image importance pass (batches 16,16,16,8 -- MAS's running average uses the
    CURRENT batch size, train_MAS.py:168-173), training on 41 images (16,16,9) and validation on 18 (16,2)."""
    import torch.optim as optim
    import methods.Finetune.train_SGD as TS
    n_prev, n_train, n_val = 56, 41, 18
    out = {}

    def ragged_loaders(seed):
        xt, yt = make_task(seed, n_train)
        xv, yv = make_task(seed + 1000, n_val)
        mk = lambda x, y: torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=BS, shuffle=False)
        return {"train": mk(xt, yt), "val": mk(xv, yv)}, {"train": n_train, "val": n_val}, (xt, yt, xv, yv)

    model = new_model()
    init = sd(model)
    loaders, sizes, data = ragged_loaders(51)
    crit = RecCE()
    opt = optim.SGD(model.parameters(), 0.05, momentum=0.9, weight_decay=5e-4)
    with quiet():
        model, best = TS.train_model(model, crit, opt, 0.05, loaders, sizes, False, 2, exp_dir=tmp, resume="",
                                     save_models_mode=False)
    out["finetune"] = dict(init=init, final=sd(model), best_acc=float(best), losses=crit.losses, lr=0.05, wd=5e-4,
                           epochs=2, data=data)
    for which in ("ewc", "mas"):
        if which == "ewc":
            import methods.EWC.main_EWC as M
            import methods.EWC.train_EWC as T
            accumulate = lambda model, path: M.accumulate_EWC_weights(None, [path], model, BS)
        else:
            import methods.MAS.main_MAS as M
            import methods.MAS.train_MAS as T
            accumulate = lambda model, path: M.accumulate_objective_based_weights(None, [path], model, BS, "L2",
                                                                                   test_set="train")
        model = new_model()
        init = sd(model)
        lam = 50.0 if which == "ewc" else 3.0
        xp, yp = make_task(61, n_prev)
        dpath = os.path.join(tmp, "%s_ragged_prev.pth" % which)
        save_dsets(dpath, xp, yp)
        with quiet():
            model = accumulate(model, dpath)
        model.reg_params["lambda"] = lam
        reg_after_pass = reg_dump(model, ("omega", "init_val"))
        torch.manual_seed(300)
        model.classifier._modules["4"] = nn.Linear(32, NCLS)
        head = {k: v.clone() for k, v in model.classifier._modules["4"].state_dict().items()}
        loaders, sizes, data = ragged_loaders(62)
        crit = RecCE()
        opt = T.Weight_Regularized_SGD(model.parameters(), 0.05, momentum=0.9, weight_decay=0.0)
        with quiet():
            model, best = T.train_model(model, crit, opt, 0.05, loaders, sizes, False, 2, exp_dir=tmp + "/", resume="")
        out[which] = dict(init=init, prev_data=(xp, yp), data=data, lam=lam, lr=0.05, wd=0.0, epochs=2,
                          reg_after_pass=reg_after_pass, new_head=head, final=sd(model), best_acc=float(best),
                          losses=crit.losses)
    torch.save(out, os.path.join(GOLDEN, "ragged.pt"))


def gen_imm(tmp):
    """IMM (groundwork for the next row): mode-IMM precision = methods.IMM.merge.diag_fisher (labels SAMPLED
    from the model's own softmax with torch.multinomial, mean NLL per batch, divisor = number of BATCHES of the phase,
    merge.py:155-186) and the mean / mode merges (merge.py:188-242) over three task models."""
    import methods.IMM.merge as MG
    models, datas = [], []
    for t in range(3):
        m = new_model()
        g = torch.Generator().manual_seed(400 + t)
        with torch.no_grad():
            for p in m.parameters():                       # three different "trained" models of the same architecture
                p.add_(torch.randn(p.shape, generator=g) * 0.05)
        models.append(m)
        datas.append((make_task(410 + t, 40), make_task(420 + t, 24)))
    head_names = ["classifier.4.weight", "classifier.4.bias"]
    precisions, seeds = [], []
    for t, m in enumerate(models):
        (xt, yt), (xv, yv) = datas[t]
        mk = lambda x, y: torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=BS, shuffle=False)
        m.params = {n: p for n, p in m.named_parameters() if p.requires_grad}
        torch.manual_seed(430 + t)                         # the multinomial draws come from the global CPU generator
        seeds.append(430 + t)
        with quiet():
            prec = MG.diag_fisher(m, {"train": mk(xt, yt), "val": mk(xv, yv)}, exclude_params=head_names)
        precisions.append({n: v.detach().clone() for n, v in prec.items()})
        del m.params
    sums = [precisions[0]]
    for t in range(1, 3):
        sums.append({n: sums[-1][n] + precisions[t][n] for n in precisions[t]})
    merged_mean, merged_mode = [], []
    for idx in (1, 2):
        with quiet():
            mm = MG.IMM_merge_models(models, idx, head_names, mean_mode=True)
            md = MG.IMM_merge_models(models, idx, head_names, precision=precisions, sum_precision=sums[idx], mean_mode=False)
        merged_mean.append(sd(mm))
        merged_mode.append(sd(md))
    torch.save(dict(states=[sd(m) for m in models], data=datas, seeds=seeds, head_names=head_names, precisions=precisions,
                    merged_mean=merged_mean, merged_mode=merged_mode), os.path.join(GOLDEN, "imm.pt"))


def gen_schedule(tmp):
    """The epoch protocol over a LONG run (A.1): with lr = 1e-9 nothing improves after the first validation, so
    val_beat_counts climbs 1, 2, ...; lr is cut at count == 5 and training stops at count > 10 (Finetune / EWC / MAS)
    resp. count >= 10 (SI, whose epoch range is also num_epochs + 1).  Recorded: criterion calls (= batches processed,
    i.e. how many epochs actually ran), the optimiser's final lr, best accuracy."""
    import torch.optim as optim
    import methods.Finetune.train_SGD as TS
    import methods.EWC.train_EWC as TE
    import methods.SI.train_SI as TI
    out = {}
    for which in ("sgd", "ewc", "si", "sgd_short"):
        model = new_model()
        init = sd(model)
        loaders, sizes, data = loaders_for(71)
        crit = RecCE()
        lr0, epochs = 1e-9, (8 if which == "sgd_short" else 30)
        with quiet():
            if which.startswith("sgd"):
                opt = optim.SGD(model.parameters(), lr0, momentum=0.9, weight_decay=0.0)
                model, best = TS.train_model(model, crit, opt, lr0, loaders, sizes, False, epochs, exp_dir=tmp, resume="",
                                             save_models_mode=False)
            elif which == "ewc":
                model.reg_params = {p: dict(omega=torch.ones_like(p), init_val=p.data.clone()) for p in model.parameters()}
                model.reg_params["lambda"] = 1.0
                opt = TE.Weight_Regularized_SGD(model.parameters(), lr0, momentum=0.9, weight_decay=0.0)
                model, best = TE.train_model(model, crit, opt, lr0, loaders, sizes, False, epochs, exp_dir=tmp + "/", resume="")
            else:
                reg = TI.initialize_reg_params(model)
                reg["lambda"] = 1.0
                model.reg_params = reg
                opt = TI.Elastic_SGD(model.parameters(), lr0, momentum=0.9, weight_decay=0.0)
                model, best = TI.train_model(model, crit, opt, lr0, loaders, sizes, False, epochs, exp_dir=tmp + "/", resume="")
        out["init"], out["data"] = init, data                 # identical for all runs (same seeds)
        out[which] = dict(lr=lr0, epochs=epochs, n_criterion_calls=len(crit.losses),
                          final_lr=float(opt.param_groups[0]["lr"]), best_acc=float(best))
    # divergence: EWC / MAS / SI abort the run when the epoch loss exceeds 1e4 or is NaN (train_EWC.py:204-205,
    # train_SI.py:242-244); Finetune has no such check (train_SGD.py) and keeps going
    for which in ("diverge_ewc", "diverge_si", "diverge_sgd"):
        model = new_model()
        loaders, sizes, data = loaders_for(71)
        crit = RecCE()
        lr0, epochs = 1e3, 4
        with quiet():
            if which == "diverge_sgd":
                opt = optim.SGD(model.parameters(), lr0, momentum=0.9, weight_decay=0.0)
                model, best = TS.train_model(model, crit, opt, lr0, loaders, sizes, False, epochs, exp_dir=tmp, resume="",
                                             save_models_mode=False)
            elif which == "diverge_ewc":
                model.reg_params = {p: dict(omega=torch.ones_like(p), init_val=p.data.clone()) for p in model.parameters()}
                model.reg_params["lambda"] = 1.0
                opt = TE.Weight_Regularized_SGD(model.parameters(), lr0, momentum=0.9, weight_decay=0.0)
                model, best = TE.train_model(model, crit, opt, lr0, loaders, sizes, False, epochs, exp_dir=tmp + "/", resume="")
            else:
                reg = TI.initialize_reg_params(model)
                reg["lambda"] = 1.0
                model.reg_params = reg
                opt = TI.Elastic_SGD(model.parameters(), lr0, momentum=0.9, weight_decay=0.0)
                model, best = TI.train_model(model, crit, opt, lr0, loaders, sizes, False, epochs, exp_dir=tmp + "/", resume="")
        out[which] = dict(lr=lr0, epochs=epochs, n_criterion_calls=len(crit.losses), best_acc=float(best),
                          final_lr=float(opt.param_groups[0]["lr"]))
    torch.save(out, os.path.join(GOLDEN, "schedule.pt"))


# ---------------------------------------------------------------------------------------------- wide (>= 64 channels)
WIDE_CFG = [64, "M", 64, 128, "M"]          # on 8x8 inputs: conv 3->64 @8x8 | pool | 64->64, 64->128 @4x4 | pool -> 128*2*2
WIDE_HW, WIDE_BS = 8, 2


class MarginMonitor:
    """Smallest relative distance of any ReLU pre-activation / max-pool runner-up from its decision boundary over every
    forward pass of a run (forward hooks on the reference model).  Gradients of a ReLU / max-pool net are discontinuous at
    these boundaries: a fixture is only a meaningful 1e-4 parity target when no unit sits within the arithmetic error of
    the implementations under test (~1e-5 of the layer's scale), so the generator searches data seeds for that."""

    def __init__(self, model):
        self.min_rel = float("inf")
        mods = list(model.features) + list(model.classifier)
        for i, m in enumerate(mods):
            nxt = mods[i + 1] if i + 1 < len(mods) else None
            if isinstance(m, (nn.Conv2d, nn.Linear)) and isinstance(nxt, nn.ReLU):
                m.register_forward_hook(self._relu)
            elif isinstance(m, nn.MaxPool2d):
                m.register_forward_hook(self._pool)

    def _relu(self, mod, inp, out):
        z = out.detach().abs()
        self.min_rel = min(self.min_rel, float(z.min() / z.max()))

    def _pool(self, mod, inp, out):
        x = inp[0].detach()
        win = x.unfold(2, 2, 2).unfold(3, 2, 2).reshape(x.shape[0], x.shape[1], -1, 4)
        top = win.topk(2, dim=-1).values
        live = top[..., 0] > 0
        if live.any():
            self.min_rel = min(self.min_rel, float((top[..., 0] - top[..., 1])[live].min() / x.max()))


def new_wide_model():
    import models.VGGSlim as V
    import utilities.utils as U
    V.cfg["wide"] = WIDE_CFG
    U.set_random(7)
    return V.VGGSlim(config="wide", num_classes=NCLS, classifier_inputdim=128 * 2 * 2, classifier_dim1=32, classifier_dim2=32)


def _wide_task(seed, n):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, 3, WIDE_HW, WIDE_HW, generator=g), torch.randint(0, NCLS, (n,), generator=g)


def _wide_loaders(seed):
    xt, yt = _wide_task(seed, 4)
    xv, yv = _wide_task(seed + 1000, 2)
    mk = lambda x, y: torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=WIDE_BS, shuffle=False)
    return {"train": mk(xt, yt), "val": mk(xv, yv)}, {"train": 4, "val": 2}, (xt, yt, xv, yv)


def _wide_ewc(tmp, seed):
    import methods.EWC.main_EWC as M
    import methods.EWC.train_EWC as T
    model = new_wide_model()
    mon = MarginMonitor(model)
    init = sd(model)
    xp, yp = _wide_task(seed, 4)
    dpath = os.path.join(tmp, "wide_prev.pth")
    save_dsets(dpath, xp, yp)
    with quiet():
        model = M.accumulate_EWC_weights(None, [dpath], model, WIDE_BS)
    model.reg_params["lambda"] = 50.0
    reg_after_pass = reg_dump(model, ("omega", "init_val"))
    torch.manual_seed(500)
    model.classifier._modules["4"] = nn.Linear(32, NCLS)
    head = {k: v.clone() for k, v in model.classifier._modules["4"].state_dict().items()}
    loaders, sizes, data = _wide_loaders(seed + 1)
    crit = RecCE()
    opt = T.Weight_Regularized_SGD(model.parameters(), 0.05, momentum=0.9, weight_decay=1e-4)
    with quiet():
        model, best = T.train_model(model, crit, opt, 0.05, loaders, sizes, False, 1, exp_dir=tmp + "/", resume="")
    return mon.min_rel, dict(init=init, prev_data=(xp, yp), data=data, lam=50.0, lr=0.05, wd=1e-4, epochs=1, seed=seed,
                             reg_after_pass=reg_after_pass, new_head=head, final=sd(model), best_acc=float(best),
                             losses=crit.losses)


def _wide_si(tmp, seed):
    import methods.SI.train_SI as T
    model = new_wide_model()
    mon = MarginMonitor(model)
    init = sd(model)
    with quiet():
        reg = T.initialize_reg_params(model)
    reg["lambda"] = 2.0
    model.reg_params = reg
    with torch.no_grad():                                  # a non-trivial omega / theta* so that the penalty is active
        g = torch.Generator().manual_seed(seed + 7)
        for p in model.parameters():
            reg[p]["omega"] = torch.rand(p.shape, generator=g)
            reg[p]["init_val"] = p.data + 0.01 * torch.randn(p.shape, generator=g)
    reg_before = reg_dump(model, ("omega", "init_val"))    # (w starts at zero)
    loaders, sizes, data = _wide_loaders(seed + 1)
    crit = RecCE()
    opt = T.Elastic_SGD(model.parameters(), 0.05, momentum=0.9, weight_decay=0.0)
    with quiet():
        model, best = T.train_model(model, crit, opt, 0.05, loaders, sizes, False, 1, exp_dir=tmp + "/", resume="")
    return mon.min_rel, dict(init=init, data=data, lam=2.0, lr=0.05, epochs=1, seed=seed, reg_before=reg_before,
                             reg_after=reg_dump(model, ("w",)), final=sd(model), best_acc=float(best), losses=crit.losses)


# data seeds found by `python -m oracle.gen_golden wide_search <ewc|si> <first seed> <count>` (several ranges in parallel; the
# smallest decision margin of the whole run is printed per improvement): the fixture is then generated from these
WIDE_SEEDS = {"ewc": 9855, "si": 29239}


def wide_search(tmp, name, start, count, want=2e-5):
    fn = {"ewc": _wide_ewc, "si": _wide_si}[name]
    best = 0.0
    for seed in range(start, start + count):
        rel, _ = fn(tmp, seed)
        if rel > best:
            best = rel
            print("wide/%s: seed %d, smallest decision margin %.2e" % (name, seed, rel), flush=True)
        if rel >= want:
            break


def gen_wide(tmp):
    """EWC (Fisher pass + penalised training) and SI (path integral) on a >= 64-channel VGGSlim, so that the fixture
    exercises the layers the tensor-core kernels take; data seeds searched for decision margins (MarginMonitor)."""
    out = {}
    for name, fn in (("ewc", _wide_ewc), ("si", _wide_si)):
        rel, fx = fn(tmp, WIDE_SEEDS[name])
        out[name] = fx
        out[name]["min_margin"] = rel
        print("wide/%s: seed %d, smallest decision margin %.2e" % (name, WIDE_SEEDS[name], rel))
    torch.save(out, os.path.join(GOLDEN, "wide.pt"))


def gen_qp():
    import scipy.optimize as so
    from oracle import qp
    rng = np.random.default_rng(7)
    cases = []
    for trial in range(24):
        k = int(rng.integers(1, 10))
        M = rng.standard_normal((k, 64))
        M[:, :4] *= 4.0
        if trial % 4 == 0 and k > 1:
            M[1] = 0.9 * M[0] + 0.1 * M[1]                  # strongly correlated memories
        g = rng.standard_normal(64)
        margin = [0.0, 0.5, 1.0][trial % 3]
        P = M @ M.T
        P = 0.5 * (P + P.T) + 1e-3 * np.eye(k)
        q = -(M @ g)
        v = qp.solve_lower_bounded_qp(P, q, np.full(k, margin))
        f = lambda z: 0.5 * z @ P @ z - q @ z
        r = so.minimize(f, np.full(k, margin + 1.0), jac=lambda z: P @ z - q, bounds=[(margin, None)] * k,
                        method="L-BFGS-B", options=dict(ftol=1e-15, gtol=1e-12, maxiter=20000))
        assert np.abs(r.x - v).max() <= 1e-6 * max(1.0, np.abs(v).max()), (trial, r.x, v)
        cases.append(dict(M=torch.from_numpy(M), g=torch.from_numpy(g), margin=margin, v=torch.from_numpy(v)))
    torch.save(cases, os.path.join(GOLDEN, "qp.pt"))


def main():
    refshim.install()
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(1)                                # fixed reduction order for the fixtures
    if len(sys.argv) > 1 and sys.argv[1] == "wide_search":
        with tempfile.TemporaryDirectory() as tmp:
            wide_search(tmp, sys.argv[2], int(sys.argv[3]), int(sys.argv[4]))
        return
    if len(sys.argv) > 1 and sys.argv[1] in ("ragged", "imm", "schedule", "wide"):   # add one fixture without touching the others
        with tempfile.TemporaryDirectory() as tmp:
            {"ragged": gen_ragged, "imm": gen_imm, "schedule": gen_schedule, "wide": gen_wide}[sys.argv[1]](tmp)
        print(sys.argv[1] + ".pt", os.path.getsize(os.path.join(GOLDEN, sys.argv[1] + ".pt")))
        return
    with tempfile.TemporaryDirectory() as tmp:
        gen_wide(tmp)
        gen_ragged(tmp)
        gen_imm(tmp)
        gen_schedule(tmp)
        gen_finetune(tmp)
        _penalty_method(tmp, "ewc")
        _penalty_method(tmp, "mas")
        gen_si(tmp)
        gen_gem(tmp)
    gen_qp()
    for f in sorted(os.listdir(GOLDEN)):
        print(f, os.path.getsize(os.path.join(GOLDEN, f)))


if __name__ == "__main__":
    main()

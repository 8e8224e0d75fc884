"""CPU restatement of the reference hot path (TEST INFRASTRUCTURE ONLY).

Plain torch-CPU fp32 / numpy-fp64 restatement of the arithmetic of Mattdl/CLsurvey's
per-task training loop and importance computations.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / --impl reference legs may import
this file; nothing under `clsurvey_b200/` does.  The product path never routes through it.

Pinning: every function here is checked in `tests/test_oracle_golden.py` against fixtures
under `tests/golden/` that were produced by running the UNMODIFIED reference functions
(imported through `oracle/refshim.py`) with `oracle/gen_golden.py`.  The fp32 layer
arithmetic (conv / linear / pooling / autograd) is torch-2.11-CPU standing in for the
reference's pinned torch 1.6 -- the reference's own L0.

Each function cites the reference file:line it restates (paths relative to
/root/reference/src).
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import qp as _qp

MOMENTUM = 0.9  # main_EWC.py:62, main_MAS.py:90, main_SI.py:86, main_SGD.py:74, gem.py:153


# ----------------------------------------------------------------------------------------------
# forward / loss / backward (a2, a3): torch autograd on CPU is the reference's own L0
# ----------------------------------------------------------------------------------------------
def forward(model, x):
    """torchvision VGG/AlexNet forward: features -> avgpool -> flatten -> classifier
    (models/VGGSlim.py:43-76 sets avgpool=Identity)."""
    return model(x)


def grads_of(model, loss):
    for p in model.parameters():
        p.grad = None
    loss.backward()
    return [p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p) for p in model.parameters()]


def loss_mean_ce(logits, y):
    """nn.CrossEntropyLoss() (main_EWC.py:55, main_MAS.py:83, main_SI.py:61, main_SGD.py:47, gem.py:119)."""
    return F.cross_entropy(logits, y)


def loss_sum_nll(logits, y):
    """nll_loss(log_softmax(out), y, size_average=False) (main_EWC.py:148)."""
    return F.nll_loss(F.log_softmax(logits, dim=1), y, reduction="sum")


def loss_sum_sq(logits):
    """MSELoss(size_average=False)(out, zeros) (train_MAS.py:552-560)."""
    return (logits ** 2).sum()


def num_correct(logits, y):
    """torch.max(outputs,1); sum(preds==labels) (train_EWC.py:182,197)."""
    return int((logits.argmax(dim=1) == y).sum().item())


# ----------------------------------------------------------------------------------------------
# optimiser steps (a1 plain SGD, a6 penalised SGD, a10 SI)
# ----------------------------------------------------------------------------------------------
def penalised_sgd_step(theta, g, omega, theta_star, buf, lam, lr, wd=0.0, mom=MOMENTUM):
    """One tensor of Weight_Regularized_SGD.step (EWC/train_EWC.py:46-84, MAS/train_MAS.py:45-93).

    omega/theta_star None  => parameter not in reg_params (plain SGD-momentum, new head).
    buf None               => first step (buf = d.clone()).
    Returns (theta_new, buf_new).  `g` is consumed like the reference consumes p.grad (in place add).
    """
    d = g.clone()
    if omega is not None:
        weight_dif = theta - theta_star                # curr.add(-1, init_val)          :62
        d = d + weight_dif * (2 * lam * omega)         # weight_dif.mul(2*lambda*omega)  :64-65
    if wd != 0:
        d = d + wd * theta                             # :70-71
    if mom != 0:
        buf = d.clone() if buf is None else buf * mom + d   # :74-78
        d = buf
    return theta - lr * d, buf                         # :84


def si_step(theta, g, omega, theta_star, buf, w, lam, lr, wd=0.0, mom=MOMENTUM):
    """One tensor of Elastic_SGD.step (SI/train_SI.py:48-125). Returns (theta_new, buf_new, w_new)."""
    g0 = g.clone()                                     # unreg_dp                         :56
    theta_old = theta.clone()                          # curr_wegiht_val                  :63
    d = g + (theta - theta_star) * (2 * lam * omega)   # :69-73
    if wd != 0:
        d = d + wd * theta                             # :81-82
    if mom != 0:
        buf = d.clone() if buf is None else buf * mom + d   # :86-91
        d = buf
    theta_new = theta - lr * d                         # :97
    w_new = w + (-1.0) * ((theta_new - theta_old) * g0)  # :98-121
    return theta_new, buf, w_new


def si_consolidate(omega, w, theta, theta_star, slak=1e-3):
    """update_reg_params (SI/train_SI.py:390-417). Returns (omega_new, w_new=0, theta_star_new=theta)."""
    path_diff = theta - theta_star
    this_omega = w / (path_diff.pow(2) + slak)
    this_omega = torch.max(this_omega, torch.zeros_like(this_omega))
    return omega + this_omega, torch.zeros_like(w), theta.clone()


# ----------------------------------------------------------------------------------------------
# importance passes (a4 EWC Fisher, a8 MAS omega) -- return per-parameter lists
# ----------------------------------------------------------------------------------------------
def fisher_pass(model, batches, data_len):
    """diag_fisher (EWC/main_EWC.py:138-157): eval mode; per batch L = SUM NLL; omega += grad**2 / data_len."""
    model.eval()
    omega = [torch.zeros_like(p) for p in model.parameters()]
    for x, y in batches:
        g = grads_of(model, loss_sum_nll(model(x), y))
        for o, gi in zip(omega, g):
            o += gi ** 2 / data_len
    return omega


def mas_pass(model, batches):
    """compute_importance_l2 + Objective_After_SGD.step (MAS/train_MAS.py:508-567, 138-181):
    omega = (omega * (b*n_b) + |grad|) / ((b+1)*n_b), n_b = CURRENT batch size."""
    model.eval()
    omega = [torch.zeros_like(p) for p in model.parameters()]
    for b, (x, y) in enumerate(batches):
        n_b = y.size(0)
        g = grads_of(model, loss_sum_sq(model(x)))
        prev_size, curr_size = b * n_b, (b + 1) * n_b
        omega = [(o * prev_size + gi.abs()) / curr_size for o, gi in zip(omega, g)]
    return omega


def imm_precision_pass(model, phase_batches, exclude_names):
    """mode-IMM precision, methods/IMM/merge.py:155-186: starts at 1e-8; eval mode; for every phase and
    batch: targets ~ multinomial(softmax(out)) drawn from the GLOBAL torch generator, L = MEAN NLL of the sampled targets,
    precision += grad**2 / (number of BATCHES of that phase).  phase_batches: {phase: [(x, y), ...]} in the reference's
    dict order.  Returns {name: tensor} without the head parameters."""
    model.eval()
    names = [n for n, _ in model.named_parameters()]
    prec = {n: torch.zeros_like(p) + 1e-8 for n, p in model.named_parameters() if n not in exclude_names}
    for phase, batches in phase_batches.items():
        # the reference iterates a DataLoader per phase; creating its iterator draws one int64 (the workers' base seed)
        # from the global generator, which shifts the multinomial stream that follows
        torch.empty((), dtype=torch.int64).random_()
        for x, _ in batches:
            out = model(x)
            targets = torch.multinomial(torch.softmax(out, dim=1).detach(), 1).squeeze()
            loss = torch.nn.functional.nll_loss(torch.log_softmax(out, dim=1), targets, reduction="mean")
            g = grads_of(model, loss)
            for n, gi in zip(names, g):
                if n in prec:
                    prec[n] += gi ** 2 / len(batches)
    return prec


def imm_merge(states, upto, head_names, precisions=None, sum_precision=None, as_reference=True):
    """IMM_merge_models (merge.py:188-242): the state of task `upto` with every non-head parameter replaced by
    sum_k precision_k / sum_precision * theta_k (mode-IMM) or -- as intended -- by the mean over tasks 0..upto (mean-IMM).

    Reference quirk (pinned by tests/golden/imm.pt): in mean mode the loop re-binds `param_value` to a state_dict tensor
    of the last merged-in model (merge.py:225-226), so the final `param_value.data = mean_param.clone()` (merge.py:239)
    lands on that temporary and the returned model is an UNCHANGED copy of task `upto`'s model.  as_reference=True
    reproduces that; as_reference=False computes the intended mean."""
    merged = {k: v.clone() for k, v in states[upto].items()}
    if precisions is None and as_reference:
        return merged
    for name in states[upto]:
        if name in head_names:
            continue
        acc = torch.zeros_like(states[upto][name])
        for k in range(upto + 1):
            if precisions is None:
                acc = acc + states[k][name]
            else:
                acc += (precisions[k][name] / sum_precision[name]) * states[k][name]
        merged[name] = acc / (upto + 1) if precisions is None else acc
    return merged


def accumulate_protocol(prev_omega, new_omega):
    """store_prev / accumelate_reg_params (EWC/main_EWC.py:177-232, MAS/train_MAS.py:710-795): omega = prev + new."""
    return [a + b for a, b in zip(prev_omega, new_omega)]


# ----------------------------------------------------------------------------------------------
# epoch protocol (A.1): set_lr + best-val bookkeeping
# ----------------------------------------------------------------------------------------------
def set_lr(lr, count, stop_ge=False):
    """set_lr (train_EWC.py:89-101, train_SGD.py:10-30, train_MAS.py:183-195; SI stops at >= 10: train_SI.py:129-141)."""
    cont = not (count >= 10 if stop_ge else count > 10)
    if count == 5:
        lr = lr * 0.1
    return lr, cont


class Trainer:
    """Restates train_model of Finetune / EWC / MAS / SI (train_SGD.py:41-189, train_EWC.py:111-234,
    train_MAS.py:208-335, train_SI.py:152-283) on an nn.Module with explicit per-parameter state.

    kind: 'sgd' | 'penalty' (EWC, MAS) | 'si'
    reg:  list (aligned with model.parameters()) of None or dict(omega=, init_val=[, w=])
    """

    def __init__(self, model, kind, lr, reg=None, lam=0.0, wd=0.0):
        self.model, self.kind, self.lr, self.lam, self.wd = model, kind, lr, lam, wd
        self.params = list(model.parameters())
        self.reg = reg if reg is not None else [None] * len(self.params)
        self.bufs = [None] * len(self.params)
        self.batch_losses = []

    def step(self, x, y):
        logits = self.model(x)
        loss = loss_mean_ce(logits, y)
        correct = num_correct(logits, y)
        g = grads_of(self.model, loss)
        with torch.no_grad():
            for i, p in enumerate(self.params):
                r = self.reg[i]
                if self.kind == "si":
                    t, b, w = si_step(p.data, g[i], r["omega"], r["init_val"], self.bufs[i], r["w"],
                                      self.lam, self.lr, self.wd)
                    r["w"] = w
                elif self.kind == "penalty" and r is not None:
                    t, b = penalised_sgd_step(p.data, g[i], r["omega"], r["init_val"], self.bufs[i],
                                              self.lam, self.lr, self.wd)
                else:
                    t, b = penalised_sgd_step(p.data, g[i], None, None, self.bufs[i], 0.0, self.lr, self.wd)
                p.data.copy_(t)
                self.bufs[i] = b
        return float(loss.item()), correct

    def evaluate(self, x, y):
        with torch.no_grad():
            logits = self.model(x)
            return float(loss_mean_ce(logits, y).item()), num_correct(logits, y)

    def train_model(self, loaders, sizes, num_epochs):
        """Epoch loop incl. the SI quirks: range(start, num_epochs + 1) (train_SI.py:182), stop at count >= 10."""
        si = self.kind == "si"
        best_acc, beat, lr = 0.0, 0, self.lr
        best_state = None
        log = []
        for epoch in range(0, num_epochs + 1 if si else num_epochs):
            for phase in ("train", "val"):
                if phase == "train":
                    lr, cont = set_lr(lr, beat, stop_ge=si)
                    self.lr = lr
                    if not cont:
                        return best_acc, log, best_state
                    self.model.train(True)
                else:
                    self.model.train(False)
                run_loss, run_corr = 0.0, 0
                for x, y in loaders[phase]:
                    l, c = self.step(x, y) if phase == "train" else self.evaluate(x, y)
                    if phase == "train":
                        self.batch_losses.append(l)
                    run_loss += l
                    run_corr += c
                ep_loss, ep_acc = run_loss / sizes[phase], run_corr / sizes[phase]
                log.append((epoch, phase, ep_loss, ep_acc))
                if self.kind in ("penalty", "si") and (ep_loss > 1e4 or math.isnan(ep_loss)):
                    return best_acc, log, best_state          # train_EWC.py:204-205, train_MAS.py:298-300, train_SI.py:242-244
                if phase == "val":
                    if ep_acc > best_acc:
                        best_acc, beat = ep_acc, 0
                        best_state = {k: v.clone() for k, v in self.model.state_dict().items()}
                    else:
                        beat += 1
        return best_acc, log, best_state


# ----------------------------------------------------------------------------------------------
# GEM (a13-a16)
# ----------------------------------------------------------------------------------------------
def compute_offsets(task_idx, cum_nc):
    """common.py:106-118."""
    return (0 if task_idx == 0 else int(cum_nc[task_idx - 1])), int(cum_nc[task_idx])


class GemOracle:
    """Restates gem.Net.observe / fill_buffer / store_grad / project2cone2 (rehearsal/model/gem.py:20-80, 206-345).

    `net` has .features and .classifier; forward bypasses avgpool exactly like gem.py:174-175 unless
    `use_avgpool` (documented deviation for AlexNet @ 64x64).
    Exemplars are stored as integer keys; `fetch(keys) -> tensor` serves them (deviation 8c-(2)).
    Memory mini-batches are taken in stored order (shuffle=False deviation).
    """

    def __init__(self, net, n_tasks, n_memories, nc_per_task, lr, margin, batch_size, fetch,
                 wd=0.0, use_avgpool=False):
        self.net, self.n_tasks, self.n_mem = net, n_tasks, n_memories
        self.cum_nc = [sum(nc_per_task[:i + 1]) for i in range(len(nc_per_task))]
        self.n_outputs = self.cum_nc[-1]
        self.lr, self.margin, self.bs, self.fetch, self.wd = lr, margin, batch_size, fetch, wd
        self.use_avgpool = use_avgpool
        self.params = list(net.parameters())
        self.P = sum(p.numel() for p in self.params)
        self.grads = torch.zeros(self.P, n_tasks)                     # gem.py:131-132  [P, n_tasks]
        self.memory_labels = torch.zeros(n_tasks, n_memories, dtype=torch.long)
        self.exemplars = {t: [None] * n_memories for t in range(n_tasks)}
        self.observed_tasks, self.old_task, self.mem_cnt = [], -1, 0
        self.bufs = [None] * len(self.params)
        self.dropout_masks = {}

    # gem.py:168-204
    def forward(self, x, t):
        feat = self.net.features(x)
        if self.use_avgpool:
            feat = self.net.avgpool(feat)
        out = feat.view(feat.size(0), -1)
        for idx, m in enumerate(self.net.classifier.children()):
            if isinstance(m, nn.Dropout):
                if m.training:
                    if idx not in self.dropout_masks:
                        self.dropout_masks[idx] = torch.bernoulli(torch.full_like(out[0], 0.5)) / 0.5
                    out = out * self.dropout_masks[idx].expand(out.shape[0], -1)
            else:
                out = m(out)
        return out

    # gem.py:322-345
    def fill_buffer(self, t, keys, y):
        bsz = y.size(0)
        endcnt = min(self.mem_cnt + bsz, self.n_mem)
        eff = endcnt - self.mem_cnt
        self.exemplars[t][self.mem_cnt:endcnt] = list(keys[:eff])
        self.memory_labels[t, self.mem_cnt:endcnt] = y[:eff]
        self.mem_cnt += eff
        if self.mem_cnt == self.n_mem:
            self.mem_cnt = 0
            return True
        return False

    def _flat_grad(self):
        return torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in self.params])

    def _zero_grad(self):
        for p in self.params:
            p.grad = None

    # gem.py:206-287
    def observe(self, x, t, y, keys, masks=None):
        """masks: optional {classifier child idx: unit mask} -- host-drawn masks handed in; when None
        they are drawn here from the torch CPU generator."""
        self.net.train()
        self.dropout_masks = {} if masks is None else dict(masks)
        stats = {"violations": 0, "dotp": None, "v": None}
        if t != self.old_task:
            self.observed_tasks.append(t)
            self.old_task = t
        self.fill_buffer(t, keys, y)
        if len(self.observed_tasks) > 1:
            for tt in range(len(self.observed_tasks) - 1):
                self._zero_grad()
                past = self.observed_tasks[tt]
                o1, o2 = compute_offsets(past, self.cum_nc)
                mem_keys = self.exemplars[past]
                mem_y = self.memory_labels[past]
                for s in range(0, self.n_mem, self.bs):          # DataLoader(batch_size=args.batch_size)
                    xb = self.fetch(mem_keys[s:s + self.bs])
                    yb = mem_y[s:s + self.bs]
                    out = self.forward(xb, past)[:, o1:o2]
                    F.cross_entropy(out, yb).backward()          # grads ACCUMULATE over memory mini-batches
                self.grads[:, past] = self._flat_grad()
        self._zero_grad()
        o1, o2 = compute_offsets(t, self.cum_nc)
        out = self.forward(x, t)[:, o1:o2]
        correct = num_correct(out, y)
        loss = F.cross_entropy(out, y)
        loss.backward()
        g = self._flat_grad()
        if len(self.observed_tasks) > 1:
            self.grads[:, t] = g
            prev = torch.tensor(self.observed_tasks[:-1], dtype=torch.long)
            mem = self.grads.index_select(1, prev)               # [P, k]
            dotp = torch.mm(g.unsqueeze(0), mem)                 # gem.py:275-276
            viol = int((dotp < 0).sum().item())
            stats["dotp"] = dotp.reshape(-1).clone()
            stats["violations"] = viol
            if viol != 0:
                xproj, v = _qp.project2cone2(g.numpy(), mem.t().contiguous().numpy(), self.margin)
                g = torch.from_numpy(xproj)
                self.grads[:, t] = g
                stats["v"] = v
        # plain SGD momentum .9 (gem.py:153,285)
        with torch.no_grad():
            off = 0
            for i, p in enumerate(self.params):
                gi = g[off:off + p.numel()].view_as(p)
                off += p.numel()
                tnew, self.bufs[i] = penalised_sgd_step(p.data, gi, None, None, self.bufs[i], 0.0, self.lr, self.wd)
                p.data.copy_(tnew)
        return float(loss.item()), correct, stats

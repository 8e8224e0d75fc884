"""Reference-named optimisers whose `step` is ONE fused CUDA launch over the engine's flat buffers.

Mirrors (same constructor / step signatures, same state_dict layout via torch.optim.SGD):
  Weight_Regularized_SGD   src/methods/EWC/train_EWC.py:12-86, src/methods/MAS/train_MAS.py:19-95
  Elastic_SGD              src/methods/SI/train_SI.py:20-126
  Objective_After_SGD      src/methods/MAS/train_MAS.py:128-181  (MAS omega running mean; never updates theta)
  SGD                      torch.optim.SGD(momentum=0.9) as used by src/methods/Finetune/main_SGD.py:74, gem.py:153

The reference issues ~10-16 elementwise kernels per parameter tensor per step; here the whole model is one launch
(28 B/param for the penalised step, 36 B/param for SI -- SURVEY.md 8d).
"""
import torch
import torch.optim as optim

from .. import dist as cdist
from .._capi import call
from ..engine import _ptr, _stream, engine_of


class _EngineSGD(optim.SGD):
    def __init__(self, params, lr=0.001, momentum=0, dampening=0, weight_decay=0, nesterov=False):
        params = list(params)
        super().__init__(params, lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay,
                         nesterov=nesterov)
        if dampening != 0 or nesterov:
            raise NotImplementedError("dampening / nesterov are not used on the reference hot path")
        self.engine = engine_of(params)
        eng = self.engine
        idx = {id(p): i for i, p in enumerate(eng.params)}
        mine = sorted(idx[id(p)] for g in self.param_groups for p in g["params"])
        assert mine == list(range(mine[0], mine[-1] + 1)), "optimised parameters must be contiguous in the model"
        self.lo = eng.offsets[mine[0]]
        self.hi = eng.offsets[mine[-1]] + (eng.numels[mine[-1]] + 3) // 4 * 4
        self.first_step = True                   # a fresh optimiser has no momentum buffers (buf = d.clone())
        self.steps = 0

    def _hyper(self):
        g = self.param_groups[0]
        return float(g["lr"]), float(g["momentum"]), float(g["weight_decay"])

    def _publish_momentum(self):
        eng = self.engine
        for g in self.param_groups:
            for p in g["params"]:
                i = next(j for j, q in enumerate(eng.params) if q is p)
                self.state[p]["momentum_buffer"] = eng.view(eng.momentum, i)

    def load_state_dict(self, sd):
        super().load_state_dict(sd)
        eng = self.engine
        loaded = False
        for g in self.param_groups:
            for p in g["params"]:
                st = self.state.get(p, {})
                if "momentum_buffer" in st and st["momentum_buffer"] is not None:
                    i = next(j for j, q in enumerate(eng.params) if q is p)
                    eng.ensure("momentum")
                    eng.view(eng.momentum, i).copy_(st["momentum_buffer"])
                    loaded = True
        self.first_step = not loaded
        if loaded:                               # state entries become views of the live engine buffer again, so that a later
            self._publish_momentum()             # state_dict() (second resume) saves the current momentum, not this snapshot

    def _plain_or_penalised(self, n_pen, two_lambda):
        eng = self.engine
        lr, mom, wd = self._hyper()
        eng.ensure("momentum")
        lo, n = self.lo, self.hi - self.lo
        n_pen = max(0, min(n_pen - lo, n))
        call("clb_sgd_penalty_step", _ptr(eng.theta[lo:]), _ptr(eng.grad[lo:]),
             _ptr(eng.omega[lo:]) if n_pen else 0, _ptr(eng.theta_star[lo:]) if n_pen else 0,
             _ptr(eng.momentum[lo:]), n, n_pen, float(two_lambda), lr, mom, wd, 1.0, int(self.first_step), _stream())
        if self.first_step and mom != 0:
            self._publish_momentum()
        self.first_step = False
        self.steps += 1
        eng.n_launch += 1


class SGD(_EngineSGD):
    """optim.SGD(params, lr, momentum=0.9, weight_decay=wd) -- Finetune (main_SGD.py:74) and GEM (gem.py:153)."""

    @torch.no_grad()
    def step(self, closure=None, reduce=True):
        if reduce:
            cdist.allreduce_grads(self.engine)
        self._plain_or_penalised(0, 0.0)


class Weight_Regularized_SGD(_EngineSGD):
    """EWC / MAS penalised SGD.  `step(reg_params)` like the reference."""

    def __init__(self, params, lr=0.001, momentum=0, dampening=0, weight_decay=0, nesterov=False, orth_reg=False,
                 L1_decay=False):
        super().__init__(params, lr, momentum, dampening, weight_decay, nesterov)
        if orth_reg or L1_decay:
            raise NotImplementedError("orth_reg / L1_decay are off on the reference hot path (method.py:737-750)")

    @torch.no_grad()
    def step(self, reg_params, closure=None):
        cdist.allreduce_grads(self.engine)
        n_pen = sync_reg_params(self.engine, reg_params, need_w=False)
        self._plain_or_penalised(n_pen, 2 * reg_params.get("lambda"))


class Elastic_SGD(_EngineSGD):
    """SI: penalised step + online path integral w (train_SI.py:48-125)."""

    @torch.no_grad()
    def step(self, reg_params, closure=None):
        eng = self.engine
        cdist.allreduce_grads(eng)
        n_pen = sync_reg_params(eng, reg_params, need_w=True)
        assert n_pen >= eng.total and self.lo == 0 and self.hi == eng.total, \
            "Elastic_SGD needs every parameter in reg_params (train_SI.py:57-62)"
        lr, mom, wd = self._hyper()
        eng.ensure("momentum")
        call("clb_si_step", _ptr(eng.theta), _ptr(eng.grad), _ptr(eng.omega), _ptr(eng.theta_star),
             _ptr(eng.momentum), _ptr(eng.w), eng.total, float(2 * reg_params.get("lambda")), lr, mom, wd, 1.0,
             int(self.first_step), _stream())
        if self.first_step and mom != 0:
            self._publish_momentum()
        self.first_step = False
        self.steps += 1
        eng.n_launch += 1


class Objective_After_SGD(_EngineSGD):
    """MAS importance accumulation: omega = (omega*prev + |grad|)/curr (train_MAS.py:138-181); theta untouched."""

    @torch.no_grad()
    def step(self, reg_params, batch_index, batch_size, closure=None):
        eng = self.engine
        sync_reg_params(eng, reg_params, need_w=False)
        prev_size = batch_index * batch_size
        curr_size = (batch_index + 1) * batch_size
        call("clb_mas_accum", _ptr(eng.omega), _ptr(eng.grad), float(prev_size), float(curr_size), eng.total, _stream())
        eng.n_launch += 1


def sync_reg_params(eng, reg_params, need_w):
    """Make the reference-format dict {Parameter: {'omega','init_val'[,'w']}} and the flat buffers one and the same
    storage (dict entries become views of eng.omega / eng.theta_star / eng.w).  Returns n_penalised (flat prefix)."""
    first = eng.omega is None
    eng.ensure("omega")
    eng.ensure("theta_star")
    if need_w:
        eng.ensure("w")
    in_reg = []
    for i, p in enumerate(eng.params):
        rp = reg_params.get(p)
        in_reg.append(rp is not None)
        if rp is None:
            continue
        for key, flat in (("omega", eng.omega), ("init_val", eng.theta_star), ("w", eng.w if need_w else None)):
            if flat is None or key not in rp:
                continue
            v = eng.view(flat, i)
            t = rp[key]
            if t.data_ptr() != v.data_ptr():
                v.copy_(t.to(eng.device, torch.float32))
                rp[key] = v
    # penalised parameters must form a prefix of the flat buffer (the fresh head is always last: main_EWC.py:49-53);
    # otherwise fall back to omega = 0 on the unpenalised tensors, which is arithmetically identical (d + x*0 = d).
    n_reg = sum(in_reg)
    if all(in_reg[:n_reg]):
        if n_reg == len(in_reg):
            return eng.total
        return eng.offsets[n_reg]
    for i, flag in enumerate(in_reg):
        if not flag:
            eng.view(eng.omega, i).zero_()
            eng.view(eng.theta_star, i).copy_(eng.view(eng.theta, i))
    return eng.total

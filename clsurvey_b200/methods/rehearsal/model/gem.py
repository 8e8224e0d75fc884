"""Mirror of src/methods/rehearsal/model/gem.py (a13-a16): GEM's Net / observe on the engine.

Same public members as the reference Net (observe, observe_FT, forward, fill_buffer, manage_memory, init_setup,
compute_offsets, memory_labels, memory_data, mem_cnt, observed_tasks, grads ...).  Differences, all documented in
DESIGN.md: the gradient memory is task-major [n_tasks, P] on the device (reference: [P, n_tasks], gem.py:131-132);
dots, Gram, the k x k QP and the projection run on the device (reference: 4.2 GB D2H + numpy + quadprog per
violation, gem.py:58-80); exemplar pixels are cached on the device; `avgpool` is applied when the wrapped net has a
real one (the reference forward bypasses it and crashes for AlexNet at 64x64, gem.py:174-175).
"""
import copy
import ctypes

import torch
import torch.nn as nn

from .... import dist as cdist
from ...._capi import call
from ....engine import LOSS_MEAN_CE, _ptr, _stream, get_engine
from ... import common as mcommon
from ...optim import SGD
from . import common

QP_EPS = 1e-3      # project2cone2(..., eps=1e-3)   gem.py:58


class Net(nn.Module):
    def __init__(self, n_inputs, n_outputs, n_tasks, args):
        super(Net, self).__init__()
        self.net = mcommon.load_model(args.prev_model_path) if isinstance(args.prev_model_path, str) \
            else args.prev_model_path
        last = str(len(self.net.classifier._modules) - 1)
        num_ftrs = self.net.classifier._modules[last].in_features
        original_head = copy.deepcopy(self.net.classifier._modules[last])
        self.net.classifier._modules[last] = nn.Linear(num_ftrs, n_outputs)          # gem.py:99-103
        self.n_outputs = n_outputs
        o = original_head.out_features
        with torch.no_grad():                                                          # gem.py:107-112
            self.net.classifier._modules[last].weight.data[:o].copy_(original_head.weight.data)
            self.net.classifier._modules[last].bias.data[:o].copy_(original_head.bias.data)
        self.n_memories = args.n_memories
        self.gpu = True
        self.memory_data = None
        self.memory_labels = torch.zeros(n_tasks, self.n_memories, dtype=torch.long)   # gem.py:125 (host, bit-exact)
        self.n_tasks = n_tasks
        self.observed_tasks = []
        self.old_task = -1
        self.mem_cnt = 0
        self.cum_nc_per_task = [sum(args.nc_per_task[:idx + 1]) for idx, nc in enumerate(args.nc_per_task)]
        self.shuffle_memory = getattr(args, "shuffle_memory", True)
        self._input_shape = tuple(getattr(args, "input_shape", (3, 64, 64)))
        self._max_batch = max(int(args.batch_size), 1)
        self.init_setup(args)

    # ---- engine plumbing (not pickled) -------------------------------------------------------
    def _engine(self):
        use_avg = isinstance(getattr(self.net, "avgpool", None), nn.AdaptiveAvgPool2d)
        eng = get_engine(self.net, self._input_shape, self._max_batch, use_avgpool=use_avg)
        if getattr(self, "_grads_for", None) is not eng:
            object.__setattr__(self, "_grads_for", eng)
            old = getattr(self, "grads", None)
            g = torch.zeros(self.n_tasks, eng.total, device=eng.device)                # task-major gradient memory
            if old is not None and tuple(old.shape) == tuple(g.shape):
                g.copy_(old)
            object.__setattr__(self, "grads", g)
            object.__setattr__(self, "_dots", torch.zeros(16, dtype=torch.float64, device=eng.device))
            object.__setattr__(self, "_gram", torch.zeros(256, dtype=torch.float64, device=eng.device))
            object.__setattr__(self, "_v", torch.zeros(16, dtype=torch.float64, device=eng.device))
            object.__setattr__(self, "_viol", torch.zeros(1, dtype=torch.int32, device=eng.device))
        return eng

    def __getstate__(self):
        st = self.__dict__.copy()
        for k in ("_grads_for", "_dots", "_gram", "_v", "_viol", "opt"):
            st.pop(k, None)
        return st

    def init_setup(self, args):
        """gem.py:146-155: fresh SGD(momentum .9) + margin; also re-run after unpickling (main_rehearsal.py:221)."""
        self.dropout_masks = {}
        self._lr, self._wd = args.lr, args.weight_decay
        self.margin = args.memory_strength
        self._max_batch = max(int(args.batch_size), 1)
        eng = self._engine()
        self.opt = SGD(self.net.parameters(), args.lr, weight_decay=args.weight_decay, momentum=0.9)
        return eng

    def compute_offsets(self, task_idx, cum_nc_per_task):
        return common.compute_offsets(task_idx, cum_nc_per_task)

    def reset_dropout_config(self):
        self.dropout_masks = {}

    def _unit_masks(self, eng, p_retain_unit=0.5):
        """gem.py:179-196: one mask per *unit*, shared by the whole batch and by every pass of this observe() call;
        drawn from the host generator."""
        masks = {}
        for op in eng.ops:
            if op["kind"] == "dropout" and op["module"].training:
                idx = op["cls_idx"]
                forced = getattr(self, "forced_masks", None)       # host-drawn masks handed in (parity tests)
                if forced is not None and idx in forced:
                    self.dropout_masks[idx] = forced[idx].to(eng.device, torch.float32)
                if idx not in self.dropout_masks:
                    self.dropout_masks[idx] = (torch.bernoulli(torch.full((op["feat"],), p_retain_unit))
                                               / p_retain_unit).to(eng.device)
                masks[idx] = self.dropout_masks[idx]
        return masks

    def forward(self, x, t, args=None, train_mode=False, p_retain_unit=0.5):
        """gem.py:168-204.  Returns logits with columns outside task t's slice set to -10e10."""
        eng = self._engine()
        x = x if x.is_cuda else x.to(eng.device)
        out = eng.forward(x, train=self.net.training, masks=self._unit_masks(eng, p_retain_unit)).clone()
        o1, o2 = self.compute_offsets(t, self.cum_nc_per_task)
        if o1 > 0:
            out[:, :o1].fill_(-10e10)
        if o2 < self.n_outputs:
            out[:, o2:self.n_outputs].fill_(-10e10)
        return out

    # ---- observe ------------------------------------------------------------------------------
    def _memory_batches(self, eng, past_task, batch_size):
        n = self.n_memories
        order = torch.randperm(n) if self.shuffle_memory else torch.arange(n)           # DataLoader(shuffle=True)
        xs = self.memory_data.pixels[past_task]
        ys = self.memory_labels[past_task]
        for s in range(0, n, batch_size):
            idx = order[s:s + batch_size]
            yield xs[idx.to(xs.device)], ys[idx]

    def observe(self, x, t, y, paths, args=None):
        """gem.py:206-287."""
        eng = self._engine()
        self.net.train()
        self.reset_dropout_config()
        if t != self.old_task:
            self.init_new_task(t, x)
        x = x if x.is_cuda else x.to(eng.device)
        self.fill_buffer(t, paths, y, x)
        masks = self._unit_masks(eng)
        s = _stream()
        bs = int(args.batch_size) if args is not None else self._max_batch
        if len(self.observed_tasks) > 1:
            for tt in range(len(self.observed_tasks) - 1):
                past = self.observed_tasks[tt]
                o1, o2 = self.compute_offsets(past, self.cum_nc_per_task)
                eng.zero_grad()
                for xb, yb in self._memory_batches(eng, past, bs):
                    eng.fwd_loss_bwd(xb, yb, LOSS_MEAN_CE, train=True, masks=masks, col_off=o1, ncols=o2 - o1,
                                     accumulate=True)
                self.grads[past].copy_(eng.grad)                                        # store_grad (gem.py:20-36)
        o1, o2 = self.compute_offsets(t, self.cum_nc_per_task)
        B = y.size(0)
        lo, hi = cdist.shard_rows(B)
        if hi > lo:
            eng.fwd_loss_bwd(x[lo:hi], y[lo:hi], LOSS_MEAN_CE, denom=B, train=True, masks=masks, col_off=o1, ncols=o2 - o1)
        else:                                           # B < world: this rank's shard is empty -> zero gradient, same collectives
            eng.backward_skip()
            eng.loss_dev.zero_()
            eng.correct_dev.zero_()
        loss, correct = eng.loss_dev.clone(), eng.correct_dev.clone()
        cdist.allreduce_grads(eng)
        viol = None
        if len(self.observed_tasks) > 1:
            k = len(self.observed_tasks) - 1
            idx = torch.tensor(self.observed_tasks[:-1], dtype=torch.int32, device=eng.device)
            self._dots.zero_()
            self._gram.zero_()
            call("clb_gem_dots_gram", _ptr(eng.grad), _ptr(self.grads), eng.total, eng.total, _ptr(idx), k,
                 _ptr(self._dots), _ptr(self._gram), s)
            call("clb_gem_solve_qp", _ptr(self._dots), _ptr(self._gram), k, float(self.margin), QP_EPS, _ptr(self._v),
                 _ptr(self._viol), s)
            call("clb_gem_project", _ptr(eng.grad), _ptr(self.grads), eng.total, eng.total, _ptr(idx), k,
                 _ptr(self._v), _ptr(self._viol), s)                                    # overwrite_grad (gem.py:39-55)
            self.grads[t].copy_(eng.grad)
            viol = self._viol.clone()
            eng.n_launch += 3
        self.opt.step(reduce=False)                     # gradient was all-reduced before the projection
        batch_stats = {'projected_grads': [viol if viol is not None else 0]}
        return loss, correct, batch_stats

    def observe_FT(self, x, t, y, paths=None, args=0):
        """gem.py:289-310: plain fine-tuning step on task t's head slice."""
        eng = self._engine()
        x = x if x.is_cuda else x.to(eng.device)
        o1, o2 = self.compute_offsets(t, self.cum_nc_per_task)
        B = y.size(0)
        lo, hi = cdist.shard_rows(B)
        if hi > lo:
            eng.fwd_loss_bwd(x[lo:hi], y[lo:hi], LOSS_MEAN_CE, denom=B, train=self.net.training,
                             masks=self._unit_masks(eng), col_off=o1, ncols=o2 - o1)
        else:
            eng.backward_skip()
            eng.loss_dev.zero_()
            eng.correct_dev.zero_()
        loss, correct = eng.loss_dev.clone(), eng.correct_dev.clone()
        self.opt.step()
        return loss, correct

    def init_new_task(self, t, batch):
        """gem.py:312-320 (mem_cnt is NOT reset -- reference quirk kept)."""
        self.observed_tasks.append(t)
        self.old_task = t
        if self.memory_data is None:
            self.memory_data = common.RehearsalMemory(self.n_tasks, self.n_memories, batch.shape[1:],
                                                      device=self._engine().device)

    def fill_buffer(self, t, paths, y, x=None):
        """gem.py:322-345: ring buffer of exemplar keys + labels (host, bit-exact) and pixels (device cache)."""
        buffer_cycle = False
        bsz = y.size(0)
        endcnt = min(self.mem_cnt + bsz, self.n_memories)
        effbsz = endcnt - self.mem_cnt
        self.memory_data.exemplars[t][self.mem_cnt:endcnt] = list(paths[:effbsz])
        if bsz == 1:
            self.memory_labels[t, self.mem_cnt] = y[0].cpu()
        else:
            self.memory_labels[t, self.mem_cnt:endcnt].copy_(y[:effbsz].cpu())
        if x is not None and self.memory_data.pixels is not None and effbsz > 0:
            self.memory_data.pixels[t, self.mem_cnt:endcnt].copy_(x[:effbsz])
        self.mem_cnt += effbsz
        if self.mem_cnt == self.n_memories:
            self.mem_cnt = 0
            buffer_cycle = True
        return buffer_cycle

    def manage_memory(self, t, args):
        """gem.py:347-373: fill the first task's buffer until the ring wraps."""
        for data in args.dset_loaders['train']:
            inputs, labels, paths = data
            if t != self.old_task:
                self.init_new_task(t, inputs)
            dev = self._engine().device
            if self.fill_buffer(t, paths, labels, inputs.to(dev)):
                return
        print("[WARNING] BUFFER WAS NOT FILLED WITH EXEMPLARS...")

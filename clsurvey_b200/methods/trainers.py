"""The per-task training loop of the reference, re-hosted on the engine (a1, a7, a12).

One implementation, four reference flavours with their quirks kept:
  'sgd' : src/methods/Finetune/train_SGD.py:41-189  (squeeze; optimizer.step(); no divergence exit; save_models_mode)
  'ewc' : src/methods/EWC/train_EWC.py:111-234      (squeeze; step(model.reg_params); exit if epoch_loss > 1e4 / NaN)
  'mas' : src/methods/MAS/train_MAS.py:208-335      (as 'ewc')
  'si'  : src/methods/SI/train_SI.py:152-283        (no squeeze; range(start, num_epochs + 1); stop at count >= 10)

What changed underneath: `model(inputs)`, the criterion and `loss.backward()` are engine calls (hand-written CUDA),
`optimizer.step` is one fused launch, and the per-batch `.item()` host syncs of the reference (train_EWC.py:196-197)
become one device->host read per phase (per-batch values are kept in a device log, summed on the host in the same
order and precision as the reference's python-float `running_loss`).
"""
import math
import os
import time

import torch

from .. import _capi
from .. import dist as cdist
from ..engine import LOSS_MEAN_CE, engine_of

LAST_RUN = {}     # diagnostics of the most recent train_model call (per-batch losses etc.); not part of the API

_builtin_print = print


def print(*a, **k):           # the reference's progress lines, once per job: data-parallel ranks > 0 stay silent
    if cdist.rank() == 0:
        _builtin_print(*a, **k)


class _CheckpointWriter:
    """best_model.pth.tar / epoch.pth.tar (train_EWC.py:207-227) written by a background thread.

    The reference pickles the whole module inline: ~0.15 s for VGG-11 with its reg_params (170 MB) -- as long as 50 training
    steps of this engine, on every validation improvement.  Here the caller takes a device-side snapshot (copy.deepcopy of the
    module: a few D2D copies, microseconds of stream time) and the thread does the D2H copies, pickling and file I/O while the
    next steps run.  The thread works on its own CUDA stream: torch.save copies every storage device->host on the calling
    thread's current stream, and on the default stream each of those ~50 copies would queue behind all the training steps the
    main thread has already enqueued (measured: 1.28 s for four checkpoints instead of 0.49 s).  Semantics kept: a newer
    snapshot of the same file waits for the older one, and train_model does not return before every file is on disk.
    CLB_ASYNC_SAVE=0 restores inline saves."""

    def __init__(self):
        import collections
        import threading
        self.enabled = os.environ.get("CLB_ASYNC_SAVE", "1") != "0"
        self.pending = collections.OrderedDict()      # path -> snapshot not yet picked up (a newer one replaces it)
        self.lock = threading.Condition()
        self.busy = False
        self.thread = None
        self.error = None
        self.stream = None

    def _worker(self):
        while True:
            with self.lock:
                if not self.pending:
                    self.busy = False
                    self.lock.notify_all()
                    return
                path, (snap, ready) = self.pending.popitem(last=False)
            try:
                t0 = time.time()
                if self.stream is not None:
                    with torch.cuda.stream(self.stream):
                        self.stream.wait_event(ready)  # the snapshot's D2D copies, queued on the trainer's stream
                        _to_host_inplace(snap)
                t1 = time.time()
                torch.save(snap, path)
                if os.environ.get("CLB_SAVE_TRACE"):
                    import sys
                    sys.stderr.write("[ckpt] %s: to host %.3f s, torch.save %.3f s\n" % (os.path.basename(path), t1 - t0, time.time() - t1))
            except Exception as e:                    # surfaced by wait()
                self.error = e
            del snap

    def submit(self, obj, path):
        import copy
        import threading
        if not self.enabled:
            torch.save(obj, path)
            return
        snap = copy.deepcopy(obj)                     # device-side snapshot, ordered on the current stream
        ready = None
        if torch.cuda.is_available():
            if self.stream is None:
                self.stream = _capi.private_stream("checkpoint")
            ready = torch.cuda.Event()
            ready.record()
        with self.lock:
            self.pending[path] = (snap, ready)
            self.pending.move_to_end(path)
            if not self.busy:
                self.busy = True
                self.thread = threading.Thread(target=self._worker, daemon=True)
                self.thread.start()

    def wait(self):
        with self.lock:
            while self.busy:
                self.lock.wait()
        if self.error is not None:
            e, self.error = self.error, None
            raise e


def _to_host_inplace(obj, _seen=None):
    """Move every CUDA tensor reachable from a checkpoint object (module / state_dict / optimizer state / reg_params, nested in
    dicts, lists and tuples) to the host, keeping object identities: Parameters stay the keys of `reg_params`."""
    seen = set() if _seen is None else _seen
    if id(obj) in seen:
        return
    seen.add(id(obj))
    if isinstance(obj, torch.Tensor):
        if obj.is_cuda:
            obj.data = obj.data.cpu()
        if obj.grad is not None and obj.grad.is_cuda:
            obj.grad.data = obj.grad.data.cpu()
    elif isinstance(obj, torch.nn.Module):
        for t in list(obj.parameters()) + list(obj.buffers()):
            _to_host_inplace(t, seen)
        for k, v in vars(obj).items():
            if k not in ("_parameters", "_buffers", "_modules"):
                _to_host_inplace(v, seen)
        for m in obj.children():
            _to_host_inplace(m, seen)
    elif isinstance(obj, dict):
        for k, v in obj.items():
            _to_host_inplace(k, seen)
            _to_host_inplace(v, seen)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            _to_host_inplace(v, seen)


def set_lr(optimizer, lr, count, stop_ge=False):
    """Early stop (count > 10; SI: >= 10) / decay x0.1 at count == 5 (train_SGD.py:10-30, train_SI.py:129-141)."""
    continue_training = True
    if (count >= 10) if stop_ge else (count > 10):
        continue_training = False
        print("training terminated")
    if count == 5:
        lr = lr * 0.1
        print("lr is set to {}".format(lr))
        for param_group in optimizer.param_groups:
            param_group["lr"] = lr
    return optimizer, lr, continue_training


def save_cuda_mem_req(out_dir, out_filename="cuda_mem_req.pth.tar"):
    """Same side file as utils.save_cuda_mem_req (src/utilities/utils.py:85-97)."""
    out_dir = os.path.dirname(out_dir)
    if not out_dir or not os.path.isdir(out_dir):
        return
    torch.save({"cuda_memory_allocated": torch.cuda.memory_allocated(), "cuda_memory_cached": torch.cuda.memory_reserved()},
               os.path.join(out_dir, out_filename))


class _StagedBatches:
    """Iterate a loader ONE batch ahead: a host batch's rows of this rank are copied host -> device on a side stream into one
    of two staging buffers while the previous step still computes (the reference copies inline, `inputs.cuda()`,
    train_EWC.py:172-173).  Device-resident batches (task cache) pass through.  Yields (B, lo, hi, x_dev, y_dev); the
    consumer calls done() after launching the step that reads the batch."""

    def __init__(self, loader, eng, squeeze, world, rk):
        self.it, self.eng, self.squeeze, self.world, self.rk = iter(loader), eng, squeeze, world, rk
        self.side = _capi.private_stream("h2d") if eng.device.type == "cuda" else None
        self.buf = [None, None]
        self.free_ev = [None, None]                  # recorded on the main stream when the step reading buffer k is launched
        self.k = 0
        self.cur_k = None
        self.pending = self._stage()

    def _stage(self):
        try:
            data = next(self.it)
        except StopIteration:
            return None
        inputs, labels = data[0], data[1]
        if self.squeeze:
            inputs = inputs.squeeze()
            if inputs.dim() == 3:                    # batch of one survives the reference's squeeze()
                inputs = inputs.unsqueeze(0)
        B = labels.size(0)
        lo, hi = cdist.shard_rows(B, self.world, self.rk)
        xs, ys = inputs[lo:hi], labels[lo:hi]
        if xs.is_cuda or self.side is None or hi == lo:
            return (B, lo, hi, xs, ys if ys.is_cuda or self.side is None else ys.to(self.eng.device), None, None)
        k = self.k
        self.k ^= 1
        if self.buf[k] is None or self.buf[k][0].shape[0] < hi - lo or self.buf[k][0].shape[1:] != xs.shape[1:]:
            self.buf[k] = (torch.empty((max(hi - lo, self.eng.max_batch),) + tuple(xs.shape[1:]), dtype=torch.float32, device=self.eng.device),
                           torch.empty(max(hi - lo, self.eng.max_batch), dtype=torch.int64, device=self.eng.device))
            self.free_ev[k] = None
        with torch.cuda.stream(self.side):
            if self.free_ev[k] is not None:
                self.side.wait_event(self.free_ev[k])             # the step that read this buffer two batches ago is done
            xd, yd = self.buf[k][0][:hi - lo], self.buf[k][1][:hi - lo]
            xd.copy_(xs, non_blocking=True)
            yd.copy_(ys, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.side)
        return (B, lo, hi, xd, yd, ready, k)

    def __iter__(self):
        return self

    def __next__(self):
        cur = self.pending
        if cur is None:
            raise StopIteration
        B, lo, hi, xd, yd, ready, k = cur
        if ready is not None:
            torch.cuda.current_stream().wait_event(ready)
        self.cur_k = k
        self.pending = self._stage()                 # next batch starts moving while this one is consumed
        return B, lo, hi, xd, yd

    def done(self):
        if self.cur_k is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self.free_ev[self.cur_k] = ev


def _to_device(t, device):
    return t if t.is_cuda else t.to(device, non_blocking=True)


def run_train_model(flavour, model, criterion, optimizer, lr, dset_loaders, dset_sizes, use_gpu, num_epochs,
                    exp_dir="./", resume="", saving_freq=5, save_models_mode=True):
    writer = _CheckpointWriter()
    try:
        return _run_train_model(writer, flavour, model, criterion, optimizer, lr, dset_loaders, dset_sizes, use_gpu, num_epochs,
                                exp_dir, resume, saving_freq, save_models_mode)
    finally:
        writer.wait()                                 # every checkpoint file is complete when train_model returns


def _run_train_model(writer, flavour, model, criterion, optimizer, lr, dset_loaders, dset_sizes, use_gpu, num_epochs,
                     exp_dir="./", resume="", saving_freq=5, save_models_mode=True):
    assert flavour in ("sgd", "ewc", "mas", "si", "l2t")        # l2t: IMM L2-transfer = step(reg_params), no divergence exit
    eng = engine_of(model.parameters())
    si = flavour == "si"
    since = time.time()
    val_beat_counts, best_acc, mem_snapshotted = 0, 0.0, False
    if resume and os.path.isfile(resume):
        checkpoint = torch.load(resume, weights_only=False)
        start_epoch = checkpoint["epoch"]
        best_acc = checkpoint["best_acc"]
        model.load_state_dict(checkpoint["state_dict"])
        optimizer.load_state_dict(checkpoint["optimizer"])
        lr = checkpoint["lr"]
        val_beat_counts = checkpoint["val_beat_counts"]
        print("=> loaded checkpoint '{}' (epoch {})".format(resume, checkpoint["epoch"]))
    else:
        start_epoch = 0
    log = dict(batch_losses=[], epochs=[], train_images=0, train_seconds=0.0)
    LAST_RUN.clear()
    LAST_RUN.update(log)
    world, rk = cdist.world_size(), cdist.rank()
    if world > 1:
        cdist.sync_replicas(eng)                     # replicas start bit-identical (fresh heads are host-drawn per rank)
    epoch_acc = 0.0
    # whole-step CUDA graphs (fwd + loss + bwd + gradient all-reduces + fused update = one graph launch).  Off for
    # models with active Dropout (host-drawn masks change every step).
    has_dropout = any(op["kind"] == "dropout" for op in eng.ops)
    use_graph = os.environ.get("CLB_CUDA_GRAPH", "1") == "1"
    seen = set()
    for epoch in range(start_epoch, num_epochs + 1 if si else num_epochs):
        print("Epoch {}/{}".format(epoch, num_epochs - 1))
        if world > 1:
            torch.manual_seed(cdist.shared_seed())   # same DataLoader shuffle / dropout draws on every rank
        for phase in ["train", "val"]:
            if phase == "train":
                optimizer, lr, continue_training = set_lr(optimizer, lr, count=val_beat_counts, stop_ge=si)
                if not continue_training:
                    _finish(since, best_acc)
                    return model, best_acc
                model.train(True)
            else:
                model.train(False)
            loader = dset_loaders[phase]
            nb = len(loader)
            loss_log = torch.zeros(max(nb, 1), dtype=torch.float32, device=eng.device)
            corr_log = torch.zeros(max(nb, 1), dtype=torch.int32, device=eng.device)
            t0 = time.time()
            n_img = 0
            staged = _StagedBatches(loader, eng, not si, world, rk)
            for i, (B, lo, hi, x_src, y_src) in enumerate(staged):    # this rank's rows, already on (or moving to) the device
                x = y = None
                graphed = False
                if hi > lo and phase == "train" and use_graph and not has_dropout and not optimizer.first_step:
                    hyper = tuple(sorted((k, v) for k, v in optimizer.param_groups[0].items() if k != "params"))
                    lam = model.reg_params.get("lambda") if flavour != "sgd" else None
                    key = ("train", flavour, hi - lo, B, hyper, lam)
                    if key in seen:                       # an eager step with this configuration has already run
                        def body(xs, ys, _B=B):
                            eng.fwd_loss_bwd(xs, ys, LOSS_MEAN_CE, denom=_B, train=True, dp_overlap=True)
                            optimizer.step() if flavour == "sgd" else optimizer.step(model.reg_params)
                        eng.graphed(key, hi - lo, body)(x_src, y_src)     # one copy: source -> the graph's static buffers
                        graphed = True
                    seen.add(key)
                if graphed:
                    loss_log[i:i + 1].copy_(eng.loss_dev)
                    corr_log[i:i + 1].copy_(eng.correct_dev)
                elif hi > lo:
                    x, y = _to_device(x_src, eng.device), _to_device(y_src, eng.device)
                    if phase == "train":
                        masks = None
                        if has_dropout and world > 1:         # draw for the GLOBAL batch, keep this rank's rows: every rank
                            masks = {op["cls_idx"]: eng.draw_dropout_mask(op, B)[lo:hi]       # consumes the same RNG stream
                                     for op in eng.ops if op["kind"] == "dropout"}
                        eng.fwd_loss_bwd(x, y, LOSS_MEAN_CE, denom=B, train=True, dp_overlap=True, masks=masks)
                    else:
                        eng.forward(x, train=False)
                        eng.loss_head(y, LOSS_MEAN_CE, denom=B, want_grad=False)
                    loss_log[i:i + 1].copy_(eng.loss_dev)
                    corr_log[i:i + 1].copy_(eng.correct_dev)
                elif phase == "train":
                    if has_dropout and world > 1:
                        for op in eng.ops:
                            if op["kind"] == "dropout":
                                eng.draw_dropout_mask(op, B)      # keep the host generator in step with the other ranks
                    eng.backward_skip(dp_overlap=True)
                if phase == "train" and not graphed:
                    if flavour == "sgd":
                        optimizer.step()
                    else:
                        optimizer.step(model.reg_params)
                staged.done()
                if not mem_snapshotted:
                    save_cuda_mem_req(exp_dir)
                    mem_snapshotted = True
                n_img += B
            losses = loss_log[:nb].tolist()                          # ONE device->host read per phase
            corrects = corr_log[:nb].tolist()
            if world > 1:
                losses = cdist.allreduce_scalars(losses)
                corrects = cdist.allreduce_scalars(corrects)
            running_loss = 0.0
            for v in losses:
                running_loss += v                                      # python-float sum, reference order
            running_corrects = int(sum(corrects))
            if phase == "train":
                log["batch_losses"].extend(losses)
                log["train_images"] += n_img
                log["train_seconds"] += time.time() - t0
            epoch_loss = running_loss / dset_sizes[phase]
            epoch_acc = running_corrects / dset_sizes[phase]
            log["epochs"].append((epoch, phase, epoch_loss, epoch_acc))
            LAST_RUN.update(log)
            print("{} Loss: {:.4f} Acc: {:.4f}".format(phase, epoch_loss, epoch_acc))
            if flavour not in ("sgd", "l2t") and (epoch_loss > 1e4 or math.isnan(epoch_loss)):
                return model, best_acc
            if phase == "val":
                if epoch_acc > best_acc:
                    best_acc = epoch_acc
                    if save_models_mode and rk == 0:
                        writer.submit(model, os.path.join(exp_dir, "best_model.pth.tar"))
                    val_beat_counts = 0
                else:
                    val_beat_counts += 1
        if save_models_mode and epoch % saving_freq == 0 and rk == 0:
            writer.submit({"epoch": epoch + 1, "lr": lr, "val_beat_counts": val_beat_counts, "epoch_acc": epoch_acc,
                           "best_acc": best_acc, "arch": "alexnet", "model": model, "state_dict": model.state_dict(),
                           "optimizer": optimizer.state_dict()}, os.path.join(exp_dir, "epoch.pth.tar"))
    _finish(since, best_acc)
    return model, best_acc


def _finish(since, best_acc):
    el = time.time() - since
    print("Training complete in {:.0f}m {:.0f}s".format(el // 60, el % 60))
    print("Best val Acc: {:4f}".format(best_acc))

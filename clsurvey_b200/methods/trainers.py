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

from .. import dist as cdist
from ..engine import LOSS_MEAN_CE, engine_of

LAST_RUN = {}     # diagnostics of the most recent train_model call (per-batch losses etc.); not part of the API


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


def _to_device(t, device):
    return t if t.is_cuda else t.to(device, non_blocking=True)


def run_train_model(flavour, model, criterion, optimizer, lr, dset_loaders, dset_sizes, use_gpu, num_epochs,
                    exp_dir="./", resume="", saving_freq=5, save_models_mode=True):
    assert flavour in ("sgd", "ewc", "mas", "si")
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
            for i, data in enumerate(loader):
                inputs, labels = data[0], data[1]
                if not si:
                    inputs = inputs.squeeze()
                    if inputs.dim() == 3:               # batch of one survives the reference's squeeze()
                        inputs = inputs.unsqueeze(0)
                B = labels.size(0)
                lo, hi = cdist.shard_rows(B, world, rk)
                x = _to_device(inputs[lo:hi], eng.device)
                y = _to_device(labels[lo:hi], eng.device)
                graphed = False
                if hi > lo and phase == "train" and use_graph and not has_dropout and not optimizer.first_step:
                    hyper = tuple(sorted((k, v) for k, v in optimizer.param_groups[0].items() if k != "params"))
                    lam = model.reg_params.get("lambda") if flavour != "sgd" else None
                    key = ("train", flavour, hi - lo, B, hyper, lam)
                    if key in seen:                       # an eager step with this configuration has already run
                        def body(xs, ys, _B=B):
                            eng.fwd_loss_bwd(xs, ys, LOSS_MEAN_CE, denom=_B, train=True, dp_overlap=True)
                            optimizer.step() if flavour == "sgd" else optimizer.step(model.reg_params)
                        eng.graphed(key, hi - lo, body)(x, y)
                        graphed = True
                    seen.add(key)
                if graphed:
                    loss_log[i:i + 1].copy_(eng.loss_dev)
                    corr_log[i:i + 1].copy_(eng.correct_dev)
                elif hi > lo:
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
            if flavour != "sgd" and (epoch_loss > 1e4 or math.isnan(epoch_loss)):
                return model, best_acc
            if phase == "val":
                if epoch_acc > best_acc:
                    best_acc = epoch_acc
                    if save_models_mode and rk == 0:
                        torch.save(model, os.path.join(exp_dir, "best_model.pth.tar"))
                    val_beat_counts = 0
                else:
                    val_beat_counts += 1
        if save_models_mode and epoch % saving_freq == 0 and rk == 0:
            torch.save({"epoch": epoch + 1, "lr": lr, "val_beat_counts": val_beat_counts, "epoch_acc": epoch_acc,
                        "best_acc": best_acc, "arch": "alexnet", "model": model, "state_dict": model.state_dict(),
                        "optimizer": optimizer.state_dict()}, os.path.join(exp_dir, "epoch.pth.tar"))
    _finish(since, best_acc)
    return model, best_acc


def _finish(since, best_acc):
    el = time.time() - since
    print("Training complete in {:.0f}m {:.0f}s".format(el // 60, el % 60))
    print("Best val Acc: {:4f}".format(best_acc))

"""Mirror of src/methods/rehearsal/train_rehearsal.py:57-199 (a17): GEM's epoch loop."""
import copy
import os
import time

import torch

from ..trainers import set_lr  # noqa: F401  (same thresholds: train_rehearsal.py:11-32)
from . import main_rehearsal

LAST_RUN = {}


def termination_protocol(since, best_acc, best_model, exp_dir):
    """train_rehearsal.py:35-50: the BEST model is written once, at the end."""
    el = time.time() - since
    print('Training complete in {:.0f}m {:.0f}s'.format(el // 60, el % 60))
    print('Best val Acc: {:4f}'.format(best_acc))
    if exp_dir and os.path.isdir(exp_dir):
        torch.save(best_model, os.path.join(exp_dir, 'best_model.pth.tar'))


def train_model(model, args, dset_sizes, resume='', save_models_mode=False, saving_freq=10):
    optimizer = model.opt
    exp_dir, lr, num_epochs = args.save_path, args.lr, args.n_epochs
    since = time.time()
    val_beat_counts, best_acc, best_model = 0, 0.0, None
    start_epoch = 0
    if resume and os.path.isfile(resume):
        checkpoint = torch.load(resume, weights_only=False)
        start_epoch = checkpoint['epoch']
        model.load_state_dict(checkpoint['state_dict'])
        optimizer.load_state_dict(checkpoint['optimizer'])
        best_acc, lr, val_beat_counts = checkpoint['best_acc'], checkpoint['lr'], checkpoint['val_beat_counts']
    log = dict(batch_losses=[], violations=[], epochs=[])
    LAST_RUN.clear()
    LAST_RUN.update(log)
    for epoch in range(start_epoch, num_epochs):
        for phase in ['train', 'val']:
            if phase == 'train':
                optimizer, lr, continue_training = set_lr(optimizer, lr, count=val_beat_counts)
                if not continue_training:
                    termination_protocol(since, best_acc, best_model, exp_dir)
                    return model, best_acc
                model.train(True)
            else:
                model.train(False)
            losses, corrects, viols = [], [], []
            for data in args.dset_loaders[phase]:
                inputs, labels, paths = data
                if phase == 'train':
                    if args.finetune:
                        loss, correct_classified = model.observe_FT(inputs, args.task_idx, labels, paths, args)
                    else:
                        loss, correct_classified, batch_stats = model.observe(inputs, args.task_idx, labels, paths, args)
                        viols.append(batch_stats['projected_grads'][0])
                else:
                    loss, correct_classified = main_rehearsal.eval_batch(model, inputs, labels, args)
                losses.append(loss)
                corrects.append(correct_classified)
            # one device->host read per phase (the reference syncs with .item() and torch.isnan every batch)
            losses = torch.cat([l.reshape(1).float() for l in losses]).tolist() if losses else []
            corrects = torch.cat([c.reshape(1).long() for c in corrects]).tolist() if corrects else []
            if any(l != l for l in losses):
                print("Canceling because Nan LOSS")
                return model, best_acc
            running_loss = 0.0
            for v in losses:
                running_loss += v
            epoch_loss = running_loss / dset_sizes[phase]
            epoch_acc = sum(corrects) / dset_sizes[phase]
            if phase == 'train':
                log['batch_losses'].extend(losses)
                log['violations'].extend(int(v.item()) if torch.is_tensor(v) else int(v) for v in viols)
            log['epochs'].append((epoch, phase, epoch_loss, epoch_acc))
            LAST_RUN.update(log)
            print('{} Loss: {:.4f} Acc: {:.4f}'.format(phase, epoch_loss, epoch_acc))
            if phase == 'val':
                if epoch_acc > best_acc:
                    best_acc = epoch_acc
                    if save_models_mode:
                        torch.save(model, os.path.join(exp_dir, 'best_model.pth.tar'))
                    val_beat_counts = 0
                    best_model = copy.deepcopy(model)        # train_rehearsal.py:177
                else:
                    val_beat_counts += 1
        if save_models_mode and epoch % saving_freq == 0:
            torch.save({'epoch': epoch + 1, 'lr': lr, 'val_beat_counts': val_beat_counts, 'epoch_acc': epoch_acc,
                        'best_acc': best_acc, 'arch': 'alexnet', 'model': model, 'state_dict': model.state_dict(),
                        'optimizer': optimizer.state_dict()}, os.path.join(exp_dir, 'epoch.pth.tar'))
    termination_protocol(since, best_acc, best_model, exp_dir)
    return model, best_acc

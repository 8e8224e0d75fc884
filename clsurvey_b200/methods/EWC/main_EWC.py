"""Mirror of src/methods/EWC/main_EWC.py (a4, a5): empirical-Fisher pass + omega bookkeeping + fine-tuning entry.

diag_fisher runs forward (eval), sum-NLL head, backward and the fused `omega += g*g / N` accumulator -- all CUDA
through the C ABI; under data parallelism whole batches are dealt round-robin to ranks and omega is all-reduced once.
"""
import os
import time

import torch
import torch.nn as nn

from ... import dist as cdist
from ..._capi import call
from ...engine import LOSS_SUM_NLL, _ptr, _stream, engine_of
from .. import common
from ..optim import sync_reg_params
from . import train_EWC as EWC_SGD


def fine_tune_EWC_acuumelation(dataset_path, previous_task_model_path, exp_dir, data_dir, reg_sets, reg_lambda=1,
                               num_epochs=100, lr=0.0008, batch_size=200, weight_decay=0, head_shared=False,
                               saving_freq=5):
    """main_EWC.py:14-76."""
    dsets = torch.load(dataset_path, weights_only=False) if isinstance(dataset_path, str) else dataset_path
    dset_loaders = common.make_loaders(dsets, batch_size, shuffle=True)
    dset_sizes = {x: len(dsets[x]) for x in ['train', 'val']}
    dset_classes = dsets['train'].classes
    start = time.time()
    model_ft = common.load_model(previous_task_model_path)
    common.bind(model_ft, dsets['train'], batch_size)
    model_ft = accumulate_EWC_weights(data_dir, reg_sets, model_ft, batch_size=batch_size)
    model_ft.reg_params['lambda'] = reg_lambda
    if not os.path.exists(exp_dir):
        os.makedirs(exp_dir)
    common.save_preprocessing_time(exp_dir, time.time() - start)
    if not head_shared:
        common.replace_last_classifier_layer(model_ft, len(dset_classes))
        common.bind(model_ft, dsets['train'], batch_size)
    criterion = nn.CrossEntropyLoss()
    optimizer_ft = EWC_SGD.Weight_Regularized_SGD(model_ft.parameters(), lr, momentum=0.9, weight_decay=weight_decay)
    resume = os.path.join(exp_dir, 'epoch.pth.tar')
    return EWC_SGD.train_model(model_ft, criterion, optimizer_ft, lr, dset_loaders, dset_sizes, True, num_epochs,
                               exp_dir, resume, saving_freq=saving_freq)


def _importance_loader(data_dir, reg_sets, batch_size, split="train"):
    if data_dir is not None:
        raise NotImplementedError("JPEG ImageFolder readers (data_dir != None) are outside the hot path (SURVEY.md 2.1 #9)")
    dset = None
    for data_path in reg_sets:                      # like the reference only the LAST loader is used (main_EWC.py:93-117)
        dset = torch.load(data_path, weights_only=False) if isinstance(data_path, str) else data_path
        dset = dset[split]
    loader = common.make_loader(dset, batch_size, shuffle=False)
    return dset, loader


def accumulate_EWC_weights(data_dir, reg_sets, model_ft, batch_size):
    """main_EWC.py:79-123."""
    dset, dset_loader = _importance_loader(data_dir, reg_sets, batch_size)
    if not hasattr(model_ft, 'reg_params'):
        model_ft.reg_params = initialize_reg_params(model_ft)
    model_ft.reg_params = store_prev_reg_params(model_ft)
    model_ft = diag_fisher(model_ft, dset_loader, len(dset))
    model_ft.reg_params = accumelate_reg_params(model_ft)
    return model_ft


def diag_fisher(model, dset_loader, data_len):
    """main_EWC.py:138-157: omega += (d sum-NLL / d theta)**2 / data_len per batch, eval mode."""
    eng = engine_of(model.parameters())
    reg_params = model.reg_params
    model.eval()
    sync_reg_params(eng, reg_params, need_w=False)
    world, rk = cdist.world_size(), cdist.rank()
    for b, (x, label) in enumerate(dset_loader):
        if b % world != rk:
            continue
        x = x if x.is_cuda else x.to(eng.device, non_blocking=True)
        # omega += g*g / data_len fused into this batch's backward pass (Engine.backward, importance=...)
        eng.fwd_loss_bwd(x, label, LOSS_SUM_NLL, train=False, importance=(1, float(data_len)))
    cdist.allreduce_flat(eng.omega)
    _zero_unregistered(eng, reg_params)
    return model


def _zero_unregistered(eng, reg_params):
    for i, p in enumerate(eng.params):
        if p not in reg_params:
            eng.view(eng.omega, i).zero_()


def initialize_reg_params(model, freeze_layers=None):
    """main_EWC.py:160-173: omega = 0, init_val = theta for every named parameter."""
    freeze_layers = [] if freeze_layers is None else freeze_layers
    reg_params = {}
    for name, param in model.named_parameters():
        if name not in freeze_layers:
            reg_params[param] = {'omega': torch.zeros_like(param.data), 'init_val': param.data.clone()}
    return reg_params


def store_prev_reg_params(model, freeze_layers=None):
    """main_EWC.py:177-201: prev_omega <- omega, omega <- 0, init_val <- theta."""
    freeze_layers = [] if freeze_layers is None else freeze_layers
    reg_params = model.reg_params
    for name, param in model.named_parameters():
        if name not in freeze_layers:
            if param in reg_params:
                reg_param = reg_params.get(param)
                reg_param['prev_omega'] = reg_param.get('omega').clone()
                reg_param['omega'] = torch.zeros_like(param.data)
                reg_param['init_val'] = param.data.clone()
        elif param in reg_params:
            del reg_params[param]
    return reg_params


def accumelate_reg_params(model, freeze_layers=None):
    """main_EWC.py:205-232: omega <- prev_omega + omega."""
    freeze_layers = [] if freeze_layers is None else freeze_layers
    reg_params = model.reg_params
    for name, param in model.named_parameters():
        if name not in freeze_layers:
            if param in reg_params:
                reg_param = reg_params.get(param)
                prev = reg_param.pop('prev_omega')
                om = reg_param['omega']
                prev = prev.to(om.device)
                if om.is_cuda and om.data_ptr() % 16 == 0 and prev.data_ptr() % 16 == 0:
                    call("clb_axpby", _ptr(om), _ptr(prev), _ptr(om), 1.0, om.numel(), _stream())   # omega = prev + omega
                else:
                    om.copy_(prev + om)
        elif param in reg_params:
            del reg_params[param]
    return reg_params


def sanitycheck(model):
    for name, param in model.named_parameters():
        if param in model.reg_params:
            omega = model.reg_params.get(param).get('omega')
            print(name, 'omega max', omega.max().item(), 'min', omega.min().item(), 'mean', omega.mean().item())

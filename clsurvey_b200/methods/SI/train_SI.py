"""Mirror of src/methods/SI/train_SI.py (a10-a12): Elastic_SGD, train_model, reg-param consolidation."""
import torch

from ..._capi import call
from ...engine import _ptr, _stream, engine_of
from ..optim import Elastic_SGD, sync_reg_params  # noqa: F401
from ..trainers import run_train_model


def set_lr(optimizer, lr, count):
    from ..trainers import set_lr as _s
    return _s(optimizer, lr, count, stop_ge=True)        # train_SI.py:129-141 stops at count >= 10


def train_model(model, criterion, optimizer, lr, dset_loaders, dset_sizes, use_gpu, num_epochs, exp_dir='./',
                resume='', saving_freq=5):
    """train_SI.py:152-283 incl. its range(start_epoch, num_epochs + 1).  Returns (model, best_val_acc)."""
    return run_train_model("si", model, criterion, optimizer, lr, dset_loaders, dset_sizes, use_gpu, num_epochs,
                           exp_dir, resume, saving_freq)


def initialize_reg_params(model):
    """train_SI.py:286-298: omega = 0, w = 0, init_val = theta for ALL parameters."""
    reg_params = {}
    for name, param in model.named_parameters():
        reg_params[param] = {'omega': torch.zeros_like(param.data), 'w': torch.zeros_like(param.data),
                             'init_val': param.data.clone(), 'name': name}
    return reg_params


def update_reg_params(model, slak=1e-3):
    """train_SI.py:367-430: omega += max(w / ((theta-theta*)^2 + slak), 0); w = 0; theta* = theta for known parameters;
    fresh zero entries for new ones (the new head).  One fused launch over the flat buffers."""
    eng = engine_of(model.parameters())
    reg_params = model.reg_params
    for param in model.parameters():
        if param not in reg_params:
            reg_params[param] = {'omega': torch.zeros_like(param.data), 'w': torch.zeros_like(param.data),
                                 'init_val': param.data.clone()}
    sync_reg_params(eng, reg_params, need_w=True)
    # new-head entries have w = 0 and theta* = theta: the kernel leaves their omega unchanged (max(0/slak, 0) = 0)
    call("clb_si_consolidate", _ptr(eng.omega), _ptr(eng.w), _ptr(eng.theta), _ptr(eng.theta_star), float(slak),
         eng.total, _stream())
    eng.n_launch += 1
    return reg_params

"""Mirror of src/methods/IMM/main_L2transfer.py: omega = 1, theta* = theta of the previous task model
(main_L2transfer.py:23-66), fresh head, penalised training."""
import os

import torch
import torch.nn as nn

from .. import common
from . import train_L2transfer as T


def update_reg_params(model, freeze_layers=None):
    """main_L2transfer.py:23-66: every parameter gets omega = ones, init_val = theta."""
    reg_params = model.reg_params
    freeze_layers = [] if freeze_layers is None else freeze_layers
    for name, param in model.named_parameters():
        if param in reg_params and name in freeze_layers:
            del reg_params[param]
            continue
        reg_params[param] = {'omega': torch.ones_like(param.data), 'init_val': param.data.clone()}
    return reg_params


def fine_tune_l2transfer(dataset_path, model_path, exp_dir, batch_size=100, num_epochs=100, lr=0.0004, reg_lambda=100,
                         init_freeze=0, weight_decay=0, saving_freq=5):
    """main_L2transfer.py:70-160."""
    dsets = torch.load(dataset_path, weights_only=False) if isinstance(dataset_path, str) else dataset_path
    dset_loaders = common.make_loaders(dsets, batch_size, shuffle=True)
    dset_sizes = {x: len(dsets[x]) for x in ['train', 'val']}
    dset_classes = dsets['train'].classes
    resume = os.path.join(exp_dir, 'epoch.pth.tar')
    model_ft = common.load_model(resume if os.path.isfile(resume) else model_path)
    if not init_freeze:
        common.replace_last_classifier_layer(model_ft, len(dset_classes))
    if not os.path.exists(exp_dir):
        os.makedirs(exp_dir)
    common.bind(model_ft, dsets['train'], batch_size)
    if not os.path.isfile(resume):
        if not hasattr(model_ft, 'reg_params'):
            model_ft.reg_params = {}
        params = list(model_ft.parameters())
        model_ft.reg_params.pop(params[-1], None)          # the fresh head starts unregistered ...
        model_ft.reg_params.pop(params[-2], None)
        model_ft.reg_params = update_reg_params(model_ft)  # ... and is registered like everything else (omega = 1)
        model_ft.reg_params['lambda'] = reg_lambda
    criterion = nn.CrossEntropyLoss()
    optimizer_ft = T.Weight_Regularized_SGD(model_ft.parameters(), lr, momentum=0.9, weight_decay=weight_decay)
    return T.train_model(model_ft, criterion, optimizer_ft, lr, dset_loaders, dset_sizes, True, num_epochs, exp_dir, resume,
                         saving_freq=saving_freq)

"""Mirror of src/methods/MAS/main_MAS.py:34-153."""
import os
import time

import torch
import torch.nn as nn

from .. import common
from ..EWC.main_EWC import _importance_loader
from . import train_MAS


def exp_lr_scheduler(optimizer, epoch, init_lr=0.0004, lr_decay_epoch=45):
    return optimizer            # main_MAS.py:13-23 only rescales an lr that the importance pass never uses


def fine_tune_objective_based_acuumelation(dataset_path, previous_task_model_path, init_model_path, exp_dir, data_dir,
                                           reg_sets, reg_lambda=1, norm='L2', num_epochs=100, lr=0.0008, batch_size=200,
                                           weight_decay=0, b1=True, L1_decay=False, head_shared=False, saving_freq=5):
    """main_MAS.py:34-106 (method.py:737-750 calls it with norm='L2', b1=False)."""
    dsets = torch.load(dataset_path, weights_only=False) if isinstance(dataset_path, str) else dataset_path
    dset_loaders = common.make_loaders(dsets, batch_size, shuffle=True)
    dset_sizes = {x: len(dsets[x]) for x in ['train', 'val']}
    dset_classes = dsets['train'].classes
    start = time.time()
    model_ft = common.load_model(previous_task_model_path)
    update_batch_size = 1 if b1 else batch_size
    common.bind(model_ft, dsets['train'], batch_size)
    model_ft = accumulate_objective_based_weights(data_dir, reg_sets, model_ft, update_batch_size, norm, test_set="train")
    model_ft.reg_params['lambda'] = reg_lambda
    if not os.path.exists(exp_dir):
        os.makedirs(exp_dir)
    common.save_preprocessing_time(exp_dir, time.time() - start)
    if not head_shared:
        last = str(len(model_ft.classifier._modules) - 1)
        if init_model_path is not None:
            init_model = common.load_model(init_model_path)
            model_ft.classifier._modules[last] = init_model.classifier._modules[last]
        else:
            common.replace_last_classifier_layer(model_ft, len(dset_classes))
        common.bind(model_ft, dsets['train'], batch_size)
    criterion = nn.CrossEntropyLoss()
    optimizer_ft = train_MAS.Weight_Regularized_SGD(model_ft.parameters(), lr, momentum=0.9, weight_decay=weight_decay,
                                                    L1_decay=L1_decay)
    resume = os.path.join(exp_dir, 'epoch.pth.tar')
    return train_MAS.train_model(model_ft, criterion, optimizer_ft, lr, dset_loaders, dset_sizes, True, num_epochs,
                                 exp_dir, resume, saving_freq=saving_freq)


def accumulate_objective_based_weights(data_dir, reg_sets, model_ft, batch_size, norm='L2', test_set="train"):
    """main_MAS.py:109-153."""
    if norm != 'L2':
        raise NotImplementedError("only the L2-norm objective is on the reference hot path (method.py:748)")
    _, loader = _importance_loader(data_dir, reg_sets, batch_size, split=test_set)
    if not hasattr(model_ft, 'reg_params'):
        model_ft.reg_params = train_MAS.initialize_reg_params(model_ft)
    model_ft.reg_params = train_MAS.initialize_store_reg_params(model_ft)
    optimizer_ft = train_MAS.Objective_After_SGD(model_ft.parameters(), lr=0.0001, momentum=0.9)
    model_ft = train_MAS.compute_importance_l2(model_ft, optimizer_ft, exp_lr_scheduler, [loader], True)
    model_ft.reg_params = train_MAS.accumelate_reg_params(model_ft)
    return model_ft

"""Mirror of src/methods/Finetune/main_SGD.py:13-82 (`fine_tune_SGD`), phase-1 trainer of every framework method."""
import os

import torch
import torch.nn as nn

from .. import common
from ..optim import SGD
from . import train_SGD as SGD_Training


def fine_tune_SGD(dset_dataloader, cumsum_dset_sizes, dset_classes, model_path, exp_dir, num_epochs=100, lr=0.0004,
                  freeze_mode=0, weight_decay=0, enable_resume=True, replace_last_classifier_layer=True,
                  save_models_mode=True, freq=5):
    resume = os.path.join(exp_dir, 'epoch.pth.tar') if enable_resume else ''
    if os.path.isfile(resume):
        model_ft = torch.load(resume, weights_only=False)['model']
    else:
        if not os.path.exists(exp_dir) and save_models_mode:
            os.makedirs(exp_dir)
        if not os.path.isfile(model_path):
            raise Exception("Model path non-existing: {}".format(model_path))
        model_ft = common.load_model(model_path)
    criterion = nn.CrossEntropyLoss()
    if freeze_mode or replace_last_classifier_layer:
        labels_per_task = [len(task_labels) for task_labels in dset_classes['train']]
        model_ft = common.replace_last_classifier_layer(model_ft, sum(labels_per_task))
    loader = dset_dataloader['train']
    common.bind(model_ft, loader.dataset, loader.batch_size or 1)
    if freeze_mode:
        last = str(len(model_ft.classifier._modules) - 1)
        optimizer_ft = SGD(model_ft.classifier._modules[last].parameters(), lr, momentum=0.9)
    else:
        optimizer_ft = SGD(model_ft.parameters(), lr, momentum=0.9, weight_decay=weight_decay)
    return SGD_Training.train_model(model_ft, criterion, optimizer_ft, lr, dset_dataloader, cumsum_dset_sizes, True,
                                    num_epochs, exp_dir, resume, save_models_mode=save_models_mode, saving_freq=freq)

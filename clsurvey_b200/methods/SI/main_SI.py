"""Mirror of src/methods/SI/main_SI.py:26-94 (`fine_tune_elastic`)."""
import os
import time

import torch
import torch.nn as nn

from .. import common
from . import train_SI


def fine_tune_elastic(dataset_path, model_path, exp_dir, batch_size=200, num_epochs=100, lr=0.0004, reg_lambda=100,
                      init_freeze=0, weight_decay=0, saving_freq=5):
    dsets = torch.load(dataset_path, weights_only=False) if isinstance(dataset_path, str) else dataset_path
    dset_loaders = common.make_loaders(dsets, batch_size, shuffle=True)
    dset_sizes = {x: len(dsets[x]) for x in ['train', 'val']}
    dset_classes = dsets['train'].classes
    resume = os.path.join(exp_dir, 'epoch.pth.tar')
    if os.path.isfile(resume):
        model_ft = torch.load(resume, weights_only=False)['model']
    else:
        if not os.path.isfile(model_path):
            raise FileNotFoundError("model path %s is empty (the reference would fall back to a pretrained AlexNet "
                                    "download, main_SI.py:44-48; no network here)" % model_path)
        model_ft = common.load_model(model_path)
        if not init_freeze:
            common.replace_last_classifier_layer(model_ft, len(dset_classes))
        if not os.path.exists(exp_dir):
            os.makedirs(exp_dir)
    common.bind(model_ft, dsets['train'], batch_size)
    criterion = nn.CrossEntropyLoss()
    start = time.time()
    if not os.path.isfile(resume):
        if not hasattr(model_ft, 'reg_params'):
            reg_params = train_SI.initialize_reg_params(model_ft)
        else:
            parameters = list(model_ft.parameters())
            model_ft.reg_params.pop(parameters[-1], None)
            model_ft.reg_params.pop(parameters[-2], None)
            reg_params = train_SI.update_reg_params(model_ft)
        reg_params['lambda'] = reg_lambda
        model_ft.reg_params = reg_params
    common.save_preprocessing_time(exp_dir, time.time() - start)
    optimizer_ft = train_SI.Elastic_SGD(model_ft.parameters(), lr, momentum=0.9, weight_decay=weight_decay)
    return train_SI.train_model(model_ft, criterion, optimizer_ft, lr, dset_loaders, dset_sizes, True, num_epochs,
                                exp_dir, resume, saving_freq=saving_freq)

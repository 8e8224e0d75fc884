"""Helpers shared by the per-method entry points (host side)."""
import os

import torch
import torch.nn as nn

from ..engine import get_engine


def load_model(path):
    """torch.load of a pickled whole nn.Module (the reference's model interchange format, main_EWC.py:39)."""
    m = torch.load(path, weights_only=False, map_location="cpu")
    if isinstance(m, dict):
        m = m["model"]
    return m


def make_loaders(dsets, batch_size, shuffle=True, workers=0):
    """The reference's per-task loaders (main_EWC.py:29-31: DataLoader, shuffle, 8 workers).  The un-augmented task is cached
    on the device once (clsurvey_b200/data.py) and served in the SAME order a DataLoader would draw."""
    from .. import data as cdata
    return cdata.make_loaders(dsets, batch_size, shuffle)


def make_loader(dset, batch_size, shuffle=False):
    from .. import data as cdata
    return cdata.make_loaders({"x": dset}, batch_size, shuffle, phases=("x",))["x"]


def sample_shape(dset):
    x = dset[0][0]
    x = x.squeeze()
    return tuple(x.shape)


def replace_last_classifier_layer(model, out_dim):
    """utils.replace_last_classifier_layer (src/utilities/utils.py:68-72): fresh nn.Linear head from the host RNG."""
    last = str(len(model.classifier._modules) - 1)
    num_ftrs = model.classifier._modules[last].in_features
    model.classifier._modules[last] = nn.Linear(num_ftrs, out_dim)
    return model


def bind(model, dset, batch_size, use_avgpool=True):
    return get_engine(model, sample_shape(dset), batch_size, use_avgpool=use_avgpool)


def save_preprocessing_time(out_dir, t, out_filename="preprocess_time.pth.tar"):
    """utils.save_preprocessing_time (src/utilities/utils.py:100-105)."""
    if os.path.isfile(out_dir):
        out_dir = os.path.dirname(out_dir)
    if os.path.isdir(out_dir):
        torch.save(t, os.path.join(out_dir, out_filename))

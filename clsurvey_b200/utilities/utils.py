"""Host helpers of src/utilities/utils.py that the hot path and its evaluation use (seeding, head swap, head fetch)."""
import copy
import random

import numpy
import torch


def set_random(seed=7):
    """utils.set_random (src/utilities/utils.py:52-58): one seed for python / numpy / torch."""
    random.seed(seed)
    numpy.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
        torch.backends.cudnn.deterministic = True


def replace_last_classifier_layer(model, out_dim):
    """utils.replace_last_classifier_layer (utils.py:68-72)."""
    from ..methods import common
    return common.replace_last_classifier_layer(model, out_dim)


def get_prev_heads(prev_head_model_paths, head_layer_idx):
    """utils.get_prev_heads (utils.py:235-262): the last classifier layer of every given model (a model trained up to task t
    holds the head of task t).  Returns deep copies, so a head can be swapped into another model (get_output_def) without
    touching the model it came from."""
    if not isinstance(prev_head_model_paths, list):
        prev_head_model_paths = [prev_head_model_paths]
    heads = []
    for head_model_path in prev_head_model_paths:
        m = head_model_path
        if isinstance(m, str):
            m = torch.load(m, weights_only=False, map_location="cpu")
        if isinstance(m, dict):
            m = m["model"]
        head = m.classifier._modules[head_layer_idx]
        assert isinstance(head, torch.nn.Linear), type(head)
        h = torch.nn.Linear(head.in_features, head.out_features, bias=head.bias is not None)
        with torch.no_grad():                     # a plain copy of the values: never a view into an engine's flat buffer
            h.weight.copy_(head.weight.detach().cpu())
            if head.bias is not None:
                h.bias.copy_(head.bias.detach().cpu())
        heads.append(h)
    return heads

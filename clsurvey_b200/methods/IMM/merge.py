"""Mirror of src/methods/IMM/merge.py: mode-IMM precision (a Fisher estimate with labels SAMPLED from the
model's own softmax) and the mean / mode merges of the task models.

Engine side: forward (eval) -> host draw of the targets from the device logits' softmax (torch.multinomial on the HOST
generator, like the dropout masks: the engine receives random draws, it does not regenerate them) -> mean-NLL head ->
backward with `precision += g*g / n_batches` fused into the backward pass (Engine.backward importance mode 1) -> mode
merge as one element-wise launch per tensor and model (clb_imm_merge_accum)."""
import copy

import torch

from ..._capi import call
from ...engine import LOSS_MEAN_CE, _ptr, _stream, engine_of


def diag_fisher(model, dataset, exclude_params=None):
    """merge.py:155-186.  dataset: {phase: loader}; precision starts at 1e-8, per batch L = MEAN NLL of targets drawn from
    softmax(out), precision += grad**2 / len(dataset[phase]) (= number of BATCHES).  Returns {name: tensor}."""
    exclude_params = [] if exclude_params is None else exclude_params
    eng = engine_of(model.parameters())
    eng.ensure("omega")
    saved = eng.omega.clone()                          # the engine's importance buffer is borrowed as the accumulator
    eng.omega.fill_(1e-8)
    model.eval()
    for phase in dataset.keys():
        n_batches = len(dataset[phase])
        for x, _ in dataset[phase]:
            x = x if x.is_cuda else x.to(eng.device, non_blocking=True)
            out = eng.forward(x, train=False)
            temp = torch.softmax(out.detach().cpu(), dim=1)
            targets = torch.multinomial(temp, 1).clone().squeeze(1)
            eng.loss_head(targets, LOSS_MEAN_CE, want_grad=True)
            eng.backward(importance=(1, float(n_batches)))
    precision = {n: eng.view(eng.omega, i).clone() for i, (n, _) in enumerate(model.named_parameters())
                 if n not in exclude_params}
    eng.omega.copy_(saved)
    return precision


def IMM_merge_models(models, task_list_idx, head_param_names, precision=None, sum_precision=None, mean_mode=True):
    """merge.py:188-242.  Mode-IMM: theta = sum_k precision_k / sum_precision * theta_k for every non-head parameter.
    Mean-IMM: REFERENCE QUIRK kept -- merge.py:225-226 re-binds the loop variable `param_value` to a state_dict tensor of the
    last merged-in model, so merge.py:239 assigns the mean to that temporary and the returned model is an unchanged copy of
    models[task_list_idx] (pinned by tests/golden/imm.pt, see oracle/restate.py imm_merge)."""
    if not mean_mode and (precision is None or sum_precision is None):
        raise Exception("Can only use precision for MODE IMM, not mean IMM")
    merged_model = copy.deepcopy(models[task_list_idx])
    if mean_mode:
        return merged_model
    dev = torch.device("cuda")
    total_task_count = task_list_idx + 1
    states = [m.state_dict() for m in models[:total_task_count]]
    for param_name, param_value in merged_model.named_parameters():
        if param_name in head_param_names:
            continue
        acc = torch.empty(param_value.shape, dtype=torch.float32, device=dev)
        sp = sum_precision[param_name].detach().to(dev, torch.float32).contiguous()
        for k in range(total_task_count):
            if states[k][param_name].size() != param_value.size():
                raise Exception("ERROR WHEN MERGING MODELS: PRECEDING MODEL PARAMS TASK", str(k), " != PARAM SIZE OF REF TASK",
                                str(task_list_idx))
            pk = precision[k][param_name].detach().to(dev, torch.float32).contiguous()
            tk = states[k][param_name].detach().to(dev, torch.float32).contiguous()
            call("clb_imm_merge_accum", _ptr(acc), _ptr(pk), _ptr(sp), _ptr(tk), acc.numel(), int(k == 0), _stream())
        param_value.data = acc.to(param_value.data.device)
    return merged_model


def preprocess_merge_IMM(method, model_paths, datasets_path, batch_size, overwrite=False, debug=False):
    """merge.py:12-150: per task the precision matrix (mode-IMM), the running sum of precisions, and the merged model of
    every task but the first, saved next to that task's best model.  Returns the list of model paths to evaluate."""
    import os
    from .. import common
    IMM_mode = method.mode
    merge_model_name = 'best_model_' + IMM_mode + '_merge.pth.tar'
    last_task_idx = len(model_paths) - 1
    models = [common.load_model(p) for p in model_paths]
    merged_model_paths = [model_paths[0]]
    last_layer_index = str(len(models[0].classifier._modules) - 1)
    head_param_names = ['classifier.{}.{}'.format(last_layer_index, name) for name, p in
                        models[0].classifier._modules[last_layer_index].named_parameters()]
    precision_matrices, sum_precision_matrices, sum_precision_matrix = [], [], None
    if IMM_mode == method.modes[1]:
        precision_name = 'precision_' + IMM_mode + '.pth.tar'
        for t in range(last_task_idx + 1):
            out_path = os.path.join(os.path.dirname(model_paths[t]), precision_name)
            sum_path = os.path.join(os.path.dirname(model_paths[t]), "sum_" + precision_name)
            if os.path.exists(out_path) and not overwrite:
                precision_matrix = torch.load(out_path, weights_only=False)
            else:
                dsets = datasets_path[t]
                dsets = torch.load(dsets, weights_only=False) if isinstance(dsets, str) else dsets
                common.bind(models[t], dsets['train'], batch_size)
                loaders = common.make_loaders(dsets, batch_size, shuffle=True)
                precision_matrix = diag_fisher(models[t], loaders, exclude_params=head_param_names)
                assert set(precision_matrix) == {n for n, _ in models[t].named_parameters() if n not in head_param_names}
                torch.save(precision_matrix, out_path)
            precision_matrices.append(precision_matrix)
            if sum_precision_matrix is None:
                sum_precision_matrix = precision_matrix
            else:
                if os.path.exists(sum_path) and not overwrite:
                    sum_precision_matrix = torch.load(sum_path, weights_only=False)
                else:
                    sum_precision_matrix = {n: p + precision_matrix[n] for n, p in sum_precision_matrix.items()}
                    torch.save(sum_precision_matrix, sum_path)
                sum_precision_matrices.append(sum_precision_matrix)
    for t in range(1, last_task_idx + 1):
        out_file_path = os.path.join(os.path.dirname(model_paths[t]), merge_model_name)
        if IMM_mode == method.modes[0]:
            merged_model = IMM_merge_models(models, t, head_param_names, mean_mode=True)
        else:
            merged_model = IMM_merge_models(models, t, head_param_names, precision=precision_matrices,
                                            sum_precision=sum_precision_matrices[t - 1], mean_mode=False)
        torch.save(merged_model, out_file_path)
        merged_model_paths.append(out_file_path)
    return merged_model_paths

"""Evaluation forward path: `test_model` of src/framework/inference.py:8-87 and the default
`inference_eval` of src/methods/method.py:1066-1087, forward-only through the engine.

What changed underneath: the batches come from the device-resident task cache (clsurvey_b200/data.py), the forward runs on
the engine's kernels, and the reference's per-sample python loop over `labels[i].item()` (inference.py:64-69, one device
sync per image) is two `index_add_` calls on device counters read back once."""
import numpy as np
import torch

from ..methods import common
from ..utilities import utils


def test_model(method, model, dataset_path, target_task_head_idx, target_head=None, batch_size=200, subset='test',
               per_class_stats=False, final_layer_idx=None, task_idx=None):
    """Accuracy (%) of `model` on `dataset[subset]` using head `target_head[target_task_head_idx]` (inference.py:8-87).
    `dataset_path` is a path to (or the object of) a {'train','val'[,'test']} dict of torch Datasets."""
    if target_head is not None:
        if not isinstance(target_head, list):
            target_head = [target_head]
        assert target_task_head_idx == 0, "Only EBLL, LWF have heads in model itself, here head idx indicates target_headlist idx"
    if hasattr(model, 'classifier'):
        final_layer_idx = str(len(model.classifier._modules) - 1)
    model.eval()
    dsets = torch.load(dataset_path, weights_only=False) if isinstance(dataset_path, str) else dataset_path
    if subset not in dsets:                                  # inference.py:28-34
        print('no test set has been found')
        subset = 'val'
    dset_classes = dsets['train'].classes
    loader = common.make_loader(dsets[subset], batch_size, shuffle=False)     # order does not matter for the counters
    holder = type("Holder", (object,), {})()
    holder.task_imgfolders, holder.batch_size, holder.model = dsets, batch_size, model
    holder.heads, holder.current_head_idx = target_head, target_task_head_idx
    holder.final_layer_idx, holder.task_idx = final_layer_idx, task_idx
    n_cls = len(dset_classes)
    class_correct = class_total = None
    for data in loader:
        images, labels = data[0], data[1]
        images = images.squeeze()
        if images.dim() == 3:
            images = images.unsqueeze(0)
        outputs = method.get_output(images, holder)
        if class_correct is None:
            class_correct = torch.zeros(n_cls, dtype=torch.float64, device=outputs.device)
            class_total = torch.zeros(n_cls, dtype=torch.float64, device=outputs.device)
        labels = labels.to(outputs.device)
        _, pred = torch.max(outputs.data, 1)
        c = (pred == labels).to(torch.float64)
        class_total.index_add_(0, labels, torch.ones_like(c))
        class_correct.index_add_(0, labels, c)
    class_correct = [0.0] * n_cls if class_correct is None else class_correct.tolist()      # ONE device->host read
    class_total = [0.0] * n_cls if class_total is None else class_total.tolist()
    if per_class_stats:
        print("For all correct-head classified:")
        for i in range(n_cls):
            print('Accuracy of %5s : %2d %%' % (dset_classes[i], 100 * class_correct[i] / max(class_total[i], 1.0)))
    accuracy = np.sum(class_correct) * 100 / max(np.sum(class_total), 1.0)
    print('Overall Accuracy: ' + str(accuracy))
    test_model.last_class_stats = (class_correct, class_total)
    return accuracy


def inference_eval_default(args, manager):
    """method.py:1066-1087: load the model under evaluation, fetch the task's head from the model trained on that task."""
    model = common.load_model(args.eval_model_path)
    head_layer_idx = str(len(model.classifier._modules) - 1)
    current_head = model.classifier._modules[head_layer_idx]
    assert isinstance(current_head, torch.nn.Linear), "NO VALID HEAD IDX"
    target_heads = utils.get_prev_heads(args.head_paths, head_layer_idx)
    print("EVAL on prev heads: ", args.head_paths)
    assert len(target_heads) == 1
    return test_model(manager.method, model, args.dset_path, 0, subset=args.test_set, target_head=target_heads,
                      batch_size=args.batch_size, task_idx=args.eval_dset_idx)

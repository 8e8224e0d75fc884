"""Evaluation forward path: `test_model` of src/framework/inference.py:8-87 and the default
`inference_eval` of src/methods/method.py:1066-1087, forward-only through the engine."""
import torch

from ..engine import get_engine
from ..methods import common


def test_model(method, model, dataset, target_head_idx, target_head=None, batch_size=200, subset='test', task_idx=None):
    """Accuracy (%) of `model` on `dataset[subset]` using head `target_head[target_head_idx]` (inference.py:8-87).
    `dataset` is a {'train','val','test'} dict of torch Datasets or a path to one."""
    dsets = torch.load(dataset, weights_only=False) if isinstance(dataset, str) else dataset
    ds = dsets[subset]
    loader = torch.utils.data.DataLoader(ds, batch_size=batch_size, shuffle=False, num_workers=0)
    args = type("Args", (), {})()
    args.model, args.heads, args.current_head_idx = model, target_head, target_head_idx
    args.final_layer_idx = str(len(model.classifier._modules) - 1) if hasattr(model, "classifier") else None
    args.task_idx = task_idx
    correct = total = 0
    corr_dev = None
    for batch in loader:
        images, labels = batch[0], batch[1]
        images = images.squeeze()
        if images.dim() == 3:
            images = images.unsqueeze(0)
        out = method.get_output(images, args)
        pred = out.argmax(dim=1)
        c = (pred.cpu() == labels.cpu()).sum()
        corr_dev = c if corr_dev is None else corr_dev + c
        total += labels.size(0)
    correct = int(corr_dev.item()) if corr_dev is not None else 0
    return 100.0 * correct / max(total, 1)


def inference_eval_default(args, manager):
    """method.py:1066-1087: load the model under evaluation, fetch the task's head from the model trained on that task."""
    model = common.load_model(args.eval_model_path)
    last = str(len(model.classifier._modules) - 1)
    heads = []
    for path in args.head_paths:
        hm = common.load_model(path)
        heads.append(hm.classifier._modules[last])
    assert len(heads) == 1
    return test_model(manager.method, model, args.dset_path, 0, subset=args.test_set, target_head=heads,
                      batch_size=args.batch_size, task_idx=args.eval_dset_idx)

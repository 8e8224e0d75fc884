"""Mirror of src/methods/rehearsal/main_rehearsal.py:19-36 (eval_batch) and :68-255 (main), GEM only.

iCaRL and the two replay baselines are outside the hot path."""
import os
import types

import torch

from .. import common as mcommon
from .model import gem as gem_model


def eval_batch(model, x_batch, y_batch, args):
    """main_rehearsal.py:19-36: validation through task `task_idx`'s head slice (eval mode)."""
    from ...engine import LOSS_MEAN_CE
    model.eval()
    eng = model._engine()
    o1, o2 = model.compute_offsets(args.task_idx, model.cum_nc_per_task)
    x = x_batch if x_batch.is_cuda else x_batch.to(eng.device)
    eng.forward(x, train=False)
    eng.loss_head(y_batch, LOSS_MEAN_CE, col_off=o1, ncols=o2 - o1, want_grad=False)
    return eng.loss_dev.clone(), eng.correct_dev.clone()


class PathRetriever(torch.utils.data.Dataset):
    """(sample, target, key) triples -- stands in for ImageFolder_Subset_PathRetriever (src/data/imgfolder.py:179-198):
    the key is the image path when the wrapped dataset exposes `.samples`/`.imgs`, else the index."""

    def __init__(self, dset):
        self.dset = dset
        self.keys = getattr(dset, "samples", None) or getattr(dset, "imgs", None)

    def __len__(self):
        return len(self.dset)

    def __getitem__(self, i):
        item = self.dset[i]
        key = self.keys[i][0] if self.keys is not None else i
        return item[0], item[1], key


DEFAULTS = dict(task_name=None, task_count=None, prev_model_path=None, save_path='results/', n_outputs=200, method='gem',
                postprocess=False, debug=False, weight_decay=0, is_scratch_model=False, n_memories=0,
                memory_strength=0, finetune=False, n_epochs=1, batch_size=70, lr=1e-3, cuda=True, n_tasks=10,
                dataset_path=None, n_inputs=-1)


def main(overwrite_args, nc_per_task):
    """main_rehearsal.py:68-255 for method == 'gem'.  Returns (model, best_val_acc) or (None, None) when postprocessing."""
    args = types.SimpleNamespace(**DEFAULTS)
    args.nc_per_task = nc_per_task
    for k, v in overwrite_args.items():
        setattr(args, k, v)
    if args.method != 'gem':
        raise NotImplementedError("only GEM is on the hot path (SURVEY.md 2.1 #7)")
    args.task_idx = args.task_count - 1
    assert args.n_outputs == sum(args.nc_per_task)
    assert args.n_tasks == len(nc_per_task)
    if args.task_count == 1:
        assert args.postprocess, "FIRST TASK WE DO ONLY POSTPROCESSING"
    dsets = torch.load(args.dataset_path, weights_only=False) if isinstance(args.dataset_path, str) else args.dataset_path
    args.task_imgfolders = dsets
    args.dset_loaders = {x: torch.utils.data.DataLoader(PathRetriever(dsets[x]), batch_size=args.batch_size, shuffle=True,
                                                        num_workers=0) for x in ['train', 'val']}
    dset_sizes = {x: len(dsets[x]) for x in ['train', 'val']}
    args.input_shape = mcommon.sample_shape(dsets['train'])
    if args.is_scratch_model:
        assert args.task_idx == 0
        model = gem_model.Net(args.n_inputs, args.n_outputs, args.n_tasks, args)
    else:
        model = torch.load(args.prev_model_path, weights_only=False)
        model._input_shape = tuple(args.input_shape)
    model.init_setup(args)
    assert model.n_tasks == args.n_tasks and model.n_outputs == args.n_outputs
    if args.postprocess:
        model.manage_memory(args.task_idx, args)
        os.makedirs(os.path.dirname(args.save_path) or ".", exist_ok=True)
        torch.save(model, args.save_path)
        return None, None
    from . import train_rehearsal
    resume = os.path.join(args.save_path, 'epoch.pth.tar')
    return train_rehearsal.train_model(model, args, dset_sizes, resume=resume)

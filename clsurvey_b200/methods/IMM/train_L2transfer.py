"""Mirror of src/methods/IMM/train_L2transfer.py: the L2-transfer trainer of IMM.

`Weight_Regularized_SGD.step` (train_L2transfer.py:35-100) is the EWC / MAS penalised step with omega == 1, i.e. the same
fused launch (clb_sgd_penalty_step); the epoch loop (train_L2transfer.py:119-230) is the common protocol (stop at > 10, lr
cut at 5, no divergence exit) = the 'sgd' flavour with `optimizer.step(model.reg_params)`."""
import torch

from ..optim import Weight_Regularized_SGD  # noqa: F401
from ..trainers import run_train_model, set_lr  # noqa: F401


def train_model(model, criterion, optimizer, lr, dset_loaders, dset_sizes, use_gpu, num_epochs, exp_dir='./', resume='',
                saving_freq=5):
    """train_L2transfer.py:119-230.  Returns (model, best_val_acc)."""
    return run_train_model("l2t", model, criterion, optimizer, lr, dset_loaders, dset_sizes, use_gpu, num_epochs, exp_dir,
                           resume, saving_freq)

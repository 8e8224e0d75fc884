"""Mirror of src/methods/EWC/train_EWC.py (a6, a7): Weight_Regularized_SGD + train_model, engine underneath."""
from ..optim import Weight_Regularized_SGD  # noqa: F401
from ..trainers import run_train_model, set_lr  # noqa: F401


def train_model(model, criterion, optimizer, lr, dset_loaders, dset_sizes, use_gpu, num_epochs, exp_dir='./',
                resume='', saving_freq=5):
    """train_EWC.py:111-234.  Returns (model, best_val_acc)."""
    return run_train_model("ewc", model, criterion, optimizer, lr, dset_loaders, dset_sizes, use_gpu, num_epochs,
                           exp_dir, resume, saving_freq)

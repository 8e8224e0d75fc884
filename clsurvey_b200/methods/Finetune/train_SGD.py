"""Mirror of src/methods/Finetune/train_SGD.py (a1): same `train_model` / `set_lr` surface, engine underneath."""
from ..trainers import run_train_model, set_lr  # noqa: F401


def train_model(model, criterion, optimizer, lr, dset_loaders, dset_sizes, use_gpu, num_epochs, exp_dir='./',
                resume='', save_models_mode=True, saving_freq=5, print_freq=100):
    """train_SGD.py:41-189.  Returns (model, best_val_acc)."""
    return run_train_model("sgd", model, criterion, optimizer, lr, dset_loaders, dset_sizes, use_gpu, num_epochs,
                           exp_dir, resume, saving_freq, save_models_mode=save_models_mode)

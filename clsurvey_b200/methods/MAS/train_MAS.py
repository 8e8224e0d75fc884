"""Mirror of src/methods/MAS/train_MAS.py (a6, a8, a9), L2-norm / b1=False path only (method.py:737-750)."""
import torch

from ... import dist as cdist
from ...engine import LOSS_SUM_SQ, engine_of
from ..EWC.main_EWC import (_zero_unregistered, accumelate_reg_params, initialize_reg_params,  # noqa: F401
                            store_prev_reg_params)
from ..optim import Objective_After_SGD, Weight_Regularized_SGD, sync_reg_params  # noqa: F401
from ..trainers import run_train_model, set_lr  # noqa: F401

initialize_store_reg_params = store_prev_reg_params        # train_MAS.py:710-733 (same protocol as EWC's)


def train_model(model, criterion, optimizer, lr, dset_loaders, dset_sizes, use_gpu, num_epochs, exp_dir='./',
                resume='', saving_freq=5):
    """train_MAS.py:208-335.  Returns (model, best_val_acc)."""
    return run_train_model("mas", model, criterion, optimizer, lr, dset_loaders, dset_sizes, use_gpu, num_epochs,
                           exp_dir, resume, saving_freq)


def compute_importance_l2(model, optimizer, lr_scheduler, dset_loaders, use_gpu):
    """train_MAS.py:508-567: per batch L = sum(out**2), backward, omega <- (omega*b*n_b + |g|)/((b+1)*n_b).

    Data parallel: whole batches are dealt round-robin; each rank accumulates sum_b |g_b| (prev=1, curr=1 form) and the
    all-reduced sum is divided once -- equal to the reference's running mean for equal batch sizes.
    The recurrence unrolls to omega = mean_b(|g_b| / n_b), so ragged batches shard too."""
    eng = engine_of(model.parameters())
    model.eval()
    world, rk = cdist.world_size(), cdist.rank()
    index = 0
    if world == 1:
        for dset_loader in dset_loaders:
            for inputs, labels in dset_loader:
                x = inputs if inputs.is_cuda else inputs.to(eng.device, non_blocking=True)
                # Objective_After_SGD.step (train_MAS.py:163-177) fused into this batch's backward pass
                sync_reg_params(eng, model.reg_params, need_w=False)
                n_b = labels.size(0)
                eng.fwd_loss_bwd(x, None, LOSS_SUM_SQ, train=False, importance=(2, float(index * n_b), float((index + 1) * n_b)))
                index += 1
    else:
        from ..._capi import call
        from ...engine import _ptr, _stream
        sync_reg_params(eng, model.reg_params, need_w=False)
        # the recurrence unrolls to omega = (1 / n_batches) * sum_b |g_b| / n_b with n_b the size of batch b -- also for
        # ragged last batches: each rank adds |g_b| / n_b for its batches ((omega*n_b + |g|) / n_b), the sums are all-reduced
        # and divided by the number of batches once
        for dset_loader in dset_loaders:
            for inputs, labels in dset_loader:
                if index % world == rk:
                    x = inputs if inputs.is_cuda else inputs.to(eng.device, non_blocking=True)
                    n_b = float(labels.size(0))
                    eng.fwd_loss_bwd(x, None, LOSS_SUM_SQ, train=False, importance=(2, n_b, n_b))
                index += 1
        cdist.allreduce_flat(eng.omega)
        eng.zero_grad()                                   # omega <- (omega*1 + |0|) / n_batches
        call("clb_mas_accum", _ptr(eng.omega), _ptr(eng.grad), 1.0, float(max(index, 1)), eng.total, _stream())
    _zero_unregistered(eng, model.reg_params)
    return model

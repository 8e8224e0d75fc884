"""Maximal-plasticity search: the learning-rate grid of src/framework/lr_grid_train.py:9-160, with its grid nodes trained as
CONCURRENT REPLICAS.

The reference trains the |lrs| x finetune_iterations nodes one after the other (lr_grid_train.py:51-83) although they are
independent trainings from the same start model: phase 1 is 5x the cost of one training, and for EWC / MAS / SI its models
are thrown away (only the best lr and its accuracy survive, framework_train.py:230-235).  At the reference's global batch
of 200 one B200 is already far from saturated by data parallelism (8 GPUs: the step is latency-bound), so the better use
of a node is one replica per GPU: every node of the grid runs in its own process pinned to one GPU
(CUDA_VISIBLE_DEVICES), at most `gpus` at a time, and the parent applies the reference's selection rule to the returned
accuracies in the reference's order -- so best_lr / best_acc, the log file, grid_checkpoint.pth, the per-node hyperparams
file and the storage policy are those of a sequential run.  Seeding is per node (utils.set_random(finetune_iteration), as
in lr_grid_train.py:73,77), which makes a node's result independent of where and when it runs.
"""
import os
import pickle
import shutil
import subprocess
import sys
import time

import torch

from ..utilities import utils


class StoragePolicy(object):
    """lr_grid_train.py:162-176."""

    def __init__(self, save_models_mode):
        if save_models_mode not in ['all', 'keep_none', 'only_keep_best']:
            raise Exception("Invalid value for save_models_mode")
        self.keep_none = save_models_mode == 'keep_none'
        self.only_keep_best = save_models_mode == 'only_keep_best'


def float_to_scientific_str(value, sig_count=1):
    """utils.float_to_scientific_str (src/utilities/utils.py): '1.0e-03' style node directory names."""
    from decimal import Decimal
    return ('%.' + str(sig_count) + 'E') % Decimal(value)


def node_dirname(lr, finetune_iterations, it):
    name = "lr=" + str(float_to_scientific_str(lr))
    if finetune_iterations > 1:
        name += "_it" + str(it)
    return name


def _visible_gpus(gpus):
    if gpus is None:
        gpus = torch.cuda.device_count() if torch.cuda.is_available() else 1
    if isinstance(gpus, int):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        ids = [g.strip() for g in vis.split(",")] if vis else [str(i) for i in range(max(gpus, 1))]
        return ids[:max(gpus, 1)]
    return [str(g) for g in gpus]


def run_nodes_concurrently(nodes, args, manager, gpus=None, timeout=None):
    """nodes: [(lr, finetune_iteration, exp_dir)].  Trains every node in its own process, one GPU each, at most len(gpus) at a
    time.  Returns {(lr, it): (acc, seconds)}."""
    gpu_ids = _visible_gpus(gpus)
    free, running, results = list(gpu_ids), [], {}
    pending = list(nodes)
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

    def launch(node, gpu):
        lr, it, exp_dir = node
        os.makedirs(exp_dir, exist_ok=True)
        payload = os.path.join(exp_dir, "grid_node_payload.pkl")
        with open(payload, "wb") as f:
            pickle.dump(dict(args=args, manager=manager, lr=lr, iteration=it, exp_dir=exp_dir), f)
        env = dict(os.environ, CUDA_VISIBLE_DEVICES=gpu, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
        for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):              # a replica is a single-process job
            env.pop(k, None)
        log = open(os.path.join(exp_dir, "grid_node.log"), "w")
        p = subprocess.Popen([sys.executable, "-m", "clsurvey_b200.framework.lr_grid_train", payload], env=env, stdout=log,
                             stderr=subprocess.STDOUT, cwd=root)
        return (p, node, gpu, time.time(), log)

    while pending or running:
        while pending and free:
            running.append(launch(pending.pop(0), free.pop(0)))
        time.sleep(0.05)
        for ent in list(running):
            p, node, gpu, t0, log = ent
            rc = p.poll()
            if rc is None:
                if timeout is not None and time.time() - t0 > timeout:
                    p.kill()
                continue
            running.remove(ent)
            free.append(gpu)
            log.close()
            res = os.path.join(node[2], "grid_node_result.pth")
            if rc != 0 or not os.path.exists(res):
                tail = open(os.path.join(node[2], "grid_node.log")).read()[-2000:]
                raise RuntimeError("grid node lr=%r it=%d failed (rc=%s):\n%s" % (node[0], node[1], rc, tail))
            r = torch.load(res, weights_only=False)
            results[(node[0], node[1])] = (r["acc"], r["seconds"])
    return results


def lr_grid_single_task(args, manager, save_models_mode='keep_none', gpus=None):
    """lr_grid_train.py:9-160 with concurrent grid nodes.  Returns (best_lr, best_acc)."""
    manager.store_policy = StoragePolicy(save_models_mode)
    if hasattr(manager, "dataset") and hasattr(manager.dataset, "get_taskname"):
        args.task_name = manager.dataset.get_taskname(args.task_counter)
    manager.ft_parent_exp_dir = os.path.join(manager.parent_exp_dir, 'task_' + str(args.task_counter), 'FT_LR_GRIDSEARCH')
    os.makedirs(manager.ft_parent_exp_dir, exist_ok=True)
    logfile_parent_dir = os.path.join(manager.ft_parent_exp_dir, 'log')
    os.makedirs(logfile_parent_dir, exist_ok=True)
    logfile = os.path.join(logfile_parent_dir, time.strftime("%Y-%m-%d_%H-%M-%S") + '_finetune_grid.log')

    def log(msg):
        with open(logfile, "a") as f:
            f.write(msg + "\n")

    log("FINETUNE GRIDSEARCH LOG: Processed LRs")
    processed_lrs = {}
    grid_checkpoint_file = os.path.join(manager.ft_parent_exp_dir, 'grid_checkpoint.pth')
    if os.path.exists(grid_checkpoint_file):
        processed_lrs = torch.load(grid_checkpoint_file, weights_only=False)['processed_lrs']
        log("STARTING FROM CHECKPOINT")
    args.presteps_elapsed_time = 0
    if hasattr(manager.method, 'grid_prestep'):
        manager.method.grid_prestep(args, manager)
    iters = args.finetune_iterations
    # every node that has no stored accuracy yet trains now, all of them concurrently
    todo = []
    for lr in args.lrs:
        done = len(processed_lrs.get(lr, {'acc': []})['acc'])
        for it in range(done, iters):
            todo.append((lr, it, os.path.join(manager.ft_parent_exp_dir, node_dirname(lr, iters, it))))
    t0 = time.time()
    fresh = run_nodes_concurrently(todo, args, manager, gpus) if todo else {}
    manager.grid_wall_seconds = time.time() - t0
    # the reference's selection loop over the (now known) accuracies, in the reference's order
    best_acc, best_lr = 0, None
    manager.best_exp_grid_node_dirname = None
    best_iteration_batch_dirs = []
    for lr in args.lrs:
        accum_acc, best_iteration_dir, best_iteration_acc, iteration_batch_dirs = 0, None, 0, []
        if lr not in processed_lrs:
            processed_lrs[lr] = {'acc': []}
        for it in range(iters):
            manager.gridsearch_exp_dir = os.path.join(manager.ft_parent_exp_dir, node_dirname(lr, iters, it))
            iteration_batch_dirs.append(manager.gridsearch_exp_dir)
            if it < len(processed_lrs[lr]['acc']):
                acc = processed_lrs[lr]['acc'][it]
            else:
                acc, seconds = fresh[(lr, it)]
                processed_lrs[lr]['acc'].append(acc)
                log("LR = {}, FT Iteration {}/{}, Acc = {}".format(lr, it + 1, iters, acc))
                if getattr(manager.method, "grid_chkpt", False) and hasattr(manager, "save_hyperparams"):
                    manager.save_hyperparams(manager.gridsearch_exp_dir, {'val_acc': acc, 'lr': lr, 'iteration_elapsed_time': seconds})
            if acc > best_iteration_acc:
                best_iteration_acc, best_iteration_dir = acc, manager.gridsearch_exp_dir
            accum_acc = accum_acc + acc
            torch.save({'processed_lrs': processed_lrs}, grid_checkpoint_file)
        avg_acc = accum_acc / iters
        if avg_acc > best_acc:
            best_lr, best_acc = lr, avg_acc
            manager.best_exp_grid_node_dirname = best_iteration_dir
            log("UPDATE best lr = {}".format(best_lr))
            log("UPDATE best lr acc= {}\n".format(best_acc))
            if manager.store_policy.only_keep_best:
                for out_dir in best_iteration_batch_dirs:
                    shutil.rmtree(out_dir, ignore_errors=True)
            best_iteration_batch_dirs = iteration_batch_dirs
        elif manager.store_policy.only_keep_best:
            for out_dir in iteration_batch_dirs:
                shutil.rmtree(out_dir, ignore_errors=True)
        if manager.store_policy.keep_none:
            for out_dir in iteration_batch_dirs:
                shutil.rmtree(out_dir, ignore_errors=True)
    print("FINETUNE DONE: best_lr={}, best_acc={}".format(best_lr, best_acc))
    if hasattr(manager.method, 'grid_poststep'):
        manager.method.grid_poststep(args, manager)
    return best_lr, best_acc


def _worker(payload_path):
    with open(payload_path, "rb") as f:
        pl = pickle.load(f)
    args, manager, lr, it, exp_dir = pl["args"], pl["manager"], pl["lr"], pl["iteration"], pl["exp_dir"]
    utils.set_random(it)                                   # lr_grid_train.py:77
    manager.gridsearch_exp_dir = exp_dir
    t0 = time.time()
    model, acc = manager.method.grid_train(args, manager, lr)
    torch.cuda.synchronize()
    from ..methods import trainers
    torch.save({"acc": acc, "seconds": time.time() - t0, "train_images": trainers.LAST_RUN.get("train_images", 0),
                "train_seconds": trainers.LAST_RUN.get("train_seconds", 0.0)}, os.path.join(exp_dir, "grid_node_result.pth"))


if __name__ == "__main__":
    _worker(sys.argv[1])

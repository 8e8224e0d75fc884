"""f-4 measurement: the 5 learning rates of phase 1 (src/framework/main.py:61: lrs) as concurrent replicas, one per GPU.
    python tools/lr_grid_bench.py [--gpus N] [--epochs E] [--train N_IMAGES]
Prints one JSON line: aggregate images/s of the grid (all replicas' training images / wall time of the whole grid), the
sequential equivalent (sum of the replicas' own training times) and the chosen lr."""
import argparse, json, os, sys, tempfile, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=None)
ap.add_argument("--epochs", type=int, default=3)
ap.add_argument("--train", type=int, default=8000)
ap.add_argument("--model", default="VGG11_cl_512_512")
a = ap.parse_args()
from clsurvey_b200.framework import lr_grid_train as G
from clsurvey_b200.methods import method as M
from clsurvey_b200.models import parse_model_name

tmp = tempfile.mkdtemp(prefix="clb_grid_")


from clsurvey_b200.data import TaskTensorDataset
g = torch.Generator().manual_seed(7)
mk = lambda n: TaskTensorDataset(torch.randn(n, 3, 64, 64, generator=g), torch.randint(0, 20, (n,), generator=g), range(20))
dpath = os.path.join(tmp, "task.pth")
torch.save({"train": mk(a.train), "val": mk(2000)}, dpath)
torch.manual_seed(7)
mpath = os.path.join(tmp, "first_task_model.pth.tar")
torch.save(parse_model_name(a.model, (64, 64), 20), mpath)
lrs = [1e-2, 5e-3, 1e-3, 5e-4, 1e-4]                         # src/framework/main.py:61
args = types.SimpleNamespace(lrs=lrs, finetune_iterations=1, task_counter=2, batch_size=200, num_epochs=a.epochs,
                             weight_decay=0.0, saving_freq=1000)
manager = types.SimpleNamespace(method=M.Finetune(), parent_exp_dir=os.path.join(tmp, "exp"), current_task_dataset_path=dpath,
                                previous_task_model_path=mpath)
t0 = time.time()
best_lr, best_acc = G.lr_grid_single_task(args, manager, save_models_mode="all", gpus=a.gpus)
wall = time.time() - t0
imgs, seq = 0, 0.0
for lr in lrs:
    r = torch.load(os.path.join(manager.ft_parent_exp_dir, G.node_dirname(lr, 1, 0), "grid_node_result.pth"), weights_only=False)
    imgs += r["train_images"]
    seq += r["train_seconds"]
n_gpus = len(G._visible_gpus(a.gpus))
print(json.dumps({"metric": "images/sec/task (LR grid, 5 replicas)", "n_gpus": n_gpus, "replicas": len(lrs), "epochs": a.epochs,
                  "train_images_per_replica": a.train * a.epochs, "grid_wall_s": wall, "nodes_wall_s": manager.grid_wall_seconds,
                  "aggregate_images_per_s_wall": imgs / manager.grid_wall_seconds,
                  "sum_of_replica_train_seconds": seq, "images_per_s_inside_training": imgs / seq * min(n_gpus, len(lrs)),
                  "best_lr": best_lr, "best_acc": best_acc}))

"""Where the end-to-end time of bench.py's e2e arm goes: train_MAS.train_model on a pinned-host task, 10 epochs x 40 batches,
(a) as benched, (b) with torch.save replaced by a no-op (loop + snapshot cost only), (c) CLB_ASYNC_SAVE=0 semantics timed."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from clsurvey_b200.data import PinnedLoader
from clsurvey_b200.engine import Engine
from clsurvey_b200.methods import trainers
from clsurvey_b200.methods.MAS import train_MAS
model = bench.build_model("VGG11_cl_512_512")
eng = Engine(model, (3, 64, 64), 200)
params = list(model.parameters())
model.reg_params = {p: {"omega": torch.rand_like(p.data) * 1e-3, "init_val": p.data.clone()} for p in params[:-2]}
model.reg_params["lambda"] = 3.0
n, E = 40, 10
g = torch.Generator().manual_seed(6)
xe = torch.randn(n * 200, 3, 64, 64, generator=g); ye = torch.randint(0, 20, (n * 200,), generator=g)
ds = torch.utils.data.TensorDataset(xe, ye)
vds = torch.utils.data.TensorDataset(xe[:200].clone(), ye[:200].clone())
tmp = tempfile.mkdtemp()
orig = torch.save
saves = []
def timed_save(*a, **k):
    t = time.time(); orig(*a, **k); saves.append(time.time() - t)
def run(tag, epochs=E):
    loaders = {"train": PinnedLoader(ds, 200), "val": PinnedLoader(vds, 200)}
    o = train_MAS.Weight_Regularized_SGD(model.parameters(), 0.01, momentum=0.9, weight_decay=0.0)
    del saves[:]
    out = sys.stdout
    sys.stdout = open(os.devnull, "w")
    torch.cuda.synchronize(); t1 = time.time()
    train_MAS.train_model(model, torch.nn.CrossEntropyLoss(), o, 0.01, loaders, {"train": len(ds), "val": 200}, True, epochs, exp_dir=tmp + "/", resume="")
    torch.cuda.synchronize(); t2 = time.time()
    sys.stdout = out
    lr = trainers.LAST_RUN
    print("%-28s total %.3f s = %.2f ms/step; inside train phases %.3f s; %d torch.save calls, %.3f s in them (%s)" % (
        tag, t2 - t1, (t2 - t1) * 1e3 / (n * epochs), lr.get("train_seconds", 0.0), len(saves), sum(saves),
        " ".join("%.2f" % s for s in saves)), flush=True)
run("warm-up", 1)
torch.save = timed_save
run("as benched")
run("as benched (again)")
torch.save = lambda *a, **k: saves.append(0.0)
run("torch.save -> no-op")
torch.save = timed_save
os.environ["CLB_ASYNC_SAVE"] = "0"
run("inline saves")

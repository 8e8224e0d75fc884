import os, sys, time, tempfile
sys.path.insert(0, "/root/repo")
import torch
import bench
from clsurvey_b200.data import PinnedLoader
from clsurvey_b200.engine import Engine
from clsurvey_b200.methods.MAS import train_MAS
model = bench.build_model("VGG11_cl_512_512")
eng = Engine(model, (3, 64, 64), 200)
params = list(model.parameters())
model.reg_params = {p: {"omega": torch.rand_like(p.data) * 1e-3, "init_val": p.data.clone()} for p in params[:-2]}
model.reg_params["lambda"] = 3.0
n = 40
g = torch.Generator().manual_seed(6)
xe = torch.randn(n * 200, 3, 64, 64, generator=g); ye = torch.randint(0, 20, (n * 200,), generator=g)
ds = torch.utils.data.TensorDataset(xe, ye)
vds = torch.utils.data.TensorDataset(xe[:200].clone(), ye[:200].clone())
tmp = tempfile.mkdtemp()
save_t = [0.0]
orig = torch.save
def timed_save(*a, **k):
    torch.cuda.synchronize(); t = time.time(); orig(*a, **k); save_t[0] += time.time() - t
torch.save = timed_save
for epochs in (1, 1, 1, 3):
    t0 = time.time()
    loaders = {"train": PinnedLoader(ds, 200), "val": PinnedLoader(vds, 200)}
    t1 = time.time()
    o = train_MAS.Weight_Regularized_SGD(model.parameters(), 0.01, momentum=0.9, weight_decay=0.0)
    save_t[0] = 0.0
    sys.stdout = open(os.devnull, "w")
    train_MAS.train_model(model, torch.nn.CrossEntropyLoss(), o, 0.01, loaders, {"train": len(ds), "val": 200}, True, epochs, exp_dir=tmp + "/", resume="")
    torch.cuda.synchronize()
    sys.stdout = sys.__stdout__
    t2 = time.time()
    print("epochs %d: loaders %.3f s, train_model %.3f s (torch.save %.3f s) -> %.2f ms/step excluding saves" % (epochs, t1 - t0, t2 - t1, save_t[0], (t2 - t1 - save_t[0]) * 1e3 / (n * epochs)), flush=True)

"""Device-resident task datasets.

The reference re-reads every task's images through `DataLoader` workers + PIL decode for every epoch
(src/data/imgfolder.py:86-128, src/methods/method.py:1058-1061, src/methods/EWC/main_EWC.py:29-31), and its training sets
are the UN-AUGMENTED ones (src/framework/main.py:163,197-202), i.e. every epoch sees the same tensors.  One TinyImagenet
task is 8000 x 3 x 64 x 64 fp32 = 393 MB: it fits in HBM thousands of times over, so a task is materialised ONCE (one pass
over the reference's Dataset object, whatever it decodes from) and every later batch is a slice / gather on the device.

`CachedLoader` is a drop-in for the `DataLoader` objects the reference's `train_model` functions iterate
(`for data in dset_loaders[phase]: inputs, labels = data`, train_EWC.py:160-166): same length, same batch composition and --
with shuffle=True -- the same permutation a `DataLoader(shuffle=True)` would draw from the host generator at that point
(torch's RandomSampler protocol is replayed draw for draw), so switching the cache on does not change a single batch.
`PinnedLoader` is the same with the tensors in pinned HOST memory (batches are views; the trainers copy them with
non-blocking H2D copies): the end-to-end arm of bench.py uses it.
"""
import os

import torch

_CACHE = {}                    # id(dataset) -> (dataset, x, y): a task's tensors live as long as its Dataset object


def _materialise(dset, device, pinned=False):
    """One pass over a map-style dataset of (image tensor, label[, ...]) -> (x [N,...] fp32, y [N] int64)."""
    key = (id(dset), str(device), pinned)
    hit = _CACHE.get(key)
    if hit is not None and hit[0] is dset:
        return hit[1], hit[2]
    tensors = getattr(dset, "tensors", None)
    if tensors is not None and len(tensors) >= 2:                     # TensorDataset: no per-sample loop
        x, y = tensors[0], tensors[1]
    else:
        xs, ys = [], []
        for i in range(len(dset)):
            item = dset[i]
            xs.append(torch.as_tensor(item[0]))
            ys.append(int(item[1]))
        x, y = torch.stack(xs), torch.tensor(ys, dtype=torch.int64)
    x, y = x.to(torch.float32), y.to(torch.int64)
    if pinned:
        x, y = x.contiguous().pin_memory(), y.contiguous().pin_memory()
    else:
        x, y = x.to(device).contiguous(), y.to(device).contiguous()
    _CACHE[key] = (dset, x, y)
    return x, y


def drop_cache():
    _CACHE.clear()


def _loader_permutation(n):
    """The index order a fresh `iter(DataLoader(dataset, shuffle=True))` yields, drawn from the host generator exactly like
    torch does: the loader iterator first draws its base seed, then RandomSampler seeds a private generator and calls
    randperm (torch/utils/data/dataloader.py, sampler.py).  tests/test_cpu_data.py pins this against a real DataLoader."""
    torch.empty((), dtype=torch.int64).random_()                      # _BaseDataLoaderIter._base_seed
    seed = int(torch.empty((), dtype=torch.int64).random_().item())   # RandomSampler.__iter__
    g = torch.Generator()
    g.manual_seed(seed)
    return torch.randperm(n, generator=g)


class CachedLoader:
    """Iterable over (inputs, labels) device batches of a cached task; len() = number of batches (like DataLoader)."""

    def __init__(self, dset, batch_size, shuffle=False, device="cuda", pinned=False):
        self.dataset, self.batch_size, self.shuffle = dset, int(batch_size), shuffle
        self.device, self.pinned = torch.device(device), pinned
        self.x, self.y = _materialise(dset, self.device, pinned)

    def __len__(self):
        return (self.x.shape[0] + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n, bs = self.x.shape[0], self.batch_size
        if not self.shuffle:
            torch.empty((), dtype=torch.int64).random_()               # a DataLoader iterator draws its base seed even unshuffled
            for i in range(0, n, bs):
                yield self.x[i:i + bs], self.y[i:i + bs]
            return
        perm = _loader_permutation(n)
        if not self.pinned:
            perm = perm.to(self.device)
        for i in range(0, n, bs):
            idx = perm[i:i + bs]
            yield self.x.index_select(0, idx), self.y.index_select(0, idx)


class PinnedLoader(CachedLoader):
    """The same protocol with the task in pinned host memory: un-shuffled batches are views, copied H2D by the trainer."""

    def __init__(self, dset, batch_size, shuffle=False):
        super().__init__(dset, batch_size, shuffle, device="cpu", pinned=True)


class TaskTensorDataset(torch.utils.data.TensorDataset):
    """An in-memory task in the shape the reference's pickled datasets have: (image, label) items and a `.classes` list
    (src/data/imgfolder.py ImageFolderTrainVal).  Importable, so it survives torch.save / worker processes."""

    def __init__(self, x, y, classes):
        super().__init__(x, y)
        self.classes = list(classes)


def cache_enabled(dset, limit_bytes=None):
    """CLB_DATA_CACHE=0 switches the cache off; tasks beyond CLB_DATA_CACHE_GB (default 16) stay on their DataLoader."""
    if os.environ.get("CLB_DATA_CACHE", "1") == "0" or not torch.cuda.is_available():
        return False
    try:
        n = len(dset)
        item = dset[0][0]
        per = item.numel() * 4
    except Exception:
        return False
    limit = float(os.environ.get("CLB_DATA_CACHE_GB", "16")) * 2 ** 30 if limit_bytes is None else limit_bytes
    return n * per <= limit


def make_loaders(dsets, batch_size, shuffle=True, phases=("train", "val"), device="cuda"):
    """{'train': loader, 'val': loader}: cached on the device when the task fits, else the reference's DataLoader."""
    out = {}
    for ph in phases:
        if cache_enabled(dsets[ph]):
            out[ph] = CachedLoader(dsets[ph], batch_size, shuffle, device)
        else:
            out[ph] = torch.utils.data.DataLoader(dsets[ph], batch_size=batch_size, shuffle=shuffle, num_workers=0)
    return out

"""Import the UNMODIFIED reference (Mattdl/CLsurvey, /root/reference/src) on CPU under a shim layer.

TEST INFRASTRUCTURE ONLY.  This module is the strongest oracle this repo has: the
reference's own Python functions, executed on the CPU by torch 2.11 (standing in for
the pinned torch 1.6).  It exists only in the authoring container (the GPU box has no
/root/reference), so it is used solely by `oracle/gen_golden.py` to produce the
fixtures under `tests/golden/` and by `tests/test_oracle_vs_reference.py` (skipped when
the reference tree is absent) to pin `oracle/restate.py` against the real thing.

Nothing under clsurvey_b200/ imports this file.

Shim set -- none of these edits reference source files:
  * identity ``Tensor.cuda`` / ``Module.cuda`` (reference hard-codes .cuda(), e.g.
    src/methods/EWC/main_EWC.py:145,154; src/methods/EWC/train_EWC.py:58-59)
  * ``transforms.Scale`` alias (src/methods/EWC/main_EWC.py:88)
  * ``torchvision.models.VGG._initialize_weights`` (src/models/VGGSlim.py:76)
  * ``torch.load(weights_only=False)`` (src/methods/EWC/main_EWC.py:39)
  * stub modules matplotlib / pylab / torchnet / quadprog (src/methods/SI/train_SI.py:8,
    src/methods/rehearsal/model/gem.py:12)
  * ``quadprog.solve_qp`` = exact fp64 active-set enumeration (oracle/qp.py)
  * ``torch.cuda.LongTensor`` / ``memory_cached`` on a CPU box
  * GEM only: RehearsalMemory.get_imagefolder/get_dataloader serve tensors by key with
    shuffle=False (the reference re-reads JPEG paths from disk, gem.py:233-237)
"""
import functools
import os
import sys
import types

import torch
import torch.nn as nn
import torchvision

REF_SRC = os.environ.get("CLSURVEY_REFERENCE_SRC", "/root/reference/src")
_installed = False


def available():
    return os.path.isdir(os.path.join(REF_SRC, "methods"))


def _vgg_init(self):
    # torchvision VGG init that VGGSlim.py:76 expects (removed from torchvision >= 0.13 as a method)
    for m in self.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.BatchNorm2d):
            nn.init.constant_(m.weight, 1)
            nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.Linear):
            nn.init.normal_(m.weight, 0, 0.01)
            nn.init.constant_(m.bias, 0)


class KeyedTensorStore(torch.utils.data.Dataset):
    """Serves (tensor, target) by exemplar key -- replaces the JPEG path reader for GEM memories."""
    store = {}

    def __init__(self, keys, targets):
        self.keys = list(keys)
        self.targets = targets

    def __len__(self):
        return len(self.keys)

    def __getitem__(self, i):
        x = KeyedTensorStore.store[self.keys[i]]
        if self.targets is None:
            return x
        return x, self.targets[i]


def install():
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF_SRC)
    sys.path.insert(0, REF_SRC)
    torch.Tensor.cuda = lambda self, *a, **k: self
    nn.Module.cuda = lambda self, *a, **k: self
    torchvision.transforms.Scale = torchvision.transforms.Resize
    if not hasattr(torchvision.models.VGG, "_initialize_weights"):
        torchvision.models.VGG._initialize_weights = _vgg_init
    if not getattr(torch.load, "_clb_shim", False):
        orig = torch.load
        ld = functools.partial(orig, weights_only=False)
        ld._clb_shim = True
        torch.load = ld
    torch.cuda.memory_allocated = lambda device=None: 0
    torch.cuda.memory_cached = lambda device=None: 0
    torch.cuda.LongTensor = torch.LongTensor

    def stub(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules.setdefault(name, m)
        return sys.modules[name]

    mpl = stub("matplotlib", rcParams={}, use=lambda *a, **k: None)
    plt = stub("matplotlib.pyplot")
    mpl.pyplot = plt
    stub("pylab")
    tn = stub("torchnet")
    tn.meter = types.SimpleNamespace(ClassErrorMeter=object)
    from oracle import qp as _qp
    stub("quadprog", solve_qp=_qp.solve_qp_quadprog_signature)
    _installed = True


def patch_gem_memory():
    """Deviation (2) of SURVEY.md 8c: serve memory exemplars from tensors, deterministic order."""
    install()
    import methods.rehearsal.model.common as common

    def get_imagefolder(self, exemplarlist, targetlist, transform):
        return KeyedTensorStore(exemplarlist, targetlist)

    def get_dataloader(self, imgfolder, batch_size=None):
        if batch_size is None:
            batch_size = self.n_memories
        return torch.utils.data.DataLoader(imgfolder, batch_size=batch_size, shuffle=False, num_workers=0)

    common.RehearsalMemory.get_imagefolder = get_imagefolder
    common.RehearsalMemory.get_dataloader = get_dataloader

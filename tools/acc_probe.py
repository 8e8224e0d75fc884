"""Worst-case errors of the default-mode engine on VGG-11 (batch from argv, default 200) against the decision-forced fp64
evaluation (tests/forced_ref.py): logits, loss, gradients, Fisher omega.  Used to compare kernel variants (CLB_LIB_PATH)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.forced_ref import forced_reference
from tests.util import rel_err
from clsurvey_b200.engine import Engine
from clsurvey_b200.models import make_vgg

B = int(sys.argv[1]) if len(sys.argv) > 1 else 200
torch.manual_seed(7)
model = make_vgg("VGG11_cl_512_512")
with torch.no_grad():
    for m in model.classifier:
        if hasattr(m, "weight"):
            m.weight.mul_(10.0)
eng = Engine(model, (3, 64, 64), B)
g = torch.Generator().manual_seed(5)
x = torch.randn(B, 3, 64, 64, generator=g).cuda()
y = torch.randint(0, 20, (B,), generator=g).cuda()
model.eval()
for mode in (0, 1, 2):
    eng.fwd_loss_bwd(x, y, mode, train=False)
    loss, _ = eng.read_loss_correct()
    logits = eng.logits.clone()
    grads = [eng.view(eng.grad, i).clone() for i in range(len(eng.params))]
    lg, lref, gref, audit = forced_reference(eng, x, y, mode, device="cuda")
    ge = [rel_err(a, b) for a, b in zip(grads, gref)]
    oe = [rel_err(a ** 2, b ** 2) for a, b in zip(grads, gref)]
    print("mode %d: logits %.2e loss %.2e  grad max %.2e (tensor %d)  g^2 max %.2e  flips %d / %d, max flip margin %.1e" % (
        mode, rel_err(logits, lg), abs(loss - lref) / abs(lref), max(ge), ge.index(max(ge)), max(oe), audit["flips"], audit["decisions"],
        audit["max_flip_margin"]), flush=True)

"""Probe the kind::f16 (bf16) operand conventions -- smem K-major rows and the A-through-TMEM packing (csrc/probe/clb_debug.cu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

def _probe_lib():
    """libclb_probe.so (clsurvey_b200/build.py build_probe): bring-up probes, not part of the product ABI."""
    import ctypes
    from clsurvey_b200 import build
    lib = ctypes.CDLL(build.build_probe() if not os.path.exists(os.path.join(build.OUT_DIR, "libclb_probe.so")) else os.path.join(build.OUT_DIR, "libclb_probe.so"))
    return lib


def _probe_call(name, *args):
    import ctypes
    fn = getattr(_probe_lib(), name)
    fn.restype = ctypes.c_int
    rc = fn(*[ctypes.c_void_p(a) if isinstance(a, int) and a > 2 ** 31 else a for a in args])
    assert rc == 0, (name, rc)
from clsurvey_b200 import _capi
_capi.lib()
g = torch.Generator().manual_seed(0)
A = torch.randn(128, 64, generator=g)
B = torch.randn(128, 64, generator=g)
ref = (A.bfloat16().double() @ B.bfloat16().double().T).float()
Ad, Bd = A.cuda(), B.cuda()
for v in (0, 1, 2):
    D = torch.zeros(128, 128, device="cuda")
    _probe_call("clb_debug_umma_bf16", Ad.data_ptr(), Bd.data_ptr(), D.data_ptr(), v, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    err = ((D.cpu() - ref).abs().max() / ref.abs().max()).item()
    print("variant %d: rel err vs bf16-rounded operands %.3e   D[0,:3]=%s ref=%s" % (v, err, D[0, :3].tolist(), ref[0, :3].tolist()), flush=True)

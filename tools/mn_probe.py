"""Probe the MN-major smem descriptor convention (see csrc/probe/clb_debug.cu)."""
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
A = torch.randn(128, 32, generator=g)          # [m][k]
B = torch.randn(128, 32, generator=g)          # [n][k]
ref = (A.double() @ B.double().T).float()
At = A.t().contiguous().cuda()                 # [k][m]  (M contiguous)
Bd = B.cuda()
for v in (4, 0, 1, 2, 3):
    D = torch.zeros(128, 128, device="cuda")
    _probe_call("clb_debug_umma_mn", At.data_ptr(), Bd.data_ptr(), D.data_ptr(), v, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    err = ((D.cpu() - ref).abs().max() / ref.abs().max()).item()
    print("D[0,:4]", D[0, :4].tolist(), "ref", ref[0, :4].tolist())
    print("variant %d (k_group_major=%d swap=%d): rel err %.3e" % (v, v & 1, (v >> 1) & 1, err), flush=True)

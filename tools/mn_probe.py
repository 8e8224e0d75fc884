"""Probe the MN-major smem descriptor convention (see csrc/clb_debug.cu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
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
    _capi.call("clb_debug_umma_mn", At.data_ptr(), Bd.data_ptr(), D.data_ptr(), v, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    err = ((D.cpu() - ref).abs().max() / ref.abs().max()).item()
    print("D[0,:4]", D[0, :4].tolist(), "ref", ref[0, :4].tolist())
    print("variant %d (k_group_major=%d swap=%d): rel err %.3e" % (v, v & 1, (v >> 1) & 1, err), flush=True)

"""Probe the kind::f16 (bf16) operand conventions -- smem K-major rows and the A-through-TMEM packing (csrc/clb_debug.cu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clsurvey_b200 import _capi
_capi.lib()
g = torch.Generator().manual_seed(0)
A = torch.randn(128, 64, generator=g)
B = torch.randn(128, 64, generator=g)
ref = (A.bfloat16().double() @ B.bfloat16().double().T).float()
Ad, Bd = A.cuda(), B.cuda()
for v in (0, 1, 2):
    D = torch.zeros(128, 128, device="cuda")
    _capi.call("clb_debug_umma_bf16", Ad.data_ptr(), Bd.data_ptr(), D.data_ptr(), v, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    err = ((D.cpu() - ref).abs().max() / ref.abs().max()).item()
    print("variant %d: rel err vs bf16-rounded operands %.3e   D[0,:3]=%s ref=%s" % (v, err, D[0, :3].tolist(), ref[0, :3].tolist()), flush=True)
